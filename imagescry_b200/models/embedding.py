"""Embedding models — drop-in for `imagescry/models/embedding.py` with stage 1 on the sm_100a kernels.

`EmbeddingModule` keeps the reference's extension point (`embedding.py:40-76`): subclasses provide
`preprocess` and `forward`; `predict_step` chains preprocess → forward → per-cell L2 normalisation.
`EfficientNetEmbedder.preprocess` (`:150-165`) is the fused CUDA stage 1.  The backbone's `forward`
(`:168-177`) stays torchvision's EfficientNetV2 `.features` — reported, not owned.
The per-cell L2 step of `predict_step` (`:74`) is the one-pass `isx_l2norm_cells` kernel; the
pipeline (`models/pipelines.py`) fuses it into the projection kernel instead.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Literal

import torch
from jaxtyping import Float, UInt8, jaxtyped
from torch import Tensor, nn

from imagescry_b200 import _lib
from imagescry_b200.data import EmbeddingBatch, ImageBatch
from imagescry_b200.image.transforms import preprocess_patches, preprocess_tiles, resized_shape
from imagescry_b200.models.decomposition import HParamsModule
from imagescry_b200.typechecking import typechecker


def l2_normalize_cells(fmap: Tensor, eps: float = 1e-12) -> Tensor:
    """`nn.functional.normalize(x, p=2, dim=1)` of a B×E×h×w feature map (`embedding.py:74`): every
    spatial cell divided by `max(||cell||_2, eps)` over the channels, in one pass over HBM."""
    _lib.require_cuda(fmap, "fmap")
    if fmap.ndim != 4:
        raise ValueError(f"fmap must be B×E×h×w, got shape {tuple(fmap.shape)}")
    x = fmap.float().contiguous()
    out = torch.empty_like(x)
    B, E, h, w = x.shape
    if x.numel():
        with _lib.on_device(x) as stream:
            rc = _lib.load().isx_l2norm_cells(x.data_ptr(), B, E, h, w, float(eps), out.data_ptr(), stream)
        _lib.check(rc, "isx_l2norm_cells")
    return out


class EmbeddingModule(ABC, HParamsModule):
    """Embedding module interface (`embedding.py:27-106`)."""

    @abstractmethod
    def preprocess(self, images: UInt8[Tensor, "B C H1 W1"]) -> Float[Tensor, "B C H2 W2"]:
        ...  # pragma: no cover

    @abstractmethod
    def forward(self, x: Float[Tensor, "B C H1 W1"]) -> Float[Tensor, "B E H2 W2"]:
        ...  # pragma: no cover

    def feature_map(self, batch: ImageBatch) -> Tensor:
        """preprocess → forward, without the L2 step (input of the fused stage-2 kernel)."""
        return self.forward(self.preprocess(batch.images))

    @jaxtyped(typechecker=typechecker)
    def predict_step(self, batch: ImageBatch) -> EmbeddingBatch:
        """Preprocess, extract the feature map, L2-normalise each cell (`embedding.py:57-76`)."""
        x = self.feature_map(batch)
        x = l2_normalize_cells(x)
        return EmbeddingBatch(indices=batch.indices, embeddings=x)

    def embed_images(self, dataloader) -> list[EmbeddingBatch]:
        """Run `predict_step` over a dataloader in eval / inference mode (what `Trainer.predict`
        does for the reference, `embedding.py:78-98`), moving each batch to the module's device."""
        self.eval()
        device = next(self.parameters()).device
        with torch.inference_mode():
            return [self.predict_step(batch.to(device)) for batch in dataloader]

    @property
    @abstractmethod
    def embedding_dim(self) -> int:
        ...  # pragma: no cover


class EfficientNetEmbedder(EmbeddingModule):
    """EfficientNetV2 backbone feature extractor (`embedding.py:108-182`)."""

    def __init__(
        self, *, backbone_size: Literal["s", "m", "l"] = "s", max_side_length: int = 640, pretrained: bool = False,
        preprocess_dtype: torch.dtype = torch.float32,
    ) -> None:
        super().__init__()
        from torchvision.models import (
            EfficientNet_V2_L_Weights,
            EfficientNet_V2_M_Weights,
            EfficientNet_V2_S_Weights,
            efficientnet_v2_l,
            efficientnet_v2_m,
            efficientnet_v2_s,
        )

        self._embedding_dim = 1_280
        self.save_hyperparameters({"backbone_size": backbone_size, "max_side_length": max_side_length})
        self.backbone_size = backbone_size
        self.max_side_length = max_side_length
        self.preprocess_dtype = preprocess_dtype
        if backbone_size == "s":
            weights, ctor = (EfficientNet_V2_S_Weights.DEFAULT if pretrained else None), efficientnet_v2_s
        elif backbone_size == "m":
            weights, ctor = (EfficientNet_V2_M_Weights.DEFAULT if pretrained else None), efficientnet_v2_m
        elif backbone_size == "l":
            weights, ctor = (EfficientNet_V2_L_Weights.DEFAULT if pretrained else None), efficientnet_v2_l
        else:
            raise ValueError(f"Invalid model size: {backbone_size}")
        self.feature_layers = ctor(weights=weights).features

    @jaxtyped(typechecker=typechecker)
    def preprocess(self, images: UInt8[Tensor, "B C H1 W1"]) -> Float[Tensor, "B C H2 W2"]:
        """Resize (long side → `max_side_length` iff larger) and normalise to [-3, 3]
        (`embedding.py:150-165`) in one fused pass pair over the uint8 tiles."""
        h, w = images.shape[-2:]
        out_hw = resized_shape(h, w, self.max_side_length, "long") if max(h, w) > self.max_side_length else None
        return preprocess_tiles(images, output_hw=out_hw, min_value=-3, max_value=3, out_dtype=self.preprocess_dtype)

    def preprocess_hwc(self, tiles: UInt8[Tensor, "B H W C"]) -> Tensor:
        """Same as `preprocess` for interleaved HWC tiles straight from a decoder (`image/io.py:41-52`
        hands over HWC; the reference permutes on the host)."""
        h, w = tiles.shape[1:3]
        out_hw = resized_shape(h, w, self.max_side_length, "long") if max(h, w) > self.max_side_length else None
        return preprocess_tiles(
            tiles, layout="nhwc", output_hw=out_hw, min_value=-3, max_value=3, out_dtype=self.preprocess_dtype
        )

    def preprocess_patches(
        self, images: UInt8[Tensor, "n H W C"], patch_size: int, *, stride: int | None = None, layout: str = "nhwc"
    ) -> Tensor:
        """Patch tiling + `preprocess` in one fused pass pair: large HWC images are cut into
        `patch_size` windows (stride defaults to the patch size) and every window is treated as a tile
        (resized iff larger than `max_side_length`, normalised with batch statistics, clipped)."""
        out_hw = (
            resized_shape(patch_size, patch_size, self.max_side_length, "long") if patch_size > self.max_side_length else None
        )
        return preprocess_patches(
            images, patch_size, stride=stride, layout=layout, output_hw=out_hw, min_value=-3, max_value=3,
            out_dtype=self.preprocess_dtype,
        )

    @jaxtyped(typechecker=typechecker)
    def forward(self, x: Float[Tensor, "B C H1 W1"]) -> Float[Tensor, "B E H2 W2"]:
        return self.feature_layers.forward(x)

    @property
    def embedding_dim(self) -> int:
        return self._embedding_dim
