"""Batch carriers of the sift path — mirror of `imagescry/data.py:29-144`.

`ImageBatch` and `EmbeddingBatch` keep the reference's fields, methods and error behaviour (frozen
slotted dataclasses; `ValueError` when the tensors live on different devices).  Datasets, samplers
and the SQLite-backed readers of the reference module are host-side plumbing outside the hot path
and are not rebuilt here.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from jaxtyping import Float, Int64, UInt8, jaxtyped
from torch import Tensor

from imagescry_b200.typechecking import typechecker


@jaxtyped(typechecker=typechecker)
@dataclass(frozen=True, slots=True)
class ImageBatch:
    """Batch of RGB uint8 images and their dataset indices (`data.py:29-76`)."""

    indices: Int64[Tensor, "B"]
    images: UInt8[Tensor, "B 3 H W"]

    def __len__(self) -> int:
        return len(self.indices)

    def __post_init__(self) -> None:
        if self.indices.device != self.images.device:
            raise ValueError(
                "Tensors must be on the same device. "
                f"Got indices on {self.indices.device} and images on {self.images.device}"
            )

    def cpu(self) -> "ImageBatch":
        return self.to("cpu")

    def to(self, device: str | torch.device) -> "ImageBatch":
        return ImageBatch(indices=self.indices.to(device), images=self.images.to(device))

    @property
    def device(self) -> torch.device:
        return self.indices.device


@jaxtyped(typechecker=typechecker)
@dataclass(frozen=True, slots=True)
class EmbeddingBatch:
    """Batch of image embeddings and their dataset indices (`data.py:79-144`)."""

    indices: Int64[Tensor, "B"]
    embeddings: Float[Tensor, "B E H W"]

    def __len__(self) -> int:
        return len(self.indices)

    def __post_init__(self) -> None:
        if self.indices.device != self.embeddings.device:
            raise ValueError(
                "Tensors must be on the same device. "
                f"Got indices on {self.indices.device} and embeddings on {self.embeddings.device}"
            )

    def cpu(self) -> "EmbeddingBatch":
        return self.to("cpu")

    def get_flat_vectors(self) -> Float[Tensor, "N E"]:
        """Flatten batch and spatial dimensions (`data.py:112-118`): permute(0,2,3,1).reshape(-1, E).
        A pure layout operation (torch view/copy); the fused kernels never need it materialised."""
        return self.embeddings.permute(0, 2, 3, 1).reshape(-1, self.embedding_dim)

    def to(self, device: str | torch.device) -> "EmbeddingBatch":
        return EmbeddingBatch(indices=self.indices.to(device), embeddings=self.embeddings.to(device))

    @property
    def device(self) -> torch.device:
        return self.indices.device

    @property
    def embedding_dim(self) -> int:
        return self.embeddings.size(1)

    @property
    def spatial_dims(self) -> tuple[int, int]:
        return self.embeddings.size(2), self.embeddings.size(3)
