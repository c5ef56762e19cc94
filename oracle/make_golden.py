"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference is mounted:

    python oracle/make_golden.py

The reference's source files are imported as they lie under /root/reference/src, behind the import
stubs in oracle/ref_stubs (base classes only — see its README).  Installed torch is 2.11.0+cu128 /
torchvision 0.26 (the reference pins 2.8.0 / 0.23.0, uv.lock:2579-2580,2637-2638); every vector is
therefore "reference source on the container's torch, CPU".

Seeds and shapes follow the reference's own fixtures:
  tests/test_image/conftest.py:22-33       uint8 3×30×45, torch.manual_seed(1234)
  tests/test_image/test_transform.py:32,53 output sizes (4,4) (5,5) (5,7) (7,5) (33,38); 16 31 46 × side_ref
  tests/test_models/test_decomposition.py:18-39  MultivariateNormal fixtures, seed 1234
  tests/test_models/test_embedding.py:29-45      EmbeddingBatch 3×128×7×10
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402
from torch.distributions import MultivariateNormal  # noqa: E402

from imagescry.data import EmbeddingBatch, ImageBatch  # noqa: E402
from imagescry.image.transforms import normalize_per_channel, resize  # noqa: E402
from imagescry.models.decomposition import PCA  # noqa: E402
from imagescry.models.embedding import EfficientNetEmbedder, EmbeddingModule  # noqa: E402
from imagescry.models.pipelines import EmbeddingPCAPipeline  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def npy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().numpy()


def gen_transforms() -> None:
    torch.manual_seed(1234)
    image = torch.randint(low=0, high=256, size=(3, 30, 45), dtype=torch.uint8)
    d: dict[str, np.ndarray] = {"image": npy(image)}

    # normalize_per_channel exactly as the reference test calls it (float input, 1-image batch)
    x = image.float().unsqueeze(0)
    d["norm_f32in"] = npy(normalize_per_channel(x))
    d["norm_mean"] = npy(x.mean(dim=(0, 2, 3), keepdim=True))
    d["norm_std"] = npy(x.std(dim=(0, 2, 3), keepdim=True))
    # uint8 input + clip (the preprocess call, embedding.py:165)
    d["norm_u8in_clip3"] = npy(normalize_per_channel(image.unsqueeze(0), min_value=-3, max_value=3))
    # supplied statistics (docstring example, transforms.py:50-56)
    m = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1) * 255
    s = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1) * 255
    d["norm_supplied_mean"], d["norm_supplied_std"] = npy(m), npy(s)
    d["norm_supplied"] = npy(normalize_per_channel(image.unsqueeze(0), channel_means=m, channel_stds=s))
    d["norm_supplied_clip1"] = npy(
        normalize_per_channel(image.unsqueeze(0), channel_means=m, channel_stds=s, min_value=-1.0, max_value=1.0)
    )

    # resize: exact sizes
    for hw in [(4, 4), (5, 5), (5, 7), (7, 5), (33, 38)]:
        d[f"resize_exact_{hw[0]}x{hw[1]}"] = npy(resize(image, output_size=hw, side_ref="height"))
    # resize: integer size × side_ref × transposed input
    for size in (16, 31, 46):
        for side_ref in ("height", "width", "long", "short"):
            for tr in (False, True):
                src = image.transpose(1, 2).contiguous() if tr else image
                d[f"resize_int_{size}_{side_ref}_{'T' if tr else 'N'}"] = npy(resize(src, size, side_ref=side_ref))
    # 2-D and 4-D inputs (to_4d, transforms.py:130-164)
    d["resize_2d_16"] = npy(resize(image[0], 16))
    d["resize_4d_16"] = npy(resize(image.unsqueeze(0), 16))
    np.savez_compressed(os.path.join(OUT, "transforms.npz"), **d)


def gen_preprocess() -> None:
    torch.manual_seed(1234)
    d: dict[str, np.ndarray] = {}
    # a batch that is NOT resized (max side 48 <= 640) and the same batch resized (max_side_length=32)
    images = torch.randint(0, 256, (5, 3, 40, 48), dtype=torch.uint8)
    d["images"] = npy(images)
    emb = EfficientNetEmbedder.__new__(EfficientNetEmbedder)  # preprocess only reads max_side_length
    nn.Module.__init__(emb)
    for msl in (640, 32, 19):
        emb.max_side_length = msl
        out = EfficientNetEmbedder.preprocess(emb, images)
        d[f"pre_msl{msl}"] = npy(out)
        x = images
        if max(x.shape[-2:]) > msl:
            x = resize(x, msl, side_ref="long")
        x = x.float()
        d[f"pre_msl{msl}_mean"] = npy(x.mean(dim=(0, 2, 3), keepdim=True))
        d[f"pre_msl{msl}_std"] = npy(x.std(dim=(0, 2, 3), keepdim=True))
    # portrait tiles, odd sizes
    images2 = torch.randint(0, 256, (2, 3, 61, 37), dtype=torch.uint8)
    d["images2"] = npy(images2)
    emb.max_side_length = 24
    d["pre2_msl24"] = npy(EfficientNetEmbedder.preprocess(emb, images2))
    # low-variance tile batch (statistics robustness)
    images3 = (torch.randint(0, 3, (3, 3, 16, 16)) + 200).to(torch.uint8)
    d["images3"] = npy(images3)
    emb.max_side_length = 640
    d["pre3"] = npy(EfficientNetEmbedder.preprocess(emb, images3))
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **d)


class _TinyEmbedder(EmbeddingModule):
    """A reference `EmbeddingModule` subclass with a small deterministic backbone so that
    `predict_step` / the pipeline run end to end without EfficientNet's 20 M parameters."""

    def __init__(self, dim: int = 64) -> None:
        super().__init__()
        self._dim = dim
        self.net = nn.Sequential(nn.Conv2d(3, dim, 8, stride=8), nn.SiLU())

    def preprocess(self, images):  # same body as EfficientNetEmbedder.preprocess with msl=640
        return normalize_per_channel(images, min_value=-3, max_value=3)

    def forward(self, x):
        return self.net(x)

    @property
    def embedding_dim(self) -> int:
        return self._dim


def gen_embed_pca() -> None:
    d: dict[str, np.ndarray] = {}
    # --- PCA fixtures (test_decomposition.py:18-39)
    locs = torch.tensor([0.0, 1.0, -1.0, 0.0])
    torch.manual_seed(1234)
    unc = MultivariateNormal(loc=locs, covariance_matrix=torch.eye(4)).sample((1000,))
    torch.manual_seed(1234)
    cov = torch.tensor([[1.0, 0.5, 0, 0], [0.5, 1.0, 0, 0], [0, 0, 1.0, -0.5], [0, 0, -0.5, 1.0]])
    cor = MultivariateNormal(loc=locs, covariance_matrix=cov).sample((1000,))
    for name, x, mev in (("unc", unc, 0.6), ("cor", cor, 0.8)):
        pca = PCA(min_explained_variance=mev).fit(x)
        d[f"pca_{name}_x"] = npy(x)
        d[f"pca_{name}_means"] = npy(pca.feature_means)
        d[f"pca_{name}_comps"] = npy(pca.component_vectors)
        d[f"pca_{name}_explained"] = npy(pca.explained_variance)
        d[f"pca_{name}_out"] = npy(pca.transform(x))
        d[f"pca_{name}_k"] = np.array(pca.num_components)

    # --- EmbeddingBatch 3×128×7×10 (test_embedding.py:29-45): L2-normalise, flatten, project
    torch.manual_seed(1234)
    fmap = torch.randn(3, 128, 7, 10) * 1.7 + 0.3
    emb = nn.functional.normalize(fmap, p=2, dim=1)  # embedding.py:74
    batch = EmbeddingBatch(indices=torch.arange(3), embeddings=emb)
    flat = batch.get_flat_vectors()
    pca = PCA(min_num_components=24, max_num_components=24).fit(flat)
    proj = pca.transform(flat)
    out = proj.reshape(3, 7, 10, pca.num_components).permute(0, 3, 1, 2)  # pipelines.py:82-84
    d["eb_fmap"], d["eb_l2"], d["eb_flat"] = npy(fmap), npy(emb), npy(flat)
    d["eb_means"], d["eb_comps"] = npy(pca.feature_means), npy(pca.component_vectors)
    d["eb_proj"], d["eb_out_nchw"] = npy(proj), npy(out)
    d["eb_out_strides"] = np.array(out.stride())

    # --- full pipeline predict_step with a tiny reference EmbeddingModule
    torch.manual_seed(1234)
    model = _TinyEmbedder(64).eval()
    images = torch.randint(0, 256, (4, 3, 64, 96), dtype=torch.uint8)
    with torch.inference_mode():
        ib = ImageBatch(indices=torch.arange(4), images=images)
        x = model.preprocess(images)
        fm = model.forward(x)
        full = model.predict_step(ib)
        pca2 = PCA(min_num_components=16, max_num_components=16).fit(full.get_flat_vectors().clone())
        pipe = EmbeddingPCAPipeline(embedding_model=model, pca=pca2)
        res = pipe.predict_step(ib)
    d["pipe_images"], d["pipe_pre"], d["pipe_fmap"] = npy(images), npy(x), npy(fm)
    d["pipe_l2"] = npy(full.embeddings)
    d["pipe_means"], d["pipe_comps"] = npy(pca2.feature_means), npy(pca2.component_vectors)
    d["pipe_out"] = npy(res.embeddings)
    d["pipe_conv_w"], d["pipe_conv_b"] = npy(model.net[0].weight), npy(model.net[0].bias)
    np.savez_compressed(os.path.join(OUT, "embed_pca.npz"), **d)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order for the fp32 statistics that get frozen
    gen_transforms()
    gen_preprocess()
    gen_embed_pca()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
