"""GPU parity of the class seams of `models/embedding.py` / `models/pipelines.py` (SURVEY.md §8 a2,
a6, a9) and of patch tiling (north_star stage 1):

* `isx_l2norm_cells` — the stand-alone `F.normalize(x, p=2, dim=1)` of `EmbeddingModule.predict_step`
  (`embedding.py:74`) — against the reference's own outputs (golden `eb_l2`, `pipe_l2`) and the oracle;
* `EfficientNetEmbedder.preprocess` / `preprocess_hwc` / `predict_step` / `embed_images` and
  `EmbeddingPCAPipeline.predict`, instantiated as a user would, with the shape expectations of
  `/root/reference/tests/test_models/test_embedding.py:78-106`;
* `preprocess_patches` against the oracle's restatement (parity unpinned: no reference code)."""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch

from conftest import ulp_diff
from oracle import oracle as O

pytestmark = pytest.mark.gpu

from imagescry_b200.data import EmbeddingBatch, ImageBatch  # noqa: E402
from imagescry_b200.image import transforms as T  # noqa: E402
from imagescry_b200.models.embedding import EfficientNetEmbedder, l2_normalize_cells  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------ a6
def test_l2_normalize_cells_golden(golden):
    """The reference's own `nn.functional.normalize(x, p=2, dim=1)` outputs.  The quotient is the IEEE
    division; the kernel's norm is within an ulp of exact (fp64 fold), torch-CPU's fp32 norm reduction
    carries a few ulp of its own, so the two differ by a few ulp (bar: 1e-3 relative)."""
    g = golden("embed_pca")
    for src, want in (("eb_fmap", "eb_l2"), ("pipe_fmap", "pipe_l2")):
        out = l2_normalize_cells(dev(g[src])).cpu().numpy()
        assert out.shape == g[want].shape
        assert ulp_diff(out, g[want]).max() <= 6, (src, ulp_diff(out, g[want]).max())
        assert np.abs(out - g[want]).max() <= 1e-6 * np.abs(g[want]).max()


@pytest.mark.parametrize(
    "B,E,h,w",
    [
        (3, 128, 7, 10),     # 70 cells: the any-shape kernel
        (2, 1280, 16, 16),   # 512² tiles: TMA slabs
        (5, 1280, 8, 8),     # 256² tiles
        (4, 200, 4, 4),      # E not a multiple of the 256-channel box
        (3, 96, 5, 4),       # hw = 20: a ragged last slab
        (1, 64, 1, 1),
        (2, 1600, 4, 4),     # more channels than the register plan holds: any-shape kernel
    ],
)
def test_l2_normalize_cells_vs_oracle(B, E, h, w):
    rng = np.random.default_rng(B * E + h)
    fmap = (rng.standard_normal((B, E, h, w)) * rng.uniform(0.01, 30, (B, 1, h, w))).astype(np.float32)
    fmap[0, :, 0, 0] = 0.0  # an all-zero cell: 0 / max(0, eps) = 0
    out = l2_normalize_cells(dev(fmap)).cpu().numpy()
    ref = O.l2_normalize(fmap)
    assert np.all(out[0, :, 0, 0] == 0.0)
    assert ulp_diff(out, ref).max() <= 2
    norms = np.sqrt((out.astype(np.float64) ** 2).sum(axis=1))
    norms[0, 0, 0] = 1.0
    assert np.abs(norms - 1.0).max() <= 1e-6
    assert l2_normalize_cells(torch.zeros((0, E, h, w)).cuda()).shape == (0, E, h, w)


# ------------------------------------------------------------------------------------ class seams
def test_embedder_preprocess_matches_reference_fixture(golden):
    """`EfficientNetEmbedder(max_side_length=...).preprocess` as a user calls it, against the
    reference class's own outputs (oracle/make_golden.py ran the unmodified reference)."""
    g = golden("preprocess")
    images = dev(g["images"])
    for msl in (640, 32, 19):
        model = EfficientNetEmbedder(max_side_length=msl).cuda()
        out = model.preprocess(images)
        want = g[f"pre_msl{msl}"]
        assert out.shape == want.shape and out.dtype == torch.float32
        assert np.abs(out.cpu().numpy() - want).max() <= 1e-6  # batch statistics computed (exact-integer sums)
        hwc = images.permute(0, 2, 3, 1).contiguous()
        assert torch.equal(model.preprocess_hwc(hwc), out)
    model = EfficientNetEmbedder(max_side_length=24).cuda()
    assert np.abs(model.preprocess(dev(g["images2"])).cpu().numpy() - g["pre2_msl24"]).max() <= 1e-6
    assert model.embedding_dim == 1280 and model.hparams.max_side_length == 24


@pytest.mark.parametrize("H", [35, 64, 128])
@pytest.mark.parametrize("W", [42, 73, 96])
def test_predict_step_shapes(H, W):
    """/root/reference/tests/test_models/test_embedding.py:78-106: output shape
    (B, 1280, ceil(H/32), ceil(W/32)) for B in {1, 2, 3}; beyond the reference's test the cells must be
    unit vectors (predict_step's F.normalize)."""
    torch.manual_seed(0)
    model = EfficientNetEmbedder().cuda().eval()
    for B in (1, 2, 3):
        batch = ImageBatch(indices=torch.arange(B), images=torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8)).to("cuda")
        with torch.inference_mode():
            out = model.predict_step(batch)
        assert isinstance(out, EmbeddingBatch)
        assert out.embeddings.shape == (B, 1280, math.ceil(H / 32), math.ceil(W / 32))
        assert torch.equal(out.indices, batch.indices)
        norms = out.embeddings.float().norm(dim=1)
        assert torch.allclose(norms, torch.ones_like(norms), atol=1e-5)


def test_embed_images_and_pipeline_predict_over_batches():
    """`embed_images` (embedding.py:78-98) and `EmbeddingPCAPipeline.predict` (pipelines.py:99-131) over
    a list of host batches of different shapes; the pipeline's fused output equals flatten ->
    PCA.transform -> reshape of predict_step's embeddings (pipelines.py:76-84)."""
    from imagescry_b200.models.decomposition import PCA
    from imagescry_b200.models.pipelines import EmbeddingPCAPipeline

    torch.manual_seed(1)
    model = EfficientNetEmbedder(max_side_length=96).cuda()
    gen = torch.Generator().manual_seed(3)
    loader = [
        ImageBatch(indices=torch.arange(4), images=torch.randint(0, 256, (4, 3, 64, 96), dtype=torch.uint8, generator=gen)),
        ImageBatch(indices=torch.arange(4, 6), images=torch.randint(0, 256, (2, 3, 128, 128), dtype=torch.uint8, generator=gen)),
    ]
    embs = model.embed_images(loader)
    assert [tuple(e.embeddings.shape) for e in embs] == [(4, 1280, 2, 3), (2, 1280, 3, 3)]  # 128 -> 96 -> 3x3
    assert all(e.embeddings.is_cuda for e in embs) and not model.training
    flat = torch.cat([e.get_flat_vectors() for e in embs])
    pca = PCA(min_num_components=16, max_num_components=16).cuda().fit(flat)
    pipe = EmbeddingPCAPipeline(embedding_model=model, pca=pca)
    outs = pipe.predict(loader)
    for e, o in zip(embs, outs):
        B, _, h, w = e.embeddings.shape
        want = pca.transform(e.get_flat_vectors()).reshape(B, h, w, 16).permute(0, 3, 1, 2)
        assert o.embeddings.shape == (B, 16, h, w) and o.embeddings.stride() == want.stride()
        assert torch.equal(o.indices, e.indices)
        scale = want.abs().max()
        assert (o.embeddings - want).abs().max() <= 2e-5 * scale
    pooled = EmbeddingPCAPipeline(embedding_model=model, pca=pca, pool="mean").predict(loader)
    assert [tuple(o.embeddings.shape) for o in pooled] == [(4, 16, 1, 1), (2, 16, 1, 1)]
    with pytest.raises(ValueError):
        EmbeddingPCAPipeline(embedding_model=model, pca=PCA().cuda())


# ------------------------------------------------------------------------------------ patch tiling
PATCH_CASES = [
    # (n, H, W, patch, stride, out_hw)
    (2, 64, 96, 32, 32, None),      # exact tiling, no resize
    (1, 100, 130, 32, 16, None),    # overlapping windows, remainder cropped (sampling kernel with scale 1)
    (1, 96, 128, 32, 16, None),     # overlapping windows on 4-pixel-aligned geometry: table-lookup apply
    (2, 70, 90, 24, 24, (12, 12)),  # 2x down-scale of every window
    (1, 128, 160, 64, 48, (37, 37)),
    (3, 40, 40, 40, 40, (16, 16)),  # one window per image
    (1, 96, 64, 20, 9, (30, 30)),   # up-scale
]


@pytest.mark.parametrize("case", PATCH_CASES)
@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
def test_preprocess_patches_vs_oracle(case, layout):
    n, H, W, patch, stride, out_hw = case
    rng = np.random.default_rng(H * W + patch)
    img = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    xin = img if layout == "nhwc" else np.ascontiguousarray(img.transpose(0, 3, 1, 2))
    lay = O.NHWC if layout == "nhwc" else O.NCHW
    windows = O.extract_patches(xin, patch, stride, lay)
    res = O.bilinear_resize(windows, *out_hw, layout=lay) if out_hw else O.to_nchw(windows, lay).astype(np.float32)
    # supplied statistics: bit-exact
    pm = rng.uniform(100, 150, (1, 3, 1, 1)).astype(np.float32)
    ps = rng.uniform(40, 80, (1, 3, 1, 1)).astype(np.float32)
    ref = O.normalize_per_channel(res, channel_means=pm, channel_stds=ps, min_value=-3, max_value=3)
    out = T.preprocess_patches(dev(xin), patch, stride=stride, layout=layout, output_hw=out_hw, channel_means=dev(pm),
                               channel_stds=dev(ps), min_value=-3, max_value=3)
    assert out.shape == ref.shape
    assert np.array_equal(out.cpu().numpy(), ref)
    # batch statistics over all windows: identical to the fused tile path on the materialised windows
    # (the same kernels, the same sums), and to the oracle
    out2 = T.preprocess_patches(dev(xin), patch, stride=stride, layout=layout, output_hw=out_hw, min_value=-3, max_value=3)
    tiles = T.preprocess_tiles(dev(windows), layout=layout, output_hw=out_hw, min_value=-3, max_value=3)
    ref2 = O.normalize_per_channel(res, min_value=-3, max_value=3)
    if out_hw is None:
        assert np.array_equal(out2.cpu().numpy(), ref2)  # exact-integer statistics
        assert torch.equal(out2, tiles)
    else:
        assert np.abs(out2.cpu().numpy() - ref2).max() <= 1e-6
        assert (out2 - tiles).abs().max() <= 1e-6
    out3 = T.preprocess_patches(dev(xin), patch, stride=stride, layout=layout, output_hw=out_hw, channel_means=dev(pm),
                                channel_stds=dev(ps), min_value=-3, max_value=3, out_dtype=torch.bfloat16)
    assert np.array_equal(out3.float().cpu().numpy(), O.bf16_round(ref))


def test_embedder_preprocess_patches():
    """The model-level entry: windows larger than `max_side_length` are resized like any tile."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8)
    model = EfficientNetEmbedder(max_side_length=24).cuda()
    out = model.preprocess_patches(dev(img), 32)
    ref = O.preprocess_patches(img, 32, max_side_length=24, layout=O.NHWC)
    assert out.shape == ref.shape == (2 * 3 * 4, 3, 24, 24)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-6
    out = EfficientNetEmbedder().cuda().preprocess_patches(dev(img), 32, stride=16)
    ref = O.preprocess_patches(img, 32, 16, layout=O.NHWC)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert T.preprocess_patches(dev(img[:0]), 32).shape == (0, 3, 32, 32)  # no images: no windows
    with pytest.raises(ValueError):
        T.preprocess_patches(dev(img), 200)
    with pytest.raises(ValueError):
        T.preprocess_patches(dev(img[..., :2]), 16)


# ------------------------------------------------------------------------------------ HWC ingest (8f.3)
def test_hwc_ingest_equals_direct_preprocess():
    """Decoder-order tiles of mixed shapes -> pinned double-buffered staging -> NHWC kernels: every
    same-shape batch equals `preprocess` of the reference-style NCHW stack of the same tiles (more
    batches than staging slots, so buffers are reused)."""
    from imagescry_b200.ingest import HwcTileIngest, preprocess_hwc_tiles, similar_shape_batches

    rng = np.random.default_rng(8)
    shapes = [(40, 48), (64, 64), (40, 48), (33, 57), (64, 64), (40, 48), (64, 64), (40, 48), (33, 57), (40, 48), (40, 48)]
    tiles = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    tiles[3] = torch.from_numpy(tiles[3])  # torch CPU tensors are accepted as well
    model = EfficientNetEmbedder(max_side_length=48).cuda()
    out = preprocess_hwc_tiles(model, tiles, max_batch_size=2)
    plan = similar_shape_batches(shapes, 2)
    assert [i.tolist() for i, _ in out] == plan and len(plan) > 2
    for idx, pre in out:
        stack = np.stack([np.asarray(tiles[i]) for i in idx.tolist()])          # B H W 3, decoder order
        nchw = torch.from_numpy(np.ascontiguousarray(stack.transpose(0, 3, 1, 2))).cuda()  # what the reference collates
        assert torch.equal(pre, model.preprocess(nchw))
        assert pre.device.type == "cuda" and idx.dtype == torch.int64
    seen = []
    for idx, batch in HwcTileIngest("cuda", 3).batches(tiles):
        assert batch.dtype == torch.uint8 and batch.shape[1:] == (*shapes[int(idx[0])], 3)
        assert np.array_equal(batch.cpu().numpy(), np.stack([np.asarray(tiles[i]) for i in idx.tolist()]))
        seen += idx.tolist()
    assert sorted(seen) == list(range(len(tiles)))
    with pytest.raises(ValueError):
        list(HwcTileIngest("cuda", 3).batches([np.zeros((4, 4), dtype=np.uint8)]))


# ------------------------------------------------------------------------------------ end-to-end sift
def test_sift_end_to_end_small(golden):
    """Config 5 in miniature on one GPU: HWC tiles -> preprocess -> (tiny conv backbone, torch, not owned)
    -> L2 + mean pool + PCA -> all-pairs graph, against the oracle's composition of the same stages."""
    from imagescry_b200.image.transforms import preprocess_tiles
    from imagescry_b200.models.decomposition import PCA
    from imagescry_b200.models.embedding import EmbeddingModule
    from imagescry_b200.sift import sift

    g = golden("embed_pca")

    class Tiny(EmbeddingModule):
        def __init__(self):
            super().__init__()
            self.net = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 8, stride=8), torch.nn.SiLU())

        def preprocess(self, images):
            return preprocess_tiles(images, min_value=-3, max_value=3)

        def preprocess_hwc(self, tiles):
            return preprocess_tiles(tiles, layout="nhwc", min_value=-3, max_value=3)

        def forward(self, x):
            return self.net(x)

        @property
        def embedding_dim(self):
            return 64

    model = Tiny()
    model.net[0].weight.data = torch.from_numpy(g["pipe_conv_w"])
    model.net[0].bias.data = torch.from_numpy(g["pipe_conv_b"])
    model = model.cuda().eval()
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 256, (300, 32, 48, 3), dtype=np.uint8)
    comps = np.linalg.qr(rng.standard_normal((64, 16)))[0].astype(np.float32)
    means = (rng.standard_normal(64) * 0.01).astype(np.float32)
    pca = PCA(num_features=64, num_components=16)
    pca.feature_means.data = torch.from_numpy(means).reshape(1, -1)
    pca.component_vectors.data = torch.from_numpy(comps)
    pca._fitted.data = torch.tensor(True)
    pca._num_features.data = torch.tensor(64)
    pca._num_components.data = torch.tensor(16)
    pca = pca.cuda()
    rows, scores, idx = sift(model, pca, dev(tiles), 5, batch_size=128)
    assert rows.shape == (300, 16) and idx.shape == (300, 5)
    # oracle composition, batch by batch (batch statistics per batch)
    ref_rows = []
    for s in range(0, 300, 128):
        pre = O.preprocess(tiles[s:s + 128], layout=O.NHWC)
        with torch.inference_mode():
            fmap = model(torch.from_numpy(pre).cuda()).cpu().numpy()
        ref_rows.append(O.pipeline_project(fmap, means, comps, pool="mean"))
    ref_rows = np.concatenate(ref_rows)
    assert np.abs(rows.cpu().numpy() - ref_rows).max() <= 1e-4 * np.abs(ref_rows).max()
    # the graph of the rows the product produced (bf16 store of those rows) against the oracle's graph
    from test_gpu_knn import check

    store_rows = O.bf16_round(rows.cpu().numpy())
    check(store_rows, None, 5, scores, idx, graph=True)
