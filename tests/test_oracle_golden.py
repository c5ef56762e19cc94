"""Pin the CPU oracle (oracle/) to golden vectors produced by the unmodified reference.

The fixtures under tests/golden were written by oracle/make_golden.py, which imports the reference
source from /root/reference/src and runs it on torch-CPU.  None of these tests needs a GPU or the
reference tree.
"""

from __future__ import annotations

import numpy as np
import pytest

from conftest import ulp_diff
from oracle import oracle as O

SIDE_REFS = ("height", "width", "long", "short")


# ------------------------------------------------------------------ stage 1: resize
@pytest.mark.parametrize("hw", [(4, 4), (5, 5), (5, 7), (7, 5), (33, 38)])
def test_resize_exact_sizes_bit_exact(golden, hw):
    g = golden("transforms")
    ref = g[f"resize_exact_{hw[0]}x{hw[1]}"]
    out = O.resize(g["image"], hw, side_ref="height")
    assert out.shape == ref.shape and out.dtype == np.float32
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("size", [16, 31, 46])
@pytest.mark.parametrize("side_ref", SIDE_REFS)
@pytest.mark.parametrize("tr", [False, True])
def test_resize_int_sizes_bit_exact(golden, size, side_ref, tr):
    g = golden("transforms")
    img = g["image"]
    src = np.ascontiguousarray(img.transpose(0, 2, 1)) if tr else img
    ref = g[f"resize_int_{size}_{side_ref}_{'T' if tr else 'N'}"]
    out = O.resize(src, size, side_ref=side_ref)
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_resize_2d_4d(golden):
    g = golden("transforms")
    assert np.array_equal(O.resize(g["image"][0], 16), g["resize_2d_16"])
    assert np.array_equal(O.resize(g["image"][None], 16), g["resize_4d_16"])


def test_bilinear_c_matches_numpy_restatement(golden):
    g = golden("preprocess")
    x = g["images"]
    for oh, ow in [(20, 24), (26, 32), (15, 19), (80, 96)]:
        a = O.bilinear_resize(x, oh, ow)
        b = O.bilinear_resize_numpy(x, oh, ow)
        assert np.array_equal(a, b)


def test_bilinear_nhwc_equals_nchw(golden):
    x = golden("preprocess")["images"]
    nhwc = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    assert np.array_equal(O.bilinear_resize(x, 26, 32), O.bilinear_resize(nhwc, 26, 32, layout=O.NHWC))


# ------------------------------------------------------------------ stage 1: normalise
def test_normalize_supplied_stats_bit_exact(golden):
    g = golden("transforms")
    img = g["image"][None]
    out = O.normalize_per_channel(img, channel_means=g["norm_supplied_mean"], channel_stds=g["norm_supplied_std"])
    assert np.array_equal(out, g["norm_supplied"])
    out = O.normalize_per_channel(
        img, channel_means=g["norm_supplied_mean"], channel_stds=g["norm_supplied_std"], min_value=-1.0, max_value=1.0
    )
    assert np.array_equal(out, g["norm_supplied_clip1"])


def test_normalize_apply_bit_exact_given_reference_stats(golden):
    """With the reference's own fp32 batch statistics, the apply stage is bit-identical."""
    g = golden("transforms")
    img = g["image"][None]
    out = O.normalize_per_channel(img.astype(np.float32), channel_means=g["norm_mean"], channel_stds=g["norm_std"])
    assert np.array_equal(out, g["norm_f32in"])
    out = O.normalize_per_channel(img, channel_means=g["norm_mean"], channel_stds=g["norm_std"], min_value=-3, max_value=3)
    assert np.array_equal(out, g["norm_u8in_clip3"])


def test_channel_stats_within_ulps_of_reference(golden):
    """The oracle's statistics are the correctly rounded ones; the reference's fp32 reduction is a
    few ulp away from them.  Tolerance written here: 4 ulp."""
    g = golden("transforms")
    m, s = O.channel_stats(g["image"][None].astype(np.float32))
    assert ulp_diff(m, g["norm_mean"]).max() <= 4
    assert ulp_diff(s, g["norm_std"]).max() <= 4
    x = g["image"].astype(np.float64)
    assert np.array_equal(m.ravel(), x.mean(axis=(1, 2)).astype(np.float32))
    assert np.array_equal(s.ravel(), x.std(axis=(1, 2), ddof=1).astype(np.float32))


def test_normalize_computed_stats_close_to_reference(golden):
    """End to end with computed stats: |y - y_ref| <= 1 ulp at the output scale (2^-23 * max(1,|y|))
    plus the few-ulp statistics offset, i.e. 1e-6 absolute on values bounded by 3."""
    g = golden("transforms")
    out = O.normalize_per_channel(g["image"][None], min_value=-3, max_value=3)
    assert np.abs(out - g["norm_u8in_clip3"]).max() <= 1e-6
    out = O.normalize_per_channel(g["image"][None].astype(np.float32))
    ch_mean = out.mean(axis=(0, 2, 3))
    ch_std = out.std(axis=(0, 2, 3), ddof=1)
    assert np.allclose(ch_mean, 0, atol=1e-4) and np.allclose(ch_std, 1, atol=1e-4)  # test_transform.py:14-24


# ------------------------------------------------------------------ stage 1: preprocess
@pytest.mark.parametrize("msl", [640, 32, 19])
def test_preprocess_apply_bit_exact(golden, msl):
    g = golden("preprocess")
    out = O.preprocess(g["images"], max_side_length=msl, channel_means=g[f"pre_msl{msl}_mean"], channel_stds=g[f"pre_msl{msl}_std"])
    assert out.shape == g[f"pre_msl{msl}"].shape
    assert np.array_equal(out, g[f"pre_msl{msl}"])


@pytest.mark.parametrize("msl", [640, 32, 19])
def test_preprocess_computed_stats(golden, msl):
    g = golden("preprocess")
    out = O.preprocess(g["images"], max_side_length=msl)
    assert np.abs(out - g[f"pre_msl{msl}"]).max() <= 1e-6
    nhwc = np.ascontiguousarray(g["images"].transpose(0, 2, 3, 1))
    assert np.array_equal(O.preprocess(nhwc, max_side_length=msl, layout=O.NHWC), out)


def test_preprocess_other_shapes(golden):
    g = golden("preprocess")
    assert np.abs(O.preprocess(g["images2"], max_side_length=24) - g["pre2_msl24"]).max() <= 1e-6
    # low-variance batch: dividing by a small std amplifies the statistics offset
    assert np.abs(O.preprocess(g["images3"]) - g["pre3"]).max() <= 2e-6


# ------------------------------------------------------------------ stage 2
def test_l2_flat_project(golden):
    g = golden("embed_pca")
    l2 = O.l2_normalize(g["eb_fmap"])
    assert np.allclose(l2, g["eb_l2"], rtol=1e-6, atol=1e-8)
    assert np.array_equal(O.flat_vectors(g["eb_l2"]), g["eb_flat"])
    proj = O.pca_transform(g["eb_flat"], g["eb_means"], g["eb_comps"])
    assert np.allclose(proj, g["eb_proj"], rtol=1e-3, atol=1e-6)
    out = O.pipeline_project(g["eb_fmap"], g["eb_means"], g["eb_comps"])
    assert out.shape == g["eb_out_nchw"].shape
    assert np.allclose(out, g["eb_out_nchw"], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name", ["unc", "cor"])
def test_pca_transform_fixtures(golden, name):
    g = golden("embed_pca")
    out = O.pca_transform(g[f"pca_{name}_x"], g[f"pca_{name}_means"], g[f"pca_{name}_comps"])
    assert out.shape[1] == int(g[f"pca_{name}_k"])
    assert np.allclose(out, g[f"pca_{name}_out"], rtol=1e-3, atol=1e-5)


def test_pipeline_predict_step(golden):
    g = golden("embed_pca")
    pre = O.preprocess(g["pipe_images"])
    assert np.abs(pre - g["pipe_pre"]).max() <= 1e-6
    out = O.pipeline_project(g["pipe_fmap"], g["pipe_means"], g["pipe_comps"])
    assert np.allclose(out, g["pipe_out"], rtol=1e-3, atol=1e-6)
    assert np.allclose(O.l2_normalize(g["pipe_fmap"]), g["pipe_l2"], rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------ stage 3 (parity unpinned)
def test_knn_oracle_self_consistency():
    rng = np.random.default_rng(0)
    store = O.bf16_round(rng.standard_normal((3000, 64)).astype(np.float32))
    q = O.bf16_round(rng.standard_normal((17, 64)).astype(np.float32))
    store[100] = store[7]  # exact duplicate → tie broken by lower index
    s, i = O.cosine_knn(store, q, 10, block=512)
    sn = store / np.linalg.norm(store, axis=1, keepdims=True)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    full = qn @ sn.T
    for r in range(q.shape[0]):
        order = np.lexsort((np.arange(full.shape[1]), -full[r]))[:10]
        assert np.allclose(full[r][order], s[r], atol=1e-5)
        assert set(order) == set(i[r]) or np.abs(np.sort(full[r])[-10] - np.sort(full[r])[-11]) < 1e-5
    # blocks of different size give the same answer
    s2, i2 = O.cosine_knn(store, q, 10, block=3000)
    assert np.array_equal(i, i2) and np.allclose(s, s2, atol=1e-6)
    # tie: querying with store[7] returns 7 before 100
    s3, i3 = O.cosine_knn(store, store[7:8], 2)
    assert list(i3[0]) == [7, 100]


def test_knn_small_store_padding_and_merge():
    rng = np.random.default_rng(1)
    store = rng.standard_normal((6, 16)).astype(np.float32)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    s, i = O.cosine_knn(store, q, 10)
    assert (i[:, 6:] == -1).all() and np.isinf(s[:, 6:]).all()
    # merge of two shards equals the global answer
    big = rng.standard_normal((500, 16)).astype(np.float32)
    s_all, i_all = O.cosine_knn(big, q, 5)
    s0, i0 = O.cosine_knn(big[:250], q, 5)
    s1, i1 = O.cosine_knn(big[250:], q, 5, index_base=250)
    sm, im = O.topk_merge(np.stack([s0, s1]), np.stack([i0, i1]), 5)
    assert np.array_equal(im, i_all) and np.array_equal(sm, s_all)


def test_bf16_helpers():
    x = np.array([1.0, 1.00390625, 3.14159, -2.71828, 1e-3, 65504.0], dtype=np.float32)
    r = O.bf16_round(x)
    assert np.array_equal(O.bf16_from_bits(O.bf16_bits(x)), r)
    assert np.all(np.abs(r - x) <= np.abs(x) * 2.0**-8)


# ------------------------------------------------------------------ torch-CPU port (timed baseline)
def test_torch_port_matches_oracle(golden):
    import torch

    from oracle import torch_port as TP

    torch.manual_seed(0)
    g = golden("preprocess")
    for msl in (640, 32):
        out = TP.preprocess(torch.from_numpy(g["images"]), msl).numpy()
        assert np.abs(out - g[f"pre_msl{msl}"]).max() <= 1e-6
    ge = golden("embed_pca")
    out = TP.l2_project(torch.from_numpy(ge["eb_fmap"]), torch.from_numpy(ge["eb_means"]), torch.from_numpy(ge["eb_comps"]))
    assert np.allclose(out.numpy(), ge["eb_out_nchw"], rtol=1e-3, atol=1e-6)
    rng = np.random.default_rng(0)
    store = O.bf16_round(rng.standard_normal((5000, 64)).astype(np.float32))
    q = O.bf16_round(rng.standard_normal((21, 64)).astype(np.float32))
    s, i = TP.cosine_knn(torch.from_numpy(store), torch.from_numpy(q), 10, block=1024)
    rs, ri = O.cosine_knn(store, q, 10)
    assert np.array_equal(i.numpy(), ri) and np.allclose(s.numpy(), rs, atol=1e-5)


@pytest.mark.parametrize("name,mev", [("unc", None), ("cor", None)])
def test_pca_fit_oracle_matches_reference_fit(golden, name, mev):
    """oracle.pca_fit against the reference's own `PCA.fit` outputs on its test fixtures
    (tests/test_models/test_decomposition.py:18-39 data; fixtures from oracle/make_golden.py)."""
    g = golden("embed_pca")
    x = g[f"pca_{name}_x"]
    k_ref = int(g[f"pca_{name}_k"])
    # the fixture was fitted with the min_explained_variance that yields k_ref components: recover it
    expl = g[f"pca_{name}_explained"]
    cum = np.cumsum(expl)
    thr = float((cum[k_ref - 2] + cum[k_ref - 1]) / 2) if k_ref > 1 else float(cum[0] / 2)
    means, comps, explained, k = O.pca_fit(x, min_explained_variance=thr)
    assert k == k_ref
    assert np.allclose(means, g[f"pca_{name}_means"], atol=1e-6)
    assert np.allclose(explained, expl, atol=1e-5)
    ref_c = g[f"pca_{name}_comps"]
    dots = np.abs(np.sum(comps * ref_c, axis=0))  # same directions up to sign
    assert np.all(dots > 1 - 1e-4)


# ---- ROI masks: the reference's own expectations, tests/test_geometry.py:10-52 ----
def test_roi_mask_reference_single_polygon():
    # test_create_roi_mask_single_polygon (and the docstring example, geometry.py:33-43)
    mask = O.create_roi_mask([(0, 0), (4, 0), (4, 3), (0, 3)], (6, 8), (3, 4))
    assert mask.shape == (3, 4) and mask.dtype == np.int64
    assert np.array_equal(mask, np.array([[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0]]))


def test_roi_mask_reference_multiple_polygons():
    # test_create_roi_mask_multiple_polygons
    roi = [[(0, 0), (1, 0), (1, 1), (0, 1)], [(2, 2), (3, 2), (3, 3), (2, 3)]]
    mask = O.create_roi_mask(roi, (4, 4), (2, 2))
    assert np.array_equal(mask, np.array([[1, 0], [0, 1]]))


def test_roi_mask_class_index_holes_and_pool():
    ring = [(2, 2), (14, 2), (14, 14), (2, 14)]
    hole = [(5, 5), (11, 5), (11, 11), (5, 11)]
    mask = O.create_roi_mask({"exterior": ring, "interiors": [hole]}, (16, 16), (8, 8), class_index=7)
    assert set(np.unique(mask)) == {0, 7}
    assert mask[3, 3] == 0 and mask[4, 4] == 0  # cells [6,8]x[6,8] and [8,10]x[8,10] lie inside the hole
    assert mask[1, 1] == 7 and mask[2, 3] == 7 and mask[0, 0] == 0
    fmap = np.arange(2 * 3 * 8 * 8, dtype=np.float32).reshape(2, 3, 8, 8)
    pooled = O.roi_pool(fmap, mask, class_index=7)
    assert pooled.shape == (2, 3)
    assert np.allclose(pooled[1, 2], fmap[1, 2][mask == 7].mean())
    assert np.array_equal(O.roi_pool(fmap, mask, class_index=3), np.zeros((2, 3), dtype=np.float32))
