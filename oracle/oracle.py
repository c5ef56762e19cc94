"""TEST INFRASTRUCTURE — CPU oracle for the imagescry sift path (numpy + a small C library).

Restates, function by function, what the reference computes on the hot path
(SURVEY.md §8a).  Citations are `/root/reference/src/imagescry/...:line`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module.  Nothing under `imagescry_b200/` does: the product path is CUDA only
and raises when its shared library is missing.

Parity status
-------------
* Stages 1 and 2 (`resize`, `normalize_per_channel`, `preprocess`, `l2_normalize`,
  `flat_vectors`, `pca_transform`, `pipeline_project(pool=None)`): **PINNED** against golden vectors
  produced by running the unmodified reference in the build container
  (`oracle/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).
* Stage 3 (`cosine_knn`, `topk_merge`), `pool="mean"` and patch tiling (`extract_patches`):
  **parity unpinned** — the reference has
  no k-NN, pooling or merge code and no test for them.  The restatement is
  the composition of the reference's own idioms: `F.normalize` (models/embedding.py:74) and
  `torch.matmul` (models/decomposition.py:91), followed by a stable (score desc, index asc) top-k.

Arithmetic that lives in a third-party dependency (torch; pinned 2.8.0 at uv.lock:2579-2580,
2.11.0 installed) is restated from its published algorithm; see each docstring.
"""

from __future__ import annotations

import ctypes
import math
import os
from typing import Literal

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NCHW = 0
NHWC = 1


def _lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libsift_oracle.so")
        if not os.path.exists(path):
            import importlib.util

            spec = importlib.util.spec_from_file_location("_isx_build_oracle", os.path.join(_HERE, "build_oracle.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            path = mod.build()
        _LIB = ctypes.CDLL(path)
        assert _LIB.isxo_abi_version() == 1
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------------------------
# Stage 1 — resize / normalise / preprocess
# ----------------------------------------------------------------------------------------------
def calc_scale_factor(height: int, width: int, output_size: int, side_ref: str) -> float:
    """image/transforms.py:168-197 `_calc_scale_factor` (python double division)."""
    if side_ref == "height":
        return output_size / height
    if side_ref == "width":
        return output_size / width
    if side_ref == "long":
        return output_size / max(height, width)
    if side_ref == "short":
        return output_size / min(height, width)
    raise ValueError(f"Invalid side_ref: {side_ref}")


def resized_shape(height: int, width: int, output_size, side_ref: str = "long") -> tuple[int, int]:
    """Output (H2, W2) of `resize` (transforms.py:106-121).

    Integer `output_size`: `interpolate(scale_factor=s, recompute_scale_factor=True)` →
    `floor(in * s)` per dim in python doubles (torch/nn/functional.py interpolate, the
    `_sym_int(input.size(i + 2) * scale_factors[i])` branch).  Tuple: taken as is.
    """
    if isinstance(output_size, int):
        s = calc_scale_factor(height, width, output_size, side_ref)
        return int(math.floor(height * s)), int(math.floor(width * s))
    return int(output_size[0]), int(output_size[1])


def bilinear_resize(x: np.ndarray, out_h: int, out_w: int, layout: int = NCHW) -> np.ndarray:
    """Bilinear resize, `align_corners=False`, no antialias (transforms.py:112-121 → ATen
    `upsample_bilinear2d`).  `x`: uint8 or float32, 4-D, NCHW or NHWC.  Returns float32 NCHW.
    Implemented in C with explicit `fmaf` (oracle/sift_oracle.c)."""
    assert x.ndim == 4
    x = np.ascontiguousarray(x)
    if layout == NCHW:
        B, C, H, W = x.shape
    else:
        B, H, W, C = x.shape
    out = np.empty((B, C, out_h, out_w), dtype=np.float32)
    if x.dtype == np.uint8:
        _lib().isxo_bilinear_u8(_p(x), layout, B, C, H, W, out_h, out_w, _p(out))
    else:
        x = np.ascontiguousarray(x, dtype=np.float32)
        _lib().isxo_bilinear_f32(_p(x), layout, B, C, H, W, out_h, out_w, _p(out))
    return out


def bilinear_resize_numpy(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Independent numpy restatement of the same bilinear formula (NCHW only), used to cross-check
    the C implementation.  `fmaf(a, b, c)` is emulated as float32(float64(a)*float64(b)+float64(c));
    for image-range operands the fp64 expression is exact, so there is a single rounding."""
    x = np.asarray(x)
    B, C, H, W = x.shape
    xf = x.astype(np.float32)

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    def coords(in_size: int, out_size: int):
        scale = np.float32(in_size) / np.float32(out_size)
        dst = np.arange(out_size, dtype=np.float32)
        src = fma(np.full_like(dst, scale), dst + np.float32(0.5), np.full_like(dst, np.float32(-0.5)))
        src = np.maximum(src, np.float32(0.0)).astype(np.float32)
        i0 = np.minimum(src.astype(np.int64), in_size - 1)
        i1 = np.minimum(i0 + 1, in_size - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        l0 = (np.float32(1.0) - l1).astype(np.float32)
        return i0, i1, l0, l1

    y0, y1, lh0, lh1 = coords(H, out_h)
    x0, x1, lw0, lw1 = coords(W, out_w)
    lh0, lh1 = lh0[None, None, :, None], lh1[None, None, :, None]
    lw0, lw1 = lw0[None, None, None, :], lw1[None, None, None, :]
    r0, r1 = xf[:, :, y0, :], xf[:, :, y1, :]
    p00, p01, p10, p11 = r0[..., x0], r0[..., x1], r1[..., x0], r1[..., x1]

    def w(a, b):
        return np.broadcast_to((a * b).astype(np.float32), p00.shape)

    w00, w01, w10, w11 = w(lh0, lw0), w(lh0, lw1), w(lh1, lw0), w(lh1, lw1)
    acc = fma(w00, p00, (w01 * p01).astype(np.float32))
    acc = fma(w10, p10, acc)
    return fma(w11, p11, acc)


def resize(x: np.ndarray, output_size, *, side_ref: str = "long") -> np.ndarray:
    """image/transforms.py:78-126 `resize`: accepts 2-/3-/4-D (phantom leading dims added by
    `to_4d` :130-164 and squeezed back :124), converts to float, bilinear-resizes."""
    x = np.asarray(x)
    nd = x.ndim
    if nd not in (2, 3, 4):
        raise ValueError(f"Invalid image tensor shape: {x.shape}")
    x4 = x.reshape((1,) * (4 - nd) + x.shape)
    H, W = x4.shape[-2:]
    oh, ow = resized_shape(H, W, output_size, side_ref)
    src = x4 if x4.dtype == np.uint8 else x4.astype(np.float32)
    out = bilinear_resize(src, oh, ow)
    return out.reshape(out.shape[4 - nd:])


def channel_stats(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Per-channel mean and unbiased std over dims (0,2,3) of an NCHW float32 array
    (transforms.py:62-65), correctly rounded to fp32 (two-pass fp64, see sift_oracle.c)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, C, H, W = x.shape
    mean = np.empty(C, dtype=np.float32)
    std = np.empty(C, dtype=np.float32)
    _lib().isxo_channel_stats_f32(_p(x), B, C, H, W, _p(mean), _p(std))
    return mean.reshape(1, C, 1, 1), std.reshape(1, C, 1, 1)


def normalize_per_channel(
    x: np.ndarray,
    *,
    channel_means: np.ndarray | None = None,
    channel_stds: np.ndarray | None = None,
    min_value: float | None = None,
    max_value: float | None = None,
    eps: float = 1e-6,
) -> np.ndarray:
    """image/transforms.py:16-74 `normalize_per_channel` on an NCHW array of any numeric dtype.

    `.float()` (:59); batch statistics over (0,2,3) unless supplied (:62-65); `(x-m)/(s+eps)` (:68)
    in fp32 with one rounding per operation; optional clip (:71-72)."""
    xf = np.ascontiguousarray(np.asarray(x).astype(np.float32))
    B, C, H, W = xf.shape
    if channel_means is None or channel_stds is None:
        m, s = channel_stats(xf)
        channel_means = m if channel_means is None else channel_means
        channel_stds = s if channel_stds is None else channel_stds
    cm = np.asarray(channel_means, dtype=np.float32)
    cs = np.asarray(channel_stds, dtype=np.float32)
    assert cm.shape[1:] == (C, 1, 1) and cs.shape[1:] == (C, 1, 1) and cm.shape[0] in (1, B) and cs.shape[0] in (1, B)
    sb = max(cm.shape[0], cs.shape[0])
    cm = np.ascontiguousarray(np.broadcast_to(cm, (sb, C, 1, 1)))
    cs = np.ascontiguousarray(np.broadcast_to(cs, (sb, C, 1, 1)))
    out = np.empty_like(xf)
    _lib().isxo_normalize_f32(
        _p(xf), B, C, H, W, _p(cm), _p(cs), sb, ctypes.c_float(eps),
        int(min_value is not None), ctypes.c_float(0.0 if min_value is None else min_value),
        int(max_value is not None), ctypes.c_float(0.0 if max_value is None else max_value),
        _p(out),
    )
    return out


def to_nchw(x_u8: np.ndarray, layout: int) -> np.ndarray:
    """HWC→CHW of `pil_to_tensor` (image/io.py:52) + stacking of `_collate_image_batch`
    (data.py:456-459): NHWC uint8 → NCHW uint8."""
    if layout == NCHW:
        return np.ascontiguousarray(x_u8)
    return np.ascontiguousarray(np.transpose(x_u8, (0, 3, 1, 2)))


def preprocess(
    images_u8: np.ndarray,
    *,
    max_side_length: int = 640,
    layout: int = NCHW,
    channel_means: np.ndarray | None = None,
    channel_stds: np.ndarray | None = None,
) -> np.ndarray:
    """models/embedding.py:150-165 `EfficientNetEmbedder.preprocess`: resize iff
    max(h, w) > max_side_length (long side → max_side_length), then normalise and clip to ±3."""
    x = to_nchw(images_u8, layout)
    h, w = x.shape[-2:]
    if max(h, w) > max_side_length:
        x = resize(x, max_side_length, side_ref="long")
    return normalize_per_channel(x, channel_means=channel_means, channel_stds=channel_stds, min_value=-3, max_value=3)


def extract_patches(images_u8: np.ndarray, patch: int, stride: int | None = None, layout: int = NHWC) -> np.ndarray:
    """Patch tiling (parity unpinned: the reference has no patch tiling).
    Restatement: every image is cut into `patch` x `patch` windows whose top-left corners lie on a
    `stride` grid; only full windows are kept (ny = (H - patch) // stride + 1, likewise nx); windows
    are ordered (image, py, px) and returned in the input's layout, i.e. exactly the tile batch
    `_collate_image_batch` (data.py:456-459) would stack had the windows been separate files."""
    stride = patch if stride is None else stride
    x = np.asarray(images_u8)
    if layout == NHWC:
        n, H, W, C = x.shape
    else:
        n, C, H, W = x.shape
    ny, nx = (H - patch) // stride + 1, (W - patch) // stride + 1
    out = []
    for i in range(n):
        for py in range(ny):
            for px in range(nx):
                y0, x0 = py * stride, px * stride
                if layout == NHWC:
                    out.append(x[i, y0:y0 + patch, x0:x0 + patch, :])
                else:
                    out.append(x[i, :, y0:y0 + patch, x0:x0 + patch])
    return np.ascontiguousarray(np.stack(out, axis=0))


def preprocess_patches(
    images_u8: np.ndarray, patch: int, stride: int | None = None, *, max_side_length: int = 640, layout: int = NHWC,
    channel_means: np.ndarray | None = None, channel_stds: np.ndarray | None = None,
) -> np.ndarray:
    """Patch tiling followed by models/embedding.py:150-165 `preprocess` on the window batch
    (statistics over all windows: a pixel covered by several windows counts once per window)."""
    return preprocess(extract_patches(images_u8, patch, stride, layout), max_side_length=max_side_length, layout=layout,
                      channel_means=channel_means, channel_stds=channel_stds)


# ----------------------------------------------------------------------------------------------
# Stage 2 — L2-normalise, flatten, (pool), project
# ----------------------------------------------------------------------------------------------
def l2_normalize(fmap: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """models/embedding.py:74 `F.normalize(x, p=2, dim=1)` on a B×E×h×w map:
    x / max(‖x‖₂, eps) per spatial cell (torch/nn/functional.py normalize)."""
    f = np.ascontiguousarray(fmap, dtype=np.float32)
    B, E, h, w = f.shape
    out = np.empty_like(f)
    _lib().isxo_l2_normalize_cells(_p(f), B, E, h * w, ctypes.c_float(eps), _p(out))
    return out


def flat_vectors(emb: np.ndarray) -> np.ndarray:
    """data.py:112-118 `EmbeddingBatch.get_flat_vectors`: permute(0,2,3,1).reshape(-1, E)."""
    B, E, h, w = emb.shape
    return np.ascontiguousarray(np.transpose(emb, (0, 2, 3, 1))).reshape(-1, E)


def pca_transform(x: np.ndarray, feature_means: np.ndarray, component_vectors: np.ndarray) -> np.ndarray:
    """models/decomposition.py:79-91 `PCA.forward`: matmul(x - feature_means, component_vectors),
    fp32 in / fp32 out.  Accumulated here in fp64 and rounded once (the reference's SGEMM
    accumulation order is library-defined; the tolerance of the comparison is 1e-3 relative)."""
    xc = x.astype(np.float32) - feature_means.astype(np.float32).reshape(1, -1)
    return (xc.astype(np.float64) @ component_vectors.astype(np.float64)).astype(np.float32)


def pipeline_project(
    fmap: np.ndarray,
    feature_means: np.ndarray,
    component_vectors: np.ndarray,
    pool: Literal[None, "mean"] = None,
) -> np.ndarray:
    """models/pipelines.py:75-84 hot part of `EmbeddingPCAPipeline.predict_step`, starting from the
    backbone's feature map: L2-normalise (embedding.py:74) → flatten (data.py:118) → PCA.transform
    (decomposition.py:165) → reshape(B,h,w,k).permute(0,3,1,2).

    pool="mean" (parity unpinned): spatial mean of the normalised map, then the
    same projection → B×k.
    """
    B, E, h, w = fmap.shape
    k = component_vectors.shape[1]
    emb = l2_normalize(fmap)
    if pool == "mean":
        pooled = emb.astype(np.float64).mean(axis=(2, 3)).astype(np.float32)
        return pca_transform(pooled, feature_means, component_vectors)
    proj = pca_transform(flat_vectors(emb), feature_means, component_vectors)
    return np.ascontiguousarray(proj.reshape(B, h, w, k).transpose(0, 3, 1, 2))


# ----------------------------------------------------------------------------------------------
# Stage 3 — cosine k-NN (parity unpinned: no reference implementation)
# ----------------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 → nearest-even bfloat16, returned as float32 (values exactly representable)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 (already bf16-representable or not) → uint16 bf16 bit patterns (round-nearest-even)."""
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


def bf16_from_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def row_rnorm(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """1 / max(‖row‖₂, eps) in fp32 (the denominator of F.normalize, embedding.py:74)."""
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    return (np.float32(1.0) / np.maximum(n, np.float32(eps))).astype(np.float32)


def cosine_scores(store: np.ndarray, queries: np.ndarray, block: int = 65536):
    """Yield (start, scores[Q, block]) of `normalize(q) @ normalize(e).T` in fp32 over store blocks."""
    q = queries.astype(np.float32)
    qn = q * row_rnorm(q)[:, None]
    for s in range(0, store.shape[0], block):
        e = store[s:s + block].astype(np.float32)
        en = e * row_rnorm(e)[:, None]
        yield s, qn @ en.T


def cosine_knn(store: np.ndarray, queries: np.ndarray, k: int, *, index_base: int = 0, block: int = 65536):
    """Exhaustive cosine top-k.  `store` N×d, `queries` Q×d (float32 holding bf16-representable
    values, or any float).  Returns (scores Q×k float32, indices Q×k int64), each row ordered by
    (score descending, index ascending) — a stable lexicographic order, not bare `topk`.
    If N < k the tail is padded with (-inf, -1)."""
    Q = queries.shape[0]
    best_s = np.full((Q, 0), -np.inf, dtype=np.float32)
    best_i = np.zeros((Q, 0), dtype=np.int64)
    for s, sc in cosine_scores(store, queries, block):
        idx = np.broadcast_to(np.arange(s, s + sc.shape[1], dtype=np.int64)[None, :], sc.shape)
        cand_s = np.concatenate([best_s, sc], axis=1)
        cand_i = np.concatenate([best_i, idx + index_base], axis=1)
        best_s, best_i = _select_topk(cand_s, cand_i, k)
    if best_s.shape[1] < k:
        pad = k - best_s.shape[1]
        best_s = np.concatenate([best_s, np.full((Q, pad), -np.inf, np.float32)], axis=1)
        best_i = np.concatenate([best_i, np.full((Q, pad), -1, np.int64)], axis=1)
    return best_s, best_i


def _select_topk(scores: np.ndarray, idx: np.ndarray, k: int):
    """Rows' top-k by (score desc, index asc)."""
    k_eff = min(k, scores.shape[1])
    if scores.shape[1] > 4 * k_eff:
        # cheap pre-selection: keep everything >= the k-th largest score (ties included)
        kth = -np.partition(-scores, k_eff - 1, axis=1)[:, k_eff - 1]
        out_s = np.empty((scores.shape[0], k_eff), np.float32)
        out_i = np.empty((scores.shape[0], k_eff), np.int64)
        for r in range(scores.shape[0]):
            m = scores[r] >= kth[r]
            s_r, i_r = scores[r][m], idx[r][m]
            o = np.lexsort((i_r, -s_r))[:k_eff]
            out_s[r], out_i[r] = s_r[o], i_r[o]
        return out_s, out_i
    order = np.lexsort((idx, -scores), axis=1)[:, :k_eff]
    return np.take_along_axis(scores, order, axis=1), np.take_along_axis(idx, order, axis=1)


def topk_merge(scores: np.ndarray, idx: np.ndarray, k: int):
    """Merge G partial results (G×Q×k each) into Q×k by (score desc, index asc); entries with
    index < 0 are padding and lose to everything."""
    G, Q, kk = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(Q, G * kk).astype(np.float32).copy()
    i = np.transpose(idx, (1, 0, 2)).reshape(Q, G * kk).astype(np.int64).copy()
    pad = i < 0
    s[pad] = -np.inf
    i_sort = np.where(pad, np.iinfo(np.int64).max, i)
    order = np.lexsort((i_sort, -s), axis=1)[:, :k]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)


# ------------------------------------------------------------------------------------------------
# Embedding-store wire format (storage/models.py:94-129) and the rows a search store is built from
# ------------------------------------------------------------------------------------------------
def blob_encode(embedding: np.ndarray) -> bytes:
    """`Embedding.create` (storage/models.py:128): the raw bytes of a float32 C-order C×H×W array."""
    return np.ascontiguousarray(embedding, dtype=np.float32).tobytes()


def blob_decode(data: bytes, dim: int, height: int, width: int) -> np.ndarray:
    """`Embedding.embedding_tensor` (storage/models.py:98-102)."""
    return np.frombuffer(data, dtype=np.float32).reshape(dim, height, width).copy()


def maps_to_rows(maps: np.ndarray, pool: str | None = None) -> np.ndarray:
    """N×C×H×W float32 maps -> bf16-rounded search rows: one per cell in `get_flat_vectors` order
    (data.py:112-118), or the spatial mean per image (pool="mean"; parity unpinned)."""
    n, c, h, w = maps.shape
    if pool == "mean":
        return bf16_round(maps.reshape(n, c, h * w).astype(np.float64).mean(axis=2).astype(np.float32))
    return bf16_round(flat_vectors(maps.astype(np.float32)))


# ------------------------------------------------------------------------------------------------
# PCA.fit (models/decomposition.py:94-148)
# ------------------------------------------------------------------------------------------------
def pca_fit(x: np.ndarray, *, min_num_components: int = 1, max_num_components: int | None = None,
            min_explained_variance: float = 0.0):
    """mean (:116) -> centre (:119) -> SVD (:122) -> eigenvalues s^2/(n-1) (:125) -> explained
    variance ratio (:128) -> component count (:131-137) -> vt[:k].T (:140).
    Returns (feature_means (1,F), component_vectors (F,k), explained_variance (F,), k).  Component
    signs are whatever the SVD returns (the reference does not fix them either)."""
    x = np.asarray(x, dtype=np.float32)
    n, f = x.shape
    means = x.mean(axis=0, keepdims=True, dtype=np.float32)
    xc = x - means
    _, s, vt = np.linalg.svd(xc.astype(np.float64), full_matrices=False)
    if vt.shape[0] < f:
        s = np.concatenate([s, np.zeros(f - s.shape[0])])
    eig = s**2 / (n - 1)
    explained = (eig / eig.sum()).astype(np.float32)
    need = int(np.sum(np.cumsum(explained) < min_explained_variance) + 1)
    k = max(min_num_components, need)
    if max_num_components is not None:
        k = min(max_num_components, k)
    k = min(k, vt.shape[0])
    return means, vt[:k].T.astype(np.float32), explained, k


def knn_graph(store: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """All-pairs similarity graph (BASELINE.json config 5; SURVEY.md §8c `exclude_self=True`): the k
    best neighbours of every store row other than the row itself, (score desc, index asc)."""
    s, i = cosine_knn(store, store, k + 1)
    n = store.shape[0]
    out_s = np.full((n, k), -np.inf, dtype=np.float32)
    out_i = np.full((n, k), -1, dtype=np.int64)
    for r in range(n):
        keep = i[r] != r
        out_s[r], out_i[r] = s[r][keep][:k], i[r][keep][:k]
    return out_s, out_i


# ------------------------------------------------------------------------------------------------
# ROI masks.  /root/reference/src/imagescry/geometry.py:14-65 `create_roi_mask`
# calls rasterio.features.rasterize(..., transform=Affine.scale(w / wf, h / hf), all_touched=True);
# rasterio 1.4 / GDAL are third-party and absent from the reference tree, so this is a restatement —
# PARITY UNPINNED beyond the reference's own expectations (tests/test_geometry.py:10-52 and the
# docstring example at geometry.py:33-43, which tests/test_oracle_golden.py checks): a cell is
# burned when the area of (polygon ∩ cell rectangle) is positive.  Formulated here by clipping every
# ring to the cell (Sutherland-Hodgman) and summing signed areas — deliberately not the edge-crossing
# / centre-parity test the CUDA kernel uses.
# ------------------------------------------------------------------------------------------------
def _polygon_rings(poly) -> list[list[tuple[float, float]]]:
    """Rings of one polygon: a shapely-like object (.exterior.coords / .interiors), a dict with
    'exterior' and optional 'interiors', or a plain vertex sequence."""
    if hasattr(poly, "exterior"):
        rings = [list(poly.exterior.coords)] + [list(r.coords) for r in poly.interiors]
    elif isinstance(poly, dict):
        rings = [list(poly["exterior"])] + [list(r) for r in poly.get("interiors", [])]
    else:
        rings = [list(poly)]
    out = []
    for r in rings:
        pts = [(float(x), float(y)) for x, y in r]
        if len(pts) > 1 and pts[0] == pts[-1]:
            pts = pts[:-1]
        out.append(pts)
    return out


def _clip_ring(ring, x0, y0, x1, y1):
    def clip(pts, inside, intersect):
        res = []
        for a, b in zip(pts, pts[1:] + pts[:1]):
            ia, ib = inside(a), inside(b)
            if ia and ib:
                res.append(b)
            elif ia and not ib:
                res.append(intersect(a, b))
            elif not ia and ib:
                res.append(intersect(a, b))
                res.append(b)
        return res

    def at_x(x):
        return lambda a, b: (x, a[1] + (b[1] - a[1]) * (x - a[0]) / (b[0] - a[0]))

    def at_y(y):
        return lambda a, b: (a[0] + (b[0] - a[0]) * (y - a[1]) / (b[1] - a[1]), y)

    pts = list(ring)
    for inside, inter in ((lambda p: p[0] >= x0, at_x(x0)), (lambda p: p[0] <= x1, at_x(x1)),
                          (lambda p: p[1] >= y0, at_y(y0)), (lambda p: p[1] <= y1, at_y(y1))):
        if not pts:
            break
        pts = clip(pts, inside, inter)
    return pts


def _ring_area(pts) -> float:
    return 0.5 * sum(a[0] * b[1] - b[0] * a[1] for a, b in zip(pts, pts[1:] + pts[:1])) if len(pts) >= 3 else 0.0


def create_roi_mask(roi, original_image_shape, feature_map_shape, class_index: int = 1) -> np.ndarray:
    """geometry.py:14-65: int64 (hf, wf) mask, `class_index` where an ROI polygon overlaps the cell."""
    h, w = original_image_shape
    hf, wf = feature_map_shape
    sx, sy = w / wf, h / hf
    polys = roi if isinstance(roi, (list, tuple)) and roi and not _is_vertex(roi[0]) else [roi]
    mask = np.zeros((hf, wf), dtype=np.int64)
    for poly in polys:
        rings = _polygon_rings(poly)
        for i in range(hf):
            for j in range(wf):
                x0, x1, y0, y1 = j * sx, (j + 1) * sx, i * sy, (i + 1) * sy
                area = 0.0
                for n, ring in enumerate(rings):
                    a = abs(_ring_area(_clip_ring(ring, x0, y0, x1, y1)))
                    area += a if n == 0 else -a
                if area > 1e-9 * sx * sy:
                    mask[i, j] = class_index
    return mask


def _is_vertex(v) -> bool:
    try:
        return len(v) == 2 and all(isinstance(c, (int, float, np.integer, np.floating)) for c in v)
    except TypeError:
        return False


def roi_pool(fmap: np.ndarray, mask: np.ndarray, class_index: int = 1) -> np.ndarray:
    """Mean of the feature-map cells whose mask value is `class_index`: (B, E, h, w) -> (B, E);
    mask (h, w) or (B, h, w).  Zero where no cell matches.  No reference code."""
    fmap = np.asarray(fmap, dtype=np.float32)
    b, e, h, w = fmap.shape
    m = np.broadcast_to(np.asarray(mask) == class_index, (b, h, w)) if np.asarray(mask).ndim == 2 else (np.asarray(mask) == class_index)
    out = np.zeros((b, e), dtype=np.float32)
    for n in range(b):
        cnt = int(m[n].sum())
        if cnt:
            out[n] = (fmap[n][:, m[n]].astype(np.float64).sum(axis=1) / cnt).astype(np.float32)
    return out
