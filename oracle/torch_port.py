"""TEST INFRASTRUCTURE — multi-threaded torch-CPU port of the sift path, used as the timed CPU baseline.

The reference is torch code that runs on the host's cores; timing a single-threaded
numpy oracle next to the GPU would understate it.  This module restates the same three stages with
the torch calls the reference itself makes, so `bench.py`'s `cpu_baseline` / `--impl reference` legs
measure what a user of the reference gets on the GPU box's CPU with all threads:

  preprocess      transforms.py:99-121 (F.interpolate bilinear) + :59-72 (mean/std/(x-m)/(s+eps)/clip)
  l2_project      embedding.py:74 (F.normalize) + data.py:118 (permute/reshape) + decomposition.py:91
  cosine_knn      no reference code: F.normalize → matmul → top-k, ties by index

Only bench.py and tests/ import it; tests/test_oracle_golden.py pins it to the numpy/C oracle.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def preprocess(images_u8: torch.Tensor, max_side_length: int = 640) -> torch.Tensor:
    """NCHW uint8 → resized (iff larger than max_side_length) → normalised, clipped to ±3."""
    x = images_u8
    h, w = x.shape[-2:]
    if max(h, w) > max_side_length:
        scale = max_side_length / max(h, w)
        x = F.interpolate(x.float(), scale_factor=scale, mode="bilinear", align_corners=False, recompute_scale_factor=True)
    x = x.float()
    m = x.mean(dim=(0, 2, 3), keepdim=True)
    s = x.std(dim=(0, 2, 3), keepdim=True)
    return ((x - m) / (s + 1e-6)).clip(-3, 3)


def l2_project(fmap: torch.Tensor, feature_means: torch.Tensor, component_vectors: torch.Tensor, pool: str | None = None) -> torch.Tensor:
    emb = F.normalize(fmap, p=2, dim=1)
    B, E, h, w = emb.shape
    if pool == "mean":
        return torch.matmul(emb.mean(dim=(2, 3)) - feature_means, component_vectors)
    flat = emb.permute(0, 2, 3, 1).reshape(-1, E)
    out = torch.matmul(flat - feature_means, component_vectors)
    return out.reshape(B, h, w, -1).permute(0, 3, 1, 2)


def prepare_store(store: torch.Tensor, block: int = 131072) -> torch.Tensor:
    """Store build (done once, not per query batch): fp32 L2-normalised copy of the store — the
    CPU counterpart of keeping inverse norms next to the device-resident bf16 store."""
    out = torch.empty(store.shape, dtype=torch.float32)
    for s0 in range(0, store.shape[0], block):
        out[s0:s0 + block] = F.normalize(store[s0:s0 + block].float(), dim=1)
    return out


def cosine_knn(store: torch.Tensor, queries: torch.Tensor, k: int, block: int = 131072, prepared: bool = False):
    """fp32 exhaustive cosine top-k with (score desc, index asc) order; store processed in row
    blocks so the Q×N score matrix is never held at once.  `prepared=True`: `store` is already the
    output of `prepare_store`."""
    qn = F.normalize(queries.float(), dim=1)
    best_s = torch.empty((queries.shape[0], 0))
    best_i = torch.empty((queries.shape[0], 0), dtype=torch.int64)
    for s0 in range(0, store.shape[0], block):
        en = store[s0:s0 + block] if prepared else F.normalize(store[s0:s0 + block].float(), dim=1)
        sc = qn @ en.T
        kk = min(k, sc.shape[1])
        # topk then a stable re-sort of the survivors by (score desc, index asc); a tie that straddles
        # the k-th place inside one block may keep a higher index (measure-zero for random data)
        v, i = sc.topk(kk, dim=1)
        cand_s = torch.cat([best_s, v], dim=1)
        cand_i = torch.cat([best_i, i + s0], dim=1)
        order = torch.argsort(cand_i, dim=1, stable=True)
        cand_s, cand_i = cand_s.gather(1, order), cand_i.gather(1, order)
        order = torch.argsort(cand_s, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = cand_s.gather(1, order), cand_i.gather(1, order)
    return best_s, best_i
