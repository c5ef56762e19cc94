"""Embedding-store format bridge: the reference's `embeddings.embedding_data` BLOBs
↔ the device-resident bf16 search store.

The reference stores every embedding map as the raw bytes of a float32, C-order `C×H×W` array
(`storage/models.py:94-129`: `Embedding.create` = `tensor.numpy().tobytes()`, `embedding_tensor` =
`np.frombuffer(...).reshape(C, H, W)`), and reads them back in bulk as `N×C×H×W`
(`storage/operations.py:108-144`).  This module decodes/encodes exactly that wire format on the
host and turns stacks of maps into search rows on the GPU (`isx_maps_to_rows_bf16`):

    pool=None    one row per feature-map cell, in `EmbeddingBatch.get_flat_vectors` order
                 (`data.py:112-118`): row = image * H*W + cell
    pool="mean"  one row per image (spatial mean)

SQLite access itself (DatabaseManager, sessions) stays the reference's code and is out of scope.
"""

from __future__ import annotations

from collections.abc import Iterable, Sequence

import numpy as np
import torch
from torch import Tensor

from imagescry_b200 import _lib
from imagescry_b200.search import EmbeddingStore


def decode_embedding_blob(data: bytes, dim: int, height: int, width: int) -> Tensor:
    """`Embedding.embedding_tensor` (`storage/models.py:94-102`): bytes → float32 `C×H×W` tensor."""
    expected = dim * height * width * 4
    if len(data) != expected:
        raise ValueError(f"embedding BLOB has {len(data)} bytes, expected {expected} for {dim}x{height}x{width} float32")
    return torch.from_numpy(np.frombuffer(data, dtype=np.float32).reshape(dim, height, width).copy())


def encode_embedding_blob(embedding: Tensor) -> tuple[bytes, int, int, int]:
    """`Embedding.create` (`storage/models.py:104-129`): float32 `C×H×W` tensor → (bytes, C, H, W)."""
    if embedding.ndim != 3:
        raise ValueError(f"embedding must be C×H×W, got shape {tuple(embedding.shape)}")
    t = embedding.detach().to(device="cpu", dtype=torch.float32).contiguous()
    c, h, w = t.shape
    return t.numpy().tobytes(), c, h, w


def stack_blobs(records: Iterable[tuple[bytes, int, int, int]]) -> Tensor:
    """Bulk read (`storage/operations.py:108-144`): equal-shaped BLOB records → pinned float32
    `N×C×H×W` host tensor ready for one H2D copy."""
    maps = [decode_embedding_blob(*r) for r in records]
    if not maps:
        raise ValueError("no embedding records")
    shape = maps[0].shape
    if any(m.shape != shape for m in maps):
        raise ValueError("all embedding maps must share one C×H×W shape (pad or group them first)")
    out = torch.stack(maps, dim=0)
    return out.pin_memory() if torch.cuda.is_available() else out


def maps_to_rows(maps: Tensor, *, pool: str | None = None) -> Tensor:
    """float32 `N×C×H×W` maps on a CUDA device → bf16 search rows (`(N·H·W)×C`, or `N×C` pooled)."""
    _lib.require_cuda(maps, "maps")
    if maps.ndim != 4:
        raise ValueError(f"maps must be N×C×H×W, got shape {tuple(maps.shape)}")
    if pool not in (None, "mean"):
        raise ValueError(f"Invalid pool: {pool}")
    x = maps.to(torch.float32).contiguous()
    n, c, h, w = x.shape
    rows = torch.empty((n if pool else n * h * w, c), dtype=torch.bfloat16, device=x.device)
    if n:
        with _lib.on_device(x) as stream:
            rc = _lib.load().isx_maps_to_rows_bf16(x.data_ptr(), n, c, h * w, int(pool is not None), rows.data_ptr(), stream)
        _lib.check(rc, "isx_maps_to_rows_bf16")
    return rows


def store_from_maps(maps: Tensor, *, pool: str | None = None, index_base: int = 0, device=None) -> EmbeddingStore:
    """Build a search store from a stack of embedding maps (host or device)."""
    if not maps.is_cuda:
        maps = maps.to(device or "cuda", non_blocking=True)
    return EmbeddingStore(maps_to_rows(maps, pool=pool), index_base=index_base)


def store_from_blobs(records: Iterable[tuple[bytes, int, int, int]], *, pool: str | None = None, index_base: int = 0, device=None) -> EmbeddingStore:
    """BLOB records (embedding_data, embedding_dim, embedding_height, embedding_width) → store."""
    return store_from_maps(stack_blobs(records), pool=pool, index_base=index_base, device=device)


def rows_to_image_cell(rows: Tensor | Sequence[int], cells_per_image: int) -> tuple[Tensor, Tensor]:
    """Search result rows of a per-cell store → (image position in the stack, cell index y*W + x)."""
    r = torch.as_tensor(rows)
    return torch.div(r, cells_per_image, rounding_mode="floor"), r % cells_per_image
