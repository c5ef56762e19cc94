"""Exhaustive cosine k-NN over a bf16 embedding store (stage 3 of the sift path).

The reference has no search entry point; this module adds one on top of what its
store holds (`storage/models.py:94-129`: float32 C×H×W BLOBs → rows of a bf16 matrix).  Semantics:
`normalize(q) @ normalize(e).T` with `F.normalize`'s eps (`models/embedding.py:74`), top-k ordered by
(score descending, index ascending).

Single GPU:   EmbeddingStore(embeddings).search(queries, k)
Several GPUs: ShardedEmbeddingStore — rows are split contiguously over the ranks of a
              `torch.distributed` group (one process per GPU); every rank searches its shard with the
              fused tcgen05 kernel, one all-gather of (score, index) pairs follows, and every rank
              merges the gathered lists.  Preprocessing and projection shard by batch and need no
              collective.
"""

from __future__ import annotations

import torch
from torch import Tensor

from imagescry_b200 import _lib

MAX_K = 128


def shard_range(num_rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [begin, end) owned by `rank`: sizes differ by at most one row."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    base, rem = divmod(num_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def row_rnorm(x: Tensor, eps: float = 1e-12) -> Tensor:
    """fp32 `1 / max(||row||, eps)` of a bf16 matrix (the denominator of `F.normalize`)."""
    _lib.require_cuda(x, "x")
    if x.dtype != torch.bfloat16 or x.ndim != 2 or not x.is_contiguous():
        raise ValueError("row_rnorm expects a contiguous 2-D bfloat16 tensor")
    n, d = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    if n:
        rc = _lib.load().isx_row_rnorm_bf16(x.data_ptr(), n, d, eps, out.data_ptr(), _lib.stream_ptr(x.device))
        _lib.check(rc, "isx_row_rnorm_bf16")
    return out


def merge_topk(scores: Tensor, indices: Tensor, k: int | None = None) -> tuple[Tensor, Tensor]:
    """Merge G partial results (G×Q×k fp32 scores, G×Q×k int32 indices; index < 0 = padding) into
    Q×k, ordered by (score desc, index asc)."""
    _lib.require_cuda(scores, "scores")
    if scores.shape != indices.shape or scores.ndim != 3:
        raise ValueError("scores and indices must both be G×Q×k")
    g, q, kk = scores.shape
    k = kk if k is None else k
    if k != kk:
        raise ValueError(f"k={k} does not match the partial lists' width {kk}")
    s = scores.contiguous().float()
    i = indices.contiguous().to(torch.int32)
    out_s = torch.empty((q, k), dtype=torch.float32, device=s.device)
    out_i = torch.empty((q, k), dtype=torch.int32, device=s.device)
    if q:
        rc = _lib.load().isx_topk_merge(
            s.data_ptr(), i.data_ptr(), g, q, k, out_s.data_ptr(), out_i.data_ptr(), _lib.stream_ptr(s.device)
        )
        _lib.check(rc, "isx_topk_merge")
    return out_s, out_i


def _as_bf16_matrix(x: Tensor, name: str) -> Tensor:
    _lib.require_cuda(x, name)
    if x.ndim != 2:
        raise ValueError(f"{name} must be 2-D (rows × features), got shape {tuple(x.shape)}")
    if not x.dtype.is_floating_point:
        raise ValueError(f"{name} must be a floating-point tensor, got {x.dtype}")
    return x.to(torch.bfloat16).contiguous()


class EmbeddingStore:
    """A device-resident N×d bf16 embedding matrix with its inverse row norms.

    Vectors are kept as given (not re-rounded after normalisation): the kernel multiplies bf16×bf16
    products exactly, accumulates in fp32 and applies both inverse norms in the epilogue."""

    def __init__(self, embeddings: Tensor, *, index_base: int = 0) -> None:
        self.embeddings = _as_bf16_matrix(embeddings, "embeddings")
        self.index_base = int(index_base)
        self.rnorm = row_rnorm(self.embeddings)
        self._workspace: Tensor | None = None

    def __len__(self) -> int:
        return self.embeddings.shape[0]

    @property
    def dim(self) -> int:
        return self.embeddings.shape[1]

    @property
    def device(self) -> torch.device:
        return self.embeddings.device

    def _ws(self, nbytes: int) -> Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    def search_raw(self, queries: Tensor, k: int, *, query_rnorm: Tensor | None = None) -> tuple[Tensor, Tensor]:
        """Local top-k: fp32 scores Q×k and int32 global indices Q×k (padding: -inf / -1)."""
        q = _as_bf16_matrix(queries, "queries")
        if q.shape[1] != self.dim:
            raise ValueError(f"queries have {q.shape[1]} features, the store has {self.dim}")
        if not (1 <= k <= MAX_K):
            raise ValueError(f"k must be between 1 and {MAX_K}, got {k}")
        if q.device != self.device:
            raise ValueError(f"queries are on {q.device}, the store is on {self.device}")
        nq = q.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        idx = torch.empty((nq, k), dtype=torch.int32, device=self.device)
        if nq == 0:
            return scores, idx
        qr = row_rnorm(q) if query_rnorm is None else query_rnorm
        lib = _lib.load()
        n = len(self)
        ws_bytes = max(int(lib.isx_knn_workspace_bytes(n, nq, self.dim, k)), 256)
        ws = self._ws(ws_bytes)
        rc = lib.isx_knn_search(
            self.embeddings.data_ptr(), self.rnorm.data_ptr(), n, q.data_ptr(), qr.data_ptr(), nq, self.dim, k,
            self.index_base, scores.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(),
            _lib.stream_ptr(self.device),
        )
        _lib.check(rc, "isx_knn_search")
        return scores, idx

    def search(self, queries: Tensor, k: int) -> tuple[Tensor, Tensor]:
        """Cosine top-k of every query row: (scores fp32 Q×k, indices int64 Q×k)."""
        scores, idx = self.search_raw(queries, k)
        return scores, idx.to(torch.int64)


def drop_self_matches(scores: Tensor, indices: Tensor, query_rows: Tensor, k: int) -> tuple[Tensor, Tensor]:
    """From Q×(k+1) results whose queries are store rows `query_rows` (global indices), remove each
    query's own row and keep the first k of the rest (order preserved).  If the row itself is not
    among the k+1 hits (k+1 exact duplicates with lower indices), the last hit is dropped."""
    q, k1 = indices.shape
    if k1 != k + 1:
        raise ValueError(f"expected k + 1 = {k + 1} columns, got {k1}")
    is_self = indices == query_rows.to(indices.dtype).reshape(-1, 1)
    # position of the column to drop: the self hit, else the last column
    drop = torch.where(is_self.any(dim=1), is_self.to(torch.int8).argmax(dim=1), torch.full((q,), k, device=indices.device))
    cols = torch.arange(k, device=indices.device).reshape(1, -1)
    take = cols + (cols >= drop.reshape(-1, 1)).to(cols.dtype)
    return scores.gather(1, take), indices.gather(1, take)


def knn_graph(store: "EmbeddingStore", k: int, *, block: int = 131072) -> tuple[Tensor, Tensor]:
    """All-pairs similarity graph over a store (BASELINE.json config 5): the k nearest other rows of
    every row.  Queries are the store's own rows, searched in blocks with k + 1 and the self match
    removed.  Returns (scores N×k fp32, indices N×k int64)."""
    n = len(store)
    out_s = torch.empty((n, k), dtype=torch.float32, device=store.device)
    out_i = torch.empty((n, k), dtype=torch.int64, device=store.device)
    for b in range(0, n, block):
        e = min(n, b + block)
        s, i = store.search_raw(store.embeddings[b:e], k + 1, query_rnorm=store.rnorm[b:e])
        rows = torch.arange(b, e, device=store.device) + store.index_base
        s, i = drop_self_matches(s, i.to(torch.int64), rows, k)
        out_s[b:e], out_i[b:e] = s, i
    return out_s, out_i


def knn_search(embeddings: Tensor, queries: Tensor, k: int) -> tuple[Tensor, Tensor]:
    """One-shot convenience: build an `EmbeddingStore` and search it."""
    return EmbeddingStore(embeddings).search(queries, k)


def gather_partials(scores: Tensor, idx: Tensor, group=None) -> tuple[Tensor, Tensor]:
    """All-gather every rank's local top-k: returns (G×Q×k scores, G×Q×k indices).  The only
    collective of the sift path; backend-agnostic (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    q, k = scores.shape
    # concatenation along dim 0 is the one output form every backend (NCCL, gloo) accepts
    all_s = torch.empty((world * q, k), dtype=scores.dtype, device=scores.device)
    all_i = torch.empty((world * q, k), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(all_s, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, idx.contiguous(), group=group)
    return all_s.view(world, q, k), all_i.view(world, q, k)


class ShardedEmbeddingStore:
    """Row-sharded store over the ranks of a process group (one process per GPU).

    `local_embeddings` are this rank's rows; their global indices start at `index_base` (by default
    the contiguous partition of `shard_range`).  Queries are replicated on every rank."""

    def __init__(self, local_embeddings: Tensor, *, total_rows: int | None = None, index_base: int | None = None, group=None) -> None:
        import torch.distributed as dist

        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if index_base is None:
            if total_rows is None:
                raise ValueError("give either index_base or total_rows")
            begin, end = shard_range(total_rows, self.world_size, self.rank)
            if end - begin != local_embeddings.shape[0]:
                raise ValueError(
                    f"rank {self.rank} should hold rows [{begin}, {end}) but got {local_embeddings.shape[0]} rows"
                )
            index_base = begin
        self.total_rows = total_rows
        self.local = EmbeddingStore(local_embeddings, index_base=index_base)

    def search(self, queries: Tensor, k: int) -> tuple[Tensor, Tensor]:
        scores, idx = self.local.search_raw(queries, k)
        all_s, all_i = gather_partials(scores, idx, self.group)
        s, i = merge_topk(all_s, all_i, k)
        return s, i.to(torch.int64)

    def replicated_rows(self) -> Tensor:
        """The whole store on every rank (bf16 N x d): one all-gather of the shards over NVLink.
        BASELINE.json config 5 replicates the 1 M x 256 store (512 MB) as the all-pairs queries.
        Shards may differ by one row (`shard_range`); they are padded to the longest for the gather."""
        import torch.distributed as dist

        sizes = [e - b for b, e in (shard_range(self.total_rows, self.world_size, r) for r in range(self.world_size))]
        per = max(sizes)
        local = self.local.embeddings
        d = local.shape[1]
        padded = torch.zeros((per, d), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        out = torch.empty((per * self.world_size, d), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=self.group)
        if all(sz == per for sz in sizes):
            return out
        return torch.cat([out[r * per: r * per + sizes[r]] for r in range(self.world_size)], dim=0)

    def knn_graph(self, k: int, *, block: int = 131072) -> tuple[Tensor, Tensor]:
        """All-pairs similarity graph over the sharded store (BASELINE.json config 5): every rank
        searches ALL rows (replicated once over NVLink) against its shard in blocks with k + 1,
        one all-gather of (score, index) + merge per block, self matches removed.  Every rank
        returns the full graph: (scores N x k fp32, indices N x k int64)."""
        if self.total_rows is None:
            raise ValueError("knn_graph needs a store built with total_rows (the contiguous partition)")
        rows_all = self.replicated_rows()
        n = rows_all.shape[0]
        rn_all = row_rnorm(rows_all)
        dev = rows_all.device
        out_s = torch.empty((n, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((n, k), dtype=torch.int64, device=dev)
        for b in range(0, n, block):
            e = min(n, b + block)
            s, i = self.local.search_raw(rows_all[b:e], k + 1, query_rnorm=rn_all[b:e])
            all_s, all_i = gather_partials(s, i, self.group)
            s, i = merge_topk(all_s, all_i, k + 1)
            s, i = drop_self_matches(s, i.to(torch.int64), torch.arange(b, e, device=dev), k)
            out_s[b:e], out_i[b:e] = s, i
        return out_s, out_i

    def replicate_queries(self, host_queries: Tensor) -> Tensor:
        """Replicated device copy of a (pinned) host query matrix without sending it over PCIe once
        per GPU: every rank uploads only its 1/G slice of the rows and one all-gather over NVLink
        completes the copy.  All ranks must pass the same `host_queries`; the returned tensor is a
        view of a staging buffer that the next call overwrites."""
        import torch.distributed as dist

        q, d = host_queries.shape
        per = (q + self.world_size - 1) // self.world_size
        dev = self.local.device
        staging = getattr(self, "_q_staging", None)
        if staging is None or staging.shape != (per * self.world_size, d) or staging.dtype != host_queries.dtype:
            staging = torch.empty((per * self.world_size, d), dtype=host_queries.dtype, device=dev)
            self._q_staging = staging
        b, e = min(q, self.rank * per), min(q, (self.rank + 1) * per)
        mine = staging[self.rank * per:self.rank * per + (e - b)]
        mine.copy_(host_queries[b:e], non_blocking=True)
        dist.all_gather_into_tensor(staging, staging[self.rank * per:(self.rank + 1) * per], group=self.group)
        return staging[:q]
