"""End-to-end sift (BASELINE.json config 5): tiles → embed → project → all-pairs similarity graph.

The three stages of the path chained as one call, for one GPU or for the ranks of a
`torch.distributed` group (every rank embeds ITS tiles — preprocessing and projection shard by batch,
no collective — and the graph is built over all ranks' rows with the query-sharded, store-rotating
search of `ShardedEmbeddingStore.knn_graph`).  The backbone forward stays the embedding model's own
torch module (`EmbeddingModule.forward`, reported, not owned).
"""

from __future__ import annotations

import torch
from torch import Tensor

from imagescry_b200.models.decomposition import PCA
from imagescry_b200.models.embedding import EmbeddingModule
from imagescry_b200.search import EmbeddingStore, ShardedEmbeddingStore


def embed_project_tiles(
    model: EmbeddingModule, pca: PCA, tiles: Tensor, *, batch_size: int = 512, layout: str = "nhwc"
) -> Tensor:
    """uint8 tiles on the model's device (`B×H×W×3` decoder order, or `B×3×H×W`) → one projected,
    mean-pooled embedding row per tile (fp32 `B×k`): stage 1 → backbone → fused L2 + pool + projection
    (`pipelines.py:63-88` with `pool="mean"`), batch by batch (batch statistics per batch, as the
    reference's `predict_step` computes them)."""
    if layout not in ("nhwc", "nchw"):
        raise ValueError(f"Invalid layout: {layout}")
    rows = []
    with torch.inference_mode():
        for s in range(0, tiles.shape[0], batch_size):
            batch = tiles[s:s + batch_size]
            x = model.preprocess_hwc(batch) if layout == "nhwc" else model.preprocess(batch)
            rows.append(pca.project_feature_map(model(x), pool="mean"))
    k = pca.component_vectors.shape[1]
    return torch.cat(rows) if rows else torch.empty((0, k), dtype=torch.float32, device=tiles.device)


def similarity_graph(rows: Tensor, k: int, *, group=None, sharded: bool | None = None, gather: bool = True) -> tuple[Tensor, Tensor]:
    """All-pairs cosine k-NN graph over embedding rows: the k nearest OTHER rows of every row.

    Single process: `rows` is the whole store.  With a process group (`sharded=True`, or any initialised
    `torch.distributed` group when `sharded` is None) `rows` are THIS rank's rows of a store that is the
    concatenation of all ranks' rows in rank order (rank r must hold rows `shard_range(total, G, r)`;
    ranks holding equal counts always do).  Returns (scores fp32, global indices int64) for all rows
    (`gather=True`) or for this rank's rows."""
    import torch.distributed as dist

    if sharded is None:
        sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not sharded:
        return EmbeddingStore(rows).knn_graph(k)
    counts = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    store = ShardedEmbeddingStore(rows, total_rows=int(counts.item()), group=group)
    return store.knn_graph(k, gather=gather)


def sift(model: EmbeddingModule, pca: PCA, tiles: Tensor, k: int, *, batch_size: int = 512, layout: str = "nhwc",
         group=None) -> tuple[Tensor, Tensor, Tensor]:
    """tiles → (embedding rows fp32 `B×k_pca`, graph scores, graph indices): `embed_project_tiles`
    followed by `similarity_graph` over every rank's rows."""
    rows = embed_project_tiles(model, pca, tiles, batch_size=batch_size, layout=layout)
    scores, idx = similarity_graph(rows, k, group=group)
    return rows, scores, idx
