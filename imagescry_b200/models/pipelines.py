"""Embed → project pipeline — drop-in for `imagescry/models/pipelines.py:22-131` (hot part).

`predict_step` (`pipelines.py:63-88`) produces the same `EmbeddingBatch` (B×k×h×w, NHWC memory) as
the reference, but L2-normalisation, flattening, centring, projection and the final layout are one
fused kernel over the backbone's feature map (stage 2).  `pool="mean"` adds the pooled variant
(B×k×1×1).  Writing to the SQLite store (`:91-97`) is outside the hot path and is not rebuilt; pass
the returned tensors to the reference's `Embedding.create` for that.
"""

from __future__ import annotations

import torch
from jaxtyping import jaxtyped
from torch import nn

from imagescry_b200.data import EmbeddingBatch, ImageBatch
from imagescry_b200.models.decomposition import PCA
from imagescry_b200.models.embedding import EmbeddingModule
from imagescry_b200.typechecking import typechecker


class EmbeddingPCAPipeline(nn.Module):
    """Embeds images, then projects the embeddings to a lower-dimensional space with PCA."""

    def __init__(self, *, embedding_model: EmbeddingModule, pca: PCA, pool: str | None = None) -> None:
        super().__init__()
        if not pca.fitted:
            raise ValueError("PCA model must be fitted before it can be used in the pipeline.")
        if pool not in (None, "mean"):
            raise ValueError(f"Invalid pool: {pool}")
        self.embedding_model = embedding_model
        self.pca = pca
        self.pool = pool

    @jaxtyped(typechecker=typechecker)
    def predict_step(self, batch: ImageBatch) -> EmbeddingBatch:
        fmap = self.embedding_model.feature_map(batch)
        out = self.pca.project_feature_map(fmap, pool=self.pool)
        if self.pool is not None:
            out = out.reshape(out.shape[0], out.shape[1], 1, 1)
        return EmbeddingBatch(indices=batch.indices, embeddings=out)

    def predict(self, dataloader) -> list[EmbeddingBatch]:
        """Eval / inference-mode loop over a dataloader (`pipelines.py:99-131` without the DB branch)."""
        self.eval()
        device = next(self.embedding_model.parameters()).device
        with torch.inference_mode():
            return [self.predict_step(batch.to(device)) for batch in dataloader]
