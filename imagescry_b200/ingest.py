"""HWC-decode ingest: decoder output → device-resident NHWC tile batches.

The reference decodes on the host (`image/io.py:41-52`: PIL → `pil_to_tensor`, an HWC→CHW permute),
groups same-shape images with `SimilarShapeBatcher` (`data.py:403-452`), stacks every group into a
contiguous NCHW uint8 `ImageBatch` (`_collate_image_batch`, `data.py:456-459`) and lets Lightning copy
that batch to the device.  Here the tiles stay in the decoder's own HWC order: they are grouped by
shape with the reference's rule, packed into pinned staging buffers, sent with ONE asynchronous
host→device copy per batch on a copy stream (double-buffered, so the copy of batch n + 1 overlaps the
kernels of batch n), and the stage-1 kernels read the interleaved bytes directly
(`EfficientNetEmbedder.preprocess_hwc`, `layout = ISX_LAYOUT_NHWC`): the permute and the `torch.stack`
copy of the reference never happen.

Decoding itself (JPEG/PNG codecs) is out of scope; any decoder that yields `H×W×3` uint8 arrays feeds
this module.
"""

from __future__ import annotations

from collections.abc import Iterable, Iterator, Sequence

import numpy as np
import torch
from torch import Tensor


def similar_shape_batches(shapes: Iterable[tuple[int, ...]], max_batch_size: int) -> list[list[int]]:
    """Index batches of same-shape items — the batches `SimilarShapeBatcher` yields (`data.py:403-452`):
    items are indexed, sorted by shape (stable, so indices ascend inside a shape group), grouped by
    equal shape, and every group is cut into chunks of at most `max_batch_size`."""
    if max_batch_size < 1:
        raise ValueError(f"max_batch_size must be positive, got {max_batch_size}")
    order = sorted(enumerate(tuple(s) for s in shapes), key=lambda t: t[1])
    batches: list[list[int]] = []
    current: list[int] = []
    current_shape = None
    for idx, shape in order:
        if shape != current_shape or len(current) == max_batch_size:
            if current:
                batches.append(current)
            current, current_shape = [], shape
        current.append(idx)
    if current:
        batches.append(current)
    return batches


class HwcTileIngest:
    """Pinned, double-buffered host→device path for HWC uint8 tiles of mixed shapes.

    `batches(tiles)` yields `(indices, device_batch)` pairs: `indices` are the positions of the batch's
    tiles in `tiles` (int64, on the device, what `ImageBatch.indices` carries) and `device_batch` is a
    `B×H×W×3` uint8 tensor on `device`.  A yielded batch stays valid until the next-but-one batch is
    requested (its device buffer is then reused); consume it — e.g. run `preprocess_hwc` — before that.
    """

    def __init__(self, device: torch.device | str, max_batch_size: int, *, depth: int = 2) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"HwcTileIngest stages tiles for a CUDA device (got {self.device}); there is no CPU path")
        if depth < 2:
            raise ValueError("depth must be at least 2 (one buffer in flight, one being filled)")
        self.max_batch_size = int(max_batch_size)
        self.depth = depth
        self._copy_stream = torch.cuda.Stream(self.device)
        self._host: list[Tensor | None] = [None] * depth
        self._dev: list[Tensor | None] = [None] * depth
        self._copied: list[torch.cuda.Event | None] = [None] * depth   # H2D of the slot finished
        self._consumed: list[torch.cuda.Event | None] = [None] * depth  # consumer kernels on the slot finished

    def _slot_buffers(self, slot: int, nbytes: int) -> tuple[Tensor, Tensor]:
        host, dev = self._host[slot], self._dev[slot]
        if host is None or host.numel() < nbytes:
            host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            dev = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._host[slot], self._dev[slot] = host, dev
        return host, dev

    def batches(self, tiles: Sequence[np.ndarray | Tensor]) -> Iterator[tuple[Tensor, Tensor]]:
        shapes = []
        for t in tiles:
            if t.ndim != 3 or t.shape[2] != 3 or str(t.dtype).replace("torch.", "") != "uint8":
                raise ValueError(f"tiles must be H×W×3 uint8 arrays, got {t.dtype} with shape {tuple(t.shape)}")
            shapes.append((int(t.shape[0]), int(t.shape[1])))
        plan = similar_shape_batches(shapes, self.max_batch_size)
        consumer = torch.cuda.current_stream(self.device)
        for n, idxs in enumerate(plan):
            slot = n % self.depth
            h, w = shapes[idxs[0]]
            nbytes = len(idxs) * h * w * 3
            # the slot's previous copy must have left the pinned buffer, and the kernels that read its
            # device buffer must be done, before either is overwritten
            if self._copied[slot] is not None:
                self._copied[slot].synchronize()
            host, dev = self._slot_buffers(slot, nbytes)
            stage = host[:nbytes].view(len(idxs), h, w, 3)
            for j, i in enumerate(idxs):
                src = tiles[i]
                stage[j].copy_(src if isinstance(src, Tensor) else torch.from_numpy(np.ascontiguousarray(src)))
            with torch.cuda.stream(self._copy_stream):
                if self._consumed[slot] is not None:
                    self._copy_stream.wait_event(self._consumed[slot])
                dev[:nbytes].copy_(host[:nbytes], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                self._copied[slot] = ev
            consumer.wait_event(ev)
            batch = dev[:nbytes].view(len(idxs), h, w, 3)
            yield torch.tensor(idxs, dtype=torch.int64, device=self.device), batch
            done = torch.cuda.Event()
            done.record(consumer)  # everything the caller enqueued on the batch so far
            self._consumed[slot] = done


def preprocess_hwc_tiles(model, tiles: Sequence[np.ndarray | Tensor], max_batch_size: int = 256) -> list[tuple[Tensor, Tensor]]:
    """Decoder output → preprocessed NCHW batches: `HwcTileIngest` + `model.preprocess_hwc` per same-shape
    batch (statistics per batch, as the reference's `predict_step` computes them).  Returns
    `[(indices, preprocessed B×3×H'×W')]` in batch order."""
    device = next(model.parameters()).device
    ingest = HwcTileIngest(device, max_batch_size)
    return [(idx, model.preprocess_hwc(batch)) for idx, batch in ingest.batches(tiles)]
