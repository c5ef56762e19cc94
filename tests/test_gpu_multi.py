"""GPU parity of the multi-GPU search path: ShardedEmbeddingStore over NCCL (one process per GPU,
row-sharded store, one all-gather of (score, index) + merge) equals the unsharded search and the
oracle.  Needs at least two visible GPUs (`gpurun --gpus 2`); skipped otherwise — the gather / merge
plumbing itself is covered on CPU with gloo in tests/test_cpu_host.py."""

from __future__ import annotations

import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ISX_REPO"])
rank, world = int(os.environ["ISX_RANK"]), int(os.environ["ISX_WORLD"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{os.environ['ISX_PORT']}", rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
from imagescry_b200.search import EmbeddingStore, ShardedEmbeddingStore, shard_range
from oracle import oracle as O
rng = np.random.default_rng(0)
n, q, d = 30011, 300, 128
store = O.bf16_round(rng.standard_normal((n, d)).astype(np.float32))
store[7] = store[29000]  # a tie across shards: the lower index must win
queries = O.bf16_round(rng.standard_normal((q, d)).astype(np.float32))
queries[0] = store[7]
b, e = shard_range(n, world, rank)
sharded = ShardedEmbeddingStore(torch.from_numpy(store[b:e]).cuda(), total_rows=n)
qd = torch.from_numpy(queries).cuda()
for k in (10, 100):
    s, i = sharded.search(qd, k)
    fs, fi = EmbeddingStore(torch.from_numpy(store).cuda()).search(qd, k)
    assert torch.equal(i, fi) and torch.allclose(s, fs, atol=1e-6), (rank, k)
    rs, ri = O.cosine_knn(store, queries, k)
    assert np.abs(s.cpu().numpy() - rs).max() <= 1e-3
    assert (i.cpu().numpy() == ri).mean() > 0.995
    assert i[0, :2].tolist() == [7, 29000]
# all-pairs graph over the sharded store (config 5) == the single-GPU graph
from imagescry_b200.search import knn_graph
gn = 5003
gstore = O.bf16_round(rng.standard_normal((gn, 64)).astype(np.float32))
gstore[11] = gstore[4000]
gb, ge = shard_range(gn, world, rank)
gs, gi = ShardedEmbeddingStore(torch.from_numpy(gstore[gb:ge]).cuda(), total_rows=gn).knn_graph(10, block=2048)
fs, fi = knn_graph(EmbeddingStore(torch.from_numpy(gstore).cuda()), 10, block=2048)
assert torch.equal(gi, fi) and torch.allclose(gs, fs, atol=1e-6), rank
assert not (gi == torch.arange(gn, device=gi.device).reshape(-1, 1)).any()
assert int(gi[11, 0]) == 4000 and int(gi[4000, 0]) == 11
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_store_nccl_equals_single_gpu(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(world):
        env = dict(os.environ, ISX_REPO=REPO, ISX_PORT=str(port), ISX_RANK=str(r), ISX_WORLD=str(world))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o[-3000:]
