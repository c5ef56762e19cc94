"""Where does the streamed end-to-end path lose time against the device-only loop?  Prints the
device-side interval between consecutive searches for (a) a plain loop on resident queries,
(b) search_host_batches, (c) the same with the uploads switched off (batches already on the device)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagescry_b200.search import EmbeddingStore, HostBatchSearch

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n, q, d, k, steps = 1_000_000, 10_000, 1280, 10, 30
store = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for s in range(0, n, 1 << 20):
    e = min(n, s + (1 << 20))
    store[s:e] = torch.randn((e - s, d), generator=g, device=dev).to(torch.bfloat16)
st = EmbeddingStore(store)
qd = torch.randn((q, d), generator=g, device=dev).to(torch.bfloat16)
qh = qd.cpu().pin_memory()

STREAMER = HostBatchSearch(st, k)

def heat():
    t0 = time.time()
    while time.time() - t0 < 2.0:
        st.search_raw(qd, k)
    torch.cuda.synchronize()

def loop_resident():
    evs = []
    for _ in range(steps):
        st.search_raw(qd, k)
        e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    torch.cuda.synchronize()
    return evs

def loop_stream():
    evs = []
    for _ in STREAMER.run(qh for _ in range(steps)):
        e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    torch.cuda.synchronize()
    return evs

for name, fn in (("resident", loop_resident), ("stream", loop_stream), ("resident", loop_resident), ("stream", loop_stream)):
    heat()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); evs = fn(); t1.record(); torch.cuda.synchronize()
    d_ = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
    print(f"{name}: total/step {t0.elapsed_time(t1) / steps:.3f} ms; between events median {sorted(d_)[len(d_)//2]:.3f} min {min(d_):.3f} max {max(d_):.3f}")
