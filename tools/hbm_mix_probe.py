"""What the HBM delivers for streams of different read:write mixes (torch elementwise kernels as neutral
probes): copy 1:1 (the MEASURED_PEAKS figure), pure read, pure write, and the 1:4 mix of the stage-1
apply pass (uint8 in, fp32 out)."""
import json, torch
dev = torch.device("cuda", 0)
n = 3 << 30  # elements
u8 = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
f32 = torch.empty((n,), dtype=torch.float32, device=dev)
u8b = torch.empty_like(u8)
def t(fn, byts, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return byts / (e0.elapsed_time(e1) / iters * 1e-3) / 1e9
out = {
    "copy_f32_1to1_GBps": t(lambda: f32[: n // 2].copy_(f32[n // 2:]), n // 2 * 8),
    "copy_u8_1to1_GBps": t(lambda: u8b.copy_(u8), n * 2),
    "read_only_sum_f32_GBps": t(lambda: f32.sum(), n * 4),
    "write_only_fill_f32_GBps": t(lambda: f32.fill_(1.0), n * 4),
    "u8_to_f32_1to4_GBps": t(lambda: f32.copy_(u8), n * 5),
    "u8_to_bf16_1to2_GBps": t(lambda: f32.view(torch.bfloat16)[:n].copy_(u8), n * 3),
}
print(json.dumps(out))
