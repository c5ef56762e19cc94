"""TEST INFRASTRUCTURE — how the oracle's bilinear FMA association was found.

torch-CPU's `upsample_bilinear2d` (what the reference's `resize` dispatches to,
image/transforms.py:112-121) is compiled with floating-point contraction, so which products are
fused is a property of the build, not of the published formula.  This script enumerates the
candidate associations against the installed torch and prints the number of mismatching outputs for
each; the one with zero mismatches is what oracle/sift_oracle.c and the CUDA kernels implement:

    src = fma(scale, dst + 0.5, -0.5)
    out = fma(w11, p11, fma(w10, p10, fma(w00, p00, w01 * p01))),  wXY = lhX * lwY (rounded)

Run in the build container only:  python oracle/probe_bilinear_forms.py
"""

import itertools

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def coords(n_in, n_out, src_fma):
    scale = f32(n_in) / f32(n_out)
    dst = np.arange(n_out, dtype=f32)
    if src_fma:
        src = fma(np.full_like(dst, scale), dst + f32(0.5), np.full_like(dst, f32(-0.5)))
    else:
        src = (scale * (dst + f32(0.5))).astype(f32) - f32(0.5)
    src = np.maximum(src, f32(0)).astype(f32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    return i0, i1, (f32(1) - l1).astype(f32), l1


def candidates(x, oh, ow, src_fma):
    B, C, H, W = x.shape
    y0, y1, lh0, lh1 = coords(H, oh, src_fma)
    x0, x1, lw0, lw1 = coords(W, ow, src_fma)
    lh0, lh1 = lh0[None, None, :, None], lh1[None, None, :, None]
    lw0, lw1 = lw0[None, None, None, :], lw1[None, None, None, :]
    r0, r1 = x[:, :, y0, :], x[:, :, y1, :]
    p00, p01, p10, p11 = r0[..., x0], r0[..., x1], r1[..., x0], r1[..., x1]
    bc = lambda w: np.broadcast_to(w.astype(f32), p00.shape)  # noqa: E731
    l_w0, l_w1, l_h0, l_h1 = bc(lw0), bc(lw1), bc(lh0), bc(lh1)
    # separable forms: horizontal then vertical
    top = fma(l_w0, p00, (l_w1 * p01).astype(f32))
    bot = fma(l_w0, p10, (l_w1 * p11).astype(f32))
    yield "separable fma(lh0,top,lh1*bot)", fma(l_h0, top, (l_h1 * bot).astype(f32))
    yield "separable fma(lh1,bot,lh0*top)", fma(l_h1, bot, (l_h0 * top).astype(f32))
    # weight-product forms
    w00, w01, w10, w11 = bc(lh0 * lw0), bc(lh0 * lw1), bc(lh1 * lw0), bc(lh1 * lw1)
    plain = (((w00 * p00).astype(f32) + (w01 * p01).astype(f32)).astype(f32) + (w10 * p10).astype(f32)).astype(f32)
    yield "products, no fma", (plain + (w11 * p11).astype(f32)).astype(f32)
    acc = fma(w01, p01, (w00 * p00).astype(f32))
    yield "products, chain from w00*p00", fma(w11, p11, fma(w10, p10, acc))
    acc = fma(w00, p00, (w01 * p01).astype(f32))
    yield "products, chain from w01*p01  <-- torch-CPU", fma(w11, p11, fma(w10, p10, acc))


if __name__ == "__main__":
    torch.manual_seed(0)
    cases = [((1, 1, 30, 45), (33, 38)), ((2, 3, 30, 45), (13, 17)), ((5, 3, 40, 48), (26, 32)), ((2, 3, 64, 64), (32, 32))]
    totals: dict[str, int] = {}
    for shape, (oh, ow) in cases:
        x = torch.randint(0, 256, shape, dtype=torch.uint8).float()
        ref = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False).numpy()
        for src_fma in (0, 1):
            for name, out in candidates(x.numpy(), oh, ow, src_fma):
                key = f"src_fma={src_fma} {name}"
                totals[key] = totals.get(key, 0) + int((out != ref).sum())
    for k, v in totals.items():
        print(f"{v:8d} mismatches  {k}")
