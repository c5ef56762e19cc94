"""GPU parity: stage 1 (tile preprocessing) through the C ABI vs the CPU oracle and the golden
vectors frozen from the reference.  Integer/byte and fp32-elementwise work is compared bit-exactly."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import ulp_diff
from oracle import oracle as O

pytestmark = pytest.mark.gpu

from imagescry_b200.image import transforms as T  # noqa: E402


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


def host(t):
    return t.detach().float().cpu().numpy() if t.dtype == torch.bfloat16 else t.detach().cpu().numpy()


# ---------------------------------------------------------------- golden vectors (reference outputs)
def test_golden_resize_bit_exact(golden):
    g = golden("transforms")
    img = dev(g["image"])
    for hw in [(4, 4), (5, 5), (5, 7), (7, 5), (33, 38)]:
        out = host(T.resize(img, hw, side_ref="height"))
        assert np.array_equal(out, g[f"resize_exact_{hw[0]}x{hw[1]}"]), hw
    for size in (16, 31, 46):
        for side_ref in ("height", "width", "long", "short"):
            for tr in (False, True):
                src = img.transpose(1, 2).contiguous() if tr else img
                ref = g[f"resize_int_{size}_{side_ref}_{'T' if tr else 'N'}"]
                out = host(T.resize(src, size, side_ref=side_ref))
                assert out.shape == ref.shape and np.array_equal(out, ref), (size, side_ref, tr)
    assert np.array_equal(host(T.resize(img[0], 16)), g["resize_2d_16"])
    assert np.array_equal(host(T.resize(img[None], 16)), g["resize_4d_16"])


def test_golden_normalize(golden):
    g = golden("transforms")
    img = dev(g["image"])[None]
    m, s = dev(g["norm_supplied_mean"]), dev(g["norm_supplied_std"])
    assert np.array_equal(host(T.normalize_per_channel(img, channel_means=m, channel_stds=s)), g["norm_supplied"])
    out = T.normalize_per_channel(img, channel_means=m, channel_stds=s, min_value=-1.0, max_value=1.0)
    assert np.array_equal(host(out), g["norm_supplied_clip1"])
    # the reference's own fp32 statistics → bit-identical apply stage
    m, s = dev(g["norm_mean"]), dev(g["norm_std"])
    assert np.array_equal(host(T.normalize_per_channel(img.float(), channel_means=m, channel_stds=s)), g["norm_f32in"])
    out = T.normalize_per_channel(img, channel_means=m, channel_stds=s, min_value=-3, max_value=3)
    assert np.array_equal(host(out), g["norm_u8in_clip3"])
    # computed statistics: within 1e-6 absolute of the reference (few-ulp statistics offset)
    out = host(T.normalize_per_channel(img, min_value=-3, max_value=3))
    assert np.abs(out - g["norm_u8in_clip3"]).max() <= 1e-6
    out = host(T.normalize_per_channel(img.float()))
    assert np.allclose(out.mean(axis=(0, 2, 3)), 0, atol=1e-4) and np.allclose(out.std(axis=(0, 2, 3), ddof=1), 1, atol=1e-4)


@pytest.mark.parametrize("msl", [640, 32, 19])
def test_golden_preprocess(golden, msl):
    from imagescry_b200.image.transforms import preprocess_tiles, resized_shape

    g = golden("preprocess")
    images = dev(g["images"])
    h, w = images.shape[-2:]
    out_hw = resized_shape(h, w, msl, "long") if max(h, w) > msl else None
    m, s = dev(g[f"pre_msl{msl}_mean"]), dev(g[f"pre_msl{msl}_std"])
    out = preprocess_tiles(images, output_hw=out_hw, channel_means=m, channel_stds=s, min_value=-3, max_value=3)
    assert np.array_equal(host(out), g[f"pre_msl{msl}"])
    out = preprocess_tiles(images, output_hw=out_hw, min_value=-3, max_value=3)
    assert np.abs(host(out) - g[f"pre_msl{msl}"]).max() <= 1e-6
    nhwc = images.permute(0, 2, 3, 1).contiguous()
    out2 = preprocess_tiles(nhwc, layout="nhwc", output_hw=out_hw, min_value=-3, max_value=3)
    assert torch.equal(out, out2)


# ---------------------------------------------------------------- oracle on seeded inputs
CASES = [
    # (B, C, H, W, out_hw)
    (3, 3, 64, 64, None),       # fast paths (plane % 16 == 0)
    (2, 3, 30, 45, None),       # odd plane → staged path
    (5, 1, 32, 48, None),
    (2, 4, 16, 16, None),
    (3, 3, 64, 64, (32, 32)),   # exact 2x down-scale: integer box path
    (2, 3, 48, 72, (24, 36)),   # box path (NHWC), rectangular; NCHW needs outW % 8 == 0
    (3, 3, 32, 80, (16, 40)),   # box path in both layouts
    (1, 3, 6, 4128, (3, 2064)), # box path, rows wider than a CTA (several passes over q)
    (2, 3, 20, 44, (10, 22)),   # 2x but outW % 4 != 0: generic sampling kernel
    (2, 3, 96, 80, (37, 31)),
    (2, 3, 40, 48, (80, 96)),   # up-scale
    (1, 3, 200, 300, (64, 96)),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_preprocess_vs_oracle_u8(case, layout):
    B, C, H, W, out_hw = case
    rng = np.random.default_rng(B * 1000 + H)
    x = rng.integers(0, 256, (B, C, H, W), dtype=np.uint8)
    xin = x if layout == "nchw" else np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    lay = O.NCHW if layout == "nchw" else O.NHWC
    res = O.bilinear_resize(xin, *out_hw, layout=lay) if out_hw else O.to_nchw(xin, lay).astype(np.float32)
    ref = O.normalize_per_channel(res, min_value=-3, max_value=3)
    rm, rs = O.channel_stats(res)
    out = T.preprocess_tiles(dev(xin), layout=layout, output_hw=out_hw, min_value=-3, max_value=3)
    # statistics: both are the correctly rounded fp32 of the exact value → identical → bit-exact output
    assert np.array_equal(host(out), ref)
    # supplied per-image statistics ("#B C 1 1" with B rows)
    pm = rng.uniform(100, 150, (B, C, 1, 1)).astype(np.float32)
    ps = rng.uniform(40, 80, (B, C, 1, 1)).astype(np.float32)
    ref2 = O.normalize_per_channel(res, channel_means=pm, channel_stds=ps)
    out2 = T.preprocess_tiles(dev(xin), layout=layout, output_hw=out_hw, channel_means=dev(pm), channel_stds=dev(ps))
    assert np.array_equal(host(out2), ref2)
    # bf16 output = round-to-nearest-even of the fp32 result
    out3 = T.preprocess_tiles(dev(xin), layout=layout, output_hw=out_hw, min_value=-3, max_value=3, out_dtype=torch.bfloat16)
    assert np.array_equal(host(out3), O.bf16_round(ref))
    del rm, rs


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_resize_normalize_division_bit_exact_many_divisors(layout):
    """The three-channel resize kernel hoists the reciprocal of div.rn.f32's fast path out of the
    pixel loop; (x - m) / (s + eps) must still be the IEEE quotient for every divisor.  96 images,
    each with its own per-image statistics over ten decades, plus divisors outside the hoisted
    path's range (they take the per-pixel IEEE division)."""
    rng = np.random.default_rng(2024)
    B, H, W, oh, ow = 96, 40, 56, 23, 31
    x = rng.integers(0, 256, (B, 3, H, W), dtype=np.uint8)
    xin = x if layout == "nchw" else np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    lay = O.NCHW if layout == "nchw" else O.NHWC
    res = O.bilinear_resize(xin, oh, ow, layout=lay)
    pm = rng.uniform(-300, 300, (B, 3, 1, 1)).astype(np.float32)
    ps = (10.0 ** rng.uniform(-5, 5, (B, 3, 1, 1))).astype(np.float32)
    ps[0] = 1e-30   # outside [2^-60, 2^60]: per-pixel IEEE division
    ps[1] = 3e25
    pm[2] = 5e7     # |mean| beyond the hoisted path's bound
    pm[3] = 0.0     # x - m can be exactly 0 and tiny
    for kw in (dict(), dict(min_value=-3, max_value=3), dict(min_value=-0.5)):
        ref = O.normalize_per_channel(res, channel_means=pm, channel_stds=ps, **kw)
        out = T.preprocess_tiles(dev(xin), layout=layout, output_hw=(oh, ow), channel_means=dev(pm), channel_stds=dev(ps), **kw)
        assert np.array_equal(host(out), ref), kw


def test_preprocess_float_input_and_stats_ulps():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((4, 3, 33, 47)).astype(np.float32) * 50 + 120
    ref = O.normalize_per_channel(x)
    out = host(T.normalize_per_channel(dev(x)))
    # float input: statistics accumulate in fp64 in one pass; allow 1 ulp on mean/std → 1e-6 abs
    assert np.abs(out - ref).max() <= 2e-6
    m, s = O.channel_stats(x)
    out = host(T.normalize_per_channel(dev(x), channel_means=dev(m), channel_stds=dev(s)))
    assert np.array_equal(out, O.normalize_per_channel(x, channel_means=m, channel_stds=s))
    # float16 / int32 inputs follow the reference's `.float()` cast
    xi = rng.integers(-1000, 1000, (2, 2, 8, 8)).astype(np.int32)
    assert np.array_equal(host(T.normalize_per_channel(dev(xi))), O.normalize_per_channel(xi))


def test_box_path_statistics_exact_large_batch():
    """2x down-scale statistics are exact integer sums of the four taps: compare with an int64 /
    fp64 truth at a size where fp32 or even naive fp64 accumulation orders could differ."""
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, (48, 128, 128, 3), dtype=np.uint8)
    S = (x[:, 0::2, 0::2].astype(np.int64) + x[:, 0::2, 1::2] + x[:, 1::2, 0::2] + x[:, 1::2, 1::2])  # B, 64, 64, 3
    n = S.shape[0] * S.shape[1] * S.shape[2]
    s1 = S.sum(axis=(0, 1, 2)).astype(object)
    s2 = (S * S).sum(axis=(0, 1, 2)).astype(object)
    mean = np.array([float(s1[c]) / (4.0 * n) for c in range(3)], dtype=np.float32)
    var = [float(n * s2[c] - s1[c] * s1[c]) / (16.0 * n * (n - 1)) for c in range(3)]
    std = np.sqrt(np.array(var, dtype=np.float64)).astype(np.float32)
    y = (S.astype(np.float32) * 0.25).transpose(0, 3, 1, 2)
    assert np.array_equal(y, O.bilinear_resize(x, 64, 64, layout=O.NHWC))  # the box identity itself
    ref = O.normalize_per_channel(y, channel_means=mean.reshape(1, 3, 1, 1), channel_stds=std.reshape(1, 3, 1, 1), min_value=-3, max_value=3)
    out = host(T.preprocess_tiles(dev(x), layout="nhwc", output_hw=(64, 64), min_value=-3, max_value=3))
    assert np.array_equal(out, ref)
    xp = np.ascontiguousarray(x.transpose(0, 3, 1, 2))
    out = host(T.preprocess_tiles(dev(xp), layout="nchw", output_hw=(64, 64), min_value=-3, max_value=3, out_dtype=torch.bfloat16))
    assert np.array_equal(out, O.bf16_round(ref))


def test_low_variance_and_constant_tiles():
    x = np.full((2, 3, 16, 16), 200, dtype=np.uint8)
    x[0, 0, 0, 0] = 201
    out = host(T.normalize_per_channel(dev(x), min_value=-3, max_value=3))
    assert np.array_equal(out, O.normalize_per_channel(x, min_value=-3, max_value=3))
    assert np.all(out[:, 1:] == 0.0)  # constant channels: (x - mean) = 0, std = 0, eps keeps it finite


def test_large_batch_statistics_exact():
    """Integer statistics stay exact where fp32 reductions drift: 64 tiles of 256x256."""
    rng = np.random.default_rng(11)
    x = rng.integers(0, 256, (64, 3, 256, 256), dtype=np.uint8)
    xd = x.astype(np.float64)
    m = xd.mean(axis=(0, 2, 3)).astype(np.float32)
    s = xd.std(axis=(0, 2, 3), ddof=1).astype(np.float32)
    ref = O.normalize_per_channel(x, channel_means=m.reshape(1, 3, 1, 1), channel_stds=s.reshape(1, 3, 1, 1), min_value=-3, max_value=3)
    for layout, xin in (("nchw", x), ("nhwc", np.ascontiguousarray(x.transpose(0, 2, 3, 1)))):
        out = host(T.preprocess_tiles(dev(xin), layout=layout, min_value=-3, max_value=3))
        assert np.array_equal(out, ref), layout
    # and against torch's own fp32 pipeline on the GPU (the reference's code path on CUDA): few ulp
    xt = dev(x).float()
    tm, ts = xt.mean((0, 2, 3)), xt.std((0, 2, 3))
    assert ulp_diff(host(tm), m).max() <= 8 and ulp_diff(host(ts), s).max() <= 8


def test_errors_and_edge_cases():
    with pytest.raises(RuntimeError):
        T.normalize_per_channel(torch.zeros(1, 3, 4, 4, dtype=torch.uint8))  # CPU tensor: no fallback
    with pytest.raises(Exception):
        T.normalize_per_channel(torch.zeros(3, 4, 4, dtype=torch.uint8).cuda())  # wrong rank (jaxtyping)
    with pytest.raises(ValueError):
        T.normalize_per_channel(torch.zeros(1, 1, 1, 1, dtype=torch.uint8).cuda())  # one pixel: no unbiased std
    empty = T.preprocess_tiles(torch.zeros(0, 3, 8, 8, dtype=torch.uint8).cuda())
    assert empty.shape == (0, 3, 8, 8)
    # round trip property at a larger size: resize to the same size is the identity cast
    x = torch.randint(0, 256, (2, 3, 50, 70), dtype=torch.uint8).cuda()
    assert torch.equal(T.resize(x, (50, 70)), x.float())


def test_full_size_batch_properties():
    """BASELINE.json config 3 at full size (4096 tiles of 512x512x3), through size-independent
    properties: the reference's own test property (per-channel mean 0 / std 1 of the normalised
    batch, tests/test_image/test_transform.py:14-24), clip bounds, layout independence, the 2x box
    identity, and a sampled bit-exact comparison with the oracle."""
    B, H, W = 4096, 512, 512
    g = torch.Generator(device="cuda").manual_seed(1234)
    tiles = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    out = T.preprocess_tiles(tiles, layout="nhwc")  # no clip: statistics are exactly testable
    assert out.shape == (B, 3, H, W) and out.dtype == torch.float32
    m = out.mean(dim=(0, 2, 3), dtype=torch.float64)
    s = out.double().std(dim=(0, 2, 3)) if False else torch.sqrt((out.double() ** 2).mean(dim=(0, 2, 3)) - m**2)
    assert torch.all(m.abs() < 1e-4) and torch.all((s - 1).abs() < 1e-4)
    # sampled tiles, bit-exact against the oracle given the batch statistics the kernel derived
    x64 = tiles[:1].double()
    xs = tiles.view(-1, 3)
    n = xs.shape[0]
    s1 = xs.sum(dim=0, dtype=torch.int64).cpu().numpy().astype(object)
    s2 = (xs.to(torch.int64) ** 2).sum(dim=0).cpu().numpy().astype(object)
    mean = np.array([float(s1[c]) / n for c in range(3)], dtype=np.float32)
    std = np.sqrt(np.array([float(n * s2[c] - s1[c] * s1[c]) / (float(n) * (n - 1)) for c in range(3)])).astype(np.float32)
    pick = [0, 1234, B - 1]
    ref = O.normalize_per_channel(
        np.ascontiguousarray(tiles[pick].cpu().numpy().transpose(0, 3, 1, 2)),
        channel_means=mean.reshape(1, 3, 1, 1), channel_stds=std.reshape(1, 3, 1, 1), min_value=-3, max_value=3,
    )
    del out, x64
    clipped = T.preprocess_tiles(tiles, layout="nhwc", min_value=-3, max_value=3)
    assert float(clipped.min()) >= -3.0 and float(clipped.max()) <= 3.0
    assert np.array_equal(host(clipped[pick]), ref)
    # planar input gives the same bits
    planar = tiles.permute(0, 3, 1, 2).contiguous()
    assert torch.equal(T.preprocess_tiles(planar, min_value=-3, max_value=3), clipped)
    del clipped
    # max_side_length = 256: the resized batch is the 2x2 box mean, normalised with ITS statistics
    small = T.preprocess_tiles(tiles, layout="nhwc", output_hw=(256, 256))
    box = tiles.view(B, 256, 2, 256, 2, 3).float().sum(dim=(2, 4)).mul_(0.25).permute(0, 3, 1, 2)
    bm = box.mean(dim=(0, 2, 3), dtype=torch.float64)
    bs = torch.sqrt(((box.double() - bm.view(1, 3, 1, 1)) ** 2).sum(dim=(0, 2, 3)) / (box.numel() // 3 - 1))
    expect = (box - bm.float().view(1, 3, 1, 1)) / (bs.float().view(1, 3, 1, 1) + 1e-6)
    assert torch.equal(small, expect)
    assert torch.equal(T.preprocess_tiles(planar, output_hw=(256, 256)), small)
