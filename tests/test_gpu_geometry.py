"""GPU parity: ROI masks and ROI query vectors through the C ABI vs the CPU oracle.

The mask is integer work: bit-exact.  The oracle decides by clipped area, the kernel by edge
crossings and centre parity; random vertices are drawn on a 1/8 grid so that neither formulation
sits on a rounding boundary."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

from imagescry_b200 import geometry as G  # noqa: E402
from imagescry_b200 import search as S  # noqa: E402


def test_reference_expectations():
    # /root/reference/tests/test_geometry.py:10-52
    m = G.create_roi_mask([(0, 0), (4, 0), (4, 3), (0, 3)], (6, 8), (3, 4))
    assert m.dtype == torch.int64 and m.is_cuda and tuple(m.shape) == (3, 4)
    assert m.tolist() == [[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0]]
    roi = [[(0, 0), (1, 0), (1, 1), (0, 1)], [(2, 2), (3, 2), (3, 3), (2, 3)]]
    assert G.create_roi_mask(roi, (4, 4), (2, 2)).tolist() == [[1, 0], [0, 1]]


def star(rng, cx, cy, r, n):
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(0.3, 1.0, n) * r
    pts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
    return (np.round(pts * 8) / 8 + 1 / 16).tolist()  # off the cell boundaries of every grid used below


@pytest.mark.parametrize("image,fmap", [((256, 256), (8, 8)), ((512, 384), (16, 12)), ((100, 60), (7, 5)), ((640, 640), (20, 20))])
def test_random_polygons_vs_oracle(image, fmap):
    rng = np.random.default_rng(image[0] + fmap[1])
    h, w = image
    for trial in range(6):
        polys = []
        for _ in range(int(rng.integers(1, 4))):
            polys.append(star(rng, rng.uniform(-0.1, 1.1) * w, rng.uniform(-0.1, 1.1) * h, rng.uniform(0.05, 0.5) * max(h, w), int(rng.integers(3, 9))))
        roi = polys if len(polys) > 1 else polys[0]
        ci = int(rng.integers(1, 5))
        got = G.create_roi_mask(roi, image, fmap, ci).cpu().numpy()
        want = O.create_roi_mask(roi, image, fmap, ci)
        assert np.array_equal(got, want), f"trial {trial}: {np.argwhere(got != want)[:5]}"


def test_polygon_with_hole_and_input_forms():
    ring = [(2, 2), (14, 2), (14, 14), (2, 14)]
    hole = [(5, 5), (11, 5), (11, 11), (5, 11)]
    poly = {"exterior": ring, "interiors": [hole]}
    got = G.create_roi_mask(poly, (16, 16), (8, 8), class_index=7).cpu().numpy()
    assert np.array_equal(got, O.create_roi_mask(poly, (16, 16), (8, 8), class_index=7))
    closed = ring + [ring[0]]  # shapely repeats the first vertex
    assert np.array_equal(G.create_roi_mask(closed, (16, 16), (8, 8)).cpu().numpy(), O.create_roi_mask(ring, (16, 16), (8, 8)))
    with pytest.raises(ValueError):
        G.create_roi_mask([(0, 0), (1, 1)], (16, 16), (8, 8))
    with pytest.raises(RuntimeError):
        G.create_roi_mask(ring, (16, 16), (8, 8), device="cpu")


@pytest.mark.parametrize("b,e,h,w", [(3, 128, 7, 10), (5, 1280, 16, 16), (2, 70, 33, 9)])
def test_roi_query_vs_oracle(b, e, h, w):
    rng = np.random.default_rng(b * e + h)
    fmap = O.l2_normalize(np.abs(rng.standard_normal((b, e, h, w))).astype(np.float32))
    mask2 = (rng.random((h, w)) < 0.3).astype(np.int64) * 2
    mask3 = (rng.random((b, h, w)) < 0.3).astype(np.int64) * 2
    mask3[0] = 0  # an image without any ROI cell: zero vector
    d_f = torch.from_numpy(fmap).cuda()
    for mask in (mask2, mask3):
        got = G.roi_query(d_f, torch.from_numpy(mask).cuda(), class_index=2).cpu().numpy()
        want = O.roi_pool(fmap, mask, class_index=2)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-6
    assert np.array_equal(G.roi_query(d_f, torch.from_numpy(mask3).cuda(), class_index=2)[0].cpu().numpy(), np.zeros(e, dtype=np.float32))
    one = G.roi_query(d_f[1], torch.from_numpy(mask2).cuda(), class_index=2)
    assert one.shape == (e,)
    with pytest.raises(ValueError):
        G.roi_query(d_f, torch.zeros((h + 1, w), dtype=torch.int64).cuda())
    with pytest.raises(RuntimeError):
        G.roi_query(torch.from_numpy(fmap), torch.from_numpy(mask2))


def test_roi_query_searches_the_store():
    """ROI -> mask -> pooled query -> cosine search: the cells the ROI covers come back first."""
    rng = np.random.default_rng(5)
    e, h, w = 128, 8, 8
    fmap = O.l2_normalize(rng.standard_normal((1, e, h, w)).astype(np.float32))
    mask = G.create_roi_mask([(33, 33), (63, 33), (63, 63), (33, 63)], (256, 256), (h, w))  # exactly cell (1, 1)
    assert int(mask.sum()) == 1 and int(mask[1, 1]) == 1
    q = G.roi_query(torch.from_numpy(fmap).cuda(), mask)
    rows = O.bf16_round(O.flat_vectors(fmap))  # one store row per cell
    st = S.EmbeddingStore(torch.from_numpy(rows).cuda())
    scores, idx = st.search(q.to(torch.bfloat16), 3)
    assert int(idx[0, 0]) == 1 * w + 1 and float(scores[0, 0]) > 0.99
