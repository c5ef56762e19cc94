"""GPU parity: stage 3 (cosine k-NN with fused top-k, merge) through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): scores within 1e-3 absolute; index sets identical except
where the score gap at the k-th boundary is below the tolerance; ties broken by lower index."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

from imagescry_b200 import search as S  # noqa: E402

SCORE_TOL = 1e-3


def make(n, q, d, seed, dup=False):
    rng = np.random.default_rng(seed)
    store = O.bf16_round(rng.standard_normal((n, d)).astype(np.float32))
    queries = O.bf16_round(rng.standard_normal((q, d)).astype(np.float32))
    if dup and n > 20:
        store[17] = store[3]
        store[n - 1] = store[3]
        queries[0] = store[3]
    return store, queries


def check(store, queries, k, scores, idx, index_base=0, graph=False):
    """Scores within SCORE_TOL; index sets identical except where EVERY differing index's exact score
    lies within SCORE_TOL of the exact k-th score (each mismatch is justified individually)."""
    ref_s, ref_i = O.knn_graph(store, k) if graph else O.cosine_knn(store, queries, k, index_base=index_base)
    if graph:
        queries = store
    scores, idx = scores.cpu().numpy(), idx.cpu().numpy()
    assert scores.shape == ref_s.shape and idx.shape == ref_i.shape
    valid = ref_i >= 0
    assert np.array_equal(idx >= 0, valid)
    assert np.all(np.isneginf(scores[~valid]))
    assert np.abs(scores[valid] - ref_s[valid]).max(initial=0.0) <= SCORE_TOL
    # rows must be ordered by (score desc, index asc)
    for r in range(scores.shape[0]):
        row_s, row_i = scores[r][valid[r]], idx[r][valid[r]]
        order = np.lexsort((row_i, -row_s))
        assert np.array_equal(order, np.arange(len(row_s))), f"row {r} not ordered"
    exact = np.array_equal(idx, ref_i)
    if not exact:
        # allowed only where the boundary gap is below the tolerance
        sn = store * O.row_rnorm(store)[:, None]
        qn = queries * O.row_rnorm(queries)[:, None]
        for r in np.nonzero((idx != ref_i).any(axis=1))[0]:
            full = qn[r] @ sn.T
            if graph:
                full[r] = -np.inf
            got, want = set(idx[r].tolist()), set(ref_i[r].tolist())
            for j in got ^ want:
                if j < 0:
                    raise AssertionError(f"row {r}: padding mismatch")
                kth = np.sort(full)[-k] if len(full) >= k else -np.inf
                assert abs(full[j - index_base] - kth) < SCORE_TOL, f"row {r}: index {j} differs beyond tolerance"
    return exact


@pytest.mark.parametrize(
    "n,q,d,k",
    [
        (300, 7, 64, 10),        # one partial tile
        (256, 128, 64, 1),
        (1000, 130, 128, 10),    # two query blocks, ragged store
        (5000, 200, 1280, 10),   # the metric's dimensionality
        (4096, 64, 256, 16),
        (3000, 50, 1280, 100),   # k = 100 path (global candidate buffers)
        (700, 33, 72, 100),      # d not a multiple of 64
        (40000, 300, 128, 10),   # several N-splits
        (2000, 257, 128, 10),    # CTA pairs: the second pair holds one query row
        (2500, 513, 64, 100),    # CTA pairs with global candidate buffers, ragged store tile
        (60000, 1000, 256, 32),  # pairs x splits x k between the two candidate-buffer regimes
        (20000, 19000, 64, 10),  # more query blocks than units: several items per unit (resident query tile reloaded)
        (9000, 700, 200, 16),    # resident query tile with a ragged last k-block
        (5000, 300, 8, 5),       # one k-block, mostly zero-filled
    ],
)
def test_knn_vs_oracle(n, q, d, k):
    store, queries = make(n, q, d, seed=n + q + d + k, dup=True)
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    scores, idx = st.search(torch.from_numpy(queries).cuda(), k)
    assert idx.dtype == torch.int64 and scores.dtype == torch.float32
    check(store, queries, k, scores, idx)
    if n > 20:
        # exact duplicates of the query: three-way tie broken by index
        assert idx[0, :3].tolist() == [3, 17, n - 1][:k]


@pytest.mark.parametrize("n,q,d,k", [(20000, 300, 256, 10), (9000, 140, 128, 100)])
def test_knn_heterogeneous_norms(n, q, d, k):
    """Store rows whose norms span four decades: raw dot products say nothing about the ranking, the
    per-column scaling by the inverse norms decides."""
    rng = np.random.default_rng(n + k)
    store, queries = make(n, q, d, seed=n * 3 + k)
    store = O.bf16_round(store * (10.0 ** rng.uniform(-2, 2, size=(n, 1))).astype(np.float32))
    store[11] = 0.0
    queries = O.bf16_round(queries * (10.0 ** rng.uniform(-1, 1, size=(q, 1))).astype(np.float32))
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    scores, idx = st.search(torch.from_numpy(queries).cuda(), k)
    check(store, queries, k, scores, idx)


@pytest.mark.parametrize("k", [1, 10, 16, 40, 100])
def test_knn_exact_ties_across_norms(k):
    """Rows that are power-of-two multiples of one another have bit-identical cosine scores but
    different norms; every query's best hits are a dozen such copies spread over the store (other
    splits, other tiles, other CTAs), and each tie must go to the lower index.  Indices must match
    the oracle exactly."""
    rng = np.random.default_rng(77 + k)
    n, q, d, base = 6000, 200, 64, 500
    proto = O.bf16_round(rng.standard_normal((base, d)).astype(np.float32))
    src = rng.integers(0, base, size=n)
    scale = (2.0 ** rng.integers(-3, 4, size=(n, 1))).astype(np.float32)
    store = O.bf16_round(proto[src] * scale)          # exact: powers of two
    queries = O.bf16_round(proto[rng.integers(0, base, size=q)] + 0.25 * rng.standard_normal((q, d)).astype(np.float32))
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    scores, idx = st.search(torch.from_numpy(queries).cuda(), k)
    ref_s, ref_i = O.cosine_knn(store, queries, k)
    assert np.array_equal(idx.cpu().numpy(), ref_i)
    assert np.abs(scores.cpu().numpy() - ref_s).max() <= SCORE_TOL
    # every query's best hits are one prototype's copies: ties by the dozen (n / base = 12 copies each)
    assert (ref_s[:, 0] == ref_s[:, min(k, 3) - 1]).mean() > 0.9
    gs, gi = st.knn_graph(min(k, 16))
    gref_s, gref_i = O.knn_graph(store, min(k, 16))
    assert np.array_equal(gi.cpu().numpy(), gref_i)


@pytest.mark.parametrize("n,q,k", [(150, 40, 128), (200, 300, 128), (20, 5, 16), (300, 50, 128), (520, 257, 128), (700, 64, 100)])
def test_knn_negative_thresholds(n, q, k):
    """k close to n: the k-th best cosine is negative or near zero, where the chunk-level bound does
    not apply; the larger stores fill the 256-entry buffers, so the in-tile selection prune (bitwise
    bisection over the order-preserving integer image of the scores) runs on mixed-sign scores."""
    store, queries = make(n, q, 64, seed=n + k)
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    scores, idx = st.search(torch.from_numpy(queries).cuda(), k)
    check(store, queries, k, scores, idx)
    if k <= n and n <= 2 * k:
        assert (scores[:, k - 1] < 0).any()


def test_knn_small_store_padding_and_index_base():
    store, queries = make(6, 5, 64, seed=1)
    st = S.EmbeddingStore(torch.from_numpy(store).cuda(), index_base=1000)
    scores, idx = st.search(torch.from_numpy(queries).cuda(), 10)
    check(store, queries, 10, scores, idx, index_base=1000)
    assert (idx[:, 6:] == -1).all()


def test_knn_zero_vectors_and_empty():
    store, queries = make(500, 9, 64, seed=2)
    store[5] = 0.0
    queries[2] = 0.0
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    scores, idx = st.search(torch.from_numpy(queries).cuda(), 10)
    check(store, queries, 10, scores, idx)
    assert torch.all(scores[2] == 0)  # zero query: every score is 0, ties → lowest indices
    assert idx[2].tolist() == list(range(10))
    s0, i0 = st.search(torch.zeros((0, 64)).cuda(), 10)
    assert s0.shape == (0, 10) and i0.shape == (0, 10)
    with pytest.raises(ValueError):
        st.search(torch.zeros((1, 32)).cuda(), 10)
    with pytest.raises(ValueError):
        st.search(torch.zeros((1, 64)).cuda(), 1000)
    with pytest.raises(RuntimeError):
        S.EmbeddingStore(torch.zeros((4, 64)))  # CPU tensor: no fallback


def test_rnorm_and_merge_vs_oracle():
    rng = np.random.default_rng(3)
    x = O.bf16_round(rng.standard_normal((1000, 1280)).astype(np.float32) * 3)
    x[10] = 0
    got = S.row_rnorm(torch.from_numpy(x).cuda().to(torch.bfloat16)).cpu().numpy()
    assert np.allclose(got, O.row_rnorm(x), rtol=2e-6)
    for g, q, k in [(13, 37, 10), (8, 20, 100), (1, 5, 3), (3, 4, 128)]:
        s = rng.standard_normal((g, q, k)).astype(np.float32)
        i = rng.permutation(g * q * k).astype(np.int32).reshape(g, q, k)
        s[0, 0, :] = s[0, 0, 0]  # ties
        i[-1, :, k // 2:] = -1  # padding
        ref_s, ref_i = O.topk_merge(s, i, k)
        out_s, out_i = S.merge_topk(torch.from_numpy(s).cuda(), torch.from_numpy(i).cuda())
        assert np.array_equal(out_i.cpu().numpy(), ref_i) and np.array_equal(out_s.cpu().numpy(), ref_s)


def test_sharded_equals_single_via_merge():
    """Row shards searched separately and merged equal the unsharded search (size-independent
    property used at full scale; here emulating 4 ranks on one device)."""
    store, queries = make(20000, 64, 128, seed=5)
    qd = torch.from_numpy(queries).cuda()
    full_s, full_i = S.EmbeddingStore(torch.from_numpy(store).cuda()).search(qd, 10)
    parts_s, parts_i = [], []
    for r in range(4):
        b, e = S.shard_range(len(store), 4, r)
        s, i = S.EmbeddingStore(torch.from_numpy(store[b:e]).cuda(), index_base=b).search_raw(qd, 10)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = S.merge_topk(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(mi.to(torch.int64), full_i) and torch.allclose(ms, full_s, atol=1e-6)


def test_self_search_property_large():
    """At a size the oracle cannot brute-force quickly: querying with store rows returns the row
    itself first with score 1, and scores are non-increasing."""
    n, d = 200_000, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    store = torch.randn((n, d), generator=g, device="cuda").to(torch.bfloat16)
    st = S.EmbeddingStore(store)
    rows = torch.arange(0, n, 997, device="cuda")[:150]
    scores, idx = st.search(store[rows], 10)
    assert torch.equal(idx[:, 0], rows)
    assert torch.allclose(scores[:, 0], torch.ones_like(scores[:, 0]), atol=1e-3)
    assert torch.all(scores[:, 1:] <= scores[:, :-1])
    # cross-check against a chunked fp32 torch brute force (library code, test-side only)
    sn = torch.nn.functional.normalize(store.float(), dim=1)
    ref = (sn[rows] @ sn.T).topk(10, dim=1)
    assert torch.allclose(scores, ref.values, atol=SCORE_TOL)


def test_store_format_bridge_blobs_to_search():
    """storage/models.py:94-129 BLOBs -> device store -> search: rows equal the oracle's bf16 rows
    bit for bit, and a query built from one cell finds that (image, cell)."""
    from imagescry_b200 import store_format as F

    rng = np.random.default_rng(9)
    maps = rng.standard_normal((7, 72, 5, 6)).astype(np.float32)  # C not a multiple of 32, hw = 30
    records = [(O.blob_encode(m), 72, 5, 6) for m in maps]
    # codec round trip equals the reference's own decode
    for (data, c, h, w), m in zip(records, maps):
        assert np.array_equal(F.decode_embedding_blob(data, c, h, w).numpy(), O.blob_decode(data, c, h, w))
        assert F.encode_embedding_blob(torch.from_numpy(m))[0] == data
    stacked = F.stack_blobs(records)
    rows = F.maps_to_rows(stacked.cuda())
    assert np.array_equal(rows.float().cpu().numpy(), O.maps_to_rows(maps))
    pooled = F.maps_to_rows(stacked.cuda(), pool="mean").float().cpu().numpy()
    ref_pooled = O.maps_to_rows(maps, pool="mean")
    assert np.abs(pooled - ref_pooled).max() <= np.abs(ref_pooled).max() * 2.0**-7  # one bf16 ulp
    store = F.store_from_blobs(records)
    assert len(store) == 7 * 30 and store.dim == 72
    q = torch.from_numpy(maps[4, :, 2, 3]).reshape(1, -1).cuda()
    scores, idx = store.search(q, 3)
    img, cell = F.rows_to_image_cell(idx[0, :1].cpu(), 30)
    assert (int(img), int(cell)) == (4, 2 * 6 + 3) and abs(float(scores[0, 0]) - 1.0) < 1e-2
    check(O.maps_to_rows(maps), O.bf16_round(maps[4, :, 2, 3].reshape(1, -1)), 3, scores, idx)
    with pytest.raises(ValueError):
        F.decode_embedding_blob(records[0][0], 72, 5, 5)


def _device_store(n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    store = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    for s0 in range(0, n, 1 << 18):
        e = min(n, s0 + (1 << 18))
        store[s0:e] = torch.randn((e - s0, d), generator=g, device="cuda").to(torch.bfloat16)
    return store, g


def _judge_sampled(st, store, queries, k, scores, idx, samples=64):
    """64 sampled queries against an exact fp32 torch brute force over the full store (library code,
    test side only); every index mismatch must be justified by a score gap below the tolerance."""
    from bench import exact_topk_local, judge_topk

    q = queries.shape[0]
    pick = torch.arange(0, q, max(1, q // samples), device="cuda")[:samples]
    qn = torch.nn.functional.normalize(queries[pick].float(), dim=1)
    ex_s, ex_i = exact_topk_local(store, 0, qn, k + 32)
    res = judge_topk(scores[pick].cpu(), idx[pick].cpu(), ex_s.cpu(), ex_i.cpu(), k, SCORE_TOL)
    assert res["index_mismatch_beyond_tol"] == 0, res
    assert res["max_score_err"] <= SCORE_TOL, res
    return res


def test_full_size_store_sampled_queries():
    """BASELINE.json config 2 at full size (1 M x 1280 bf16 store, 10 k queries): every result row is
    ordered and in range; sampled queries are checked against an fp32 brute force over the full store
    for k = 10 (shared-memory candidate buffers), k = 32 and k = 100 (global 256-entry buffers, the
    lockstep throttle, several items per running list); merging 8 row shards — two-array and packed
    form — reproduces the unsharded answer (the size-independent property behind the multi-GPU path)."""
    n, d, q = 1_000_000, 1280, 10_000
    store, g = _device_store(n, d, 1234)
    queries = torch.randn((q, d), generator=g, device="cuda").to(torch.bfloat16)
    st = S.EmbeddingStore(store)
    for k in (10, 32, 100):
        scores, idx = st.search(queries, k)
        assert scores.shape == (q, k) and int(idx.min()) >= 0 and int(idx.max()) < n
        assert torch.all(scores[:, 1:] <= scores[:, :-1])
        _judge_sampled(st, store, queries, k, scores, idx)
    k = 10
    scores, idx = st.search(queries, k)
    parts_s, parts_i, parts_r = [], [], []
    for r in range(8):
        b, e = S.shard_range(n, 8, r)
        shard = S.EmbeddingStore(store[b:e], index_base=b)
        ps, pi = shard.search_raw(queries, k)
        parts_s.append(ps)
        parts_i.append(pi)
        parts_r.append(shard.search_packed(queries, k))
    ms, mi = S.merge_topk(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(mi.to(torch.int64), idx) and torch.allclose(ms, scores, atol=1e-6)
    ps, pi = S.merge_topk_packed(torch.stack(parts_r))
    assert torch.equal(pi, mi) and torch.equal(ps, ms)
    # the same store searched block by block with the running lists kept in the workspace
    # (ISX_KNN_CONTINUE: how a rotating / chunked store is searched without gathering partial results)
    from imagescry_b200 import _lib

    qr = S.row_rnorm(queries)
    out_s = torch.empty((q, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((q, k), dtype=torch.int32, device="cuda")
    bounds = [0, 300_000, 300_001, 650_000, n]
    for j in range(len(bounds) - 1):
        b, e = bounds[j], bounds[j + 1]
        last = j == len(bounds) - 2
        flags = (_lib.KNN_CONTINUE if j else 0) | (0 if last else _lib.KNN_NO_FINALIZE)
        st.search_block(store[b:e], st.rnorm[b:e], b, queries, qr, k, flags=flags, out_a=out_s if last else None,
                        out_b=out_i if last else None)
    assert torch.equal(out_i.to(torch.int64), idx) and torch.allclose(out_s, scores, atol=1e-6)


def test_full_size_store_d256_resident_query_path():
    """1 M x 256 (config 5's dimensionality; the resident-query-tile kernel for k <= 16): k = 10, 32 and
    100 against the sampled brute force."""
    n, d, q = 1_000_000, 256, 10_000
    store, g = _device_store(n, d, 77)
    queries = torch.randn((q, d), generator=g, device="cuda").to(torch.bfloat16)
    st = S.EmbeddingStore(store)
    for k in (10, 32, 100):
        scores, idx = st.search(queries, k)
        assert torch.all(scores[:, 1:] <= scores[:, :-1])
        _judge_sampled(st, store, queries, k, scores, idx)


def test_knn_feature_width_not_a_multiple_of_8():
    """PCA picks `num_components` from `min_explained_variance` (decomposition.py:128-137): widths
    like 37 are normal.  The store zero-pads rows to the next multiple of 8 (scores unchanged)."""
    store, queries = make(3000, 70, 37, seed=37, dup=True)
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    assert st.num_features == 37 and st.dim == 40
    scores, idx = st.search(torch.from_numpy(queries).cuda(), 10)
    check(store, queries, 10, scores, idx)
    s, i = st.knn_graph(5)
    check(store, None, 5, s, i, graph=True)
    with pytest.raises(ValueError):
        st.search(torch.zeros((1, 36)).cuda(), 10)


def test_packed_records_round_trip():
    store, queries = make(5000, 130, 64, seed=8)
    st = S.EmbeddingStore(torch.from_numpy(store).cuda(), index_base=77)
    qd = torch.from_numpy(queries).cuda()
    s, i = st.search_raw(qd, 10)
    rec = st.search_packed(qd, 10)
    us, ui = S.unpack_records(rec)
    assert torch.equal(us, s) and torch.equal(ui, i)
    assert torch.equal(S.pack_records(s, i), rec)
    ms, mi = S.merge_topk_packed(rec.unsqueeze(0))
    assert torch.equal(ms, s) and torch.equal(mi, i)


def test_knn_graph_excludes_self():
    """All-pairs graph (config 5 shape, d = 256): k nearest OTHER rows of every row, duplicates kept;
    the row's own entry is skipped inside the kernel's selection."""
    store, _ = make(3000, 1, 256, seed=11, dup=True)  # rows 3, 17 and n-1 are identical
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    for k in (5, 40):
        s, i = S.knn_graph(st, k)
        check(store, None, k, s, i, graph=True)
        assert not (i.cpu() == torch.arange(3000).reshape(-1, 1)).any()
        assert i[3, :2].tolist() == [17, 2999] and i[17, :2].tolist() == [3, 2999]
    # fewer other rows than k: the tail is padding (-inf, -1), never the row itself
    tiny = S.EmbeddingStore(torch.from_numpy(store[:5]).cuda())
    ts, ti = tiny.knn_graph(10)
    check(store[:5], None, 10, ts, ti, graph=True)
    assert (ti[:, 4:] == -1).all() and torch.isneginf(ts[:, 4:]).all() and (ti[:, :4] >= 0).all()
    assert S.EmbeddingStore(torch.zeros((0, 64)).cuda()).knn_graph(3)[1].shape == (0, 3)
    # a store with an index base: neighbours carry global indices and self is still excluded
    st2 = S.EmbeddingStore(torch.from_numpy(store).cuda(), index_base=500)
    s2, i2 = st2.knn_graph(5)
    s1, i1 = st.knn_graph(5)
    assert torch.equal(i2, i1 + 500) and torch.equal(s2, s1)


def test_knn_graph_large_sampled():
    """200 k x 256 graph: sampled rows against the exact brute force with the row itself excluded."""
    from bench import exact_topk_local, judge_topk

    n, d, k = 200_000, 256, 10
    store, _ = _device_store(n, d, 5)
    st = S.EmbeddingStore(store)
    s, i = st.knn_graph(k)
    pick = torch.arange(0, n, n // 64, device="cuda")[:64]
    qn = torch.nn.functional.normalize(store[pick].float(), dim=1)
    ex_s, ex_i = exact_topk_local(store, 0, qn, k + 32, exclude=pick)
    res = judge_topk(s[pick].cpu(), i[pick].cpu(), ex_s.cpu(), ex_i.cpu(), k, SCORE_TOL)
    assert res["index_mismatch_beyond_tol"] == 0 and res["max_score_err"] <= SCORE_TOL, res
    assert not (i == torch.arange(n, device="cuda").reshape(-1, 1)).any()


def test_concurrent_searches_on_two_streams():
    """One store searched from two CUDA streams at once (the workspace — running lists, locks, bounds —
    is cached per stream), k = 10 and k = 100.  Two persistent 148-CTA kernels cannot be co-resident:
    the second one's CTA pairs start as the first one's finish, i.e. staggered by milliseconds, which
    is exactly the situation the lockstep throttle's 2 ms give-up exists for.  Both results must
    equal the serial ones bit for bit."""
    n, d = 400_000, 256
    store, g = _device_store(n, d, 21)
    st = S.EmbeddingStore(store)
    qa = torch.randn((6000, d), generator=g, device="cuda").to(torch.bfloat16)
    qb = torch.randn((6000, d), generator=g, device="cuda").to(torch.bfloat16)
    for k in (10, 100):
        ref_a = st.search_raw(qa, k)
        ref_b = st.search_raw(qb, k)
        torch.cuda.synchronize()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        outs = {}
        for rep in range(3):
            with torch.cuda.stream(s1):
                outs["a"] = st.search_raw(qa, k)
            with torch.cuda.stream(s2):
                outs["b"] = st.search_raw(qb, k)
        torch.cuda.synchronize()
        assert torch.equal(outs["a"][1], ref_a[1]) and torch.equal(outs["a"][0], ref_a[0])
        assert torch.equal(outs["b"][1], ref_b[1]) and torch.equal(outs["b"][0], ref_b[0])
    assert len(st._workspaces) >= 3  # default stream + the two side streams


def test_store_from_mixed_shape_blobs():
    """A database whose maps differ in H×W (what the reference's reader zero-pads, data.py:378-399):
    nothing is padded, image i's cells are rows [offsets[i], offsets[i + 1]) and equal the oracle's
    rows of that map bit for bit; a query built from one cell finds that (image, cell)."""
    from imagescry_b200 import store_format as F

    rng = np.random.default_rng(13)
    shapes = [(3, 4), (5, 6), (3, 4), (2, 7), (5, 6), (3, 4)]
    maps = [rng.standard_normal((72, h, w)).astype(np.float32) for h, w in shapes]
    records = [(O.blob_encode(m), 72, m.shape[1], m.shape[2]) for m in maps]
    store, offsets = F.store_from_mixed_blobs(records)
    assert offsets.tolist() == [0, 12, 42, 54, 68, 98, 110] and len(store) == 110
    rows = store.embeddings.float().cpu().numpy()
    for i, m in enumerate(maps):
        assert np.array_equal(rows[int(offsets[i]):int(offsets[i + 1])], O.maps_to_rows(m[None]))
    q = torch.from_numpy(maps[3][:, 1, 5]).reshape(1, -1).cuda()
    s, idx = store.search(q, 2)
    img, cell = F.rows_to_image_cell(idx[0, :1].cpu(), offsets)
    assert (int(img), int(cell)) == (3, 1 * 7 + 5)
    pooled, poff = F.store_from_mixed_blobs(records, pool="mean")
    assert poff.tolist() == list(range(7)) and len(pooled) == 6
    ref = np.stack([O.maps_to_rows(m[None], pool="mean")[0] for m in maps])
    assert np.abs(pooled.embeddings.float().cpu().numpy() - ref).max() <= np.abs(ref).max() * 2.0**-7
    with pytest.raises(ValueError):
        F.store_from_mixed_blobs(records + [(O.blob_encode(maps[0][:8]), 8, 3, 4)])


def test_search_host_batches_streams_ragged_batches():
    """Streamed search from pinned host batches (copy stream around the search) returns, batch by
    batch, what the direct search returns; ragged and unpinned batches, and an early exit."""
    store, queries = make(20000, 1000, 128, seed=5)
    st = S.EmbeddingStore(torch.from_numpy(store).cuda())
    cuts = [0, 300, 600, 617, 617 + 256, 1000]
    host = [torch.from_numpy(queries[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    host = [h.pin_memory() if j % 2 == 0 else h for j, h in enumerate(host)]
    want_s, want_i = st.search_raw(torch.from_numpy(queries).cuda(), 10)
    got = []
    for s_h, i_h in S.search_host_batches(st, host, 10):
        assert s_h.device.type == "cpu" and s_h.is_pinned() and i_h.dtype == torch.int32
        got.append((s_h.clone(), i_h.clone()))
    assert [g[0].shape[0] for g in got] == [b - a for a, b in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat([g[1] for g in got]), want_i.cpu())
    assert torch.equal(torch.cat([g[0] for g in got]), want_s.cpu())
    assert list(S.search_host_batches(st, [], 10)) == []
    gen = S.search_host_batches(st, host, 10)
    first = next(gen)
    assert torch.equal(first[1], want_i[:300].cpu())
    gen.close()  # abandoning the stream mid-way must leave nothing in flight
    with pytest.raises(ValueError):
        list(S.search_host_batches(st, [torch.zeros(5)], 10))
    check(store, queries[:300], 10, got[0][0], got[0][1].to(torch.int64))
    # a persistent streamer keeps its staging buffers across runs
    hs = S.HostBatchSearch(st, 10)
    for _ in range(2):
        again = [(s_h.clone(), i_h.clone()) for s_h, i_h in hs.run(host)]
        assert torch.equal(torch.cat([g[1] for g in again]), want_i.cpu())
