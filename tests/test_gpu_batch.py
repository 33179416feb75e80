"""GPU parity tests for the batched brute-force top-k (tcgen05 filter + exact re-rank, pqv_tc.cuh): every query of a batch
must return exactly what its own single-query reference loop returns (src/ivf/search.rs:112-141 with PQV_SQRT,
src/df_vector/exec.rs:257-277 with PQV_SUM_SEQ) -- row ids, order and distance bits."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SQRT, SEQ, BYPOS = 2, 1, 4


@pytest.fixture(scope="module")
def ctx():
    import pq_vector_b200 as P
    c = P.Context()
    yield c
    c.close()


class batch_off:
    def __enter__(self):
        self.old = os.environ.get("PQV_BATCH")
        os.environ["PQV_BATCH"] = "off"

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("PQV_BATCH", None)
        else:
            os.environ["PQV_BATCH"] = self.old


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def check_batch(ctx, data, queries, k, flags, expect_used=True, nan_payload_free=False):
    ds = ctx.dataset_from(data)
    rows, dist, cnt = ds.l2_topk(queries, k, flags)
    t = ctx.last_batch_timing()
    order = 1 if flags & SEQ else 0
    for i, q in enumerate(queries):
        er, ed = O.topk_rerank(q, data, None, k, order, bool(flags & SQRT))
        assert cnt[i] == er.size, (i, cnt[i], er.size, t)
        assert rows[i, :cnt[i]].tolist() == er.tolist(), (i, t)
        if nan_payload_free and np.isnan(ed).any():   # a NaN query: every distance is NaN, payload bits are not pinned
            assert np.isnan(dist[i, :cnt[i]]).all()
            continue
        assert bits(dist[i, :cnt[i]]).tolist() == bits(ed).tolist(), (i, t)
    if expect_used:
        assert t["queries"] == len(queries) and t["declined"] == 0, t
    ds.drop()
    return t


@pytest.mark.parametrize("n,dim,nq,k,flags", [
    (20000, 768, 24, 10, SQRT), (20000, 768, 24, 100, SQRT), (30000, 128, 300, 10, SEQ), (5001, 100, 7, 33, SQRT),
    (4096, 32, 4, 1, SEQ | 0), (3000, 1536, 40, 100, SEQ), (700, 64, 9, 1024, SQRT), (50, 64, 5, 100, SQRT),
])
def test_batch_matches_the_single_query_reference(ctx, n, dim, nq, k, flags):
    rng = np.random.default_rng(n + dim + nq + k)
    data = rng.random((n, dim), dtype=np.float32)
    queries = rng.random((nq, dim), dtype=np.float32)
    queries[0] = data[n // 2]          # an exact hit (distance 0)
    t = check_batch(ctx, data, queries, k, flags)
    assert t["tie_queries"] <= nq


def test_clustered_and_unit_norm_data(ctx):
    rng = np.random.default_rng(3)
    cent = rng.standard_normal((50, 256)).astype(np.float32)
    data = (cent[rng.integers(0, 50, 40000)] + 0.05 * rng.standard_normal((40000, 256))).astype(np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    queries = data[rng.integers(0, 40000, 64)] + 0.01 * rng.standard_normal((64, 256)).astype(np.float32)
    check_batch(ctx, data, queries.astype(np.float32), 20, SQRT)
    check_batch(ctx, data, queries.astype(np.float32), 20, SEQ)


def test_duplicate_rows_fall_back_to_the_heap_replay(ctx):
    """bit-equal distances inside the top-k: the order is the reference heap's layout -> those queries are re-run exactly"""
    rng = np.random.default_rng(4)
    base = rng.random((3000, 64), dtype=np.float32)
    data = np.concatenate([base, base[:1500], base[:700]])
    queries = rng.random((12, 64), dtype=np.float32)
    t = check_batch(ctx, data, queries, 50, SQRT)
    assert t["tie_queries"] >= 1
    # grid data: many exactly equal distances
    grid = rng.integers(0, 3, (20000, 32)).astype(np.float32)
    tq = check_batch(ctx, grid, rng.integers(0, 3, (8, 32)).astype(np.float32), 10, SEQ)
    assert tq["tie_queries"] >= 1


def test_ties_by_position_flag_needs_no_fallback(ctx):
    rng = np.random.default_rng(5)
    grid = rng.integers(0, 3, (20000, 32)).astype(np.float32)
    queries = rng.integers(0, 3, (8, 32)).astype(np.float32)
    ds = ctx.dataset_from(grid)
    rows, dist, cnt = ds.l2_topk(queries, 10, SEQ | BYPOS)
    t = ctx.last_batch_timing()
    assert t["queries"] == 8 and t["tie_queries"] == 0 and t["declined"] == 0
    with batch_off():
        r1, d1, c1 = ds.l2_topk(queries, 10, SEQ | BYPOS)
    assert np.array_equal(rows, r1) and np.array_equal(bits(dist), bits(d1)) and np.array_equal(cnt, c1)
    ds.drop()


def test_non_finite_inputs_decline_the_batch(ctx):
    rng = np.random.default_rng(6)
    data = rng.random((6000, 64), dtype=np.float32)
    queries = rng.random((6, 64), dtype=np.float32)
    bad = data.copy()
    bad[17, 3] = np.inf
    bad[4000, 0] = np.nan
    t = check_batch(ctx, bad, queries, 10, SQRT, expect_used=False)
    assert t["declined"] == 1
    qbad = queries.copy()
    qbad[2, 5] = np.nan
    t = check_batch(ctx, data, qbad, 10, SQRT, expect_used=False, nan_payload_free=True)
    assert t["declined"] == 1


def test_batch_equals_single_scans_at_scale(ctx):
    n, dim, nq, k = 400_000, 768, 200, 100
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, 1234)
    qd = ctx.dataset(dim, nq)
    qd.fill_synthetic(nq, 7)
    queries = qd.read(0, nq)
    qd.drop()
    for flags in (SQRT, SEQ):
        rows, dist, cnt = ds.l2_topk(queries, k, flags)
        t = ctx.last_batch_timing()
        assert t["queries"] == nq and t["declined"] == 0, t
        with batch_off():
            r1, d1, c1 = ds.l2_topk(queries, k, flags)
        assert ctx.last_batch_timing()["queries"] == 0
        assert np.array_equal(cnt, c1)
        assert np.array_equal(rows, r1)
        assert np.array_equal(bits(dist), bits(d1))
        print(flags, t)
    ds.drop()


def test_small_batches_use_the_single_query_path(ctx):
    rng = np.random.default_rng(8)
    data = rng.random((5000, 64), dtype=np.float32)
    ds = ctx.dataset_from(data)
    ds.l2_topk(rng.random((3, 64), dtype=np.float32), 5, SQRT)
    assert ctx.last_batch_timing()["queries"] == 0
    ds.drop()


@pytest.mark.parametrize("n,dim,nq,k,flags,grid_data", [
    (30000, 256, 32, 10, SEQ, False), (20001, 768, 12, 100, SQRT, False), (9000, 64, 16, 10, SQRT, True),
    (600, 32, 6, 1000, SEQ, False),
])
def test_sharded_batch_keys_merge_to_the_whole_table_answer(ctx, n, dim, nq, k, flags, grid_data):
    """Config C5's per-rank half (pqv_l2_topk_batch_keys) + pqv_merge_batch_keys, three uneven 'ranks' in one process:
    final queries must equal the single-query reference loop over the WHOLE table; queries flagged for replay must be
    exactly answerable by the candidate exchange (pqv_l2_topk_candidates + pqv_replay_candidates)."""
    import pq_vector_b200 as P
    rng = np.random.default_rng(n + k)
    if grid_data:   # small-integer grid: exact ties across slices
        data = rng.integers(0, 3, (n, dim)).astype(np.float32)
        queries = rng.integers(0, 3, (nq, dim)).astype(np.float32)
    else:
        data = rng.random((n, dim), dtype=np.float32)
        queries = rng.random((nq, dim), dtype=np.float32)
        queries[0] = data[n - 1]
    cuts = [0, n // 5, n // 5 + n // 2, n]
    parts = [ctx.dataset_from(data[cuts[i]:cuts[i + 1]]) for i in range(3)]
    keys, counts, tie_cand = [], [], []
    for i, p in enumerate(parts):
        kk, cc = p.l2_topk_batch_keys(queries, k, flags, cuts[i])
        keys.append(kk)
        counts.append(cc)
        # what the pass left on the device for flagged queries (only valid until the next batched call)
        tie_cand.append([p.l2_topk_batch_tie_candidates(qi, q, cap=64) for qi, q in enumerate(queries)])
    keys, counts = np.stack(keys), np.stack(counts)
    assert (counts != 0xFFFFFFFF).all()
    with pytest.raises(P.PqvError):   # the state belongs to the last batched call only
        parts[0].l2_topk_batch_tie_candidates(0, queries[0])
    rows, dist, cnt, need = P.merge_batch_keys(keys, counts, k, flags)
    order = 1 if flags & SEQ else 0
    if grid_data:
        assert need.any()
    else:
        assert not need.all()
    for i, q in enumerate(queries):
        er, ed = O.topk_rerank(q, data, None, k, order, bool(flags & SQRT))
        if need[i]:
            cand = np.concatenate([p.l2_topk_candidates(q, k, flags, cuts[j], cap=1 << 20) for j, p in enumerate(parts)])
            r, d = P.replay_candidates(cand, k, flags)
            r2, d2 = P.replay_candidates(np.concatenate([tc[i] for tc in tie_cand]), k, flags)
            assert r2.tolist() == er.tolist() and bits(d2).tolist() == bits(ed).tolist(), i
        else:
            r, d = rows[i, :cnt[i]], dist[i, :cnt[i]]
        assert r.tolist() == er.tolist(), (i, bool(need[i]))
        assert bits(d).tolist() == bits(ed).tolist(), (i, bool(need[i]))
    for p in parts:
        p.drop()


def test_more_queries_than_one_pass_takes(ctx):
    """n_queries > BATCH_MAX_QUERIES (4096): several passes, a short tail through the single-query scans."""
    rng = np.random.default_rng(4)
    n, dim, nq, k = 6000, 32, 4096 + 4096 + 3, 5
    data = rng.random((n, dim), dtype=np.float32)
    queries = rng.random((nq, dim), dtype=np.float32)
    ds = ctx.dataset_from(data)
    rows, dist, cnt = ds.l2_topk(queries, k, SQRT)
    assert ctx.last_batch_timing()["queries"] == 8192
    for i in list(range(0, nq, 397)) + [4095, 4096, 8191, 8192, nq - 1]:
        er, ed = O.topk_rerank(queries[i], data, None, k, 0, True)
        assert cnt[i] == k and rows[i].tolist() == er.tolist(), i
        assert bits(dist[i]).tolist() == bits(ed).tolist()
    ds.drop()


class tie_batch_off:
    """PQV_TIE_BATCH=off: every tie query takes the one-by-one path (an exact scan of the sample prefix each)"""
    def __enter__(self):
        self.old = os.environ.get("PQV_TIE_BATCH")
        os.environ["PQV_TIE_BATCH"] = "off"

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("PQV_TIE_BATCH", None)
        else:
            os.environ["PQV_TIE_BATCH"] = self.old


class nullcontext:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def test_tie_queries_resolved_together(ctx):
    """The tie queries of a batch share ONE exact pass over the sample prefix (pqv_tie.cuh) instead of one
    prefix scan each; every answer must still be the reference loop's, bit for bit (row order among equal distances
    included), in both summation orders and for prefixes shorter and longer than one 2048-row chunk."""
    rng = np.random.default_rng(14)
    with nullcontext():
        base = rng.random((3000, 64), dtype=np.float32)
        data = np.concatenate([base, base[:1500], base[:700]])
        queries = rng.random((12, 64), dtype=np.float32)
        for flags, k in ((SQRT, 50), (SEQ, 50), (SQRT, 1), (SEQ, 1024)):
            t = check_batch(ctx, data, queries, k, flags)
            assert t["tie_queries"] >= 1 and t["tie_batched"] == t["tie_queries"], t
        # small-integer grid: many exactly equal distances, ties across the k boundary; prefix = several chunks
        grid = rng.integers(0, 3, (70000, 32)).astype(np.float32)
        gq = rng.integers(0, 3, (8, 32)).astype(np.float32)
        for flags in (SEQ, SQRT):
            t = check_batch(ctx, grid, gq, 10, flags)
            assert t["tie_queries"] >= 1 and t["tie_batched"] == t["tie_queries"], t
        # a table shorter than k and shorter than one chunk; dim not a multiple of 32 (zero-padded last column block)
        small = rng.integers(0, 2, (40, 36)).astype(np.float32)
        t = check_batch(ctx, small, rng.integers(0, 2, (5, 36)).astype(np.float32), 100, SQRT)
        assert t["tie_batched"] == t["tie_queries"], t
        # descending distances: every row enters the heap -> the entrant region overflows -> one-by-one path, same answer
        n = 30000
        ramp = np.zeros((n, 32), np.float32)
        ramp[:, 0] = np.arange(n, 0, -1, dtype=np.float32) / np.float32(n)
        ramp[-7:, 0] = ramp[-8, 0]              # bit-equal distances inside the top-k
        t = check_batch(ctx, ramp, np.zeros((4, 32), np.float32), 10, SQRT, expect_used=False)
        if t["queries"] and not t["declined"]:
            assert t["tie_queries"] == 4 and t["tie_batched"] == 0, t


def test_tie_batch_equals_one_by_one_at_scale(ctx):
    """400k x 768 synthetic rows, 200 queries, k = 100 with sqrt (about 5 % of the queries collapse two squared distances
    onto one returned value): resolving the ties together returns what the one-by-one path returns."""
    n, dim, nq, k = 400_000, 768, 200, 100
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, 1234)
    qd = ctx.dataset(dim, nq)
    qd.fill_synthetic(nq, 7)
    queries = qd.read(0, nq)
    qd.drop()
    for flags in (SQRT, SEQ):
        with tie_batch_off():
            r0, d0, c0 = ds.l2_topk(queries, k, flags)
            t0 = ctx.last_batch_timing()
        r1, d1, c1 = ds.l2_topk(queries, k, flags)
        t1 = ctx.last_batch_timing()
        assert t0["tie_batched"] == 0 and t1["tie_batched"] == t1["tie_queries"] == t0["tie_queries"], (t0, t1)
        assert np.array_equal(c0, c1) and np.array_equal(r0, r1) and np.array_equal(bits(d0), bits(d1))
        print(flags, t1)
    ds.drop()
