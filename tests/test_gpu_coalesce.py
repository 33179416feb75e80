"""GPU tests for the coalescing front door (pqv_l2_topk_coalesced, SURVEY section 8b "Threading"): many threads issue
single-query calls; whatever batches form, every caller must receive exactly its own single-query result
(src/ivf/search.rs:112-141), bit for bit."""
import threading

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SQRT, SEQ = 2, 1


@pytest.fixture(scope="module")
def ctx():
    import pq_vector_b200 as P
    c = P.Context()
    yield c
    c.close()


def run_threads(n_threads, fn):
    out, errs = [None] * n_threads, []
    gate = threading.Barrier(n_threads)

    def work(i):
        try:
            gate.wait()
            out[i] = fn(i)
        except Exception as e:  # noqa: BLE001
            errs.append((i, e))

    th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return out, errs


@pytest.mark.parametrize("window_us", [0, 30000])
def test_concurrent_callers_get_their_own_result(ctx, window_us):
    n, dim, k, nthreads, per = 60000, 256, 10, 48, 3
    data = O.synth(n, dim, 1234)
    queries = O.synth(nthreads * per, dim, 7)
    ds = ctx.dataset_from(data)
    ctx.coalesce_config(1024, window_us)
    before = ctx.coalesce_stats()

    def fn(i):
        return [ds.l2_topk_coalesced(queries[i * per + j], k, SQRT) for j in range(per)]

    out, errs = run_threads(nthreads, fn)
    assert not errs, errs
    after = ctx.coalesce_stats()
    assert after["queries"] - before["queries"] == nthreads * per
    assert after["batches"] - before["batches"] <= nthreads * per
    if window_us:
        assert after["max_batch"] >= 4, after       # the lingering leader saw the burst
    for i in range(nthreads):
        for j in range(per):
            q = queries[i * per + j]
            er, ed = O.topk_rerank(q, data, None, k, 0, True)
            r, d = out[i][j]
            assert r.tolist() == er.tolist(), (i, j)
            assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist(), (i, j)
    ctx.coalesce_config(1024, 0)
    ds.drop()


def test_mixed_parameters_are_not_mixed_into_one_batch(ctx):
    n, dim, nthreads = 20000, 128, 24
    data = O.synth(n, dim, 99)
    queries = O.synth(nthreads, dim, 3)
    ds = ctx.dataset_from(data)
    ctx.coalesce_config(8, 20000)     # small max_batch: a burst is split over several leaders
    cfg = [(10, SQRT), (100, SQRT), (10, SEQ)]

    def fn(i):
        k, flags = cfg[i % 3]
        return ds.l2_topk_coalesced(queries[i], k, flags)

    out, errs = run_threads(nthreads, fn)
    assert not errs, errs
    for i in range(nthreads):
        k, flags = cfg[i % 3]
        er, ed = O.topk_rerank(queries[i], data, None, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
        assert out[i][0].tolist() == er.tolist(), i
        assert out[i][1].view(np.uint32).tolist() == ed.view(np.uint32).tolist(), i
    assert ctx.coalesce_stats()["max_batch"] <= 1024
    ctx.coalesce_config(1024, 0)
    ds.drop()


def test_errors_stay_with_their_caller(ctx):
    import pq_vector_b200 as P
    ds = ctx.dataset_from(O.synth(100, 8, 1))
    with pytest.raises(P.PqvError, match="k must be"):
        ds.l2_topk_coalesced(np.zeros(8, np.float32), 0)
    r, d = ds.l2_topk_coalesced(np.zeros(8, np.float32), 5)      # the front door still works afterwards
    assert r.size == 5
    ds.drop()


def test_concurrent_ivf_searches_are_coalesced(ctx):
    """TopkBuilder::search callers (search.rs:76-80 is async): concurrent single-query IVF searches through
    pqv_ivf_search_coalesced get exactly their own result; brute-force and IVF requests never share a batch."""
    n, dim, C, k, nprobe, nthreads = 50000, 64, 40, 10, 5, 40
    rng = np.random.default_rng(12)
    data = rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy() + 0.01
    offsets, ids = O.inverted_lists(O.assign(data, cent, workers=2), C)
    ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    ds = ctx.dataset_from(data)
    queries = rng.random((nthreads, dim), dtype=np.float32)
    ctx.coalesce_config(1024, 30000)
    before = ctx.coalesce_stats()

    def fn(i):
        if i % 4 == 3:
            return ds.l2_topk_coalesced(queries[i], k, SQRT)
        return ix.search_coalesced(ds, queries[i], k, nprobe, SQRT)

    out, errs = run_threads(nthreads, fn)
    assert not errs, errs
    after = ctx.coalesce_stats()
    assert after["queries"] - before["queries"] == nthreads and after["max_batch"] >= 4
    for i in range(nthreads):
        if i % 4 == 3:
            er, ed = O.topk_rerank(queries[i], data, None, k, 0, True)
        else:
            er, ed = O.topk_rerank_gather(queries[i], data, O.candidate_rows(queries[i], cent, offsets, ids, nprobe), k, 0, True)
        assert out[i][0].tolist() == er.tolist(), i
        assert out[i][1].view(np.uint32).tolist() == ed.view(np.uint32).tolist(), i
    import pq_vector_b200 as P
    with pytest.raises(P.PqvError, match="nprobe must be > 0"):
        ix.search_coalesced(ds, queries[0], k, 0)
    ctx.coalesce_config(1024, 0)
    ix.drop()
    ds.drop()
