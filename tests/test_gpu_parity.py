"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SQRT, SEQ, BYPOS = 2, 1, 4


@pytest.fixture(scope="module")
def P():
    import pq_vector_b200 as P
    return P


@pytest.fixture(scope="module")
def ctx(P):
    c = P.Context()
    yield c
    c.close()


def bits(a):
    """bit patterns, with every NaN mapped to one pattern (the payload of a propagated NaN is hardware-specific)"""
    a = np.asarray(a, np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))


def check_topk(ds, data, q, k, flags, row_ids=None):
    order = 1 if flags & SEQ else 0
    do_sqrt = bool(flags & SQRT)
    if row_ids is None:
        r, d = ds.l2_topk(q, k, flags)
        er, ed = O.topk_rerank(q, data, None, k, order, do_sqrt)
    else:
        r, d = ds.l2_topk_gather(q, row_ids, k, flags)
        er, ed = O.topk_rerank_gather(q, data, row_ids, k, order, do_sqrt)
    assert r.tolist() == er.tolist()
    assert bits(d).tolist() == bits(ed).tolist()


# ---- C1: the reference's own dataset ---------------------------------------------------------------
VLDB_IDS = {0: [0, 126, 81, 265, 315, 464, 322, 269, 169, 140], 1: [1, 177, 57, 19, 36, 16, 450, 179, 9, 140],
            100: [100, 181, 400, 352, 448, 476, 36, 198, 370, 213]}


def test_c1_vldb_top10(ctx, vldb):
    ds = ctx.dataset_from(vldb)
    for qrow, ids in VLDB_IDS.items():
        r, d = ds.l2_topk(vldb[qrow], 10, SQRT)
        assert r.tolist() == ids
        check_topk(ds, vldb, vldb[qrow], 10, SQRT)
        check_topk(ds, vldb, vldb[qrow], 10, SEQ)          # VectorTopKExec order, squared distances
    for qrow in range(0, 496, 37):                         # duplicates in the data included
        check_topk(ds, vldb, vldb[qrow], 10, SQRT)
        check_topk(ds, vldb, vldb[qrow], 496, SQRT)        # k == n
        check_topk(ds, vldb, vldb[qrow], 600, SEQ)         # k > n -> 496 results
    ds.drop()


def test_reference_kats_through_the_abi(ctx):
    # src/df_vector/tests.rs:31-39,99 and :166-174,235 (filter applied by the caller: ids >= min_id)
    for rows, min_id, expect in [([(0, 0), (1, 0), (0, 2), (5, 5), (2, 2), (0.1, 0.1)], 2, [5, 2]),
                                 ([(0, 0), (.05, .05), (.2, .2), (1, 1), (1.1, 1.1), (1.4, 1.4)], 3, [3, 4])]:
        rows = np.array(rows, np.float32)
        ds = ctx.dataset_from(rows)
        r, _ = ds.l2_topk_gather(np.zeros(2, np.float32), np.arange(min_id, 6, dtype=np.uint32), 2, SEQ)
        assert r.tolist() == expect
        st = ctx.topk_stream(np.zeros(2, np.float32), 2, SEQ)
        st.push(rows[min_id:])
        r, _ = st.finish()
        assert (r + min_id).tolist() == expect
        ds.drop()
    # src/ivf/index.rs:487-493
    ds = ctx.dataset_from(np.array([[4, 5, 6]], np.float32))
    _, d = ds.l2_topk(np.array([1, 2, 3], np.float32), 1, 0)
    assert d.tolist() == [27.0]
    ds.drop()


# ---- seeded random inputs vs oracle ------------------------------------------------------------------
@pytest.mark.parametrize("n,dim", [(1, 4), (31, 8), (32, 3), (33, 5), (257, 1), (1000, 7), (4096, 128), (5000, 130),
                                   (3000, 768), (2000, 1000), (1500, 1536), (700, 4096), (50_000, 64), (70, 16384)])
def test_bruteforce_matches_oracle(ctx, n, dim):
    rng = np.random.default_rng(n * 31 + dim)
    data = rng.random((n, dim), dtype=np.float32)
    ds = ctx.dataset_from(data)
    for k in (1, 10, 100, 1024):
        q = rng.random(dim, dtype=np.float32)
        for flags in (SQRT, SEQ, 0, SEQ | SQRT):
            check_topk(ds, data, q, k, flags)
    ds.drop()


def test_heavy_ties_follow_the_reference_heap(ctx):
    # quantised coordinates -> thousands of bit-equal distances; order must equal the BinaryHeap replay
    rng = np.random.default_rng(99)
    for n, dim, levels in [(20_000, 4, 3), (5000, 8, 2), (70_000, 2, 5)]:
        data = rng.integers(0, levels, (n, dim)).astype(np.float32)
        ds = ctx.dataset_from(data)
        q = np.zeros(dim, np.float32)
        for k in (1, 7, 100, 1000):
            check_topk(ds, data, q, k, SQRT)
            check_topk(ds, data, q, k, SEQ)
        ds.drop()


def test_adversarial_descending_order(ctx):
    # every row improves on all earlier ones: every row enters the reference heap -> entrant buffer regrow path
    n, dim = 200_000, 4
    data = np.zeros((n, dim), np.float32)
    data[:, 0] = np.linspace(1000, 1, n, dtype=np.float32)
    ds = ctx.dataset_from(data)
    q = np.zeros(dim, np.float32)
    check_topk(ds, data, q, 10, SQRT)
    assert ctx.last_timing()["entrants"] == n
    check_topk(ds, data, q, 100, SEQ)
    ds.drop()


def test_non_finite_distances(ctx):
    rng = np.random.default_rng(5)
    data = rng.random((1000, 8), dtype=np.float32)
    data[500:, 3] = np.inf            # distance +inf: only admitted while the heap is not full (search.rs:119)
    ds = ctx.dataset_from(data)
    q = rng.random(8, dtype=np.float32)
    for k in (10, 600):
        check_topk(ds, data, q, k, SQRT)
    ds.drop()


def test_nan_distances_follow_the_reference_loop(ctx):
    """A NaN distance inside the reference heap (search.rs:119-126: `d < top` is false against a NaN root, and
    partial_cmp -> Equal keeps a NaN wherever the sifts leave it) changes which later rows are admitted.  The GPU path
    answers such a query by replaying the loop over every candidate; rows with a NaN coordinate sit below k, right at
    k, and far behind it."""
    rng = np.random.default_rng(77)
    for n, dim, nan_rows in [(5000, 8, [3]), (5000, 8, [0, 1, 2]), (3000, 16, [9, 10, 2999]), (40_000, 4, [5, 700, 20_000]),
                             (2000, 64, list(range(0, 2000, 97))), (300, 8, list(range(300)))]:
        data = rng.random((n, dim), dtype=np.float32)
        for r in nan_rows:
            data[r, rng.integers(0, dim)] = np.nan
        ds = ctx.dataset_from(data)
        q = rng.random(dim, dtype=np.float32)
        for k in (1, 4, 10, 100):
            for flags in (SQRT, SEQ, 0):
                check_topk(ds, data, q, k, flags)
        ids = rng.permutation(n)[: n // 2].astype(np.uint32)
        ids[:3] = nan_rows[0]                                 # the NaN row among the first candidates (and repeated)
        for k in (2, 10, 64):
            check_topk(ds, data, q, k, SQRT, row_ids=ids)
            check_topk(ds, data, q, k, SEQ, row_ids=ids)
        # the streaming form (VectorTopKExec heap, exec.rs:467-482), NaN rows spread over several pushes
        for k in (3, 10, 100):
            st = ctx.topk_stream(q, k, SEQ)
            for a in range(0, n, 1111):
                st.push(data[a:a + 1111])
            r, d = st.finish()
            er, ed = O.topk_rerank(q, data, None, k, 1, False)
            assert r.tolist() == er.tolist() and bits(d).tolist() == bits(ed).tolist()
        ds.drop()
    # a NaN in the query: every distance is NaN, the heap keeps the first k rows it was given
    data = rng.random((500, 8), dtype=np.float32)
    ds = ctx.dataset_from(data)
    q = rng.random(8, dtype=np.float32)
    q[2] = np.nan
    for k in (1, 10, 600):
        check_topk(ds, data, q, k, SQRT)
        check_topk(ds, data, q, k, SEQ)
    ds.drop()


def test_gather_matches_oracle(ctx):
    rng = np.random.default_rng(11)
    data = rng.random((20_000, 96), dtype=np.float32)
    ds = ctx.dataset_from(data)
    q = rng.random(96, dtype=np.float32)
    for m in (1, 33, 5000, 20_000):
        ids = rng.permutation(20_000)[:m].astype(np.uint32)
        for k in (1, 10, 100):
            check_topk(ds, data, q, k, SQRT, ids)
            check_topk(ds, data, q, k, SEQ, ids)
    ids = rng.integers(0, 20_000, 3000).astype(np.uint32)     # duplicates allowed
    check_topk(ds, data, q, 50, SQRT, ids)
    r, d = ds.l2_topk_gather(q, np.zeros(0, np.uint32), 5, SQRT)
    assert r.size == 0
    ds.drop()


def test_k_above_the_in_kernel_selection(ctx):
    """the reference takes any k (search.rs:112-141); above PQV_MAX_K = 1024 the library logs every candidate's distance and
    replays the reference loop on the host: dense, gathered, quantised data (ties), both summation orders, both tie modes"""
    rng = np.random.default_rng(77)
    data = rng.random((6000, 24), dtype=np.float32)
    ds = ctx.dataset_from(data)
    q = rng.random(24, dtype=np.float32)
    for k in (1025, 2000, 5999, 6000, 9000):
        check_topk(ds, data, q, k, SQRT)
        check_topk(ds, data, q, k, SEQ)
    ids = rng.permutation(6000)[:3000].astype(np.uint32)
    check_topk(ds, data, q, 1500, SQRT, ids)
    check_topk(ds, data, q, 4000, SEQ | SQRT, ids)
    ds.drop()
    grid = rng.integers(0, 3, (20_000, 4)).astype(np.float32)   # thousands of bit-equal distances
    ds = ctx.dataset_from(grid)
    q0 = np.zeros(4, np.float32)
    check_topk(ds, grid, q0, 3000, SQRT)
    check_topk(ds, grid, q0, 1100, SEQ)
    r, d = ds.l2_topk(q0, 2500, BYPOS)
    dist = O.distances(grid, q0, 0)
    order = np.lexsort((np.arange(20_000), dist))[:2500]
    assert r.tolist() == order.tolist() and bits(d).tolist() == bits(dist[order]).tolist()
    # several queries per call take the same path one by one
    qs = rng.integers(0, 3, (5, 4)).astype(np.float32)
    rows, dd, cnt = ds.l2_topk(qs, 1200, SQRT)
    for i in range(5):
        er, ed = O.topk_rerank(qs[i], grid, None, 1200, 0, True)
        assert rows[i, :cnt[i]].tolist() == er.tolist() and bits(dd[i, :cnt[i]]).tolist() == bits(ed).tolist()
    ds.drop()


def test_ties_by_position_mode(ctx):
    rng = np.random.default_rng(12)
    data = rng.integers(0, 3, (10_000, 4)).astype(np.float32)
    ds = ctx.dataset_from(data)
    q = np.zeros(4, np.float32)
    r, d = ds.l2_topk(q, 100, BYPOS)
    dist = O.distances(data, q, 0)
    order = np.lexsort((np.arange(10_000), dist))[:100]
    assert r.tolist() == order.tolist()
    assert bits(d).tolist() == bits(dist[order]).tolist()
    ds.drop()


def test_synthetic_fill_matches_oracle_stream(ctx):
    ds = ctx.dataset(48, 5000)
    ds.fill_synthetic(5000, 1234)
    got = ds.read(0, 5000)
    assert np.array_equal(got, O.synth(5000, 48, 1234))
    assert np.array_equal(ds.read(1234, 10), O.synth(10, 48, 1234, first_row=1234))
    ds.drop()


# ---- streaming (VectorTopKExec) ----------------------------------------------------------------------
def test_stream_matches_oracle(ctx):
    rng = np.random.default_rng(21)
    dim = 64
    batches = [rng.random((m, dim), dtype=np.float32) for m in (100, 8192, 1, 33, 8192, 5000)]
    allrows = np.vstack(batches)
    q = rng.random(dim, dtype=np.float32)
    for k, flags in [(10, SEQ), (100, SEQ), (5, SQRT), (1000, SEQ)]:
        st = ctx.topk_stream(q, k, flags)
        for b in batches:
            st.push(b)
        r, d = st.finish()
        er, ed = O.topk_rerank(q, allrows, None, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
        assert r.tolist() == er.tolist() and bits(d).tolist() == bits(ed).tolist()
    # ties across batches + f64 batches narrowed to f32 first (exec.rs:542)
    tb = [rng.integers(0, 3, (m, 4)).astype(np.float64) for m in (3000, 3000, 10)]
    st = ctx.topk_stream(np.zeros(4, np.float32), 50, SEQ)
    for b in tb:
        st.push(b)
    r, d = st.finish()
    er, ed = O.topk_rerank(np.zeros(4, np.float32), np.vstack(tb).astype(np.float32), None, 50, 1, False)
    assert r.tolist() == er.tolist() and bits(d).tolist() == bits(ed).tolist()
    st = ctx.topk_stream(q, 3, SEQ)          # nothing pushed -> empty result (exec.rs:553-555)
    r, d = st.finish()
    assert r.size == 0


# ---- k-means pieces ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dim,c", [(1, 4, 1), (100, 3, 5), (1000, 8, 64), (777, 130, 100), (3000, 768, 256),
                                     (500, 7, 3), (2048, 64, 1024)])
def test_kmeans_assign_matches_oracle(ctx, n, dim, c):
    rng = np.random.default_rng(n + dim + c)
    data = rng.random((n, dim), dtype=np.float32)
    cent = data[rng.integers(0, n, c)].copy()        # duplicates likely -> first-min tie rule exercised
    got, sizes = ctx.kmeans_assign(data, cent, want_sizes=True)
    exp = O.assign(data, cent, workers=8)
    assert np.array_equal(got, exp)
    assert np.array_equal(sizes, np.bincount(exp, minlength=c).astype(np.uint64))
    ds = ctx.dataset_from(data)                      # resident-dataset form
    assert np.array_equal(ctx.kmeans_assign(ds, cent), exp)
    ds.drop()


def test_kmeans_assign_ties_and_nan(ctx):
    cent = np.array([[1, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [5, 5, 0, 0]], np.float32)
    data = np.array([[0, 0, 0, 0], [np.nan, 0, 0, 0], [np.inf, 0, 0, 0], [4, 4, 0, 0], [1, 1, 0, 0]], np.float32)
    got = ctx.kmeans_assign(data, cent)
    assert got.tolist() == [1, 0, 0, 3, 0] == O.assign(data, cent).tolist()


def test_min_dist_update_matches_oracle(ctx):
    rng = np.random.default_rng(31)
    data = rng.random((5000, 100), dtype=np.float32)
    sel = rng.permutation(5000)[:1777].astype(np.uint64)
    md = ctx.min_dist_update(data, sel, data[3])
    emd = O.min_dist_init(data, sel, data[3])
    assert bits(md).tolist() == bits(emd).tolist()
    for c in (10, 20, 30):
        ctx.min_dist_update(data, sel, data[c], md)
        O.min_dist_update(data, sel, data[c], emd, workers=4)
        assert bits(md).tolist() == bits(emd).tolist()
    ds = ctx.dataset_from(data)                      # resident form, no selection
    md2 = ctx.min_dist_update(ds, None, data[3])
    assert bits(md2).tolist() == bits(O.min_dist_init(data, None, data[3])).tolist()
    ds.drop()


def test_centroid_rank_matches_oracle(ctx):
    rng = np.random.default_rng(41)
    cent = rng.random((300, 72), dtype=np.float32)
    cent[50] = cent[10]
    cent[200] = cent[10]                              # exact ties -> stable order by index
    qs = rng.random((5, 72), dtype=np.float32)
    for nprobe in (1, 16, 300, 1000):
        got = ctx.centroid_rank(cent, qs, nprobe)
        for i in range(5):
            assert got[i].tolist() == O.find_closest_centroids(qs[i], cent, nprobe).tolist()
    assert ctx.centroid_rank(cent, cent[10], 3)[0].tolist() == [10, 50, 200]


@pytest.mark.parametrize("C,dim,nq", [(1024, 768, 256), (23, 4096, 40), (777, 50, 33), (4097, 8, 7)])
def test_batched_centroid_rank(ctx, C, dim, nq):
    """Row a6 batched (config C5): one launch for all queries x centroids, ranked on the device -- identical to the
    per-query reference ranking (index.rs:130-149), incl. exact ties (stable by index) and NaN centroids (host comparator)."""
    import os
    rng = np.random.default_rng(C + dim)
    cent = rng.random((C, dim), dtype=np.float32)
    cent[C // 2] = cent[3]
    cent[C - 1] = cent[3]
    qs = rng.random((nq, dim), dtype=np.float32)
    qs[1] = cent[3]
    for nprobe in (1, 32, C):
        got = ctx.centroid_rank(cent, qs, nprobe)
        for i in range(0, nq, max(1, nq // 16)):
            assert got[i].tolist() == O.find_closest_centroids(qs[i], cent, nprobe).tolist(), (i, nprobe)
    assert ctx.centroid_rank(cent, qs, 3)[1].tolist() == [3, C // 2, C - 1]
    os.environ["PQV_RANK_BATCH_OFF"] = "1"
    try:
        one_by_one = ctx.centroid_rank(cent, qs, 32)
    finally:
        del os.environ["PQV_RANK_BATCH_OFF"]
    assert np.array_equal(ctx.centroid_rank(cent, qs, 32), one_by_one)
    cent[5, 0] = np.nan                                   # NaN distance for every query: partial_cmp -> Equal
    got = ctx.centroid_rank(cent, qs[:4], 16)
    for i in range(4):
        assert got[i].tolist() == O.find_closest_centroids(qs[i], cent, 16).tolist()


def test_ivf_search_pipeline_matches_oracle(ctx, vldb):
    # C1 through the IVF path: rank centroids -> candidate rows in rank order -> gathered re-rank
    n, dim = vldb.shape
    c = O.build_sizes(n)[0]
    rng = np.random.default_rng(0)
    cent = vldb[rng.choice(n, c, replace=False)].copy()
    assign = ctx.kmeans_assign(vldb, cent)
    assert np.array_equal(assign, O.assign(vldb, cent))
    offsets, ids = O.inverted_lists(assign, c)
    ds = ctx.dataset_from(vldb)
    for nprobe in (1, 5, 32):
        for qrow in (0, 1, 100, 333):
            q = vldb[qrow]
            cl = ctx.centroid_rank(cent, q, nprobe)[0]
            cand = np.concatenate([ids[int(offsets[j]):int(offsets[j + 1])] for j in cl]).astype(np.uint32)
            assert cand.tolist() == O.candidate_rows(q, cent, offsets, ids, nprobe).tolist()
            check_topk(ds, vldb, q, 10, SQRT, cand)
    ds.drop()


# ---- error behaviour ---------------------------------------------------------------------------------
def test_errors(ctx, P):
    ds = ctx.dataset_from(np.zeros((10, 4), np.float32))
    with pytest.raises(P.PqvError, match="k must be > 0"):
        ds.l2_topk(np.zeros(4, np.float32), 0)
    with pytest.raises(P.PqvError, match="Query dimension mismatch: expected 4, got 5"):
        ds.l2_topk(np.zeros(5, np.float32), 1)
    r, d = ds.l2_topk(np.zeros(4, np.float32), 5000)   # k above the in-kernel selection: answered by the full replay (10 rows)
    assert r.size == 10
    with pytest.raises(P.PqvError) as ei:               # the streaming form keeps the limit
        ctx.topk_stream(np.zeros(4, np.float32), 5000)
    assert ei.value.code == 6
    with pytest.raises(P.PqvError, match="out of range"):
        ds.l2_topk_gather(np.zeros(4, np.float32), np.array([10], np.uint32), 1)
    with pytest.raises(P.PqvError, match="nprobe must be > 0"):
        ctx.centroid_rank(np.zeros((3, 4), np.float32), np.zeros(4, np.float32), 0)
    ds.drop()
    with pytest.raises(P.PqvError) as ei:
        ds2 = P.Dataset(ctx, 12345, 4)
        ds2.l2_topk(np.zeros(4, np.float32), 1)
    assert ei.value.code == 5
    empty = ctx.dataset(4, 0)
    r, d = empty.l2_topk(np.zeros(4, np.float32), 3)
    assert r.size == 0
    empty.drop()


# ---- BASELINE sizes: size-independent properties ---------------------------------------------------------
def test_c2_scale_properties(ctx):
    """1M x 768 vs the full oracle, then 10M x 768 (config C2) through properties the oracle can check from a
    sample: returned distances are bit-exact for the returned rows, ascending, and no sampled row beats the k-th."""
    dim, seed, k = 768, 1234, 100
    q = O.synth(1, dim, 7)[0]
    n1 = 1_000_000
    ds = ctx.dataset(dim, n1)
    ds.fill_synthetic(n1, seed)
    r, d = ds.l2_topk(q, k, SQRT)
    host = O.synth(n1, dim, seed)
    er, ed = O.scan_topk_mt(host, q, k, 0, workers=1)
    assert r.tolist() == er.tolist()
    assert bits(d).tolist() == bits(np.sqrt(ed)).tolist()
    del host
    ds.drop()

    n2 = 10_000_000
    ds = ctx.dataset(dim, n2)
    ds.fill_synthetic(n2, seed)
    r, d = ds.l2_topk(q, k, 0)
    assert r.size == k and np.all(np.diff(d) >= 0) and len(set(r.tolist())) == k
    for i, row in enumerate(r.tolist()):
        v = O.synth(1, dim, seed, first_row=row)[0]
        assert O.squared_l2_unroll4(q, v).view(np.uint32) == d[i].view(np.uint32)
    rng = np.random.default_rng(3)
    kth = d[-1]
    rset = set(r.tolist())
    for start in rng.integers(0, n2 - 20_000, 8):
        blk = O.synth(20_000, dim, seed, first_row=int(start))
        dd = O.distances(blk, q, 0)
        better = np.nonzero(dd < kth)[0]
        assert all(int(start + b) in rset for b in better)
    # the block that holds the winner must reproduce it
    blk0 = int(r[0]) - int(r[0]) % 1000
    dd = O.distances(O.synth(1000, dim, seed, first_row=blk0), q, 0)
    assert int(np.argmin(dd)) + blk0 == int(r[0])
    ds.drop()


def _oracle_distances_full(n, dim, seed, q, order, block=500_000):
    """squared distances of q to every one of the n synthetic rows, in the oracle's arithmetic: blocks generated and scored
    on all host threads (the loop body is per-row independent; ctypes releases the GIL), one f32 per row kept"""
    import ctypes as C
    import os
    import threading
    out = np.empty(n, dtype=np.float32)
    L = O.lib()
    nthreads = os.cpu_count() or 1
    starts = list(range(0, n, block))
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                if not starts:
                    return
                s = starts.pop()
            m = min(block, n - s)
            blk = np.empty((m, dim), dtype=np.float32)
            L.pqo_synth_fill(blk.ctypes.data_as(C.POINTER(C.c_float)), s * dim, m * dim, seed)
            L.pqo_distances(blk.ctypes.data_as(C.POINTER(C.c_float)), m, dim, q.ctypes.data_as(C.POINTER(C.c_float)), order,
                            out[s:s + m].ctypes.data_as(C.POINTER(C.c_float)))

    th = [threading.Thread(target=work) for _ in range(nthreads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return out


def test_full_scale_c2_equals_the_oracle_top100(ctx):
    """BASELINE configs[1] at its full size: the GPU's top-100 of 10M x 768 is the reference loop's top-100 -- same row ids
    in the same order, same f32 bits -- for both summation orders.  Oracle side: exact distances of ALL 10M rows (blocks on
    every host core), then the reference's bounded BinaryHeap + sqrt + stable sort over the 10M values
    (pqo_heap_topk = src/ivf/search.rs:112-141)."""
    n, dim, seed, k = 10_000_000, 768, 1234, 100
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, seed)
    qs = O.synth(3, dim, 7)
    for qi, flags in ((0, SQRT), (1, SQRT), (2, SEQ)):
        q = np.ascontiguousarray(qs[qi])
        r, d = ds.l2_topk(q, k, flags)
        dist = _oracle_distances_full(n, dim, seed, q, 1 if flags & SEQ else 0)
        er, ed = O.heap_topk(dist, None, k, bool(flags & SQRT))
        assert r.tolist() == er.tolist()
        assert bits(d).tolist() == bits(ed).tolist()
    # the batched tensor-core pass over the same table: every query equals its own oracle loop too
    qb = np.ascontiguousarray(O.synth(8, dim, 11))
    rows, dd, cnt = ds.l2_topk(qb, 10, SEQ)
    for i in (0, 7):
        dist = _oracle_distances_full(n, dim, seed, np.ascontiguousarray(qb[i]), 1)
        er, ed = O.heap_topk(dist, None, 10, False)
        assert cnt[i] == 10 and rows[i].tolist() == er.tolist() and bits(dd[i]).tolist() == bits(ed).tolist()
    ds.drop()


def test_peer_exchange_world_of_one(ctx, P):
    """pqv_peer_exchange_* + pqv_l2_topk_candidates_p2p with a single rank (the buffer is its own peer): publish / flag /
    wait kernels, sequence parity over several searches, and the slot-overflow signal.  The multi-rank exchange is
    checked by benchmarks/check_p2p_exchange.py under torchrun."""
    from pq_vector_b200.sharded import ShardedTopk
    rng = np.random.default_rng(21)
    n, dim, k = 50000, 96, 20
    data = rng.random((n, dim), dtype=np.float32)
    ds = ctx.dataset_from(data)
    st = ShardedTopk(lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), 0, "cpu", cap=2048)
    st.enable_p2p(ctx, ds)
    for i in range(5):                                   # both buffer parities, several sequence numbers
        q = rng.random(dim, dtype=np.float32)
        r, d = st.search(q, k, SQRT)
        er, ed = O.topk_rerank(q, data, None, k, 0, True)
        assert r.tolist() == er.tolist() and bits(d).tolist() == bits(ed).tolist()
    keys = ds.l2_topk_candidates_p2p(data[3], k, SQRT, 1000)
    assert keys is not None and ((keys & np.uint64(0xFFFFFFFF)) >= 1000).all()      # pos_base applied
    # descending distances: every row enters the heap -> far more candidates than the slot holds -> overflow signal,
    # and ShardedTopk falls back to the collective path with the right answer
    desc = np.zeros((4000, 4), np.float32)
    desc[:, 0] = np.arange(4000, 0, -1)
    ds2 = ctx.dataset_from(desc)
    st2 = ShardedTopk(lambda q, k_, f_, pb: ds2.l2_topk_candidates(q, k_, f_, pb), 0, "cpu", cap=64)
    st2.enable_p2p(ctx, ds2)
    assert ds2.l2_topk_candidates_p2p(np.zeros(4, np.float32), 5, SQRT) is None
    r, d = st2.search(np.zeros(4, np.float32), 5, SQRT)
    er, ed = O.topk_rerank(np.zeros(4, np.float32), desc, None, 5, 0, True)
    assert r.tolist() == er.tolist() and bits(d).tolist() == bits(ed).tolist()
    ds.drop(); ds2.drop()


def test_large_pageable_append_goes_through_the_staging_lanes_intact(ctx):
    """>= 32 MB from pageable memory: multi-lane pinned staging (append_staged); the table must hold the same bytes as
    with the plain copy, across chunk borders and for a ragged tail, and searches over it must agree with the oracle."""
    import os
    rng = np.random.default_rng(77)
    dim, n = 96, 120_001                       # 46 MB: 5.5 staging chunks of 8 MB, ragged tail
    data = rng.random((n, dim), dtype=np.float32)
    ds = ctx.dataset(dim, 16)                  # small hint: the shard also has to grow
    ds.append(data[:7])
    ds.append(data[7:])                        # staged
    assert ds.rows == n
    for lo in (0, 7, 21845 - 3, 60000, n - 50):          # 21845.33 rows per 8 MB chunk
        assert np.array_equal(ds.read(lo, 50 if lo + 50 <= n else n - lo), data[lo:lo + 50])
    q = rng.random(dim, dtype=np.float32)
    check_topk(ds, data, q, 10, SQRT)
    os.environ["PQV_APPEND_DIRECT"] = "1"
    try:
        ds2 = ctx.dataset_from(data)
    finally:
        del os.environ["PQV_APPEND_DIRECT"]
    r1, d1 = ds.l2_topk(q, 25, SEQ)
    r2, d2 = ds2.l2_topk(q, 25, SEQ)
    assert r1.tolist() == r2.tolist() and d1.view(np.uint32).tolist() == d2.view(np.uint32).tolist()
    ds.drop()
    ds2.drop()
