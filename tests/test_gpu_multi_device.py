"""One context driving several devices through the C ABI (pqv_init with n_devices > 1): a dataset's rows are split into
contiguous ranges, one per device, and every entry point of the path must answer exactly as over a single device.

The shards live behind one pqv_ctx, so the same code runs with the SAME GPU listed twice (two device states, two streams,
two sets of scratch): that is what a single-GPU box exercises; with >= 2 GPUs visible the test also uses two real devices
(peer copies, concurrent passes)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SQRT, SEQ, BYPOS = 2, 1, 4


def _device_sets():
    import torch
    sets = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append([0, 1])
    if torch.cuda.device_count() >= 4:
        sets.append([0, 1, 2, 3])
    return sets


@pytest.fixture(scope="module")
def P():
    import pq_vector_b200 as P
    return P


@pytest.fixture(scope="module", params=_device_sets(), ids=lambda s: "dev" + "".join(map(str, s)))
def mctx(P, request):
    c = P.Context(request.param)
    assert c.device_count == len(request.param)
    yield c
    c.close()


@pytest.fixture(scope="module")
def sctx(P):
    c = P.Context([0])
    yield c
    c.close()


def nbits(a):
    a = np.asarray(a, np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))


def same(a, b):
    (r, d), (er, ed) = a, b
    assert r.tolist() == er.tolist()
    assert nbits(d).tolist() == nbits(ed).tolist()


def test_dense_and_gathered_single_queries(mctx):
    rng = np.random.default_rng(3)
    for n, dim, grid in [(5000, 64, False), (7001, 8, True), (3, 4, False), (40_000, 128, False)]:
        data = rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32)
        ds = mctx.dataset_from(data)
        assert ds.rows == n
        assert np.array_equal(ds.read(0, n), data)
        for k, flags in [(1, SQRT), (10, SQRT), (100, SEQ), (10, SEQ | SQRT), (7, BYPOS)]:
            q = data[rng.integers(n)] if grid else rng.random(dim, dtype=np.float32)
            if flags & BYPOS:
                r, d = ds.l2_topk(q, k, flags)
                dd = O.distances(data, q, 0)
                order = np.lexsort((np.arange(n), dd))[:k]
                assert r.tolist() == order.tolist()
            else:
                same(ds.l2_topk(q, k, flags), O.topk_rerank(q, data, None, k, 1 if flags & SEQ else 0, bool(flags & SQRT)))
            # candidate lists that cross the shard borders back and forth, with repeats
            ids = rng.integers(0, n, max(1, n // 3)).astype(np.uint32)
            same(ds.l2_topk_gather(q, ids, k, flags & ~BYPOS),
                 O.topk_rerank_gather(q, data, ids, k, 1 if flags & SEQ else 0, bool(flags & SQRT)))
        sel = rng.integers(0, n, 57).astype(np.uint32)
        assert np.array_equal(ds.read_rows(sel), data[sel])
        ds.drop()


def test_nan_rows_over_several_shards(mctx):
    rng = np.random.default_rng(4)
    n, dim = 6000, 8
    data = rng.random((n, dim), dtype=np.float32)
    for r in (2, 2999, 3000, 5999):
        data[r, 1] = np.nan
    ds = mctx.dataset_from(data)
    q = rng.random(dim, dtype=np.float32)
    for k in (3, 10, 100):
        same(ds.l2_topk(q, k, SQRT), O.topk_rerank(q, data, None, k, 0, True))
        ids = rng.permutation(n)[:2500].astype(np.uint32)
        ids[1] = 3000
        same(ds.l2_topk_gather(q, ids, k, SEQ), O.topk_rerank_gather(q, data, ids, k, 1, False))
    ds.drop()


@pytest.mark.parametrize("n,dim,nq,k,flags,grid", [(30_000, 64, 64, 10, SEQ, False), (20_000, 128, 40, 100, SQRT, False),
                                                   (9000, 32, 33, 5, SQRT, True), (5000, 8, 12, 5, SQRT, False)])
def test_batched_queries_over_several_shards(mctx, n, dim, nq, k, flags, grid):
    """pqv_l2_topk with a batch: every shard's tensor-core pass on its own host thread, per-shard k + 1 keys merged;
    every query must equal its own single-query reference (grid data: ties everywhere -> the multi-shard replay)"""
    rng = np.random.default_rng(n + nq)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32)
    ds = mctx.dataset_from(data)
    qs = (data[rng.integers(0, n, nq)] + (0 if grid else 0.01)).astype(np.float32)
    rows, dist, cnt = ds.l2_topk(qs, k, flags)
    t = mctx.last_batch_timing()
    assert t["queries"] == (nq if dim >= 32 else 0) and not t["declined"]     # short rows take the single-query scans
    for i in range(nq):
        er, ed = O.topk_rerank(qs[i], data, None, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
        assert cnt[i] == er.size and rows[i, :cnt[i]].tolist() == er.tolist(), i
        assert nbits(dist[i, :cnt[i]]).tolist() == nbits(ed).tolist()
    ds.drop()


def test_assignment_and_index_build_equal_the_single_device_ones(mctx, sctx):
    rng = np.random.default_rng(11)
    n, dim, C = 30_000, 64, 40
    data = rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)] + 0.01
    mds, sds = mctx.dataset_from(data), sctx.dataset_from(data)
    a = mctx.kmeans_assign(mds, cent)
    assert np.array_equal(a, O.assign(data, cent, workers=4))
    a2, sizes = mctx.kmeans_assign(mds, cent, n=n - 777, want_sizes=True)     # a prefix that ends inside the last shard
    assert np.array_equal(a2, a[:n - 777]) and sizes.sum() == n - 777
    # build: sample gathered from every shard, trained on the first device, rows assigned where they live
    for n_clusters in (C, None):
        mi = mctx.ivf_build(mds, n_clusters=n_clusters, max_iters=4, seed=9)
        si = sctx.ivf_build(sds, n_clusters=n_clusters, max_iters=4, seed=9)
        assert mi.to_bytes() == si.to_bytes()
        qs = rng.random((6, dim), dtype=np.float32)
        mask = rng.random(n) < 0.4
        for q in qs:
            for nprobe in (1, 5, 10_000):
                same(mi.search(mds, q, 10, nprobe, SQRT), si.search(sds, q, 10, nprobe, SQRT))
                mr = mi.vector_topk(mds, q, 20, nprobe, SEQ, 5000, mask)
                sr = si.vector_topk(sds, q, 20, nprobe, SEQ, 5000, mask)
                same(mr[:2], sr[:2])
                assert mr[2:] == sr[2:]
        br, bd, bc = mi.search_batch(mds, qs, 10, 5, SQRT)
        xr, xd, xc = si.search_batch(sds, qs, 10, 5, SQRT)
        assert bc.tolist() == xc.tolist() and br.tolist() == xr.tolist() and nbits(bd).tolist() == nbits(xd).tolist()
        vr, vd, vc = mi.vector_topk_batch(mds, qs, 10, 5, SEQ, mask)
        yr, yd, yc = si.vector_topk_batch(sds, qs, 10, 5, SEQ, mask)
        assert vc.tolist() == yc.tolist() and vr.tolist() == yr.tolist() and nbits(vd).tolist() == nbits(yd).tolist()
        mi.drop(); si.drop()
    # a tiny table: the "sample" is the whole table
    tiny = rng.random((50, 16), dtype=np.float32)
    mt, stn = mctx.dataset_from(tiny), sctx.dataset_from(tiny)
    mi, si = mctx.ivf_build(mt, max_iters=3, seed=1), sctx.ivf_build(stn, max_iters=3, seed=1)
    assert mi.to_bytes() == si.to_bytes()
    for d_ in (mds, sds, mt, stn):
        d_.drop()


def test_array_distance_arm_over_several_shards(mctx):
    """the un-indexed arm (pqv_array_distance, _topk, _topk_filtered) over a table spread over several device states: the
    Float64 column and the exact (distance, row) top-k against the oracle, with a filter whose bits cross the shard
    boundaries (row counts that are no multiples of 8, quantised data with ties across shards)"""
    rng = np.random.default_rng(31)
    for n, dim, grid in [(10_007, 48, False), (5_003, 16, True), (3, 8, False)]:
        data = (rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32))
        q = (np.zeros(dim) if grid else rng.random(dim) + 1e-9)
        ds = mctx.dataset_from(data)
        for metric in (0, 1):
            got = ds.array_distance(q, metric)
            exp = O.array_distance_column(data, q, metric)
            assert np.asarray(got).view(np.uint64).tolist() == np.asarray(exp).view(np.uint64).tolist()
        for k in (1, 7, 100, 1024):
            rows, dist = ds.array_distance_topk(q, k)
            er, ed = O.array_distance_topk(data, q, k)
            assert rows.tolist() == er.tolist() and np.asarray(dist).view(np.uint64).tolist() == np.asarray(ed).view(np.uint64).tolist()
        for share in (0.5, 0.01, 0.0):
            mask = rng.random(n) < share
            rows, dist = ds.array_distance_topk(q, 50, row_mask=mask)
            keep = np.nonzero(mask)[0]
            if keep.size == 0:
                assert rows.size == 0
                continue
            er, ed = O.array_distance_topk(data[keep], q, 50)
            assert rows.tolist() == keep[er].tolist()
            assert np.asarray(dist).view(np.uint64).tolist() == np.asarray(ed).view(np.uint64).tolist()
        ds.drop()


def test_kmeans_train_and_min_dist_update_over_several_shards(mctx, sctx):
    """the two k_means helpers that need rows from everywhere (pqv_kmeans_train: every row in table order;
    pqv_min_dist_update: a selection of rows) collect them on the first device: same bits as over one device / the oracle"""
    rng = np.random.default_rng(41)
    n, dim, C = 9_001, 24, 17
    data = rng.random((n, dim), dtype=np.float32)
    mds, sds = mctx.dataset_from(data), sctx.dataset_from(data)
    mc = mctx.kmeans_train(mds, C, max_iters=3, seed=5, sum_workers=4)
    sc = sctx.kmeans_train(sds, C, max_iters=3, seed=5, sum_workers=4)
    assert mc.view(np.uint32).tolist() == sc.view(np.uint32).tolist()
    sel = rng.permutation(n)[:4000].astype(np.uint64)
    c0, c1 = data[17], data[4242]
    md = mctx.min_dist_update(mds, sel, c0)
    assert md.view(np.uint32).tolist() == O.distances(data[sel.astype(np.int64)], c0, 0).view(np.uint32).tolist()
    mctx.min_dist_update(mds, sel, c1, md)
    exp = np.minimum(O.distances(data[sel.astype(np.int64)], c0, 0), O.distances(data[sel.astype(np.int64)], c1, 0))
    assert md.view(np.uint32).tolist() == exp.view(np.uint32).tolist()
    full = mctx.min_dist_update(mds, None, c0)          # no selection: every row, in table order
    assert full.view(np.uint32).tolist() == O.distances(data, c0, 0).view(np.uint32).tolist()
    mds.drop(); sds.drop()
