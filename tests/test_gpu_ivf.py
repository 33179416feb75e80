"""GPU parity of the IVF layer (build / blob / candidate_rows / search) against the oracle pipeline."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
SQRT, SEQ = 2, 1
M64 = (1 << 64) - 1


class SplitMix64:
    """Python twin of the build's draw stream (pqv_ivf_impl.cuh) so the oracle pipeline sees the same draws."""

    def __init__(self, seed):
        self.s = seed & M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    def below(self, n):
        lim = M64 - (M64 % n)
        while True:
            v = self.next()
            if v < lim:
                return v % n

    def unit_f32(self):
        return np.float32(self.next() >> 40) * np.float32(1.0 / 16777216.0)


def sample_indices(rng, n, amount):
    if amount * 2 >= n:
        a = list(range(n))
        for i in range(amount):
            j = i + rng.below(n - i)
            a[i], a[j] = a[j], a[i]
        return a[:amount]
    seen, out = set(), []
    for j in range(n - amount, n):
        t = rng.below(j + 1)
        if t in seen:
            seen.add(j); out.append(j)
        else:
            seen.add(t); out.append(t)
    for i in range(amount, 1, -1):
        r = rng.below(i)
        out[i - 1], out[r] = out[r], out[i - 1]
    return out


def oracle_build(data, n_clusters, max_iters, seed, workers):
    """build_ivf_index + k_means (src/ivf/index.rs:152-214, 323-457) from oracle pieces and the shared draws."""
    n, dim = data.shape
    C, sample_size, _ = O.build_sizes(n, n_clusters)
    sample = data if sample_size == n else data[np.array(sample_indices(SplitMix64(seed), n, sample_size))]
    ns = sample.shape[0]
    rng = SplitMix64(seed)
    init_n = max(min(ns, 50000), C)
    init_idx = list(range(ns)) if init_n == ns else sample_indices(rng, ns, init_n)
    sel = np.array(init_idx, dtype=np.uint64)
    cent = np.zeros((C, dim), np.float32)
    cent[0] = sample[init_idx[rng.below(init_n)]]
    md = O.min_dist_init(sample, sel, cent[0])
    for i in range(1, C):
        total = O.min_dist_update(sample, sel, cent[i - 1], md, workers=workers)
        if total > 0:
            thr = np.float32(rng.unit_f32() * total)
            s = O.kmeanspp_pick(md, thr)
            if s < init_n:
                cent[i] = sample[init_idx[s]]
        else:
            cent[i] = sample[init_idx[rng.below(init_n)]]
    assign = np.zeros(ns, np.uint32)
    iters = 0
    for _ in range(max_iters):
        iters += 1
        changed, sizes = O.lloyd_assign(sample, cent, assign, workers=8)
        if changed == 0:
            break
        cent = O.centroid_update(sample, assign, sizes, C)
    full = O.assign(data, cent, workers=8)
    offsets, ids = O.inverted_lists(full, C)
    return cent, offsets, ids, iters


@pytest.fixture(scope="module")
def ctx():
    import pq_vector_b200 as P
    c = P.Context()
    yield c
    c.close()


@pytest.mark.parametrize("n,dim,C,iters", [(496, 64, None, 20), (3000, 32, 16, 5), (60, 5, 60, 3), (5000, 48, None, 20),
                                           (2100, 24, 7, 1)])
def test_build_matches_oracle_pipeline(ctx, n, dim, C, iters):
    rng = np.random.default_rng(n + dim)
    # clustered data so Lloyd moves centroids for a few rounds
    centers = rng.random((12, dim), dtype=np.float32) * 4
    data = (centers[rng.integers(0, 12, n)] + rng.standard_normal((n, dim)).astype(np.float32) * 0.3).astype(np.float32)
    ds = ctx.dataset_from(data)
    ix = ctx.ivf_build(ds, n_clusters=C, max_iters=iters, seed=42, sum_workers=4)
    blob = ix.to_bytes()
    d2, cent, offsets, ids = O.index_from_bytes(blob)
    ecent, eoff, eids, eiters = oracle_build(data, C, iters, 42, 4)
    assert d2 == dim and cent.shape == ecent.shape
    assert ix.build_stats()["lloyd_iters"] == eiters
    assert np.array_equal(cent.view(np.uint32), ecent.view(np.uint32))       # centroids bit-exact given equal draws
    assert np.array_equal(offsets, eoff) and np.array_equal(ids, eids)
    assert blob == O.index_to_bytes(dim, ecent, eoff, eids)                    # byte-identical blob
    # blob round trip through the library
    ix2 = ctx.ivf_from_bytes(blob)
    assert ix2.to_bytes() == blob and (ix2.dim, ix2.n_clusters, ix2.n_ids) == (dim, cent.shape[0], n)
    # candidate_rows + search parity (TopkBuilder semantics)
    for qi in (0, n // 2):
        q = data[qi] + np.float32(0.01)
        for nprobe in (1, 3, 1000):
            cand = ix2.candidate_rows(q, nprobe)
            assert cand.tolist() == O.candidate_rows(q, ecent, eoff, eids, nprobe).tolist()
            for k in (1, 10):
                r, d = ix.search(ds, q, k, nprobe, SQRT)
                er, ed = O.topk_rerank_gather(q, data, cand, k, 0, True)
                assert r.tolist() == er.tolist()
                assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    ix.drop(); ix2.drop(); ds.drop()


@pytest.mark.parametrize("case", ["identical", "two-points", "tiny", "huge", "grid", "big-init", "nan-row", "one-worker",
                                  "forty-workers", "six-hundred-workers", "odd-rows", "short-chunks"])
def test_kmeanspp_on_the_device_handles_awkward_tables(ctx, case):
    """the device-side k-means++ pick (pqv_kmeanspp.cuh: chunk sums, draw, block-parallel exact walk) against the oracle's
    literal chains on tables that leave its integer model: zero totals (uniform draws), tiny / huge distances, distances on a
    coarse grid (rounding ties all along the walk), an init set of the full 50 000 rows, non-finite distances"""
    rng = np.random.default_rng(len(case))
    workers, iters = 4, 2
    if case == "identical":
        data, C = np.tile(rng.random((1, 8), dtype=np.float32), (400, 1)), 9
    elif case == "two-points":
        data, C = np.repeat(rng.random((2, 8), dtype=np.float32), 300, axis=0), 6
    elif case == "tiny":
        data, C = (rng.random((3000, 8)) * 1e-19).astype(np.float32), 12
    elif case == "huge":
        data, C = (rng.random((3000, 8)) * 3e17).astype(np.float32), 12
    elif case == "grid":
        data, C = (rng.integers(0, 5, (20000, 4)) * 1024.5).astype(np.float32), 30
    elif case == "big-init":
        data, C, workers = rng.random((60000, 8), dtype=np.float32) * 7, 40, 3
    elif case == "nan-row":
        data, C = rng.random((2500, 8), dtype=np.float32), 10
        data[7, 3] = np.nan
        data[900, 0] = np.inf
    elif case == "forty-workers":     # chunk chains in two warps; chunk boundaries off the 16-byte grid of the bulk copy
        data, C, workers = rng.random((9001, 8), dtype=np.float32), 20, 40
    elif case == "six-hundred-workers":  # more chain warps than the two-wave copy serves: plain loads
        data, C, workers = rng.random((7003, 8), dtype=np.float32), 15, 600
    elif case == "odd-rows":          # row count % 4 != 0: the array's tail comes by plain loads
        data, C, workers = rng.random((2999, 8), dtype=np.float32), 14, 5
    elif case == "short-chunks":      # chunks shorter than the first wave of the copy
        data, C, workers = rng.random((1000, 8), dtype=np.float32), 10, 16
    else:
        data, C, workers = rng.random((9000, 16), dtype=np.float32), 25, 1
    ds = ctx.dataset_from(data)
    ix = ctx.ivf_build(ds, n_clusters=C, max_iters=iters, seed=7, sum_workers=workers)
    _, cent, offsets, ids = O.index_from_bytes(ix.to_bytes())
    ecent, eoff, eids, _ = oracle_build(data, C, iters, 7, workers)
    assert np.array_equal(np.isnan(cent), np.isnan(ecent))
    assert np.array_equal(np.nan_to_num(cent, nan=-1.0).view(np.uint32), np.nan_to_num(ecent, nan=-1.0).view(np.uint32))
    assert np.array_equal(offsets, eoff) and np.array_equal(ids, eids)
    ix.drop(); ds.drop()


def test_reference_index_kat_blob(ctx):
    # src/ivf/index.rs:495-511
    blob = (np.array([3, 2], "<u4").tobytes() + np.arange(1, 7, dtype="<f4").tobytes()
            + np.array([3, 0, 2, 4], "<u4").tobytes() + np.array([2, 1, 3], "<u4").tobytes())
    ix = ctx.ivf_from_bytes(blob)
    assert (ix.dim, ix.n_clusters, ix.n_ids) == (3, 2, 5) and ix.to_bytes() == blob
    assert ix.candidate_rows(np.array([1, 2, 3], np.float32), 1).tolist() == [0, 2, 4]
    assert ix.candidate_rows(np.array([4, 5, 6], np.float32), 2).tolist() == [1, 3, 0, 2, 4]
    ix.drop()


def test_ivf_errors(ctx):
    import pq_vector_b200 as P
    ds = ctx.dataset_from(np.zeros((5, 4), np.float32))
    with pytest.raises(P.PqvError, match="n_clusters cannot exceed number of vectors"):
        ctx.ivf_build(ds, n_clusters=6)
    with pytest.raises(P.PqvError, match="max_iters must be > 0"):
        ctx.ivf_build(ds, max_iters=0)
    with pytest.raises(P.PqvError, match="IVF index buffer too small"):
        ctx.ivf_from_bytes(b"\0" * 7)
    empty = ctx.dataset(4, 0)
    with pytest.raises(P.PqvError, match="Cannot build IVF index with zero vectors"):
        ctx.ivf_build(empty)
    ix = ctx.ivf_build(ds, n_clusters=2, max_iters=2)
    with pytest.raises(P.PqvError, match="nprobe must be > 0"):
        ix.search(ds, np.zeros(4, np.float32), 1, 0)
    with pytest.raises(P.PqvError, match="Query dimension mismatch"):
        ix.search(ds, np.zeros(5, np.float32), 1, 1)
    ix.drop(); ds.drop(); empty.drop()


def test_vldb_c1_through_the_index(ctx, vldb):
    """Config C1: nprobe=32 >= C=23 -> every row is a candidate -> the SURVEY 8c top-10 ids."""
    ds = ctx.dataset_from(vldb)
    ix = ctx.ivf_build(ds)                      # defaults: C = ceil(sqrt(496)) = 23, 20 iters, seed 42
    assert ix.n_clusters == 23
    for qrow, ids in {0: [0, 126, 81, 265, 315, 464, 322, 269, 169, 140],
                      100: [100, 181, 400, 352, 448, 476, 36, 198, 370, 213]}.items():
        assert ix.candidate_rows(vldb[qrow], 32).size == 496     # snapshot vector_topk_vldb_tree.snap:29
        r, d = ix.search(ds, vldb[qrow], 10, 32, SQRT)
        # with all rows as candidates but in list order the heap history differs from row order; the SET and the
        # distances are those of the table, and on this data there are no ties inside the top 10 -> same ids
        assert r.tolist() == ids
    ix.drop(); ds.drop()


@pytest.mark.parametrize("n,dim,C", [(30000, 64, 40), (2500, 32, None), (120, 8, 120)])
def test_sharded_build_pieces_equal_the_monolithic_build(ctx, n, dim, C):
    """pqv_ivf_sample_rows + pqv_kmeans_train + pqv_kmeans_assign, glued by ShardedIvfBuild (world of one here; the
    two-rank exchange is covered by the gloo test and benchmarks/check_sharded_ivf.py), give the blob of pqv_ivf_build."""
    from pq_vector_b200.sharded import ShardedIvfBuild
    rng = np.random.default_rng(n)
    cent0 = rng.standard_normal((25, dim)).astype(np.float32)
    data = (cent0[rng.integers(0, 25, n)] + 0.3 * rng.standard_normal((n, dim))).astype(np.float32)
    ds = ctx.dataset_from(data)
    want = ctx.ivf_build(ds, n_clusters=C, max_iters=7, seed=5).to_bytes()

    def train(sample, c, max_iters, seed):
        sd = ctx.dataset_from(sample)
        out = ctx.kmeans_train(sd, c, max_iters, seed)
        sd.drop()
        return out

    sb = ShardedIvfBuild(ds.read_rows, train, lambda cent: ctx.kmeans_assign(ds, cent), n, 0, n, dim)
    assert np.array_equal(ds.read_rows([n - 1, 0, 7]), data[[n - 1, 0, 7]])
    assert sb.build(C, 7, 5) == want
    ds.drop()


def _search_both_ways(ctx, ix, ds, q, k, nprobe, flags=SQRT):
    """the one-round-trip device path and the host-ranked path (PQV_IVF_FUSED is read once per process, so the host
    path is reached the way a caller reaches it: candidate_rows + gather)"""
    r, d = ix.search(ds, q, k, nprobe, flags)
    cand = ix.candidate_rows(q, nprobe)
    r2, d2 = ds.l2_topk_gather(q, cand, k, flags) if cand.size else (np.empty(0, np.uint32), np.empty(0, np.float32))
    assert r.tolist() == r2.tolist() and d.view(np.uint32).tolist() == d2.view(np.uint32).tolist()
    return r, d, cand


def test_search_with_empty_lists_ties_and_duplicates(ctx):
    """hand-made index: empty clusters, duplicate centroids (stable rank: lowest cluster index first), duplicate rows
    (exact distance ties across lists) -- device ranking + expansion + scan against the oracle's walk"""
    rng = np.random.default_rng(5)
    dim, n = 16, 4000
    data = rng.integers(0, 3, (n, dim)).astype(np.float32)
    data[100:200] = data[0:100]                                   # duplicates -> bit-equal distances
    cent = rng.integers(0, 3, (12, dim)).astype(np.float32)
    cent[7] = cent[2]                                             # duplicate centroid: equal distances, rank 2 before 7
    cent[11] = cent[2]
    assign = O.assign(data, cent, workers=2)                      # first-min: clusters 7 and 11 stay empty
    offsets, ids = O.inverted_lists(assign, 12)
    assert offsets[8] == offsets[7] and offsets[12] == offsets[11]
    ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    ds = ctx.dataset_from(data)
    for qi in (0, 150, 3999):
        q = data[qi]
        for nprobe in (1, 2, 5, 12, 99):
            for k in (1, 10, 300):
                for flags in (SQRT, SEQ):
                    r, d, cand = _search_both_ways(ctx, ix, ds, q, k, nprobe, flags)
                    assert cand.tolist() == O.candidate_rows(q, cent, offsets, ids, nprobe).tolist()
                    er, ed = O.topk_rerank_gather(q, data, cand, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
                    assert r.tolist() == er.tolist() and d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    # k above the in-kernel selection (PQV_MAX_K = 1024): distance log + reference loop on the host
    for k in (1025, 3000):
        for flags in (SQRT, SEQ):
            r, d = ix.search(ds, data[150], k, 5, flags)
            cand = O.candidate_rows(data[150], cent, offsets, ids, 5)
            er, ed = O.topk_rerank_gather(data[150], data, cand, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
            assert r.tolist() == er.tolist() and d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    # an index whose probed lists are all empty
    off2 = np.zeros(13, np.uint64)
    off2[12:] = 0
    ix2 = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, off2, np.empty(0, np.uint32)))
    r, d = ix2.search(ds, data[0], 5, 3, SQRT)
    assert r.size == 0 and d.size == 0
    ix.drop(); ix2.drop(); ds.drop()


def test_search_with_non_finite_centroids_takes_the_host_ranking(ctx):
    """a NaN / inf centroid distance: the device ranking declines (NaN) or ranks it last (inf); the result must be the
    oracle's either way (index.rs:142-146: partial_cmp -> Equal for NaN)"""
    rng = np.random.default_rng(8)
    dim, n = 8, 1500
    data = rng.random((n, dim), dtype=np.float32)
    cent = rng.random((9, dim), dtype=np.float32)
    assign = O.assign(data, cent, workers=2)
    offsets, ids = O.inverted_lists(assign, 9)
    ds = ctx.dataset_from(data)
    for bad in (np.inf, np.nan):
        c2 = cent.copy()
        c2[4, 3] = bad
        ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, c2, offsets, ids))
        for nprobe in (1, 4, 9):
            q = data[7]
            r, d = ix.search(ds, q, 10, nprobe, SQRT)
            cand = O.candidate_rows(q, c2, offsets, ids, nprobe)
            er, ed = O.topk_rerank_gather(q, data, cand, 10, 0, True)
            assert r.tolist() == er.tolist() and d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
        ix.drop()
    ds.drop()


def _nbits(a):
    a = np.asarray(a, np.float32)
    return np.where(np.isnan(a), np.uint32(0x7FC00000), a.view(np.uint32))   # NaN payloads are hardware-specific


def test_search_with_nan_rows_replays_the_reference_loop(ctx):
    """table rows with a NaN coordinate (assigned to cluster 0 by the strict `<` of nearest_centroid, index.rs:244-257):
    once such a row sits in the re-rank heap the reference stops behaving like a threshold (search.rs:119-126); single
    searches, the batched search and the one-call operator must all give the oracle's answer"""
    rng = np.random.default_rng(123)
    dim, n, C = 16, 6000, 10
    data = rng.random((n, dim), dtype=np.float32)
    for r in (0, 5, 17, 400, 4000, 5999):
        data[r, r % dim] = np.nan
    cent = rng.random((C, dim), dtype=np.float32)
    assign = O.assign(data, cent, workers=2)
    assert assign[5] == 0
    offsets, ids = O.inverted_lists(assign, C)
    ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    ds = ctx.dataset_from(data)
    qs = np.vstack([cent[0] + 0.01, cent[3], rng.random((6, dim), dtype=np.float32)]).astype(np.float32)
    for nprobe in (1, 3, C):
        for k, flags in ((1, SQRT), (10, SQRT), (100, SEQ)):
            expect = []
            for q in qs:
                r, d, cand = _search_both_ways(ctx, ix, ds, q, k, nprobe, flags)
                er, ed = O.topk_rerank_gather(q, data, cand, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
                assert r.tolist() == er.tolist() and _nbits(d).tolist() == _nbits(ed).tolist()
                expect.append((er, ed))
                r, d, total, scored = ix.vector_topk(ds, q, k, nprobe, flags)
                xr, xd, xtotal, xscored = _expected_vector_topk(q, data, cent, offsets, ids, k, nprobe,
                                                                1 if flags & SEQ else 0, bool(flags & SQRT), None, None)
                assert (total, scored) == (xtotal, xscored)
                assert r.tolist() == xr.tolist() and _nbits(d).tolist() == _nbits(xd).tolist()
            br, bd, bc = ix.search_batch(ds, qs, k, nprobe, flags)
            for i, (er, ed) in enumerate(expect):
                assert bc[i] == er.size and br[i, :bc[i]].tolist() == er.tolist()
                assert _nbits(bd[i, :bc[i]]).tolist() == _nbits(ed).tolist()
    ix.drop()
    ds.drop()


def test_host_list_builder_equals_the_device_one(ctx):
    """PQV_CSR=host (read once per process) cannot be flipped here, so the equality is checked at the blob level: the
    device-built lists of pqv_ivf_build against lists rebuilt on the host from the same assignment (oracle)"""
    rng = np.random.default_rng(3)
    n, dim, C = 70000, 16, 300
    data = rng.random((n, dim), dtype=np.float32)
    ds = ctx.dataset_from(data)
    ix = ctx.ivf_build(ds, n_clusters=C, max_iters=2, seed=9)
    d2, cent, offsets, ids = O.index_from_bytes(ix.to_bytes())
    a = O.assign(data, cent, workers=4)
    eoff, eids = O.inverted_lists(a, C)
    assert np.array_equal(offsets, eoff) and np.array_equal(ids, eids)
    assert all(np.all(np.diff(ids[int(offsets[c]):int(offsets[c + 1])].astype(np.int64)) > 0) for c in range(C))
    ix.drop(); ds.drop()


# ---- VectorTopKExec over a resident indexed table in one call (pqv_vector_topk_indexed) -------------------------------
def _expected_vector_topk(q, data, cent, offsets, ids, k, nprobe, order, do_sqrt, cap, mask):
    cand = O.candidate_rows(q, cent, offsets, ids, nprobe)        # index_exec.rs:159-163, rank order
    total = cand.size
    if cap is not None:
        cand = cand[:cap]                                         # CandidateCursor over one file = a prefix (access.rs:193-243)
    rows = np.sort(cand)                                          # RowSelection: file order (access.rs:107-176)
    if mask is not None:
        rows = rows[mask[rows]]                                   # FilterExec before scoring (tests.rs:151-241)
    er, ed = O.topk_rerank_gather(q, data, rows, k, order, do_sqrt) if rows.size else (np.empty(0, np.uint32), np.empty(0, np.float32))
    return er, ed, total, rows.size


@pytest.mark.parametrize("n,dim,C,grid", [(20000, 64, 64, False), (6000, 8, 37, True), (3000, 130, 5, False)])
def test_vector_topk_indexed_matches_the_operator_semantics(ctx, n, dim, C, grid):
    rng = np.random.default_rng(n + C)
    # grid data: few distinct coordinates -> many bit-equal distances, so the ORDER in which rows reach the heap matters
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)] + (0.0 if grid else 0.01)
    assign = O.assign(data, cent, workers=2)
    offsets, ids = O.inverted_lists(assign, C)
    ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    ds = ctx.dataset_from(data)
    masks = [None, rng.random(n) < 0.5, np.zeros(n, bool), np.arange(n) >= n // 3]
    for qi in range(3):
        q = data[rng.integers(n)] if grid else rng.random(dim, dtype=np.float32)
        for nprobe in (1, max(2, C // 4), C + 5):
            for cap in (None, 1, 777, 10 ** 9):
                for mask in masks:
                    for k, flags in ((10, SEQ), (100, SQRT)):
                        r, d, total, scored = ix.vector_topk(ds, q, k, nprobe, flags, cap, mask)
                        er, ed, etotal, escored = _expected_vector_topk(q, data, cent, offsets, ids, k, nprobe,
                                                                        1 if flags & SEQ else 0, bool(flags & SQRT), cap, mask)
                        assert (total, scored) == (etotal, escored), (nprobe, cap)
                        assert r.tolist() == er.tolist(), (qi, nprobe, cap, k, flags)
                        assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    ix.drop()
    ds.drop()


def test_vector_topk_indexed_reference_examples_and_fallback(ctx):
    import pq_vector_b200 as P
    # df_vector/tests.rs:31-39 and :166-174 with their filters as row masks
    for rows, min_id, expect, fetched in [([(0, 0), (1, 0), (0, 2), (5, 5), (2, 2), (.1, .1)], 2, [5, 2], 4),
                                          ([(0, 0), (.05, .05), (.2, .2), (1, 1), (1.1, 1.1), (1.4, 1.4)], 3, [3, 4], 3)]:
        data = np.array(rows, np.float32)
        cent = data[[0, 3, 4]].copy()
        offsets, ids = O.inverted_lists(O.assign(data, cent), 3)
        ix = ctx.ivf_from_bytes(O.index_to_bytes(2, cent, offsets, ids))
        ds = ctx.dataset_from(data)
        r, d, total, scored = ix.vector_topk(ds, np.zeros(2, np.float32), 2, 64, SEQ, None, np.arange(6) >= min_id)
        assert r.tolist() == expect and (total, scored) == (6, fetched)
        # a NaN centroid distance: the device ranking declines, the host-ranked path must give the same selection
        c2 = cent.copy()
        c2[1, 0] = np.nan
        ix2 = ctx.ivf_from_bytes(O.index_to_bytes(2, c2, offsets, ids))
        r2, _, total2, _ = ix2.vector_topk(ds, np.zeros(2, np.float32), 2, 64, SEQ, None, np.arange(6) >= min_id)
        assert r2.tolist() == expect and total2 == 6
        with pytest.raises(P.PqvError, match="row_mask has"):
            ix.vector_topk(ds, np.zeros(2, np.float32), 2, 64, SEQ, None, np.ones(5, bool))
        with pytest.raises(P.PqvError, match="nprobe must be > 0"):
            ix.vector_topk(ds, np.zeros(2, np.float32), 2, 0)
        ix.drop()
        ix2.drop()
        ds.drop()


# ---- IVF search with the rows sharded over ranks (pqv_ivf_search_candidates + ShardedIvfSearch.translate) --------------
@pytest.mark.parametrize("n,dim,C,grid", [(30000, 64, 48, False), (9000, 8, 21, True)])
def test_sharded_ivf_search_pieces_equal_the_whole_table_search(ctx, n, dim, C, grid):
    """Three slices of one table on one GPU stand in for three ranks: per-slice entrant candidates, positions translated
    to the global candidate sequence, union replayed -> exactly TopkBuilder's answer over the whole table."""
    import pq_vector_b200 as P
    from pq_vector_b200.sharded import ShardedIvfSearch, index_to_bytes, shard_counts, shard_index
    rng = np.random.default_rng(n)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy()
    offsets, ids = O.inverted_lists(O.assign(data, cent, workers=2), C)
    bounds = [0, n // 5, n // 2, n]
    counts = shard_counts(offsets, ids, bounds)
    parts = []
    for s in range(3):
        lo, hi = bounds[s], bounds[s + 1]
        l_off, l_ids = shard_index(offsets, ids, lo, hi)
        ix = ctx.ivf_from_bytes(index_to_bytes(cent, l_off, l_ids))
        ds = ctx.dataset_from(data[lo:hi])
        parts.append((ix, ds, ShardedIvfSearch(None, counts, s, lo)))
    whole_ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    whole_ds = ctx.dataset_from(data)
    for qi in range(4):
        q = data[rng.integers(n)] if grid else rng.random(dim, dtype=np.float32)
        for nprobe in (1, 5, C):
            for k, flags in ((10, SQRT), (100, SEQ)):
                k_all, r_all, total = [], [], 0
                for ix, ds, sh in parts:
                    keys, rows, probe = ix.search_candidates(ds, q, k, nprobe, flags)
                    gk, total = sh.translate(keys, probe)
                    k_all.append(gk)
                    r_all.append(rows.astype(np.int64) + sh.lo)
                k_all, r_all = np.concatenate(k_all), np.concatenate(r_all).astype(np.uint32)
                row_of = np.zeros(max(total, 1), np.uint32)
                row_of[(k_all & np.uint64(0xFFFFFFFF)).astype(np.int64)] = r_all
                r, d = P.replay_candidates(k_all, k, flags, row_ids=row_of)
                cand = O.candidate_rows(q, cent, offsets, ids, nprobe)
                assert total == cand.size
                er, ed = O.topk_rerank_gather(q, data, cand, k, 1 if flags & SEQ else 0, bool(flags & SQRT))
                assert r.tolist() == er.tolist(), (qi, nprobe, k)
                assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
                wr, wd = whole_ix.search(whole_ds, q, k, nprobe, flags)
                assert wr.tolist() == r.tolist()
    for ix, ds, _ in parts:
        ix.drop()
        ds.drop()
    whole_ix.drop()
    whole_ds.drop()


# ---- batched IVF search (pqv_ivf_search_batch): one masked tensor-core pass for all queries ----------------------------
@pytest.mark.parametrize("n,dim,C,nq,grid", [(40000, 64, 48, 40, False), (30000, 128, 100, 300, False), (12000, 8, 21, 17, True)])
def test_batched_ivf_search_equals_single_searches(ctx, n, dim, C, nq, grid):
    import pq_vector_b200 as P
    rng = np.random.default_rng(n + nq)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy() + (0.0 if grid else 0.01)
    offsets, ids = O.inverted_lists(O.assign(data, cent, workers=2), C)
    ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    ds = ctx.dataset_from(data)
    queries = (data[rng.integers(0, n, nq)] if grid else rng.random((nq, dim), dtype=np.float32)).copy()
    used = 0
    for nprobe in (1, max(2, C // 8), C):
        for k, flags in ((10, SQRT), (100, SEQ), (10, SEQ | P.PQV_ROW_ORDER)):
            rows, dist, cnt = ix.search_batch(ds, queries, k, nprobe, flags)
            bt = ctx.last_batch_timing()
            used += bt["queries"] > 0 and not bt["declined"]
            order, do_sqrt = (1 if flags & SEQ else 0), bool(flags & SQRT)
            for i in range(0, nq, max(1, nq // 25)):
                cand = O.candidate_rows(queries[i], cent, offsets, ids, nprobe)
                if flags & P.PQV_ROW_ORDER:
                    cand = np.sort(cand)
                er, ed = O.topk_rerank_gather(queries[i], data, cand, k, order, do_sqrt)
                assert cnt[i] == er.size, (i, nprobe, k, flags, bt)
                assert rows[i, :cnt[i]].tolist() == er.tolist(), (i, nprobe, k, flags, bt)
                assert dist[i, :cnt[i]].view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    if not grid:
        assert used > 0            # the masked tensor-core pass really ran (grid data may decline it: ties everywhere)
    # a NaN centroid: that ranking is not an order -> those queries take the host-ranked single-query path
    c2 = cent.copy()
    c2[C // 2, 0] = np.nan
    ix2 = ctx.ivf_from_bytes(O.index_to_bytes(dim, c2, offsets, ids))
    rows, dist, cnt = ix2.search_batch(ds, queries[:8], 10, 3, SQRT)
    for i in range(8):
        r1, d1 = ix2.search(ds, queries[i], 10, 3, SQRT)
        assert rows[i, :cnt[i]].tolist() == r1.tolist() and dist[i, :cnt[i]].view(np.uint32).tolist() == d1.view(np.uint32).tolist()
    ix.drop()
    ix2.drop()
    ds.drop()


def test_sharded_batched_ivf_keys_merge_to_the_whole_table_answer(ctx):
    """Three slices on one GPU stand in for three ranks: per-slice k + 1 keys from the masked batched pass, merged
    (pqv_merge_batch_keys); flagged queries answered by the single-query route -> the whole-table batched answer."""
    import pq_vector_b200 as P
    from pq_vector_b200.sharded import index_to_bytes, shard_index
    rng = np.random.default_rng(2024)
    n, dim, C, nq, k, nprobe = 45000, 64, 40, 64, 10, 6
    data = rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy() + 0.01
    offsets, ids = O.inverted_lists(O.assign(data, cent, workers=2), C)
    queries = rng.random((nq, dim), dtype=np.float32)
    bounds = [0, 9000, 30000, n]
    keys, counts = [], []
    used = 0
    for s in range(3):
        lo, hi = bounds[s], bounds[s + 1]
        l_off, l_ids = shard_index(offsets, ids, lo, hi)
        ix = ctx.ivf_from_bytes(index_to_bytes(cent, l_off, l_ids))
        ds = ctx.dataset_from(data[lo:hi])
        kq, cq = ix.search_batch_keys(ds, queries, k, nprobe, SQRT, pos_base=lo)
        used += ctx.last_batch_timing()["queries"] > 0
        keys.append(kq)
        counts.append(cq)
        ix.drop()
        ds.drop()
    assert used == 3
    rows, dd, cnt, need = P.merge_batch_keys(np.stack(keys), np.stack(counts), k, SQRT)
    for i in range(nq):
        er, ed = O.topk_rerank_gather(queries[i], data, O.candidate_rows(queries[i], cent, offsets, ids, nprobe), k, 0, True)
        if need[i]:
            continue
        assert cnt[i] == er.size and rows[i, :cnt[i]].tolist() == er.tolist(), i
        assert dd[i, :cnt[i]].view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    assert int(need.sum()) < nq // 2


@pytest.mark.parametrize("n,dim,C,nq,grid", [(40000, 64, 48, 48, False), (9000, 8, 21, 13, True)])
def test_batched_vector_topk_with_a_shared_filter(ctx, n, dim, C, nq, grid):
    rng = np.random.default_rng(n * 3 + nq)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if grid else rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy() + (0.0 if grid else 0.01)
    offsets, ids = O.inverted_lists(O.assign(data, cent, workers=2), C)
    ix = ctx.ivf_from_bytes(O.index_to_bytes(dim, cent, offsets, ids))
    ds = ctx.dataset_from(data)
    queries = (data[rng.integers(0, n, nq)] if grid else rng.random((nq, dim), dtype=np.float32)).copy()
    used = 0
    for mask in (None, rng.random(n) < 0.3, np.arange(n) >= n // 2, np.zeros(n, bool)):
        for nprobe in (2, C):
            for k, flags in ((10, SEQ), (50, SQRT)):
                rows, dist, cnt = ix.vector_topk_batch(ds, queries, k, nprobe, flags, mask)
                bt = ctx.last_batch_timing()
                used += bt["queries"] > 0 and not bt["declined"]
                for i in range(0, nq, max(1, nq // 12)):
                    er, ed, _, _ = _expected_vector_topk(queries[i], data, cent, offsets, ids, k, nprobe,
                                                         1 if flags & SEQ else 0, bool(flags & SQRT), None, mask)
                    assert cnt[i] == er.size, (i, nprobe, k, bt)
                    assert rows[i, :cnt[i]].tolist() == er.tolist(), (i, nprobe, k, bt)
                    assert dist[i, :cnt[i]].view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    if not grid:
        assert used > 0
    ix.drop()
    ds.drop()
