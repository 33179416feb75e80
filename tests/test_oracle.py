"""Pins oracle/ (the C restatement of the reference hot path) against every known answer the
reference's own tests hold for this path, the SURVEY section 8c vldb table, and an independent numpy
float32 emulation of the two summation orders.  CPU only."""
import math

import numpy as np
import pytest

import oracle as O


# ---- numpy float32 step-by-step emulations (independent of the C code) -------------------------
def np_unroll4(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    s = np.float32(0)
    n4 = (a.size // 4) * 4
    d = (a[:n4] - b[:n4]).astype(np.float32)
    sq = (d * d).astype(np.float32).reshape(-1, 4)
    c = ((sq[:, 0] + sq[:, 1]).astype(np.float32) + sq[:, 2]).astype(np.float32)
    c = (c + sq[:, 3]).astype(np.float32)
    for x in c:
        s = np.float32(s + x)
    for i in range(n4, a.size):
        dd = np.float32(a[i] - b[i])
        s = np.float32(s + np.float32(dd * dd))
    return s


def np_seq(v, q):
    v = np.asarray(v, np.float32); q = np.asarray(q, np.float32)
    d = (v - q).astype(np.float32)
    sq = (d * d).astype(np.float32)
    s = np.float32(0)
    for x in sq:
        s = np.float32(s + x)
    return s


class RustHeap:
    """Second restatement (Python) of std BinaryHeap<HeapItem> used to cross-check the C one."""

    def __init__(self):
        self.d = []

    @staticmethod
    def le(a, b):
        return not (a[0] > b[0])

    def sift_up(self, start, pos):
        elt = self.d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if self.le(elt, self.d[parent]):
                break
            self.d[pos] = self.d[parent]
            pos = parent
        self.d[pos] = elt

    def push(self, it):
        self.d.append(it)
        self.sift_up(0, len(self.d) - 1)

    def pop(self):
        last = self.d.pop()
        if self.d:
            self.d[0] = last
            end = len(self.d)
            pos = 0
            elt = self.d[0]
            child = 1
            while child <= max(end - 2, 0) and end >= 2:
                if self.le(self.d[child], self.d[child + 1]):
                    child += 1
                self.d[pos] = self.d[child]
                pos = child
                child = 2 * pos + 1
            if child == end - 1:
                self.d[pos] = self.d[child]
                pos = child
            self.d[pos] = elt
            self.sift_up(0, pos)


def py_heap_topk(dist, rows, k, do_sqrt):
    h = RustHeap()
    for d, r in zip(dist, rows):
        if len(h.d) < k:
            h.push((np.float32(d), int(r)))
        elif d < h.d[0][0]:
            h.pop()
            h.push((np.float32(d), int(r)))
    items = [((np.sqrt(d, dtype=np.float32) if do_sqrt else d), r) for d, r in h.d]
    items = sorted(items, key=lambda t: t[0])  # python sort is stable
    return [r for _, r in items], [d for d, _ in items]


# ---- reference KATs ----------------------------------------------------------------------------
def test_kat_squared_l2_27():
    # src/ivf/index.rs:487-493
    assert abs(float(O.squared_l2_unroll4([1, 2, 3], [4, 5, 6])) - 27.0) < 1e-6
    assert abs(float(O.squared_l2_seq([1, 2, 3], [4, 5, 6])) - 27.0) < 1e-6


@pytest.mark.parametrize("rows,min_id,expect", [
    # src/df_vector/tests.rs:31-39,99  -> ids [5, 2]
    ([(0, 0), (1, 0), (0, 2), (5, 5), (2, 2), (0.1, 0.1)], 2, [5, 2]),
    # src/df_vector/tests.rs:166-174,235 -> ids [3, 4]
    ([(0, 0), (.05, .05), (.2, .2), (1, 1), (1.1, 1.1), (1.4, 1.4)], 3, [3, 4]),
])
def test_kat_vector_topk_ids(rows, min_id, expect):
    rows = np.array(rows, np.float32)
    ids = np.arange(min_id, len(rows), dtype=np.uint32)
    r, d = O.topk_rerank(np.zeros(2, np.float32), rows[min_id:], ids, 2, order=1, do_sqrt=False)
    assert r.tolist() == expect
    r0, _ = O.topk_rerank(np.zeros(2, np.float32), rows[min_id:], ids, 2, order=0, do_sqrt=True)
    assert r0.tolist() == expect


def test_kat_index_blob_roundtrip():
    # src/ivf/index.rs:495-511
    cent = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    offsets = np.array([0, 3, 5], np.uint64)
    ids = np.array([0, 2, 4, 1, 3], np.uint32)
    blob = O.index_to_bytes(3, cent, offsets, ids)
    expect = (np.array([3, 2], "<u4").tobytes() + cent.astype("<f4").tobytes()
              + np.array([3, 0, 2, 4], "<u4").tobytes() + np.array([2, 1, 3], "<u4").tobytes())
    assert blob == expect
    dim, c2, o2, i2 = O.index_from_bytes(blob)
    assert dim == 3 and np.array_equal(c2, cent) and o2.tolist() == [0, 3, 5] and i2.tolist() == ids.tolist()
    with pytest.raises(ValueError, match="too small"):
        O.index_from_bytes(b"\x00" * 7)


def test_build_sizes_rules():
    # src/ivf/index.rs:161-174, 332
    assert O.build_sizes(496) == (23, 24, 24)              # ceil(sqrt 496)=23; 496/20=24
    assert O.build_sizes(10_000_000, 1024) == (1024, 100_000, 50_000)
    assert O.build_sizes(1_000_000) == (1000, 50_000, 50_000)
    assert O.build_sizes(3) == (2, 2, 2)
    with pytest.raises(ValueError):
        O.build_sizes(0)
    with pytest.raises(ValueError):
        O.build_sizes(5, 6)


# ---- SURVEY section 8c vldb table -----------------------------------------------------------------
VLDB_GOLD = {
    0: ([0, 126, 81, 265, 315, 464, 322, 269, 169, 140],
        [0.0, 0.76749963, 0.77252311, 0.78625035, 0.79400265, 0.79552329, 0.81042033, 0.81212163, 0.81365204,
         0.81549746]),
    1: ([1, 177, 57, 19, 36, 16, 450, 179, 9, 140],
        [0.0, 0.77718174, 0.80194002, 0.80471569, 0.81710893, 0.82245362, 0.82514495, 0.83085144, 0.83195382,
         0.83236438]),
    100: ([100, 181, 400, 352, 448, 476, 36, 198, 370, 213],
          [0.0, 0.69029856, 0.73427695, 0.74319357, 0.74727768, 0.75444263, 0.76359934, 0.76441282, 0.77556115,
           0.78026402]),
}


@pytest.mark.parametrize("qrow", [0, 1, 100])
def test_vldb_top10_table(vldb, qrow):
    ids, dist = VLDB_GOLD[qrow]
    r, d = O.topk_rerank(vldb[qrow], vldb, None, 10, order=0, do_sqrt=True)
    assert r.tolist() == ids
    np.testing.assert_allclose(d, np.array(dist, np.float32), rtol=2e-7, atol=0)
    r1, _ = O.topk_rerank(vldb[qrow], vldb, None, 10, order=1, do_sqrt=False)
    assert r1.tolist() == ids


def test_vldb_nprobe_all_is_bruteforce(vldb):
    # snapshot vector_topk_vldb_tree.snap: nprobe=32 >= C=23 -> candidate_rows 496
    n = vldb.shape[0]
    c, sample, _ = O.build_sizes(n)
    assert c == 23
    rng = np.random.default_rng(0)
    cent = vldb[rng.choice(n, c, replace=False)].copy()
    a = O.assign(vldb, cent, workers=3)
    offsets, ids = O.inverted_lists(a, c)
    cand = O.candidate_rows(vldb[0], cent, offsets, ids, 32)
    assert cand.size == 496 and sorted(cand.tolist()) == list(range(496))
    r, d = O.topk_rerank_gather(vldb[0], vldb, cand, 10)
    assert r.tolist() == VLDB_GOLD[0][0]


# ---- bit-exactness of the C code vs the numpy emulation ---------------------------------------
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 7, 8, 127, 128, 768, 1000, 1536, 4096])
def test_distance_bits_match_numpy_emulation(dim):
    rng = np.random.default_rng(dim)
    a = rng.random(dim, dtype=np.float32)
    b = (rng.standard_normal(dim) * 3).astype(np.float32)
    assert O.squared_l2_unroll4(a, b).view(np.uint32) == np_unroll4(a, b).view(np.uint32)
    assert O.squared_l2_seq(a, b).view(np.uint32) == np_seq(a, b).view(np.uint32)
    assert O.squared_l2_seq_f64(a.astype(np.float64) + 1e-12, b).view(np.uint32) == np_seq(a, b).view(np.uint32)


def test_orders_differ_somewhere():
    # the two loops are different functions of the same input
    rng = np.random.default_rng(1)
    diff = 0
    for _ in range(50):
        a = rng.random(768, dtype=np.float32); b = rng.random(768, dtype=np.float32)
        diff += O.squared_l2_unroll4(a, b).view(np.uint32) != O.squared_l2_seq(b, a).view(np.uint32)
    assert diff > 0


def test_distances_vectorised_entry(vldb):
    q = vldb[7]
    d0 = O.distances(vldb[:64], q, 0)
    d1 = O.distances(vldb[:64], q, 1)
    for i in range(64):
        assert d0[i].view(np.uint32) == np_unroll4(q, vldb[i]).view(np.uint32)
        assert d1[i].view(np.uint32) == np_seq(vldb[i], q).view(np.uint32)


# ---- heap semantics ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,k,levels", [(1, 1, 0), (5, 10, 0), (64, 8, 0), (500, 10, 6), (2000, 100, 40),
                                        (3000, 37, 3), (257, 256, 5)])
def test_heap_matches_python_restatement(n, k, levels):
    rng = np.random.default_rng(n * 1000 + k)
    if levels:   # heavy ties: quantised distances
        dist = (rng.integers(0, levels, n) / 4).astype(np.float32)
    else:
        dist = rng.random(n, dtype=np.float32)
    rows = rng.permutation(n).astype(np.uint32)
    for do_sqrt in (False, True):
        r, d = O.heap_topk(dist, rows, k, do_sqrt)
        pr, pd = py_heap_topk(dist, rows, k, do_sqrt)
        assert r.tolist() == pr
        assert [x.view(np.uint32) for x in d] == [np.float32(x).view(np.uint32) for x in pd]
        # size-independent properties: sorted ascending; multiset of distances = k smallest
        assert np.all(np.diff(d) >= 0)
        ks = np.sort(dist)[: min(k, n)]
        if do_sqrt:
            ks = np.sqrt(ks)
        assert np.array_equal(np.sort(d), ks.astype(np.float32))


def test_heap_tie_at_boundary_keeps_earlier():
    # strict '<' (search.rs:121-122): a later candidate equal to the current max never enters
    dist = np.array([1, 1, 1, 1], np.float32)
    r, _ = O.heap_topk(dist, np.array([10, 11, 12, 13], np.uint32), 2, False)
    assert sorted(r.tolist()) == [10, 11]


def test_empty_and_short_inputs():
    r, d = O.heap_topk(np.zeros(0, np.float32), np.zeros(0, np.uint32), 5, True)
    assert r.size == 0 and d.size == 0
    r, d = O.heap_topk(np.array([4.0, 1.0], np.float32), None, 5, True)
    assert r.tolist() == [1, 0] and d.tolist() == [1.0, 2.0]


# ---- IVF pieces -------------------------------------------------------------------------------
def test_nearest_centroid_first_min_and_nan():
    cent = np.array([[1, 1], [0, 0], [0, 0], [5, 5]], np.float32)
    assert O.nearest_centroid([0, 0], cent) == 1          # tie -> lowest index (index.rs:251 strict <)
    assert O.nearest_centroid([np.nan, 0], cent) == 0      # NaN never wins -> cluster 0
    assert O.nearest_centroid([np.inf, 0], cent) == 0


def test_assign_independent_of_worker_split():
    rng = np.random.default_rng(3)
    data = rng.random((1000, 24), dtype=np.float32)
    cent = data[:17].copy()
    a1 = O.assign(data, cent, 1)
    for w in (2, 3, 8, 2000):
        assert np.array_equal(a1, O.assign(data, cent, w))
    for i in range(0, 1000, 97):
        assert a1[i] == O.nearest_centroid(data[i], cent)
    offsets, ids = O.inverted_lists(a1, 17)
    for c in range(17):
        l = ids[int(offsets[c]):int(offsets[c + 1])]
        assert np.all(np.diff(l.astype(np.int64)) > 0)       # ascending row ids (index.rs:202-206)
        assert np.all(a1[l] == c)


def test_find_closest_centroids_stable():
    cent = np.array([[2, 0], [1, 0], [1, 0], [0, 0], [1, 0]], np.float32)
    got = O.find_closest_centroids([0, 0], cent, 10)          # nprobe clamped to C (index.rs:131)
    assert got.tolist() == [3, 1, 2, 4, 0]                    # stable among the three ties
    assert O.find_closest_centroids([0, 0], cent, 2).tolist() == [3, 1]


def test_kmeans_pieces():
    rng = np.random.default_rng(5)
    data = rng.random((300, 10), dtype=np.float32)
    sel = rng.permutation(300)[:120].astype(np.uint64)
    md = O.min_dist_init(data, sel, data[5])
    for s in range(120):
        assert md[s].view(np.uint32) == np_unroll4(data[sel[s]], data[5]).view(np.uint32)
    md1 = md.copy(); md8 = md.copy()
    t1 = O.min_dist_update(data, sel, data[9], md1, workers=1)
    t8 = O.min_dist_update(data, sel, data[9], md8, workers=8)
    assert np.array_equal(md1, md8)
    # total for w workers = in-order sum of per-chunk in-order sums (index.rs:356-370)
    exp = np.float32(0)
    for x in md1:
        exp = np.float32(exp + x)
    assert t1.view(np.uint32) == exp.view(np.uint32)
    chunk = -(-120 // 8); tot = np.float32(0)
    for s0 in range(0, 120, chunk):
        loc = np.float32(0)
        for x in md8[s0:s0 + chunk]:
            loc = np.float32(loc + x)
        tot = np.float32(tot + loc)
    assert t8.view(np.uint32) == tot.view(np.uint32)
    pick = O.kmeanspp_pick(md1, 0.5 * float(t1))
    cs = np.float32(0)
    for i, x in enumerate(md1):
        cs = np.float32(cs + x)
        if cs >= np.float32(0.5 * float(t1)):
            assert pick == i
            break
    # Lloyd step + update incl. empty cluster -> origin
    cent = np.vstack([data[:4], np.full((1, 10), 100, np.float32)])
    assign = np.zeros(300, np.uint32)
    changed, sizes = O.lloyd_assign(data, cent, assign, workers=4)
    assert sizes.sum() == 300 and sizes[4] == 0 and changed == int((assign != 0).sum())
    newc = O.centroid_update(data, assign, sizes, 5)
    assert np.all(newc[4] == 0)
    j = int(np.argmax(sizes))
    acc = np.zeros(10, np.float32)
    for i in np.nonzero(assign == j)[0]:
        acc = (acc + data[i]).astype(np.float32)
    assert np.array_equal(newc[j], (acc / np.float32(sizes[j])).astype(np.float32))


# ---- synthetic generator ----------------------------------------------------------------------
def test_synth_generator_distribution_and_addressing():
    a = O.synth(100, 16, 1234)
    b = O.synth(10, 16, 1234, first_row=50)
    assert np.array_equal(a[50:60], b)                       # keyed by absolute element index
    assert a.min() >= 0 and a.max() < 1
    assert np.all((a * 16777216.0) == np.floor(a * 16777216.0))   # 24-bit grid, as rand's gen::<f32>()
    big = O.synth(2000, 64, 7)
    assert abs(big.mean() - 0.5) < 0.01 and abs(big.var() - 1 / 12) < 0.005
    assert not np.array_equal(O.synth(4, 4, 1), O.synth(4, 4, 2))
    # known answers (pin the stream so the CUDA generator can be checked against constants too)
    assert O.synth(1, 4, 1234).view(np.uint32).ravel().tolist() == SYNTH_KAT


SYNTH_KAT = None  # filled below at import


def _fill_kat():
    global SYNTH_KAT
    def u32(seed, idx):
        M = (1 << 64) - 1
        z = (seed + (idx + 1) * 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        return z >> 32
    SYNTH_KAT = [int(np.float32((u32(1234, i) >> 8) * (1.0 / 16777216.0)).view(np.uint32)) for i in range(4)]


_fill_kat()


# ---- the un-indexed array_distance arm (DataFusion built-in; parity unpinned) ---------------------
def test_array_distance_published_examples_and_independent_fold():
    # DataFusion's documented example: array_distance([1, 2], [1, 4]) = 2.0; the reference's own L2 KAT vectors give sqrt(27)
    assert O.array_distance_column(np.array([[1, 2]], np.float32), [1.0, 4.0]).tolist() == [2.0]
    assert O.array_distance_column(np.array([[1, 2, 3]], np.float32), [4.0, 5.0, 6.0]).tolist() == [np.sqrt(27.0)]
    rng = np.random.default_rng(11)
    for dim in (1, 7, 64, 769):
        x = rng.random((40, dim), dtype=np.float32)
        q = rng.random(dim)                                   # f64 literal, not f32-representable
        t = (x.astype(np.float64) - q) ** 2
        ref = np.sqrt(np.cumsum(t, axis=1)[:, -1])            # numpy cumsum = the same left-to-right f64 fold
        assert O.array_distance_column(x, q).view(np.uint64).tolist() == ref.view(np.uint64).tolist()
        xd = x.astype(np.float64)
        dot = np.cumsum(xd * q, axis=1)[:, -1]
        na = np.cumsum(xd * xd, axis=1)[:, -1]
        nb = np.cumsum(q * q)[-1]
        cos = 1.0 - dot / (np.sqrt(na) * np.sqrt(nb))
        assert O.array_distance_column(x, q, 1).view(np.uint64).tolist() == cos.view(np.uint64).tolist()


def test_array_distance_topk_order():
    x = np.array([[0, 0], [1, 0], [0, 2], [5, 5], [2, 2], [0.1, 0.1]], np.float32)    # df_vector/tests.rs:31-39 rows
    r, d = O.array_distance_topk(x, [0.0, 0.0], 3)
    assert r.tolist() == [0, 5, 1] and d[0] == 0.0 and d[2] == 1.0
    dup = np.concatenate([x, x])
    r, _ = O.array_distance_topk(dup, [0.0, 0.0], 4)
    assert r.tolist() == [0, 6, 5, 11]                                           # equal keys: ascending row
    x2 = x.copy()
    x2[0, 0] = np.inf
    r, d = O.array_distance_topk(x2, [np.inf, 0.0], 6)
    assert r.tolist()[-1] == 0 and np.isnan(d[-1]) and np.isinf(d[:-1]).all()    # NaN sorts last
    r, _ = O.array_distance_topk(x, [0.0, 0.0], 100)
    assert r.size == 6
