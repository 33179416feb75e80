"""GPU parity tests for the tcgen05 assignment filter (pqv_tc.cuh): the emitted u32 argmin must equal the reference's
strict-'<' scan (src/ivf/index.rs:244-257, 395-430) bit for bit, whatever the tensor cores round away."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import pq_vector_b200 as P
    c = P.Context()
    yield c
    c.close()


class forced:
    """PQV_ASSIGN=simt|tc for the duration of a block (read with getenv at call time)."""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        self.old = os.environ.get("PQV_ASSIGN")
        os.environ["PQV_ASSIGN"] = self.path

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("PQV_ASSIGN", None)
        else:
            os.environ["PQV_ASSIGN"] = self.old


class operand_kind:
    """PQV_TC_KIND=tf32 keeps the filter on the f32 rows (kind::tf32) instead of their fp16 shadow (read at call time)."""

    def __init__(self, kind):
        self.kind = kind

    def __enter__(self):
        self.old = os.environ.get("PQV_TC_KIND")
        if self.kind == "tf32":
            os.environ["PQV_TC_KIND"] = "tf32"
        else:
            os.environ.pop("PQV_TC_KIND", None)

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("PQV_TC_KIND", None)
        else:
            os.environ["PQV_TC_KIND"] = self.old


def kmeans_like_centroids(data, c, rng, per=40):
    """centroids as means of random row subsets: close together, like Lloyd's on unclustered data"""
    idx = rng.integers(0, data.shape[0], (c, per))
    return data[idx].mean(axis=1, dtype=np.float64).astype(np.float32)


def check(ctx, data, cent, expect_path=1):
    with forced("tc"):
        got = ctx.kmeans_assign(data, cent)
    t = ctx.last_assign_timing()
    exp = O.assign(data, cent, workers=8)
    bad = np.flatnonzero(got != exp)
    assert bad.size == 0, (bad[:10], got[bad[:10]], exp[bad[:10]], t)
    assert t["path"] == expect_path, t
    assert t["rows"] == data.shape[0]
    return t


@pytest.mark.parametrize("n,dim,c", [(20000, 768, 1024), (5001, 100, 300), (4096, 32, 16), (2500, 1536, 257),
                                     (3000, 36, 1000), (129, 64, 8)])
def test_uniform_rows_near_tied_centroids(ctx, n, dim, c):
    rng = np.random.default_rng(n * 7 + dim + c)
    data = rng.random((n, dim), dtype=np.float32)
    cent = kmeans_like_centroids(data, c, rng)
    t = check(ctx, data, cent)
    assert t["ambiguous_rows"] + t["overflow_rows"] <= n


@pytest.mark.parametrize("kind,code", [("f16", 1), ("tf32", 0)])
def test_both_operand_kinds(ctx, kind, code):
    """the same sweeps through the fp16-shadow filter (kind::f16) and the f32/tf32 one: identical assignments, and the
    narrower fp16 residual leaves fewer rows for the exact re-check"""
    rng = np.random.default_rng(77)
    data = rng.random((30000, 768), dtype=np.float32)
    cent = kmeans_like_centroids(data, 1024, rng)
    with operand_kind(kind):
        t = check(ctx, data, cent)
        assert t["kind"] == code, t
        # scale invariance of the fp16 operands: the shadow is scaled by a power of two, tiny / large magnitudes keep their window
        for scale in (2.0 ** -20, 2.0 ** 15):
            t2 = check(ctx, (data * np.float32(scale)).astype(np.float32), (cent * np.float32(scale)).astype(np.float32))
            assert t2["kind"] == code and abs(t2["ambiguous_rows"] - t["ambiguous_rows"]) <= 300, (t, t2)
    print(kind, t)


def test_duplicate_centroids_lowest_index_wins(ctx):
    rng = np.random.default_rng(5)
    data = rng.random((6000, 128), dtype=np.float32)
    base = kmeans_like_centroids(data, 40, rng)
    cent = np.concatenate([base, base[::-1], base[:17]])  # every centroid appears 2-3 times -> exact distance ties
    t = check(ctx, data, cent)
    assert t["ambiguous_rows"] + t["overflow_rows"] == 6000  # no row can be decided by the filter alone


def test_clustered_rows_are_decided_by_the_filter(ctx):
    rng = np.random.default_rng(6)
    cent = (rng.standard_normal((512, 256)) * 4).astype(np.float32)
    lab = rng.integers(0, 512, 30000)
    data = (cent[lab] + rng.standard_normal((30000, 256)) * 0.1).astype(np.float32)
    t = check(ctx, data, cent)
    assert t["ambiguous_rows"] + t["overflow_rows"] < 300, t
    got = ctx.kmeans_assign(data, cent)
    assert np.array_equal(got, lab.astype(np.uint32))


def test_non_finite_and_huge_rows_take_the_exact_scan(ctx):
    rng = np.random.default_rng(8)
    data = rng.random((4000, 64), dtype=np.float32)
    cent = kmeans_like_centroids(data, 64, rng)
    data[5, 3] = np.nan
    data[77, 0] = np.inf
    data[78, 1] = -np.inf
    data[300] = 1e20
    data[301, 7] = 3e38
    data[2000] = 0.0
    data[3999] = -1e19
    t = check(ctx, data, cent)
    assert t["overflow_rows"] >= 6


@pytest.mark.parametrize("m", [1, 7, 255, 256, 257, 1500])
def test_overflow_list_lengths_around_the_few_rows_switch(ctx, m):
    """rows the filter hands to the full exact scan: lists of up to 256 rows take few_rows_assign_kernel (one warp per row and
    32 centroids), longer ones the 64-row tiles -- the device-side count picks, both must give the reference's argmin"""
    rng = np.random.default_rng(100 + m)
    n, dim, C = 6000, 96, 70
    data = rng.random((n, dim), dtype=np.float32)
    cent = kmeans_like_centroids(data, C, rng)
    bad = rng.permutation(n)[:m]
    data[bad, rng.integers(0, dim, m)] = 1e20           # huge norm: the window is meaningless, the row goes to the scan
    data[bad[: max(1, m // 3)], 0] = np.inf              # ... and some with no finite distance at all (default cluster 0)
    t = check(ctx, data, cent)
    assert t["overflow_rows"] >= m


def test_non_finite_centroid_table(ctx):
    rng = np.random.default_rng(9)
    data = rng.random((2500, 48), dtype=np.float32)
    cent = kmeans_like_centroids(data, 33, rng)
    cent[4, 2] = np.inf
    cent[9, 0] = np.nan
    t = check(ctx, data, cent)
    assert t["overflow_rows"] == 2500
    cent = np.zeros((20, 48), np.float32)  # all-zero table: every distance ties -> cluster 0
    with forced("tc"):
        assert not ctx.kmeans_assign(data, cent).any()


def test_scaled_and_shifted_data(ctx):
    rng = np.random.default_rng(10)
    for scale, shift in [(1e-3, 0.0), (1e4, 0.0), (1.0, 1000.0), (1e-18, 0.0), (3e12, -1e12)]:
        data = (rng.random((3000, 96)) * scale + shift).astype(np.float32)
        cent = kmeans_like_centroids(data, 200, rng)
        check(ctx, data, cent)


def test_unit_norm_rows(ctx):
    rng = np.random.default_rng(11)
    data = rng.standard_normal((10000, 384)).astype(np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    cent = kmeans_like_centroids(data, 100, rng, per=8)
    check(ctx, data, cent)


def test_resident_dataset_and_simt_agree_at_scale(ctx):
    """300k x 768 against 1024 centroids: the tensor-core path and the exact SIMT kernel (itself pinned to the oracle)."""
    n, dim, c = 300_000, 768, 1024
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, 1234)
    rng = np.random.default_rng(12)
    sample = ds.read(0, 50_000)
    cent = kmeans_like_centroids(sample, c, rng, per=48)
    with forced("simt"):
        exp = ctx.kmeans_assign(ds, cent)
        ts = ctx.last_assign_timing()
    with forced("tc"):
        got = ctx.kmeans_assign(ds, cent)
        tt = ctx.last_assign_timing()
    assert ts["path"] == 0 and tt["path"] == 1
    assert np.array_equal(got, exp)
    # a slice against the CPU oracle as well
    assert np.array_equal(got[:4000], O.assign(sample[:4000], cent, workers=8))
    print("simt", ts, "\ntc", tt)
    d, a = ctx.bench_assign(ds, cent, iters=2, want_assign=True)
    assert np.array_equal(a, exp)
    ds.drop()


def test_default_path_selection(ctx):
    rng = np.random.default_rng(13)
    data = rng.random((4096, 64), dtype=np.float32)
    cent = data[:32].copy()
    os.environ.pop("PQV_ASSIGN", None)
    ctx.kmeans_assign(data, cent)
    assert ctx.last_assign_timing()["path"] == 1
    ctx.kmeans_assign(data[:100], cent)               # tiny: exact SIMT
    assert ctx.last_assign_timing()["path"] == 0
    ctx.kmeans_assign(np.ascontiguousarray(data[:, :63]), np.ascontiguousarray(cent[:, :63]))  # dim % 4 != 0
    assert ctx.last_assign_timing()["path"] == 0


@pytest.mark.parametrize("n,dim,c", [(12000, 64, 2048), (9000, 128, 3000), (6000, 32, 4096)])
def test_wide_tables_take_the_unstaged_constants(ctx, n, dim, c):
    """more centroids than the epilogue stages in shared memory (norms / weights / group maxima then come from global
    memory, and the 60-bit group stack runs at its widest indices)"""
    rng = np.random.default_rng(n + c)
    data = rng.random((n, dim), dtype=np.float32)
    cent = kmeans_like_centroids(data, c, rng, per=3)
    check(ctx, data, cent)


def test_heavy_tailed_centroid_spread(ctx):
    """a table as k-means++ + a few Lloyd rounds leave it: most centroids near the data mean, some single-member clusters
    (raw rows, |c - mu| ten times larger), an empty cluster at the origin (index.rs:446-453) -- the far-out columns inflate
    the error weight of their chunk, and used to keep the window of 31 neighbours open (group store overflow)"""
    rng = np.random.default_rng(99)
    n, dim, c = 40000, 256, 1024
    data = rng.random((n, dim), dtype=np.float32)
    cent = kmeans_like_centroids(data, c, rng, per=64)
    far = rng.choice(c, 120, replace=False)
    cent[far] = data[rng.choice(n, 120, replace=False)]          # single-member clusters
    cent[far[:3]] = 0.0                                           # empty clusters
    t = check(ctx, data, cent)
    assert t["overflow_rows"] < n // 100, t                       # the store copes: almost nothing goes to the exact scan
