"""GPU parity tests for the un-indexed `array_distance` arm (pqv_adist.cuh): the Float64
distance column and its top-k must equal the oracle's sequential f64 fold bit for bit.  PARITY UNPINNED against the
reference itself: the UDF is DataFusion's (datafusion-functions-nested 52.1.0, not under /root/reference) and no test of
the reference reaches it -- the oracle restates the published algorithm."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

L2, COS = 0, 1


@pytest.fixture(scope="module")
def ctx():
    import pq_vector_b200 as P
    c = P.Context()
    yield c
    c.close()


def bits(a):
    return np.asarray(a, np.float64).view(np.uint64)


@pytest.mark.parametrize("n,dim", [(5000, 768), (1000, 13), (33, 2), (4097, 66), (300, 130), (64, 4096), (1, 1), (2049, 1536)])
@pytest.mark.parametrize("metric", [L2, COS])
def test_column_bit_exact(ctx, n, dim, metric):
    data = O.synth(n, dim, 1234)
    q = O.synth(1, dim, 7)[0].astype(np.float64) + 1e-9    # a literal that is NOT f32-representable
    ds = ctx.dataset_from(data)
    got = ds.array_distance(q, metric)
    exp = O.array_distance_column(data, q, metric)
    assert bits(got).tolist() == bits(exp).tolist()
    ds.drop()


def test_docs_known_answer(ctx):
    # DataFusion docs: array_distance([1, 2], [1, 4]) = 2.0; and the reference's own KAT vectors (index.rs:487-493): sqrt(27)
    ds = ctx.dataset_from(np.array([[1, 2]], np.float32))
    assert ds.array_distance([1.0, 4.0]).tolist() == [2.0]
    ds.drop()
    ds = ctx.dataset_from(np.array([[1, 2, 3]], np.float32))
    assert ds.array_distance([4.0, 5.0, 6.0]).tolist() == [np.sqrt(27.0)]
    ds.drop()


@pytest.mark.parametrize("n,dim,k", [(20000, 768, 10), (20000, 768, 100), (1000, 13, 1000), (50, 8, 100), (70000, 64, 1024)])
@pytest.mark.parametrize("metric", [L2, COS])
def test_topk_matches_oracle(ctx, n, dim, k, metric):
    data = O.synth(n, dim, 1234)
    q = O.synth(1, dim, 7)[0].astype(np.float64) * 0.999
    ds = ctx.dataset_from(data)
    rows, dist = ds.array_distance_topk(q, k, metric)
    er, ed = O.array_distance_topk(data, q, k, metric)
    assert rows.tolist() == er.tolist()
    assert bits(dist).tolist() == bits(ed).tolist()
    assert rows.size == min(k, n)
    ds.drop()


def test_topk_ties_by_row_and_nan_last(ctx):
    rng = np.random.default_rng(5)
    base = rng.random((40, 16), dtype=np.float32)
    data = np.concatenate([base, base, base[:7]])            # every distance appears 2-3 times
    data[11, 3] = np.inf                                      # inf - inf -> NaN distance for rows 11 (query has inf there)
    data[60, 3] = np.inf
    q = rng.random(16)
    ds = ctx.dataset_from(data)
    for k in (1, 5, 40, 87):
        rows, dist = ds.array_distance_topk(q, k)
        er, ed = O.array_distance_topk(data, q, k)
        assert rows.tolist() == er.tolist(), k
        assert bits(dist).tolist() == bits(ed).tolist()
    q2 = q.copy()
    q2[3] = np.inf
    rows, dist = ds.array_distance_topk(q2, 87)
    er, ed = O.array_distance_topk(data, q2, 87)
    assert rows.tolist() == er.tolist()
    assert np.isnan(dist[-2:]).all() and rows[-2:].tolist() == [11, 60]      # NaN sorts last, ties by row
    assert np.isinf(dist[:-2]).all()
    ds.drop()


def test_all_rows_identical(ctx):
    data = np.tile(np.arange(32, dtype=np.float32), (5000, 1))
    ds = ctx.dataset_from(data)
    rows, dist = ds.array_distance_topk(np.zeros(32), 17)
    assert rows.tolist() == list(range(17)) and len(set(dist.tolist())) == 1
    ds.drop()


def test_errors(ctx):
    import pq_vector_b200 as P
    ds = ctx.dataset_from(O.synth(10, 8, 1))
    with pytest.raises(P.PqvError, match="same length"):
        ds.array_distance(np.zeros(7))
    with pytest.raises(P.PqvError, match="k must be"):
        ds.array_distance_topk(np.zeros(8), 0)
    with pytest.raises(P.PqvError, match="metric"):
        ds.array_distance(np.zeros(8), 9)
    ds.drop()
    empty = ctx.dataset(8, 0)
    r, d = empty.array_distance_topk(np.zeros(8), 3)
    assert r.size == 0 and empty.array_distance(np.zeros(8)).size == 0
    empty.drop()


def test_large_table_properties(ctx):
    """1 M x 768 (3 GB): ascending output; every returned distance recomputed by the oracle from the regenerated row;
    the k-th distance bounds a random sample of non-returned rows."""
    n, dim, k = 1_000_000, 768, 100
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, 1234)
    q = O.synth(1, dim, 7)[0].astype(np.float64)
    rows, dist = ds.array_distance_topk(q, k)
    assert rows.size == k and np.all(np.diff(dist) >= 0) and len(set(rows.tolist())) == k
    for r, d in zip(rows.tolist(), dist):
        v = O.synth(1, dim, 1234, first_row=r)
        assert bits(O.array_distance_column(v, q)).tolist() == bits([d]).tolist()
    rng = np.random.default_rng(0)
    chosen = set(rows.tolist())
    for r in rng.integers(0, n, 200).tolist():
        if r not in chosen:
            assert O.array_distance_column(O.synth(1, dim, 1234, first_row=r), q)[0] >= dist[-1]
    t = ctx.last_timing()
    assert t["scan_bytes"] == n * dim * 4
    ds.drop()


def test_topk_under_a_filter(ctx):
    """WHERE below the sort: rows whose mask bit is clear never reach the top-k (pqv_array_distance_topk_filtered)."""
    import pq_vector_b200 as P
    rng = np.random.default_rng(21)
    n, dim = 33_333, 48
    data = rng.integers(0, 4, (n, dim)).astype(np.float32)          # many equal distances: ties by row under the mask too
    q = rng.random(dim)
    ds = ctx.dataset_from(data)
    for mask in (rng.random(n) < 0.5, np.arange(n) % 97 == 3, np.zeros(n, bool), np.ones(n, bool)):
        keep = np.nonzero(mask)[0]
        for k in (1, 10, 500):
            rows, dist = ds.array_distance_topk(q, k, row_mask=mask)
            if keep.size == 0:
                assert rows.size == 0
                continue
            er, ed = O.array_distance_topk(data[keep], q, k)
            assert rows.tolist() == keep[er].tolist(), k
            assert bits(dist).tolist() == bits(ed).tolist()
    with pytest.raises(P.PqvError, match="row_mask has"):
        ds.array_distance_topk(q, 3, row_mask=np.ones(5, bool))
    ds.drop()
