"""GPU: the reference-interface mirror (pq_vector_b200/builders.py) end to end, written after the reference's own tests:

  test_build_index_inplace_appends_footer              src/ivf/parquet.rs:623-660
  vector_topk_end_to_end                               src/df_vector/tests.rs:16-104      ids [5, 2]
  vector_topk_applies_filters_after_candidate_pruning  src/df_vector/tests.rs:151-241     ids [3, 4]
  vector_topk_vldb_tree_snapshot                       src/df_vector/tests.rs:106-149     496 candidates at nprobe 32
plus the C1 config of BASELINE.json (TopkBuilder over the vldb table, nprobe = all) against the golden top-10."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
VLDB_IDS = {0: [0, 126, 81, 265, 315, 464, 322, 269, 169, 140], 1: [1, 177, 57, 19, 36, 16, 450, 179, 9, 140],
            100: [100, 181, 400, 352, 448, 476, 36, 198, 370, 213]}


@pytest.fixture(scope="module")
def B():
    from pq_vector_b200 import builders
    yield builders
    builders.set_context(None)


def _write(path, rows, typ=pa.float32(), **extra):
    cols = {"id": pa.array(list(range(len(rows))), pa.int32()), "embedding": pa.array(rows, pa.list_(typ))}
    cols.update(extra)
    pq.write_table(pa.table(cols), path, compression="NONE")


def test_build_index_inplace_appends_footer(B, tmp_path):
    path = str(tmp_path / "vectors.parquet")
    _write(path, [[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])                    # parquet.rs:626-636
    size0 = os.path.getsize(path)
    B.IndexBuilder(path, "embedding").n_clusters(2).max_iters(5).seed(7).build_inplace()
    assert os.path.getsize(path) > size0                                  # parquet.rs:649-650
    assert B.has_pq_vector_index(path)
    blob, column = B.read_index_payload(path)
    dim, cent, offsets, ids = O.index_from_bytes(blob)
    assert column == "embedding" and dim == 2                             # parquet.rs:652-655
    assert cent.shape == (2, 2) and sorted(ids.tolist()) == [0, 1, 2]
    assert pq.read_table(path).column("id").to_pylist() == [0, 1, 2]      # still a valid parquet file
    q = np.array([0.9, 0.1], np.float32)
    res = B.TopkBuilder(path, q).k(2).nprobe(2).search()                  # nprobe = all clusters
    data = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]], np.float32)
    er, ed = O.topk_rerank_gather(q, data, O.candidate_rows(q, cent, offsets, ids, 2), 2, 0, True)
    assert [r.row_idx for r in res] == er.tolist() and sorted(er.tolist()) == [0, 1]
    assert np.array([r.distance for r in res], np.float32).view(np.uint32).tolist() == ed.view(np.uint32).tolist()


def test_c1_vldb_through_the_builders(B, vldb, tmp_path):
    src = str(tmp_path / "vldb.parquet")
    _write(src, vldb.tolist())
    out = str(tmp_path / "vldb_indexed.parquet")
    B.IndexBuilder(src, "embedding").build_new(out)                       # defaults: C = ceil(sqrt(496)) = 23
    assert not B.has_pq_vector_index(src) and B.has_pq_vector_index(out)
    assert pq.read_table(out).column("id").to_pylist() == list(range(496))
    blob, _ = B.read_index_payload(out)
    dim, cent, offsets, ids = O.index_from_bytes(blob)
    assert dim == 4096 and cent.shape[0] == 23 and sorted(ids.tolist()) == list(range(496))
    B.drop_resident()                                                     # force the file path: index + column re-read
    for qrow, want in VLDB_IDS.items():
        res = B.TopkBuilder(out, vldb[qrow]).k(10).nprobe(32).search()    # nprobe 32 >= 23 clusters: brute force
        assert [r.row_idx for r in res] == want
        er, ed = O.topk_rerank_gather(vldb[qrow], vldb, O.candidate_rows(vldb[qrow], cent, offsets, ids, 32), 10, 0, True)
        assert [r.row_idx for r in res] == er.tolist()
        assert np.array([r.distance for r in res], np.float32).view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    # a pruned search (nprobe 3) equals the oracle's walk over the same embedded index
    for qrow in (0, 17, 100, 333):
        res = B.TopkBuilder(out, vldb[qrow]).k(10).nprobe(3).search()
        cand = O.candidate_rows(vldb[qrow], cent, offsets, ids, 3)
        er, ed = O.topk_rerank_gather(vldb[qrow], vldb, cand, 10, 0, True)
        assert [r.row_idx for r in res] == er.tolist()
    with pytest.raises(B.PqVectorError, match="Query dimension mismatch: expected 4096, got 3"):
        B.TopkBuilder(out, [1.0, 2.0, 3.0]).k(1).nprobe(1).search()


def _batches(rows, ids, typ=pa.float32(), list_type=None, batch=2):
    lt = list_type or pa.list_(typ)
    t = pa.table({"id": pa.array(ids, pa.int32()), "embedding": pa.array(rows, lt)})
    return t.to_batches(max_chunksize=batch)


def test_vector_topk_end_to_end(B):
    rows = [[0.0, 0.0], [1.0, 0.0], [0.0, 2.0], [5.0, 5.0], [2.0, 2.0], [0.1, 0.1]]      # tests.rs:31-39
    keep = [i for i in range(6) if i >= 2]                                              # WHERE id >= 2
    out = B.vector_topk(_batches([rows[i] for i in keep], keep), "embedding", [0.0, 0.0], 2)
    assert out.column(0).to_pylist() == [5, 2]                                          # tests.rs:99
    assert out.schema.names == ["id", "embedding"]                                      # whole rows, scan schema
    assert np.array(out.column(1).to_pylist(), np.float32).tolist() == np.array([[0.1, 0.1], [0.0, 2.0]], np.float32).tolist()


def test_vector_topk_applies_filters_after_candidate_pruning(B):
    rows = [[0.0, 0.0], [0.05, 0.05], [0.2, 0.2], [1.0, 1.0], [1.1, 1.1], [1.4, 1.4]]    # tests.rs:166-174
    keep = [i for i in range(6) if i >= 3]
    out = B.vector_topk(_batches([rows[i] for i in keep], keep), "embedding", [0.0, 0.0], 2)
    assert out.column(0).to_pylist() == [3, 4]                                          # tests.rs:235


@pytest.mark.parametrize("flavour", ["list32", "list64", "large64", "fixed32"])
def test_vector_topk_matches_the_oracle_on_every_list_flavour(B, flavour):
    rng = np.random.default_rng(11)
    n, dim, k = 3000, 24, 17
    data = rng.integers(0, 4, (n, dim)).astype(np.float64) + (rng.random((n, dim)) if flavour != "fixed32" else 0)
    typ = pa.float64() if flavour.endswith("64") else pa.float32()
    lt = {"list32": pa.list_(typ), "list64": pa.list_(typ), "large64": pa.large_list(typ), "fixed32": pa.list_(typ, dim)}[flavour]
    rows = data.astype(np.float32 if typ == pa.float32() else np.float64).tolist()
    skip = set()
    if flavour != "fixed32":
        rows[5] = rows[5][:-1]          # wrong length -> skipped (exec.rs:526-528)
        skip.add(5)
    rows[9] = None                      # null row -> skipped (exec.rs:496-498)
    skip.add(9)
    q = rng.integers(0, 4, dim).astype(np.float32)
    out = B.vector_topk(_batches(rows, list(range(n)), list_type=lt, batch=700), "embedding", q, k)
    live = [i for i in range(n) if i not in skip]
    vals = np.array([rows[i] for i in live], dtype=np.float64).astype(np.float32)      # `value as f32` (exec.rs:542)
    er, _ = O.topk_rerank(q, vals, np.array(live, np.uint32), k, 1, False)
    assert out.column(0).to_pylist() == er.tolist()
    # fewer live rows than k, and no rows at all
    few = B.vector_topk(_batches(rows[:12], list(range(12)), list_type=lt), "embedding", q, 50)
    assert few.num_rows == 12 - len(skip)
    none = B.vector_topk([], "embedding", q, 3, schema=out.schema)
    assert none.num_rows == 0 and none.schema == out.schema


def test_resident_tables_are_per_column_and_follow_rewrites(B, tmp_path):
    """the residency cache: one HBM block per (file state, vector column) -- a second vector column of the same file must
    not be answered from the first one's rows -- and a rewritten file releases the blocks of its old state"""
    rng = np.random.default_rng(3)
    a, b = rng.random((300, 8), dtype=np.float32), rng.random((300, 8), dtype=np.float32) + 5
    path = str(tmp_path / "two.parquet")
    _write(path, a.tolist(), other=pa.array(b.tolist(), pa.list_(pa.float32())))
    da, _, _ = B._resident_table(path, "embedding")
    db, _, _ = B._resident_table(path, "other")
    assert da.handle != db.handle
    assert np.array_equal(da.read(0, 300), a) and np.array_equal(db.read(0, 300), b)
    assert B._resident_table(path, "embedding")[0].handle == da.handle          # a hit, not a reload
    n_before = len(B._tables)
    c = rng.random((400, 8), dtype=np.float32)
    _write(path, c.tolist(), other=pa.array((c + 1).tolist(), pa.list_(pa.float32())))   # new size / mtime
    dc, rows, _ = B._resident_table(path, "embedding")
    assert rows == 400 and np.array_equal(dc.read(0, 400), c)
    assert len(B._tables) == n_before - 1                                        # both old blocks gone, one new
    with pytest.raises(Exception):
        da.read(0, 1)                                                           # the stale block was dropped on the device
