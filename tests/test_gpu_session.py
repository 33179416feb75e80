"""GPU: the DataFusion-surface mirror (pq_vector_b200/session.py) end to end, written after the reference's own SQL tests:

  vector_topk_end_to_end                               src/df_vector/tests.rs:16-104    ids [5, 2]; candidate_rows 6, embeddings_fetched 4
  vector_topk_applies_filters_after_candidate_pruning  src/df_vector/tests.rs:151-241   ids [3, 4]; embeddings_fetched 3
  vector_topk_vldb_tree_snapshot                       src/df_vector/tests.rs:106-149   candidate_rows 496 at nprobe 32
plus the stock plan (no optimizer rule: the built-in array_distance under SortExec(TopK), benches/query.rs:76-103)."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from pq_vector_b200 import builders, session
    yield session
    session.drop_resident()
    builders.set_context(None)


def _write(path, rows, **extra):
    cols = {"id": pa.array(list(range(len(rows))), pa.int32()), "vec": pa.array(rows, pa.list_(pa.float32()))}
    cols.update(extra)
    pq.write_table(pa.table(cols), path, compression="NONE")


def _indexed(tmp_path, rows, name="indexed.parquet", **kw):
    from pq_vector_b200 import IndexBuilder
    src, out = str(tmp_path / ("src_" + name)), str(tmp_path / name)
    _write(src, rows, **kw)
    IndexBuilder(src, "vec").build_new(out)
    return src, out


def _ctx(S, nprobe=64, max_candidates=None):
    from pq_vector_b200 import VectorTopKOptions
    return S.SessionStateBuilder().with_physical_optimizer_rule(VectorTopKOptions(nprobe, max_candidates)).build()


def test_vector_topk_end_to_end(S, tmp_path):
    rows = [[0.0, 0.0], [1.0, 0.0], [0.0, 2.0], [5.0, 5.0], [2.0, 2.0], [0.1, 0.1]]          # tests.rs:31-39
    _, indexed = _indexed(tmp_path, rows)
    ctx = _ctx(S)
    ctx.register_parquet("t", indexed)
    df = ctx.sql("SELECT id, vec FROM t WHERE id >= 2 ORDER BY array_distance(vec, [0.0, 0.0]) LIMIT 2")
    batches = df.collect()
    ids = [i for b in batches for i in b.column(0).to_pylist()]
    assert ids == [5, 2]                                                                      # tests.rs:99
    assert batches[0].schema.names == ["id", "vec"] and batches[0].column(1).to_pylist()[0] == pytest.approx([0.1, 0.1])
    m = df.metrics                                                                            # ...plan_tree.snap:10,20
    assert (m["candidate_rows"], m["embeddings_fetched"], m["batches_fetched"], m["files_scanned"]) == (6, 4, 1, 1)
    assert (m["k"], m["nprobe"], m["query_dim"], m["column"]) == (2, 64, 2, "vec")


def test_vector_topk_applies_filters_after_candidate_pruning(S, tmp_path):
    rows = [[0.0, 0.0], [0.05, 0.05], [0.2, 0.2], [1.0, 1.0], [1.1, 1.1], [1.4, 1.4]]        # tests.rs:166-174
    _, indexed = _indexed(tmp_path, rows)
    ctx = _ctx(S)
    ctx.register_parquet("t", indexed)
    df = ctx.sql("SELECT id FROM t WHERE id >= 3 ORDER BY array_distance(vec, [0.0, 0.0]) LIMIT 2")
    assert df.to_table().column("id").to_pylist() == [3, 4]                                   # tests.rs:235
    assert df.metrics["embeddings_fetched"] == 3 and df.metrics["candidate_rows"] == 6        # ...filter_plan_tree.snap:15


def test_vector_topk_vldb(S, vldb, tmp_path):
    _, indexed = _indexed(tmp_path, vldb.tolist(), name="vldb_indexed.parquet",
                          title=pa.array([f"paper {i}" for i in range(len(vldb))]))
    ctx = _ctx(S, nprobe=32, max_candidates=2048)                                             # tests.rs:114-117
    ctx.register_parquet("t", indexed)
    lit = "[" + ", ".join(repr(float(v)) for v in vldb[0]) + "]"
    df = ctx.sql(f"SELECT title FROM t ORDER BY array_distance(vec, {lit}) LIMIT 3")
    got = df.to_table().column("title").to_pylist()
    assert df.metrics["candidate_rows"] == 496 and df.metrics["query_dim"] == 4096            # ...vldb_tree.snap:22,29
    er, _ = O.topk_rerank(vldb[0], vldb, None, 3, 1, False)                                   # sequential order, exec.rs:529-533
    assert got == [f"paper {i}" for i in er.tolist()] and er.tolist() == [0, 126, 81]         # SURVEY 8c golden ids


def test_max_candidates_caps_the_scored_rows(S, tmp_path):
    data = O.synth(3000, 16, 1234)
    _, indexed = _indexed(tmp_path, data.tolist())
    q = O.synth(1, 16, 7)[0]
    lit = "[" + ", ".join(repr(float(v)) for v in q) + "]"
    ctx = _ctx(S, nprobe=4, max_candidates=100)
    ctx.register_parquet("t", indexed)
    df = ctx.sql(f"SELECT id FROM t ORDER BY array_distance(vec, {lit}) LIMIT 5")
    ids = df.to_table().column("id").to_pylist()
    from pq_vector_b200 import builders as B
    blob, _ = B.read_index_payload(indexed)
    dim, cent, offsets, lists = O.index_from_bytes(blob)
    cand = O.candidate_rows(q, cent, offsets, lists, 4)
    assert df.metrics["candidate_rows"] == cand.size and df.metrics["embeddings_fetched"] == 100
    rows = np.sort(cand[:100])                                                                # first 100 in rank order, then file order
    er, _ = O.topk_rerank_gather(q, data, rows, 5, 1, False)
    assert ids == er.tolist()


def test_two_files_one_heap(S, tmp_path):
    a, b = O.synth(400, 8, 1), O.synth(300, 8, 2)
    _, ia = _indexed(tmp_path, a.tolist(), name="a.parquet")
    _, ib = _indexed(tmp_path, b.tolist(), name="b.parquet")
    ctx = _ctx(S, nprobe=1000)
    ctx.register_parquet("t", [ia, ib])
    q = O.synth(1, 8, 3)[0]
    lit = "[" + ", ".join(repr(float(v)) for v in q) + "]"
    df = ctx.sql(f"SELECT id, vec FROM t ORDER BY array_distance(vec, {lit}) LIMIT 7")
    t = df.to_table()
    both = np.concatenate([a, b])
    er, _ = O.topk_rerank(q, both, None, 7, 1, False)
    got_vecs = np.array(t.column("vec").to_pylist(), np.float32)
    assert np.array_equal(got_vecs, both[er])
    assert df.metrics["candidate_rows"] == 700 and df.metrics["files"] == 2


def test_rule_on_a_file_without_index_is_an_error(S, tmp_path):
    src = str(tmp_path / "plain.parquet")
    _write(src, [[0.0, 0.0], [1.0, 1.0]])
    from pq_vector_b200 import PqVectorError
    ctx = _ctx(S)
    ctx.register_parquet("t", src)
    with pytest.raises(PqVectorError, match="Missing pq-vector index metadata"):               # index_exec.rs:116-121
        ctx.sql("SELECT id FROM t ORDER BY array_distance(vec, [0.0, 0.0]) LIMIT 1").collect()
    _, indexed = _indexed(tmp_path, [[0.0, 0.0], [1.0, 1.0]])
    ctx.register_parquet("u", indexed)
    with pytest.raises(PqVectorError, match="Query dimension mismatch: expected 2, got 3"):    # index_exec.rs:152-158
        ctx.sql("SELECT id FROM u ORDER BY array_distance(vec, [0.0, 0.0, 1.0]) LIMIT 1").collect()


def test_stock_plan_without_the_rule(S, tmp_path):
    """No with_pq_vector: DataFusion's own plan -- f64 array_distance over every row that passes the filter, top-k."""
    data = O.synth(5000, 24, 1234)
    src = str(tmp_path / "plain.parquet")
    _write(src, data.tolist())
    q = O.synth(1, 24, 7)[0].astype(np.float64) + 1e-9
    lit = "[" + ", ".join(repr(float(v)) for v in q) + "]"
    ctx = S.SessionStateBuilder().build()
    ctx.register_parquet("t", src)
    df = ctx.sql(f"SELECT id FROM t ORDER BY array_distance(vec, {lit}) LIMIT 10")
    assert df.explain()["operator"] == "SortExec(TopK)"
    er, _ = O.array_distance_topk(data, q, 10)
    assert df.to_table().column("id").to_pylist() == er.tolist()
    df = ctx.sql(f"SELECT id FROM t WHERE id >= 2500 AND id < 4000 ORDER BY array_distance(vec, {lit}) LIMIT 10")
    er, _ = O.array_distance_topk(data[2500:4000], q, 10)
    assert df.to_table().column("id").to_pylist() == (er + 2500).tolist()
    rows = [[0.0, 0.0], [1.0, 0.0], [0.0, 2.0], [5.0, 5.0], [2.0, 2.0], [0.1, 0.1]]
    _write(str(tmp_path / "six.parquet"), rows)
    ctx.register_parquet("six", str(tmp_path / "six.parquet"))
    got = ctx.sql("SELECT id FROM six WHERE id >= 2 ORDER BY array_distance(vec, [0.0, 0.0]) LIMIT 2").to_table()
    assert got.column("id").to_pylist() == [5, 2]              # both arms agree on the reference's example
