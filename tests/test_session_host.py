"""CPU-only: the host logic of the DataFusion-surface mirror (pq_vector_b200/session.py) -- the query shape the
reference's optimizer rule recognises (src/df_vector/physical.rs:134-229), the candidate round-robin
(src/df_vector/access.rs:193-243) and the filter mask.  No distance is computed here."""
import numpy as np
import pyarrow as pa
import pytest

from pq_vector_b200.builders import PqVectorError, VectorTopKOptions
from pq_vector_b200.session import CandidateCursor, ParsedQuery, SessionStateBuilder, _filter_mask


def test_reference_test_queries_parse():
    q = ParsedQuery("SELECT id, vec FROM t WHERE id >= 2 ORDER BY array_distance(vec, [0.0, 0.0]) LIMIT 2")   # tests.rs:77-80
    assert (q.columns, q.table, q.column, q.k) == (["id", "vec"], "t", "vec", 2)
    assert q.literal.tolist() == [0.0, 0.0] and q.predicates == [("id", ">=", 2)]
    q = ParsedQuery("select title from t\n order by array_distance([1, -2.5e-1, 3.], embedding) asc limit 3;")  # either order, physical.rs:204-211
    assert q.column == "embedding" and q.literal.tolist() == [1.0, -0.25, 3.0] and q.columns == ["title"]
    q = ParsedQuery("SELECT * FROM docs WHERE a < 3 AND b = 'x' and c <> 1.5 ORDER BY array_distance(v, [1]) LIMIT 10")
    assert q.columns is None and q.predicates == [("a", "<", 3), ("b", "=", "x"), ("c", "<>", 1.5)]


@pytest.mark.parametrize("sql", [
    "SELECT id FROM t ORDER BY id LIMIT 2",                                      # not array_distance: the rule does not fire
    "SELECT id FROM t ORDER BY array_distance(vec, [0.0]) DESC LIMIT 2",         # physical.rs:143
    "SELECT id FROM t ORDER BY array_distance(vec, [0.0])",                      # no fetch: physical.rs:163-168
    "SELECT id FROM t ORDER BY array_distance(vec, other) LIMIT 2",              # column x column
    "SELECT id FROM t WHERE id + 1 > 2 ORDER BY array_distance(vec, [0.0]) LIMIT 2",
    "SELECT id FROM t ORDER BY array_distance(vec, [a, b]) LIMIT 2",
])
def test_unsupported_shapes_are_refused(sql):
    with pytest.raises(PqVectorError):
        ParsedQuery(sql)


def test_candidate_cursor_round_robin():
    c = CandidateCursor(3)                      # access.rs:193-243, stepped by hand
    c.add_candidates(0, [1, 2, 3])
    c.add_candidates(1, [10])
    c.add_candidates(2, [20, 21])
    assert c.next_batch(4) == [(0, 1), (1, 10), (2, 20), (0, 2)]
    assert c.next_batch(10) == [(2, 21), (0, 3)]         # resumes at the file after the last one served
    assert c.next_batch(5) == []
    assert CandidateCursor(0).next_batch(3) == [] and CandidateCursor(2).next_batch(0) == []


def test_filter_mask_drops_null_predicates():
    t = pa.table({"id": pa.array([0, 1, None, 3], pa.int32()), "tag": ["a", "b", "a", None]})
    assert _filter_mask(t, []) is None
    assert _filter_mask(t, [("id", ">=", 1)]).tolist() == [False, True, False, True]
    assert _filter_mask(t, [("id", ">=", 0), ("tag", "=", "a")]).tolist() == [True, False, False, False]
    with pytest.raises(PqVectorError, match="No field named nope"):
        _filter_mask(t, [("nope", "=", 1)])


def test_builder_defaults_and_errors(tmp_path):
    ctx = SessionStateBuilder().with_pq_vector().build()
    assert ctx.options == VectorTopKOptions(nprobe=5, max_candidates=None)      # options.rs:13-19
    assert SessionStateBuilder().build().options is None                         # no rule: stock plan
    with pytest.raises(PqVectorError, match="not found"):
        ctx.register_parquet("t", str(tmp_path / "missing.parquet"))
    with pytest.raises(PqVectorError, match="table 'u' not found"):
        ctx.sql("SELECT id FROM u ORDER BY array_distance(vec, [0.0]) LIMIT 1").to_table()
    assert ctx.sql("SELECT id FROM u ORDER BY array_distance(vec, [0.0]) LIMIT 1").explain()["operator"] == "VectorTopKExec"
    assert np.isfinite(ParsedQuery("SELECT a FROM t ORDER BY array_distance(v, [1e3]) LIMIT 1").literal).all()
