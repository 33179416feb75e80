#!/usr/bin/env python
"""Vector top-k search through the SQL surface (the reference's examples/datafusion_sql.rs).

Optional env vars: PQ_VECTOR_SOURCE, PQ_VECTOR_INDEXED, PQ_VECTOR_QUERY_ROW (default 0)."""
from common import INDEXED, QUERY_ROW, SOURCE, ensure_indexed, read_embedding_at_row

from pq_vector_b200 import SessionStateBuilder, VectorTopKOptions

ensure_indexed(SOURCE, INDEXED)
ctx = SessionStateBuilder().with_pq_vector(VectorTopKOptions(nprobe=8, max_candidates=None)).build()
ctx.register_parquet("t", INDEXED)
query = read_embedding_at_row(INDEXED, "embedding", QUERY_ROW)
literal = "[" + ", ".join(f"{v:.6f}" for v in query) + "]"
df = ctx.sql(f"SELECT title FROM t ORDER BY array_distance(embedding, {literal}) LIMIT 5")
print(df.to_table().to_pandas().to_string(index=False))
print(df.metrics)
