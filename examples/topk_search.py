#!/usr/bin/env python
"""Run a top-k search against an indexed Parquet file (the reference's examples/topk_search.rs).

Optional env vars: PQ_VECTOR_SOURCE, PQ_VECTOR_INDEXED, PQ_VECTOR_QUERY_ROW (default 0)."""
from common import INDEXED, QUERY_ROW, SOURCE, ensure_indexed, read_embedding_at_row

from pq_vector_b200 import TopkBuilder

ensure_indexed(SOURCE, INDEXED)
query = read_embedding_at_row(INDEXED, "embedding", QUERY_ROW)
results = TopkBuilder(INDEXED, query).k(5).nprobe(5).search()
print(f"Top 5 neighbors for row {QUERY_ROW}:")
for rank, r in enumerate(results, 1):
    print(f"{rank}. row {r.row_idx} distance {r.distance:.4f}")
