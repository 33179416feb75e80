#!/usr/bin/env python
"""Build an IVF index and write it to a new Parquet file (the reference's examples/build_index.rs).

Optional env vars: PQ_VECTOR_SOURCE (default data/vldb_2025.parquet), PQ_VECTOR_INDEXED (default data/vldb_2025_indexed.parquet)."""
from common import INDEXED, SOURCE, ensure_source

from pq_vector_b200 import IndexBuilder

ensure_source(SOURCE)
print(f"Building IVF index from {SOURCE}...")
IndexBuilder(SOURCE, "embedding").build_new(INDEXED)
print(f"Wrote indexed parquet to {INDEXED}")
