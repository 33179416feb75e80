"""Helpers shared by the examples (the role of the reference's examples/common/mod.rs): locate or create the demo Parquet
file and make sure an indexed copy exists."""
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SOURCE = os.environ.get("PQ_VECTOR_SOURCE", os.path.join(ROOT, "data", "vldb_2025.parquet"))
INDEXED = os.environ.get("PQ_VECTOR_INDEXED", os.path.join(ROOT, "data", "vldb_2025_indexed.parquet"))
QUERY_ROW = int(os.environ.get("PQ_VECTOR_QUERY_ROW", "0"))


def ensure_source(path: str = SOURCE) -> str:
    """The reference ships data/vldb_2025.parquet (496 x 4096 f32).  This repository carries only the embedding column of
    that table as a test fixture; when the Parquet file is absent a stand-in with `id`, `title`, `embedding` is written."""
    if os.path.exists(path):
        return path
    z = np.load(os.path.join(ROOT, "tests", "golden", "vldb_2025_embeddings.npz"))
    emb = np.ascontiguousarray(z["embedding"], dtype=np.float32)
    n, dim = emb.shape
    os.makedirs(os.path.dirname(path), exist_ok=True)
    col = pa.ListArray.from_arrays(pa.array(np.arange(0, (n + 1) * dim, dim, dtype=np.int32)), pa.array(emb.reshape(-1)))
    pq.write_table(pa.table({"id": pa.array(np.arange(n, dtype=np.int32)), "title": pa.array([f"paper {i}" for i in range(n)]),
                             "embedding": col}), path, compression="NONE")
    return path


def ensure_indexed(source: str = SOURCE, indexed: str = INDEXED) -> str:
    from pq_vector_b200 import IndexBuilder, has_pq_vector_index
    ensure_source(source)
    if not (os.path.exists(indexed) and has_pq_vector_index(indexed)):
        IndexBuilder(source, "embedding").build_new(indexed)
    return indexed


def read_embedding_at_row(path: str, column: str, row: int) -> np.ndarray:
    return np.asarray(pq.read_table(path, columns=[column]).column(0)[row].as_py(), dtype=np.float32)
