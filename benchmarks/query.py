#!/usr/bin/env python
"""The reference's benches/query.rs (query latency with / without the index + recall@k) run through the interface mirror:
synthetic Parquet (uniform [0,1) f32, `id` + `embedding` List<Float32>, batches of 2048 rows -- benches/bench_util.rs:12-58;
the value stream is this repo's counter-based generator, data seed 1234, query seed 7), `IndexBuilder(..).build_inplace()`
(default C = ceil(sqrt(N)), benches/query.rs:105-129), then the same query

  * without index : exhaustive scan  -- `vector_topk` over the file's record batches (the un-indexed plan of
                    benches/query.rs:76-103 computes array_distance over every row)
  * with index    : `TopkBuilder(path, query).k(K).nprobe(NPROBE).search()` (benches/query.rs:131-193)

and recall@k of the second against the first (benches/query.rs:192-193, 562-569).  Flags follow benches/query.rs:214-282.
Defaults are the reference's K = 100, NPROBE = 16; ROWS x DIM default to 200 000 x 1024 (the reference's 1 M x 1024 is
`--rows 1000000`, a 4 GB Parquet file)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402
from pq_vector_b200 import builders as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--path", default=None, help="existing Parquet file (default: generate synthetic data)")
ap.add_argument("--vector-column", default="embedding")
ap.add_argument("--id-column", default="id")
ap.add_argument("--rows", type=int, default=200_000)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--nprobe", type=int, default=16)
ap.add_argument("--n-clusters", type=int, default=None)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--query-row", type=int, default=None)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()

BATCH_ROWS = 2048
ctx = B.context()
tmpdir = None
if a.path is None:
    tmpdir = tempfile.mkdtemp(prefix="pqv_query_bench_")
    path = os.path.join(tmpdir, "bench_vectors.parquet")
    t0 = time.perf_counter()
    gen = ctx.dataset(a.dim, BATCH_ROWS)
    schema = pa.schema([pa.field("id", pa.int32(), False), pa.field("embedding", pa.list_(pa.field("item", pa.float32())), False)])
    with pq.ParquetWriter(path, schema, compression="NONE") as w:
        for r0 in range(0, a.rows, BATCH_ROWS):
            cnt = min(BATCH_ROWS, a.rows - r0)
            gen.fill_synthetic(cnt, 1234, stream_first_row=r0)
            vals = gen.read(0, cnt).reshape(-1)
            emb = pa.ListArray.from_arrays(pa.array(np.arange(0, (cnt + 1) * a.dim, a.dim, dtype=np.int32)), pa.array(vals))
            w.write_batch(pa.record_batch([pa.array(np.arange(r0, r0 + cnt, dtype=np.int32)), emb], schema=schema))
    gen.drop()
    gen_s = time.perf_counter() - t0
else:
    path, gen_s = a.path, None
size0 = os.path.getsize(path)
md = pq.read_metadata(path)
rows = md.num_rows

if a.query_row is not None:
    query = np.asarray(pq.read_table(path, columns=[a.vector_column]).column(0)[a.query_row].as_py(), dtype=np.float32)
else:
    dim = len(pq.ParquetFile(path).read_row_group(0, columns=[a.vector_column]).column(0)[0])
    qd = ctx.dataset(dim, 1)
    qd.fill_synthetic(1, 7)
    query = qd.read(0, 1)[0]
    qd.drop()

# ---- no index: exhaustive scan over the file's batches (cold: file -> host -> HBM every time, as the un-indexed plan)
t0 = time.perf_counter()
exact = B.vector_topk(pq.ParquetFile(path).iter_batches(batch_size=65536), a.vector_column, query, a.k)
noindex_s = time.perf_counter() - t0
exact_ids = exact.column(exact.schema.get_field_index(a.id_column)).to_pylist()

# ---- build in place
t0 = time.perf_counter()
ib = B.IndexBuilder(path, a.vector_column)
if a.n_clusters:
    ib = ib.n_clusters(a.n_clusters)
ib.build_inplace()
build_s = time.perf_counter() - t0
size1 = os.path.getsize(path)

# ---- with index (first call: the table and index are already resident from the build; drop them to time a cold start)
B.drop_resident()
t0 = time.perf_counter()
res = B.TopkBuilder(path, query).k(a.k).nprobe(a.nprobe).search()
cold_s = time.perf_counter() - t0
lat = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    res = B.TopkBuilder(path, query).k(a.k).nprobe(a.nprobe).search()
    lat.append(time.perf_counter() - t0)
ids = pq.read_table(path, columns=[a.id_column]).column(0).to_numpy()
got_ids = [int(ids[r.row_idx]) for r in res]
recall = len(set(got_ids) & set(exact_ids)) / max(len(exact_ids), 1)
# nprobe = all clusters must reproduce the exhaustive answer exactly (same rows; TopkBuilder sums in unroll-4 order)
allp = B.TopkBuilder(path, query).k(a.k).nprobe(1 << 30).search()
print(json.dumps({
    "config": f"{rows} x {query.size} f32 Parquet, k={a.k}, nprobe={a.nprobe}, clusters={a.n_clusters or int(np.ceil(np.sqrt(rows)))}",
    "generate_seconds": gen_s, "parquet_mb": size0 / 1e6, "index_overhead_mb": (size1 - size0) / 1e6,
    "no_index_query_seconds": noindex_s, "build_inplace_seconds": build_s,
    "indexed_query_first_call_seconds": cold_s, "indexed_query_seconds": float(np.median(lat)),
    "indexed_qps": 1.0 / float(np.median(lat)), "recall_at_k": recall,
    "nprobe_all_equals_exhaustive_ids": sorted(int(ids[r.row_idx]) for r in allp) == sorted(exact_ids)}))
if tmpdir:
    os.remove(path)
    os.rmdir(tmpdir)
