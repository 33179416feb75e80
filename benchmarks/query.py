#!/usr/bin/env python
"""The reference's benches/query.rs, flow for flow, through the interface mirror (pq_vector_b200/session.py + builders.py):
synthetic Parquet (uniform [0,1) f32, `id` + `embedding` List<Float32>, batches of 2048 rows -- benches/bench_util.rs:12-58;
the value stream is this repo's counter-based generator, data seed 1234, query seed 7), then the SAME SQL

    SELECT id FROM t ORDER BY array_distance(embedding, [..]) LIMIT k                         benches/query.rs:76-81

  * without index : a plain SessionContext -> the stock plan: Float64 array_distance over every row + TopK
                    (benches/query.rs:83-103; here pqv_array_distance_topk on the HBM-resident column)
  * build         : IndexBuilder(..).build_new / build_inplace, default C = ceil(sqrt(N))   (benches/query.rs:105-150)
  * with index    : SessionStateBuilder().with_pq_vector(VectorTopKOptions{nprobe, max_candidates}) -> VectorTopKExec
                    (benches/query.rs:152-187)

and recall@k of the second against the first (benches/query.rs:189-190).  Flags follow benches/query.rs:214-282.  Defaults
are the reference's K = 100, NPROBE = 16; ROWS x DIM default to 200 000 x 1024 (the reference's 1 M x 1024 is
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
from pq_vector_b200 import session as S  # noqa: E402

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
ap.add_argument("--max-candidates", type=int, default=None)
ap.add_argument("--build-mode", default="inplace", choices=["inplace", "rewrite", "both"])
ap.add_argument("--metrics", action="store_true")
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

lit = "[" + ", ".join(repr(float(v)) for v in query) + "]"
sql = f"SELECT {a.id_column} FROM t ORDER BY array_distance({a.vector_column}, {lit}) LIMIT {a.k}"


def run(ctx_, reps):
    t0 = time.perf_counter()
    df = ctx_.sql(sql)
    first = df.to_table()
    first_s = time.perf_counter() - t0
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        df = ctx_.sql(sql)
        df.to_table()
        lat.append(time.perf_counter() - t0)
    return first.column(0).to_pylist(), first_s, float(np.median(lat)) if lat else first_s, df.metrics


# ---- no index: plain session, stock plan (first call loads the column into HBM; the reference re-reads the file per query)
plain = S.SessionStateBuilder().build()
plain.register_parquet("t", path)
plain_keys, plain_first_s, plain_s, plain_metrics = run(plain, a.reps)
out = {"config": f"{rows} x {query.size} f32 Parquet, k={a.k}, nprobe={a.nprobe}, clusters={a.n_clusters or int(np.ceil(np.sqrt(rows)))}",
       "generate_seconds": gen_s, "parquet_mb": size0 / 1e6,
       "no_index_query_first_call_seconds": plain_first_s, "no_index_query_seconds": plain_s, "no_index_qps": 1.0 / plain_s}
if a.metrics:
    out["no_index_metrics"] = plain_metrics

# ---- build (rewrite and / or in place), then the same SQL with the optimizer rule registered
targets = []
if a.build_mode in ("rewrite", "both"):
    dst = path + ".indexed.parquet"
    t0 = time.perf_counter()
    ib = B.IndexBuilder(path, a.vector_column)
    if a.n_clusters:
        ib = ib.n_clusters(a.n_clusters)
    ib.build_new(dst)
    out["build_new_seconds"] = time.perf_counter() - t0
    out["index_overhead_mb_rewrite"] = (os.path.getsize(dst) - size0) / 1e6
    targets.append(("rewrite", dst))
if a.build_mode in ("inplace", "both"):
    t0 = time.perf_counter()
    ib = B.IndexBuilder(path, a.vector_column)
    if a.n_clusters:
        ib = ib.n_clusters(a.n_clusters)
    ib.build_inplace()
    out["build_inplace_seconds"] = time.perf_counter() - t0
    out["index_overhead_mb"] = (os.path.getsize(path) - size0) / 1e6
    targets.append(("inplace", path))
opts = B.VectorTopKOptions(nprobe=a.nprobe, max_candidates=a.max_candidates)
for label, p in targets:
    S.drop_resident()                         # time a cold first call: file -> HBM (table + index)
    ictx = S.SessionStateBuilder().with_pq_vector(opts).build()
    ictx.register_parquet("t", p)
    keys, first_s, med_s, metrics = run(ictx, a.reps)
    out[f"indexed_query_first_call_seconds_{label}"] = first_s
    out[f"indexed_query_seconds_{label}"] = med_s
    out[f"indexed_qps_{label}"] = 1.0 / med_s
    out[f"recall_at_k_{label}"] = len(set(keys) & set(plain_keys)) / max(len(plain_keys), 1)     # benches/query.rs:562-569
    if a.metrics:
        out[f"indexed_metrics_{label}"] = metrics
    # nprobe = all clusters must return the exhaustive answer (same rows; f32 sequential sums vs the stock plan's f64)
    actx = S.SessionStateBuilder().with_pq_vector(B.VectorTopKOptions(nprobe=1 << 30)).build()
    actx.register_parquet("t", p)
    out[f"nprobe_all_equals_exhaustive_ids_{label}"] = sorted(actx.sql(sql).to_table().column(0).to_pylist()) == sorted(plain_keys)
    if p != path:
        os.remove(p)
print(json.dumps(out))
if tmpdir:
    os.remove(path)
    os.rmdir(tmpdir)
