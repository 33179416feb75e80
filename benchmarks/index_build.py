#!/usr/bin/env python
"""The reference's benches/index_build.rs through the interface mirror: synthetic Parquet (benches/bench_util.rs:12-58
shape: `id` + `embedding` List<Float32>, batches of 2048 rows), copied, then IndexBuilder(..).build_inplace() with the
defaults; prints build time and index overhead.  ROWS x DIM default to 200 000 x 1024 (the reference's 1 M x 1024 is
`--rows 1000000`, a 4 GB file)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pq_vector_b200 import builders as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200_000)
ap.add_argument("--dim", type=int, default=1024)
a = ap.parse_args()
BATCH_ROWS = 2048
ctx = B.context()
tmp = tempfile.mkdtemp(prefix="pqv_index_build_")
src, work = os.path.join(tmp, "index_build_source.parquet"), os.path.join(tmp, "index_build_work.parquet")
t0 = time.perf_counter()
gen = ctx.dataset(a.dim, BATCH_ROWS)
schema = pa.schema([pa.field("id", pa.int32(), False), pa.field("embedding", pa.list_(pa.field("item", pa.float32())), False)])
with pq.ParquetWriter(src, schema, compression="NONE") as w:
    for r0 in range(0, a.rows, BATCH_ROWS):
        cnt = min(BATCH_ROWS, a.rows - r0)
        gen.fill_synthetic(cnt, 1234, stream_first_row=r0)
        emb = pa.ListArray.from_arrays(pa.array(np.arange(0, (cnt + 1) * a.dim, a.dim, dtype=np.int32)), pa.array(gen.read(0, cnt).reshape(-1)))
        w.write_batch(pa.record_batch([pa.array(np.arange(r0, r0 + cnt, dtype=np.int32)), emb], schema=schema))
gen.drop()
gen_s = time.perf_counter() - t0
shutil.copy(src, work)
t0 = time.perf_counter()
B.IndexBuilder(work, "embedding").build_inplace()
build_s = time.perf_counter() - t0
s0, s1 = os.path.getsize(src), os.path.getsize(work)
print(json.dumps({"rows": a.rows, "dim": a.dim, "generate_seconds": gen_s, "source_mb": s0 / 1e6, "index_build_seconds": build_s,
                  "indexed_mb": s1 / 1e6, "index_overhead_mb": (s1 - s0) / 1e6, "index_overhead_pct": (s1 - s0) / s0 * 100.0,
                  "device_build_stats_ms": B._resident_index(work)[0].build_stats()}))
shutil.rmtree(tmp)
