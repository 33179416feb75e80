#!/usr/bin/env python
"""Assignment sweep of config C3 (src/ivf/index.rs:189-206): N x 768 rows against 1024 centroids taken from a short
IVF build, device-resident loop (pqv_bench_assign), both paths: the tcgen05 tf32 filter + exact re-check and the exact
SIMT kernel.  Checks the two agree on every row.  Prints one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--clusters", type=int, default=1024)
ap.add_argument("--lloyd-iters", type=int, default=3)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--simt-rows", type=int, default=1_000_000, help="rows for the exact SIMT comparison sweep")
a = ap.parse_args()

ctx = P.Context([0])
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234)
ix = ctx.ivf_build(ds, n_clusters=a.clusters, max_iters=a.lloyd_iters, seed=42)
cent = ix.centroids()
out = {"config": f"{a.rows} x {a.dim} f32 vs {a.clusters} centroids (after {a.lloyd_iters} Lloyd iterations), resident in HBM",
       "build_breakdown_ms": ix.build_stats()}
os.environ["PQV_ASSIGN"] = "tc"
ctx.bench_assign(ds, cent, iters=1)
t, got = ctx.bench_assign(ds, cent, iters=a.iters, want_assign=True)
flops = 2.0 * a.rows * a.clusters * a.dim
t["filter_tflops_tf32"] = flops / (t["filter_ms"] * 1e-3) / 1e12
t["rows_gbs_total"] = a.rows * a.dim * 4 / (t["total_ms"] * 1e-3) / 1e9
out["tcgen05"] = t
os.environ["PQV_ASSIGN"] = "simt"
ns = min(a.simt_rows, a.rows)
ts, exp = ctx.bench_assign(ds, cent, iters=1, n=ns, want_assign=True)
ts["f32_ops_per_s"] = 3.0 * ns * a.clusters * a.dim / (ts["total_ms"] * 1e-3)
out["simt_exact"] = ts
out["paths_agree_on_rows"] = int(ns) if np.array_equal(got[:ns], exp) else -1
out["speedup_same_rows"] = (ts["total_ms"] / ns) / (t["total_ms"] / a.rows)
del os.environ["PQV_ASSIGN"]
print(json.dumps(out))
