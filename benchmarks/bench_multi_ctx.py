#!/usr/bin/env python
"""One process, ONE pqv_ctx over several GPUs (pqv_init with n_devices > 1) -- the form a single Rust host process uses:
rows split into contiguous ranges, one per device.  Measures the public calls with host buffers in and out:

  * pqv_l2_topk, 1 query      (config C2's table cut over the devices: every shard scans its rows, one replay)
  * pqv_l2_topk, 1024 queries (config C5's shape: every shard's tensor-core pass on its own host thread, keys merged)
  * pqv_kmeans_assign over the resident table (every shard sweeps its rows)

  python benchmarks/bench_multi_ctx.py --devices 0,1 --rows 10000000"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--devices", default="0")
ap.add_argument("--rows", type=int, default=10_000_000, help="rows of the whole table")
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
devs = [int(x) for x in a.devices.split(",")]
ctx = P.Context(devs)
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234)
qd = P.Context([devs[0]])
tmp = qd.dataset(a.dim, a.queries)
tmp.fill_synthetic(a.queries, 7)
qs = tmp.read(0, a.queries)
tmp.drop()
qd.close()
out = {"devices": devs, "rows": a.rows, "dim": a.dim}


def timed(fn, reps):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


dt, (r1, d1) = timed(lambda: ds.l2_topk(qs[0], 100, P.PQV_SQRT), a.reps * 3)
out["single_query_top100"] = {"ms": dt * 1e3, "qps": 1 / dt, "gbs": a.rows * a.dim * 4 / dt / 1e9}
dt, (rb, db, cb) = timed(lambda: ds.l2_topk(qs, 10, P.PQV_SUM_SEQ), a.reps)
t = ctx.last_batch_timing()
out["batch_top10"] = {"queries": a.queries, "ms": dt * 1e3, "qps": a.queries / dt, "device_ms": t["total_ms"], "filter_ms": t["filter_ms"],
                      "unresolved_by_the_merge": t["tie_queries"]}
cent = ds.read(0, 1024) + 0.001
dt, asg = timed(lambda: ctx.kmeans_assign(ds, cent), 3)
out["assign_1024"] = {"ms": dt * 1e3, "rows_per_s": a.rows / dt, "device_sweep_ms": ctx.last_assign_timing()["total_ms"]}
# identical to a single-device context on a slice the oracle-checked tests cover: first query, top-100 ids
out["first_ids"] = r1[:5].tolist()
out["batch_q0_ids"] = rb[0, :5].tolist()
out["assign_checksum"] = int(asg.astype(np.uint64).sum())
print(json.dumps(out))
