#!/usr/bin/env python
"""The un-indexed `array_distance` arm (benches/query.rs:79-81 ground-truth arm): N x 768 resident
rows, `ORDER BY array_distance(col, literal) LIMIT k` = f64 distance column + exact f64 top-k, through
pqv_array_distance_topk with host query/result buffers; the oracle's same loop timed on a row sample beside it."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402   (cpu baseline only)
import pq_vector_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--cpu-rows", type=int, default=500_000)
a = ap.parse_args()

ctx = P.Context([0])
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234)
queries = O.synth(a.reps + 2, a.dim, 7).astype(np.float64)
out = {"config": f"{a.rows} x {a.dim} f32 resident, f64 array_distance + top-{a.k}"}
for metric, name in ((P.PQV_METRIC_L2, "l2"), (P.PQV_METRIC_COSINE, "cosine")):
    for i in range(2):
        ds.array_distance_topk(queries[i], a.k, metric)
    ts, scan, sel = [], [], []
    for i in range(a.reps):
        t0 = time.perf_counter()
        rows, dist = ds.array_distance_topk(queries[2 + i], a.k, metric)
        ts.append(time.perf_counter() - t0)
        t = ctx.last_timing()
        scan.append(t["scan_ms"])
        sel.append(t["post_ms"])
    # spot check: recompute the winners of the last query from regenerated rows
    ok = all(O.array_distance_column(O.synth(1, a.dim, 1234, first_row=int(r)), queries[a.reps + 1], metric)[0] == d
             for r, d in zip(rows.tolist(), dist.tolist()))
    e2e = float(np.median(ts))
    sm = float(np.median(scan))
    out[name] = {"e2e_ms": e2e * 1e3, "e2e_qps": 1.0 / e2e, "distance_kernel_ms": sm, "select_ms": float(np.median(sel)),
                 "distance_kernel_gbs": a.rows * a.dim * 4 / (sm * 1e-3) / 1e9, "winners_bit_exact_vs_oracle": ok}
host = O.synth(a.cpu_rows, a.dim, 1234)
t0 = time.perf_counter()
O.array_distance_topk(host, queries[0], a.k)
dt = time.perf_counter() - t0
out["cpu_oracle"] = {"rows": a.cpu_rows, "seconds": dt, "qps_extrapolated": 1.0 / (dt * a.rows / a.cpu_rows), "cores": 1}
print(json.dumps(out))
