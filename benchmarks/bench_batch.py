#!/usr/bin/env python
"""Batched brute-force top-k (BASELINE config 5 shape on one GPU): N x 768 resident rows, nq queries per call, through
pqv_l2_topk with host query/result buffers (e2e) -- the tcgen05 filter + exact re-rank path -- against the same queries
sent one by one (single-query scan path).  Verifies a sample of queries bit for bit between the two paths."""
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
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--flags", type=int, default=P.PQV_SUM_SEQ, help="1 = VectorTopKExec semantics (exec.rs), 2 = TopkBuilder")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=16)
a = ap.parse_args()

ctx = P.Context([0])
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234)
qd = ctx.dataset(a.dim, a.queries)
qd.fill_synthetic(a.queries, 7)
queries = qd.read(0, a.queries)
qd.drop()

ds.l2_topk(queries, a.k, a.flags)  # warm-up (allocations)
ts, tm = [], None
for _ in range(a.reps):
    t0 = time.perf_counter()
    rows, dist, cnt = ds.l2_topk(queries, a.k, a.flags)
    ts.append(time.perf_counter() - t0)
    tm = ctx.last_batch_timing()
e2e = float(np.median(ts))
flops = 2.0 * a.rows * a.queries * a.dim
out = {"config": f"{a.rows} x {a.dim} f32 resident, {a.queries} queries per call, k={a.k}, flags={a.flags}",
       "e2e_seconds_per_batch": e2e, "e2e_qps": a.queries / e2e, "timing": tm,
       "filter_tflops_tf32": flops / (tm["filter_ms"] * 1e-3) / 1e12 if tm["filter_ms"] else None,
       "device_qps": a.queries / (tm["total_ms"] * 1e-3) if tm["total_ms"] else None,
       "rows_gbs_equivalent": a.rows * a.dim * 4 * a.queries / e2e / 1e9}
os.environ["PQV_BATCH"] = "off"
nchk = min(a.check, a.queries)
t0 = time.perf_counter()
r1, d1, c1 = ds.l2_topk(queries[:nchk], a.k, a.flags)
single = (time.perf_counter() - t0) / nchk
del os.environ["PQV_BATCH"]
out["single_query_seconds"] = single
out["single_query_qps"] = 1.0 / single
out["speedup_vs_single_query_scans"] = single * a.queries / e2e
out["paths_identical_on"] = int(nchk) if (np.array_equal(rows[:nchk], r1) and np.array_equal(cnt[:nchk], c1) and
                                           np.array_equal(dist[:nchk].view(np.uint32), d1.view(np.uint32))) else -1
# the reference's CPU loop for the same batch: nq independent serial scans (exec.rs:257-277 / search.rs:112-141), oracle
# port on a bounded row sample, 1 thread per query as the reference runs it and all cores side by side as a charitable figure
import oracle as O  # noqa: E402  (cpu baseline only)
cores = os.cpu_count() or 1
samp = min(200_000, a.rows)
host = O.synth(samp, a.dim, 1234)
order = 1 if a.flags & P.PQV_SUM_SEQ else 0
t0 = time.perf_counter()
O.scan_topk_mt(host, queries[0], a.k, order, 1)
dt1 = time.perf_counter() - t0
per_query = dt1 * a.rows / samp
out["cpu_reference"] = {"kind": "port", "cores": 1,
                        "sample": f"one query over {samp} x {a.dim} rows, 1 thread: {dt1:.3f} s, scaled x{a.rows / samp:g} in rows",
                        "seconds_per_query": per_query, "qps": 1.0 / per_query,
                        "qps_if_one_query_per_core": cores / per_query, "host_cores": cores,
                        "gpu_speedup_vs_one_core": out["e2e_qps"] * per_query}
print(json.dumps(out))
