#!/usr/bin/env python
"""The gathered sequential-order scan behind pqv_vector_topk_indexed (VectorTopKExec arithmetic, config C3 shape): wall time
per call and the scan kernel's time / bandwidth; run with PQV_SEQ_GATHER_VARIANT=n to compare launch shapes."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P

rows, dim, C, nprobe, k = int(os.environ.get("ROWS", 10_000_000)), 768, 1024, 32, 100
ctx = P.Context([0])
ds = ctx.dataset(dim, rows); ds.fill_synthetic(rows, 1234)
qd = ctx.dataset(dim, 64); qd.fill_synthetic(64, 7); queries = qd.read(0, 64)
ix = ctx.ivf_build(ds, n_clusters=C, max_iters=3, seed=42)
for q in queries[:4]: ix.vector_topk(ds, q, k, nprobe, P.PQV_SUM_SEQ, None, None)
wall, scan, cand = [], [], []
for q in queries:
    t0 = time.perf_counter(); ix.vector_topk(ds, q, k, nprobe, P.PQV_SUM_SEQ, None, None); wall.append(time.perf_counter() - t0)
    t = ctx.last_timing(); scan.append(t["scan_ms"]); cand.append(t["scan_bytes"] // (dim * 4))
m = lambda x: float(np.mean(x))
print(json.dumps({"variant": os.environ.get("PQV_SEQ_GATHER_VARIANT", "0"), "wall_ms": m(wall) * 1e3, "scan_ms": m(scan), "cands": m(cand),
                  "scan_gbs": m(cand) * dim * 4 / (m(scan) * 1e-3) / 1e9}))
