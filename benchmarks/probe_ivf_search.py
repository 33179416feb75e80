#!/usr/bin/env python
"""Breakdown of one IVF search (config C3 shape): wall time per query vs the gather-scan kernel time."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P

rows, dim, C, nprobe, k = int(os.environ.get("ROWS", 10_000_000)), 768, 1024, 32, 100
ctx = P.Context([0])
ds = ctx.dataset(dim, rows); ds.fill_synthetic(rows, 1234)
qd = ctx.dataset(dim, 64); qd.fill_synthetic(64, 7); queries = qd.read(0, 64)
ix = ctx.ivf_build(ds, n_clusters=C, max_iters=3, seed=42)
print(json.dumps(ix.build_stats()))
for q in queries[:4]: ix.search(ds, q, k, nprobe)
wall, scan, post, cand, ent = [], [], [], [], []
for q in queries:
    t0 = time.perf_counter(); ix.search(ds, q, k, nprobe); wall.append(time.perf_counter() - t0)
    t = ctx.last_timing(); scan.append(t["scan_ms"]); post.append(t["post_ms"]); cand.append(t["scan_bytes"] // (dim * 4)); ent.append(t["entrants"])
m = lambda x: float(np.mean(x))
print(json.dumps({"wall_ms": m(wall) * 1e3, "scan_ms": m(scan), "post_ms": m(post), "cands": m(cand), "entrants": m(ent),
                  "scan_gbs": m(cand) * dim * 4 / (m(scan) * 1e-3) / 1e9}))
# candidate-rows only
t0 = time.perf_counter()
for q in queries: ix.candidate_rows(q, nprobe)
print(json.dumps({"candidate_rows_ms": (time.perf_counter() - t0) / len(queries) * 1e3}))
