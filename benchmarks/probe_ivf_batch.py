#!/usr/bin/env python
"""Where a batched IVF search (config C3 shape, 1024 queries, top-100 with sqrt) spends its wall time: the masked
tensor-core pass, the tie queries replayed from a short prefix, whatever is left for the single-query pipeline."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P

rows, dim, C, nprobe = int(os.environ.get("ROWS", 10_000_000)), 768, 1024, 32
ctx = P.Context([0])
ds = ctx.dataset(dim, rows); ds.fill_synthetic(rows, 1234)
qd = ctx.dataset(dim, 1024); qd.fill_synthetic(1024, 7); queries = qd.read(0, 1024); qd.drop()
ix = ctx.ivf_build(ds, n_clusters=C, max_iters=20, seed=42)
out = {}
for k, flags in ((10, P.PQV_SQRT), (100, P.PQV_SQRT), (100, 0)):
    ix.search_batch(ds, queries, k, nprobe, flags)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r, d, c = ix.search_batch(ds, queries, k, nprobe, flags); ts.append(time.perf_counter() - t0)
    t = ctx.last_batch_timing()
    out[f"k{k}_flags{flags}"] = {"ms": min(ts) * 1e3, "device_ms": t["total_ms"], "tie_queries": t["tie_queries"], "tie_batched": t["tie_batched"]}
    same = 0
    for i in range(0, 1024, 64):
        rr, dd = ix.search(ds, queries[i], k, nprobe, flags)
        same += int(rr.tolist() == r[i, :c[i]].tolist() and dd.view(np.uint32).tolist() == d[i, :c[i]].view(np.uint32).tolist())
    out[f"k{k}_flags{flags}"]["identical_of_16"] = same
print(json.dumps(out))
