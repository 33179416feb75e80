import sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import pq_vector_b200 as P
ctx = P.Context([0])
n, dim = 10_000_000, 768
ds = ctx.dataset(dim, n); ds.fill_synthetic(n, 1234)
qd = ctx.dataset(dim, 1); qd.fill_synthetic(1, 7); q = qd.read(0, 1)[0]
out = {}
for name, flags in (("unroll4", P.PQV_SQRT), ("seq", P.PQV_SUM_SEQ)):
    for k in (10, 100):
        ds.bench_scan(q, k, flags, 3)
        ms = ds.bench_scan(q, k, flags, 10)
        t = ctx.last_timing()
        out[f"{name}_k{k}"] = {"ms": t["scan_ms"], "gbs": n * dim * 4 / t["scan_ms"] / 1e6, "grid": t["grid"]}
print(json.dumps(out))
