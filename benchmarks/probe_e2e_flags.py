import sys, json, time
sys.path.insert(0, '/root/repo')
import numpy as np
import pq_vector_b200 as P
ctx = P.Context([0])
n, dim = 10_000_000, 768
ds = ctx.dataset(dim, n); ds.fill_synthetic(n, 1234)
qd = ctx.dataset(dim, 40); qd.fill_synthetic(40, 7); qs = qd.read(0, 40)
out = {}
for name, flags in (("sqrt", P.PQV_SQRT), ("seq", P.PQV_SUM_SEQ), ("seq_sqrt", P.PQV_SUM_SEQ | P.PQV_SQRT)):
    for k in (10, 100):
        for i in range(3): ds.l2_topk(qs[i], k, flags)
        ts, tm = [], []
        for i in range(3, 33):
            t0 = time.perf_counter(); ds.l2_topk(qs[i], k, flags); ts.append(time.perf_counter() - t0)
            tm.append(ctx.last_timing())
        out[f"{name}_k{k}"] = {"e2e_ms": float(np.median(ts)) * 1e3, "scan_ms": float(np.median([t["scan_ms"] for t in tm])),
                               "post_ms": float(np.median([t["post_ms"] for t in tm])), "entrants": float(np.mean([t["entrants"] for t in tm]))}
print(json.dumps(out, indent=1))
