import sys, os, time, json
sys.path.insert(0, '/root/repo')
import pq_vector_b200 as P
ctx = P.Context([0])
ds = ctx.dataset(768, 10_000_000); ds.fill_synthetic(10_000_000, 1234)
import torch; torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); ix = ctx.ivf_build(ds, n_clusters=1024, max_iters=20, seed=42); t = time.perf_counter() - t0
    print("build", i, round(t * 1e3, 1), "ms", json.dumps(ix.build_stats()), flush=True)
    ix.drop()
