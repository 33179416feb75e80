"""One IVF build over 10 M x 768 (C = 1024, 20 Lloyd iterations) -- the target of ncu launch lists of the Lloyd phase and the
final assignment (the k-means++ phase in front of them is 2046 launches: use --launch-skip)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P
rows = int(os.environ.get("ROWS", 10_000_000))
ctx = P.Context([0]); ds = ctx.dataset(768, rows); ds.fill_synthetic(rows, 1234)
ix = ctx.ivf_build(ds, n_clusters=1024, max_iters=20, seed=42)
print(ix.build_stats())
