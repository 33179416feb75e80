"""One k-means++ init (100 000 x 768 sample, 1024 clusters) -- the target of ncu captures of the sweep / pick kernels."""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P
ctx = P.Context([0]); ds = ctx.dataset(768, 100000); ds.fill_synthetic(100000, 1234)
ctx.kmeans_train(ds, 1024, max_iters=1, seed=42, sum_workers=16)
