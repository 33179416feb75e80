import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
import pq_vector_b200 as P
ctx = P.Context([0]); ds = ctx.dataset(768, 2_000_000); ds.fill_synthetic(2_000_000, 1234)
ix = ctx.ivf_build(ds, n_clusters=1024, max_iters=3, seed=42)
c = ix.centroids()
n2 = (c.astype(np.float64)**2).sum(1)
mu = c.mean(0)
bn = np.sqrt(((c - mu).astype(np.float64)**2).sum(1))
print("centroid |c|^2: min %.3f med %.3f max %.3f; zero centroids: %d" % (n2.min(), np.median(n2), n2.max(), int((n2 < 1).sum())))
print("|c-mu|: min %.3f med %.3f p99 %.3f max %.3f" % (bn.min(), np.median(bn), np.quantile(bn, .99), bn.max()))
st = ix.build_stats(); print(st)
# spread of s_j for some rows
x = ds.read(0, 2000).astype(np.float64)
s = n2[None, :] - 2 * (x @ (c - mu).T.astype(np.float64))
ss = np.sort(s, axis=1)
print("per-row std of s_j: %.3f; gap 1-2 mean %.4f, P(gap<0.04)=%.3f; gap 1-6 mean %.4f" % (s.std(1).mean(), (ss[:,1]-ss[:,0]).mean(), ((ss[:,1]-ss[:,0])<0.04).mean(), (ss[:,5]-ss[:,0]).mean()))
print("bn quantiles", np.quantile(bn, [0, .1, .5, .9, .99, 1]))
