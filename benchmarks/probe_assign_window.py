#!/usr/bin/env python
"""Probe of the tcgen05 assignment filter's ambiguity counts: same sweep repeated, two centroid tables alternated, both
operand kinds -- the counts must depend on the inputs only (no stale scratch), and sit near the window model of
DESIGN.md section 4.4."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

n, dim, C = int(os.environ.get("ROWS", 1_000_000)), 768, 1024
ctx = P.Context([0])
ds = ctx.dataset(dim, n)
ds.fill_synthetic(n, 1234)
host = ds.read(0, 100_000)
rng = np.random.default_rng(1)
lab = rng.integers(0, C, host.shape[0])
cent_mean = np.stack([host[lab == c].mean(0) for c in range(C)]).astype(np.float32)   # |c - mu| ~ 0.8
cent_rows = host[:C].copy()                                                           # |c - mu| ~ 8
cent_mixed = cent_mean.copy()                                                         # empty clusters at the origin + tiny clusters
cent_mixed[::97] = 0.0
cent_mixed[5::61] = host[2000:2000 + len(cent_mixed[5::61])]
ix = ctx.ivf_build(ds, n_clusters=C, max_iters=3, seed=42)
cent_ivf = ix.centroids()
print("ivf build", ix.build_stats(), flush=True)
out = []
for kind in ("f16", "tf32"):
    if kind == "tf32":
        os.environ["PQV_TC_KIND"] = "tf32"
    else:
        os.environ.pop("PQV_TC_KIND", None)
    for name, cent in (("mean", cent_mean), ("rows", cent_rows), ("mixed", cent_mixed), ("ivf3", cent_ivf), ("mean", cent_mean)):
        t = ctx.bench_assign(ds, cent, iters=1)
        out.append({"kind": kind, "table": name, "amb": t["ambiguous_rows"], "ovf": t["overflow_rows"],
                    "filter_ms": round(t["filter_ms"], 3), "recheck_ms": round(t["recheck_ms"], 3), "k": t["kind"]})
        print(out[-1], flush=True)
