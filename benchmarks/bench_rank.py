#!/usr/bin/env python
"""Batched centroid ranking (SURVEY row a6 at config C5: 1024 queries x 1024 centroids x 768): pqv_centroid_rank with host
buffers in and out, batched launch vs the per-query loop."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

C, dim, nq, nprobe = 1024, 768, 1024, 32
rng = np.random.default_rng(0)
cent = rng.random((C, dim), dtype=np.float32)
qs = rng.random((nq, dim), dtype=np.float32)
ctx = P.Context([0])


def timed(reps=5):
    ctx.centroid_rank(cent, qs, nprobe)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = ctx.centroid_rank(cent, qs, nprobe)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


tb, ob = timed()
os.environ["PQV_RANK_BATCH_OFF"] = "1"
tl, ol = timed(2)
print(json.dumps({"config": f"{nq} queries x {C} centroids x {dim}, nprobe {nprobe}", "batched_ms": tb * 1e3, "per_query_loop_ms": tl * 1e3,
                  "speedup": tl / tb, "identical": bool(np.array_equal(ob, ol)), "queries_per_s": nq / tb}))
