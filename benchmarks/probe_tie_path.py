#!/usr/bin/env python
"""Smallest input that drives the batched tie path (pqv_tie.cuh) -- for compute-sanitizer runs:
    compute-sanitizer --tool racecheck python benchmarks/probe_tie_path.py
Duplicate rows give bit-equal distances inside the top-k, so every query's order hinges on the reference heap and the
tie kernels (prefix_dist_matrix_kernel, prefix_entrants_kernel) run; the answers are checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker)
import pq_vector_b200 as P  # noqa: E402

rng = np.random.default_rng(4)
base = rng.random((3000, 64), dtype=np.float32)
data = np.concatenate([base, base[:1500], base[:700]])
queries = rng.random((12, 64), dtype=np.float32)
ctx = P.Context([0])
ds = ctx.dataset_from(data)
for flags, order, sq in ((P.PQV_SQRT, 0, True), (P.PQV_SUM_SEQ, 1, False)):
    rows, dist, cnt = ds.l2_topk(queries, 50, flags)
    t = ctx.last_batch_timing()
    assert t["tie_queries"] >= 1 and t["tie_batched"] == t["tie_queries"], t
    for i, q in enumerate(queries):
        er, ed = O.topk_rerank(q, data, None, 50, order, sq)
        assert rows[i, :cnt[i]].tolist() == er.tolist()
        assert dist[i, :cnt[i]].view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    print("flags", flags, t["tie_queries"], "tie queries resolved together, bit-exact vs oracle")
ds.drop()
ctx.close()
