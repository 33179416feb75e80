#!/usr/bin/env python
"""Coalescing front door (pqv_l2_topk_coalesced): T concurrent callers each issue single-query top-k calls against one
resident table -- the way concurrent TopkBuilder::search / VectorTopKExec plans reach the FFI -- vs the
same calls issued one after the other.  Every call is a plain single-query call for its caller."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--threads", type=int, default=256)
ap.add_argument("--per-thread", type=int, default=8)
ap.add_argument("--flags", type=int, default=P.PQV_SQRT)
ap.add_argument("--ivf", action="store_true", help="also: concurrent TopkBuilder-style IVF searches (pqv_ivf_search_coalesced)")
ap.add_argument("--clusters", type=int, default=1024)
ap.add_argument("--nprobe", type=int, default=32)
a = ap.parse_args()

ctx = P.Context([0])
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234)
nq = a.threads * a.per_thread
qd = ctx.dataset(a.dim, nq)
qd.fill_synthetic(nq, 7)
queries = qd.read(0, nq)
qd.drop()
ds.l2_topk(queries[:64], a.k, a.flags)      # warm-up: allocations + row-norm cache
ds.l2_topk(queries[0], a.k, a.flags)

n_serial = 16
t0 = time.perf_counter()
serial = [ds.l2_topk(queries[i], a.k, a.flags) for i in range(n_serial)]
t_serial = (time.perf_counter() - t0) / n_serial

res = [None] * nq
gate = threading.Barrier(a.threads + 1)


def work(t):
    gate.wait()
    for j in range(a.per_thread):
        i = t * a.per_thread + j
        res[i] = ds.l2_topk_coalesced(queries[i], a.k, a.flags)


th = [threading.Thread(target=work, args=(t,)) for t in range(a.threads)]
for t in th:
    t.start()
s0 = ctx.coalesce_stats()
gate.wait()
t0 = time.perf_counter()
for t in th:
    t.join()
wall = time.perf_counter() - t0
s1 = ctx.coalesce_stats()
same = all(np.array_equal(res[i][0], serial[i][0]) and np.array_equal(res[i][1].view(np.uint32), serial[i][1].view(np.uint32))
           for i in range(n_serial))
out_ivf = None
if a.ivf:
    ix = ctx.ivf_build(ds, n_clusters=a.clusters, max_iters=20, seed=42)
    ix.search_batch(ds, queries[:64], a.k, a.nprobe, a.flags)
    t0 = time.perf_counter()
    ser = [ix.search(ds, queries[i], a.k, a.nprobe, a.flags) for i in range(64)]
    t_ser = (time.perf_counter() - t0) / 64
    res2 = [None] * nq
    gate2 = threading.Barrier(a.threads + 1)

    def work2(t):
        gate2.wait()
        for j in range(a.per_thread):
            i = t * a.per_thread + j
            res2[i] = ix.search_coalesced(ds, queries[i], a.k, a.nprobe, a.flags)

    th2 = [threading.Thread(target=work2, args=(t,)) for t in range(a.threads)]
    for t in th2:
        t.start()
    b0 = ctx.coalesce_stats()
    gate2.wait()
    t0 = time.perf_counter()
    for t in th2:
        t.join()
    wall2 = time.perf_counter() - t0
    b1 = ctx.coalesce_stats()
    same2 = all(np.array_equal(res2[i][0], ser[i][0]) and np.array_equal(res2[i][1].view(np.uint32), ser[i][1].view(np.uint32))
                for i in range(64))
    out_ivf = {"clusters": a.clusters, "nprobe": a.nprobe, "coalesced_qps": nq / wall2, "serial_single_search_qps": 1.0 / t_ser,
               "speedup": (nq / wall2) * t_ser, "batches": b1["batches"] - b0["batches"],
               "identical_to_serial_on": 64 if same2 else -1}
print(json.dumps({
    "ivf": out_ivf,
    "config": f"{a.rows} x {a.dim} f32 resident, {a.threads} threads x {a.per_thread} single-query calls, k={a.k}, flags={a.flags}",
    "coalesced_qps": nq / wall, "wall_seconds": wall, "serial_single_query_qps": 1.0 / t_serial,
    "speedup": (nq / wall) * t_serial, "batches": s1["batches"] - s0["batches"], "queries": s1["queries"] - s0["queries"],
    "max_batch": s1["max_batch"], "identical_to_serial_on": n_serial if same else -1}))
