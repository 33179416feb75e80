#!/usr/bin/env python
"""Config C3 of BASELINE.json: 10M x 768 synthetic f32, IVF build with n_clusters=1024 (k-means assign kernel)
+ nprobe=32 search, 1 B200.  Prints one JSON object (build seconds, final-assign GB/s and f32 op rate, search
QPS, recall@k against the brute-force scan).  Secondary bench; bench.py carries the headline metric."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--clusters", type=int, default=1024)
ap.add_argument("--max-iters", type=int, default=20)
ap.add_argument("--nprobe", type=int, default=32)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--queries", type=int, default=50)
ap.add_argument("--builds", type=int, default=3, help="builds to run; the first pays the one-time scratch allocations")
a = ap.parse_args()

ctx = P.Context([0])
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234)
qd = ctx.dataset(a.dim, a.queries)
qd.fill_synthetic(a.queries, 7)
queries = qd.read(0, a.queries)

builds = []
ix = None
for _ in range(max(a.builds, 1)):
    if ix is not None:
        ix.drop()
    t0 = time.perf_counter()
    ix = ctx.ivf_build(ds, n_clusters=a.clusters, max_iters=a.max_iters, seed=42)
    builds.append({"seconds": time.perf_counter() - t0, **ix.build_stats()})
best = min(builds, key=lambda b: b["seconds"])
build_s = best["seconds"]
st = {k: v for k, v in best.items() if k != "seconds"}
fa_s = st["final_assign_ms"] * 1e-3
out = {
    "config": f"{a.rows} x {a.dim} f32, IVF C={a.clusters}, max_iters={a.max_iters}, seed=42; search nprobe={a.nprobe} k={a.k}",
    "build_seconds": build_s, "build_breakdown_ms": st, "first_build_seconds": builds[0]["seconds"],
    "all_builds_seconds": [b["seconds"] for b in builds],
    "final_assign": {"seconds_incl_d2h_and_list_build": fa_s, "rows_gbs": a.rows * a.dim * 4 / fa_s / 1e9,
                     "f32_ops_per_s": 3.0 * a.rows * a.clusters * a.dim / fa_s,
                     "note": "3*N*C*dim non-fusable f32 ops (sub, mul, add), exact reference order"},
}
t0 = time.perf_counter()
blob = ix.to_bytes()   # IvfIndex::to_bytes (index.rs:65-83): includes the one-time fetch of the device-built lists
out["to_bytes_seconds"] = time.perf_counter() - t0
out["blob_bytes"] = len(blob)
del blob
# search
for q in queries[:3]:
    ix.search(ds, q, a.k, a.nprobe)
lat, cands, recall = [], [], []
for q in queries:
    t0 = time.perf_counter()
    r, d = ix.search(ds, q, a.k, a.nprobe)
    lat.append(time.perf_counter() - t0)
    cands.append(ctx.last_timing()["scan_bytes"] // (a.dim * 4))
for q in queries[:10]:
    r, _ = ix.search(ds, q, a.k, a.nprobe)
    br, _ = ds.l2_topk(q, a.k)
    recall.append(len(set(r.tolist()) & set(br.tolist())) / max(len(br), 1))
t = ctx.last_timing()
out["search"] = {"qps_e2e": 1.0 / float(np.mean(lat)), "ms_per_query_e2e": float(np.mean(lat)) * 1e3,
                 "mean_candidates": float(np.mean(cands)), "recall_at_k_vs_bruteforce": float(np.mean(recall)),
                 "gather_gbs_e2e": float(np.mean(cands)) * (a.dim * 4 + 4) / float(np.mean(lat)) / 1e9}
# VectorTopKExec over the same resident table + index (pqv_vector_topk_indexed): candidates in row order, sequential-order
# sums, with and without a filter bitmap (here: id >= N/2) -- the reference's operator re-reads the rows from Parquet
rng_mask = np.arange(a.rows) >= a.rows // 2
for name, mask in (("vector_topk_exec", None), ("vector_topk_exec_filtered", rng_mask)):
    for q in queries[:3]:
        ix.vector_topk(ds, q, a.k, a.nprobe, P.PQV_SUM_SEQ, None, mask)
    lat, scored, kern = [], [], []
    for q in queries:
        t0 = time.perf_counter()
        r, d, total, ns = ix.vector_topk(ds, q, a.k, a.nprobe, P.PQV_SUM_SEQ, None, mask)
        lat.append(time.perf_counter() - t0)
        scored.append(ns)
        kern.append(ctx.last_timing()["scan_ms"])
    same = None
    if mask is None:  # same candidate set as TopkBuilder's search; only summation order / visiting order differ
        r0, _ = ix.search(ds, queries[-1], a.k, a.nprobe, P.PQV_SUM_SEQ)
        same = sorted(r0.tolist()) == sorted(r.tolist())
    out[name] = {"qps_e2e": 1.0 / float(np.mean(lat)), "ms_per_query_e2e": float(np.mean(lat)) * 1e3,
                 "mean_rows_scored": float(np.mean(scored)), "gather_scan_ms": float(np.mean(kern)),
                 "gather_scan_gbs": float(np.mean(scored)) * a.dim * 4 / (float(np.mean(kern)) * 1e-3) / 1e9,
                 "same_rows_as_search": same,
                 "mask_bytes_h2d_per_query": 0 if mask is None else (a.rows + 7) // 8}
# batched IVF search: 1024 independent searches in one masked tensor-core pass (pqv_ivf_search_batch)
nqb = 1024
qb = ctx.dataset(a.dim, nqb)
qb.fill_synthetic(nqb, 7)
bq = qb.read(0, nqb)
qb.drop()
out["search_batch"] = {}
for kk in (10, a.k):
    ix.search_batch(ds, bq, kk, a.nprobe)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        br, bd, bc = ix.search_batch(ds, bq, kk, a.nprobe)
        ts.append(time.perf_counter() - t0)
    bt = ctx.last_batch_timing()
    same = True
    for i in range(0, nqb, 64):
        r1, d1 = ix.search(ds, bq[i], kk, a.nprobe)
        same &= bool(bc[i] == r1.size and np.array_equal(br[i, :bc[i]], r1) and
                     np.array_equal(bd[i, :bc[i]].view(np.uint32), d1.view(np.uint32)))
    tb = float(np.median(ts))
    out["search_batch"][f"k{kk}"] = {"queries": nqb, "seconds_per_batch": tb, "qps_e2e": nqb / tb, "timing": bt,
                                     "identical_to_single_searches_on": 16 if same else -1,
                                     "speedup_vs_single_searches": tb and (nqb * out["search"]["ms_per_query_e2e"] * 1e-3) / tb}
# the reference's CPU loops (oracle port) on bounded samples of the same work, timed on this box's host cores
import oracle as O  # noqa: E402  (cpu baseline only)
cores = os.cpu_count() or 1
cent = ix.centroids()
samp = min(20_000, a.rows)
host = O.synth(samp, a.dim, 1234)
t0 = time.perf_counter()
O.assign(host, cent, workers=cores)              # index.rs:193-201 runs on available_parallelism() threads
dt_assign = time.perf_counter() - t0
samp_c = min(100_000, a.rows)
host_c = O.synth(samp_c, a.dim, 1234)
t0 = time.perf_counter()
O.topk_rerank(queries[0], host_c, None, a.k, 0, True)   # search.rs:112-141 is serial
dt_rerank = time.perf_counter() - t0
mean_c = out["search"]["mean_candidates"]
out["cpu_reference"] = {
    "kind": "port", "cores_assign": cores, "cores_rerank": 1,
    "final_assign_seconds_extrapolated": dt_assign * a.rows / samp,
    "final_assign_sample": f"{samp} rows x {a.clusters} centroids x {a.dim} on {cores} threads: {dt_assign:.3f} s, scaled x{a.rows / samp:g}",
    "search_rerank_seconds_extrapolated": dt_rerank * mean_c / samp_c,
    "search_sample": f"{samp_c} candidate rows x {a.dim}, 1 thread: {dt_rerank:.3f} s, scaled to {mean_c:.0f} candidates "
                     "(distance loop only; the reference also re-reads index and rows from Parquet per query)",
    "gpu_final_assign_speedup": dt_assign * a.rows / samp / fa_s,
    "gpu_search_speedup": dt_rerank * mean_c / samp_c / (out["search"]["ms_per_query_e2e"] * 1e-3)}
print(json.dumps(out))
