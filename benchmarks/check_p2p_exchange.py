#!/usr/bin/env python
"""torchrun check of the NVLink peer-memory candidate exchange (pqv_peer.cuh) against the NCCL all-gather path: same
queries through both, results must be identical on every rank; prints the per-query latencies of both.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/check_p2p_exchange.py"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402
from pq_vector_b200.sharded import ShardedTopk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000, help="rows per GPU")
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--queries", type=int, default=40)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = P.Context([local])
pos_base = rank * a.rows
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234, stream_first_row=pos_base)
qd = ctx.dataset(a.dim, a.queries)
qd.fill_synthetic(a.queries, 7)
queries = qd.read(0, a.queries)
qd.drop()
scan = lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb)  # noqa: E731
nccl = ShardedTopk(scan, pos_base, dev)
p2p = ShardedTopk(scan, pos_base, dev)
p2p.enable_p2p(ctx, ds)


def run(st):
    out, lat = [], []
    for i, q in enumerate(queries):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out.append(st.search(q, a.k, P.PQV_SQRT))
        lat.append(time.perf_counter() - t0)
    return out, float(np.median(lat[3:]))


r_nccl, t_nccl = run(nccl)
r_p2p, t_p2p = run(p2p)
same = all(x[0].tolist() == y[0].tolist() and x[1].view(np.uint32).tolist() == y[1].view(np.uint32).tolist()
           for x, y in zip(r_nccl, r_p2p))
# a slot that is too small: every rank must take the collective path together and still agree
tiny = ShardedTopk(scan, pos_base, dev, cap=64)
tiny.enable_p2p(ctx, ds)
r_tiny, _ = run(tiny)
same_tiny = all(x[0].tolist() == y[0].tolist() for x, y in zip(r_nccl, r_tiny))
# grid data: thousands of bit-equal distances across the ranks (the threshold that drops later ranks' entrants is strict),
# plus a slice holding a NaN row (every rank must fall back together and agree with the collective path)
same_grid = True
for trial, (gn, gd, with_nan) in enumerate([(150_000, 16, False), (40_000, 8, True)]):
    g = np.random.default_rng(1000 + rank + 17 * trial).integers(0, 3, (gn, gd)).astype(np.float32)
    if with_nan and rank == world - 1:
        g[5, 1] = np.nan
    gds = ctx.dataset_from(g)
    gscan = lambda q, k_, f_, pb, _d=gds: _d.l2_topk_candidates(q, k_, f_, pb)  # noqa: E731
    g_nccl = ShardedTopk(gscan, rank * gn, dev)
    g_p2p = ShardedTopk(gscan, rank * gn, dev)
    g_p2p.enable_p2p(ctx, gds)
    gq = np.random.default_rng(5 + trial).integers(0, 3, (12, gd)).astype(np.float32)
    for k_ in (1, 10, 100):
        for q in gq:
            x, y = g_nccl.search(q, k_, P.PQV_SQRT), g_p2p.search(q, k_, P.PQV_SQRT)
            xn, yn = np.nan_to_num(x[1], nan=-1.0), np.nan_to_num(y[1], nan=-1.0)
            same_grid &= x[0].tolist() == y[0].tolist() and xn.view(np.uint32).tolist() == yn.view(np.uint32).tolist()
    gds.drop()
flag = torch.tensor([float(same and same_tiny and same_grid)], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"config": f"{a.rows * world} x {a.dim} over {world} GPU(s), k={a.k}, {a.queries} queries",
                      "nccl_ms_per_query": t_nccl * 1e3, "p2p_ms_per_query": t_p2p * 1e3,
                      "identical_results_on_all_ranks": bool(flag.item() == 1.0)}))
if world > 1:
    dist.destroy_process_group()
