#!/usr/bin/env python
"""BASELINE config C5 shape, one process per GPU (run under torchrun): rows sharded over the ranks (6.25 M x 768 per GPU by
default = 50 M rows on 8 GPUs), 1024 queries per call, k = 10, VectorTopKExec semantics (PQV_SUM_SEQ).  Every rank answers
the batch over its slice (pqv_l2_topk_batch_keys), ONE all-gather, host merge, tie queries through the candidate exchange.
Times whole batches end to end (host queries in, host results out), max over ranks; rank 0 prints one JSON object and
checks a few queries against their single-query sharded search.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      benchmarks/bench_batch_sharded.py"""
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
from pq_vector_b200.sharded import ShardedBatchTopk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=6_250_000, help="rows per GPU")
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--flags", type=int, default=P.PQV_SUM_SEQ)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", type=int, default=4)
a = ap.parse_args()

rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


pos_base = rank * a.rows
ctx = P.Context([local])
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234, stream_first_row=pos_base)
qd = ctx.dataset(a.dim, a.queries)
qd.fill_synthetic(a.queries, 7)
queries = qd.read(0, a.queries)
qd.drop()
sb = ShardedBatchTopk(lambda q, k_, f_, pb: ds.l2_topk_batch_keys(q, k_, f_, pb),
                      lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), pos_base, dev,
                      tie_fn=lambda qi, q: ds.l2_topk_batch_tie_candidates(qi, q))
for _ in range(2):
    sb.search(queries, a.k, a.flags)  # warm-up: allocations, row-norm cache, NCCL channels
ts = []
for _ in range(a.reps):
    barrier()
    t0 = time.perf_counter()
    rows, dd, cnt = sb.search(queries, a.k, a.flags)
    barrier()
    ts.append(max_over_ranks(time.perf_counter() - t0))
tm = ctx.last_batch_timing()
e2e = float(np.median(ts))
ok = True
for q in range(min(a.check, a.queries)):
    r, d = sb.single.search(queries[q], a.k, a.flags)
    ok &= rows[q, :cnt[q]].tolist() == r.tolist() and dd[q, :cnt[q]].view(np.uint32).tolist() == d.view(np.uint32).tolist()
if rank == 0:
    n_glob = a.rows * world
    print(json.dumps({
        "config": f"{n_glob} x {a.dim} f32 over {world} GPU(s) ({a.rows} rows each), {a.queries} queries per call, k={a.k}, flags={a.flags}",
        "n_gpus": world, "e2e_seconds_per_batch": e2e, "e2e_qps": a.queries / e2e,
        "tf32_tflops_aggregate": 2.0 * n_glob * a.queries * a.dim / e2e / 1e12,
        "gather_bytes_per_batch": sb.last_gather_bytes, "replayed_queries": sb.last_replayed,
        "rank0_batch_timing": tm, "matches_single_query_sharded_search_on": int(min(a.check, a.queries)) if ok else -1}))
if world > 1:
    dist.destroy_process_group()
