#!/usr/bin/env python
"""torchrun check of pqv_l2_topk_batch_p2p (one native call per rank: tensor-core pass + NVLink exchange of the key lists +
merge + tie replays) against the collective path (all-gather + pqv_merge_batch_keys) and against single-query searches:
results must be identical on every rank.  Uniform data (no ties) and grid data (ties in most queries).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/check_p2p_batch.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402
from pq_vector_b200.sharded import ShardedBatchTopk, ShardedTopk  # noqa: E402

rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = P.Context([local])
ok, times = True, {}
for name, n, dim, nq, k, flags, grid in [("uniform", 400_000, 768, 256, 10, P.PQV_SUM_SEQ, False), ("uniform-k100", 200_000, 128, 96, 100, P.PQV_SQRT, False),
                                         ("grid", 60_000, 64, 48, 10, P.PQV_SQRT, True)]:
    rng = np.random.default_rng(100 + rank)
    data = (rng.integers(0, 3, (n, dim)) if grid else rng.random((n, dim))).astype(np.float32)
    ds = ctx.dataset_from(data)
    qrng = np.random.default_rng(7)
    queries = (qrng.integers(0, 3, (nq, dim)) if grid else qrng.random((nq, dim))).astype(np.float32)
    pos_base = rank * n
    mk = lambda: ShardedBatchTopk(lambda q, k_, f_, pb: ds.l2_topk_batch_keys(q, k_, f_, pb),  # noqa: E731
                                  lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), pos_base, dev,
                                  tie_fn=lambda qi, q: ds.l2_topk_batch_tie_candidates(qi, q))
    coll, p2p = mk(), mk()
    p2p.enable_p2p(ctx, ds, nq, k)
    r0, d0, c0 = coll.search(queries, k, flags)
    r1, d1, c1 = p2p.search(queries, k, flags)
    native = "native_call" in p2p.last_phase_ms
    same = native and np.array_equal(c0, c1)
    for q in range(nq):
        same &= r0[q, :c0[q]].tolist() == r1[q, :c1[q]].tolist() and d0[q, :c0[q]].view(np.uint32).tolist() == d1[q, :c1[q]].view(np.uint32).tolist()
    single = ShardedTopk(lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), pos_base, dev)
    for q in range(0, nq, max(1, nq // 6)):
        rr, dd = single.search(queries[q], k, flags)
        same &= rr.tolist() == r1[q, :c1[q]].tolist() and dd.view(np.uint32).tolist() == d1[q, :c1[q]].view(np.uint32).tolist()
    ts = {}
    for label, s in (("collective", coll), ("p2p", p2p)):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            s.search(queries, k, flags)
        ts[label] = (time.perf_counter() - t0) / 3 * 1e3
    times[name] = {"ms": ts, "replayed": int(p2p.last_replayed), "native": bool(native)}
    ok &= bool(same)
    ds.drop()
flag = torch.tensor([float(ok)], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "identical_on_all_ranks": bool(flag.item() == 1.0), "cases": times}))
if world > 1:
    dist.destroy_process_group()
