#!/usr/bin/env python
"""torchrun check + timing of the sharded IVF build (ShardedIvfBuild): rows split over the ranks, training on rank 0, final
assignment on every rank's slice.  Rank 0 also holds the whole table and builds it alone; the two blobs must be identical.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/check_sharded_ivf.py"""
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
from pq_vector_b200.sharded import ShardedIvfBuild  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000, help="rows per GPU")
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--clusters", type=int, default=1024)
ap.add_argument("--max-iters", type=int, default=20)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = P.Context([local])
pos_base, n_glob = rank * a.rows, a.rows * world
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234, stream_first_row=pos_base)


def train(sample, c, max_iters, seed):
    sd = ctx.dataset_from(sample)
    out = ctx.kmeans_train(sd, c, max_iters, seed)
    sd.drop()
    return out


sb = ShardedIvfBuild(ds.read_rows, train, lambda cent: ctx.kmeans_assign(ds, cent), a.rows, pos_base, n_glob, a.dim, dev)
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
blob = sb.build(a.clusters, a.max_iters, 42)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    whole = ctx.dataset(a.dim, n_glob)
    whole.fill_synthetic(n_glob, 1234)
    t0 = time.perf_counter()
    ix = ctx.ivf_build(whole, n_clusters=a.clusters, max_iters=a.max_iters, seed=42)
    mono = ix.to_bytes()
    t1 = time.perf_counter() - t0
    print(json.dumps({"config": f"{n_glob} x {a.dim} over {world} GPU(s), C={a.clusters}", "sharded_build_seconds": dt,
                      "single_gpu_build_plus_to_bytes_seconds": t1, "blob_bytes": len(blob), "identical_blobs": blob == mono}))
if world > 1:
    dist.destroy_process_group()
