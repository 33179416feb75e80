#!/usr/bin/env python
"""torchrun check + timing of the sharded SEARCH paths on real GPUs: rows split over the ranks, index replicated.

  * ShardedIvfSearch            (TopkBuilder::search, src/ivf/search.rs:83-142)  per-rank fused pipeline over the slice +
                                one all-gather + heap replay
  * ShardedArrayDistanceTopk    (the un-indexed array_distance arm)             per-rank f64 top-k + one all-gather
Rank 0 also holds the whole table and answers alone; results must be identical (row ids and distance bits).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/check_sharded_search.py"""
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
from pq_vector_b200.sharded import (ShardedArrayDistanceTopk, ShardedBatchIvfSearch, ShardedIvfBuild, ShardedIvfSearch,  # noqa: E402
                                    index_to_bytes,
                                    shard_counts, shard_index)

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000, help="rows per GPU")
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--clusters", type=int, default=1024)
ap.add_argument("--nprobe", type=int, default=32)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--queries", type=int, default=30)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = P.Context([local])
lo, n_glob = rank * a.rows, a.rows * world
ds = ctx.dataset(a.dim, a.rows)
ds.fill_synthetic(a.rows, 1234, stream_first_row=lo)
qd = ctx.dataset(a.dim, a.queries)
qd.fill_synthetic(a.queries, 7)
queries = qd.read(0, a.queries)
qd.drop()


def train(sample, c, max_iters, seed):
    sd = ctx.dataset_from(sample)
    out = ctx.kmeans_train(sd, c, max_iters, seed)
    sd.drop()
    return out


# the index of the whole table, built sharded (every rank ends up with the same blob)
blob = ShardedIvfBuild(ds.read_rows, train, lambda cent: ctx.kmeans_assign(ds, cent), a.rows, lo, n_glob, a.dim, dev).build(
    a.clusters, 10, 42)
dim, C = np.frombuffer(blob, "<u4", 2)
cent = np.frombuffer(blob, "<f4", int(dim) * int(C), 8).reshape(int(C), int(dim))
off, ids_parts, p = [0], [], 8 + int(dim) * int(C) * 4
for _ in range(int(C)):
    ln = int(np.frombuffer(blob, "<u4", 1, p)[0])
    ids_parts.append(np.frombuffer(blob, "<u4", ln, p + 4))
    off.append(off[-1] + ln)
    p += 4 + 4 * ln
offsets, ids = np.asarray(off, np.uint64), np.concatenate(ids_parts)
bounds = [s * a.rows for s in range(world + 1)]
l_off, l_ids = shard_index(offsets, ids, lo, lo + a.rows)
local_ix = ctx.ivf_from_bytes(index_to_bytes(cent, l_off, l_ids))
ivf = ShardedIvfSearch(lambda q, k, nprobe, flags: local_ix.search_candidates(ds, q, k, nprobe, flags),
                       shard_counts(offsets, ids, bounds), rank, lo, dev)
adist = ShardedArrayDistanceTopk(lambda q, k: ds.array_distance_topk(q, k), lo, dev)


def timed(fn):
    for q in queries[:3]:
        fn(q)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    res, t0 = [], time.perf_counter()
    for q in queries:
        res.append(fn(q))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return res, (time.perf_counter() - t0) / len(queries)


r_ivf, t_ivf = timed(lambda q: ivf.search(q, a.k, a.nprobe, P.PQV_SQRT))
# batched: 256 searches in one masked pass per rank + one all-gather
nqb, kb = 256, 10
qb = ctx.dataset(a.dim, nqb)
qb.fill_synthetic(nqb, 9)
bq = qb.read(0, nqb)
qb.drop()
sbatch = ShardedBatchIvfSearch(lambda qs, k, nprobe, flags, pb: local_ix.search_batch_keys(ds, qs, k, nprobe, flags, pb), ivf, lo, dev)
sbatch.search(bq, kb, a.nprobe, P.PQV_SQRT)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
b_rows, b_dist, b_cnt = sbatch.search(bq, kb, a.nprobe, P.PQV_SQRT)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t_batch = time.perf_counter() - t0
r_ad, t_ad = timed(lambda q: adist.search(q.astype(np.float64), a.k))
if rank == 0:
    whole = ctx.dataset(a.dim, n_glob)
    whole.fill_synthetic(n_glob, 1234)
    wix = ctx.ivf_from_bytes(blob)
    same_ivf = same_ad = True
    t0 = time.perf_counter()
    for q, (r, d) in zip(queries, r_ivf):
        wr, wd = wix.search(whole, q, a.k, a.nprobe, P.PQV_SQRT)
        same_ivf &= wr.tolist() == r.tolist() and wd.view(np.uint32).tolist() == d.view(np.uint32).tolist()
    t_whole = (time.perf_counter() - t0) / len(queries)
    for q, (r, d) in zip(queries, r_ad):
        wr, wd = whole.array_distance_topk(q.astype(np.float64), a.k)
        same_ad &= wr.tolist() == r.tolist() and wd.view(np.uint64).tolist() == d.view(np.uint64).tolist()
    w_rows, w_dist, w_cnt = wix.search_batch(whole, bq, kb, a.nprobe, P.PQV_SQRT)
    same_batch = bool(np.array_equal(w_cnt, b_cnt) and all(
        np.array_equal(w_rows[i, :w_cnt[i]], b_rows[i, :b_cnt[i]]) and
        np.array_equal(w_dist[i, :w_cnt[i]].view(np.uint32), b_dist[i, :b_cnt[i]].view(np.uint32)) for i in range(nqb)))
    print(json.dumps({"sharded_batch_ivf": {"queries": nqb, "k": kb, "seconds_per_batch": t_batch, "qps": nqb / t_batch,
                                            "replayed_queries": sbatch.last_replayed, "identical_to_single_gpu_batch": same_batch},
                      "config": f"{n_glob} x {a.dim} over {world} GPU(s), C={a.clusters}, nprobe={a.nprobe}, k={a.k}",
                      "sharded_ivf_search_ms": t_ivf * 1e3, "single_gpu_ivf_search_ms": t_whole * 1e3,
                      "sharded_array_distance_topk_ms": t_ad * 1e3, "ivf_identical_to_single_gpu": bool(same_ivf),
                      "array_distance_identical_to_single_gpu": bool(same_ad), "queries": len(queries)}))
if world > 1:
    dist.destroy_process_group()
