"""CPU, world_size 2, gloo: the N>1 host logic (candidate exchange + replay) gives the single-process answer.
The per-rank 'scan' here is the oracle's distance sweep turned into candidate keys (every row of the slice is
a trivially valid candidate superset); on the GPU box the same class is driven by Dataset.l2_topk_candidates."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_cases(rank, world, ports, fn_name, cases, out_dir):
    """all cases of one test in ONE pair of processes (a spawn costs ~10 s of interpreter + torch start-up): every case
    gets its own rendezvous port, process group and output directory"""
    fn = globals()[fn_name]
    for i, case in enumerate(cases):
        d = os.path.join(out_dir, f"case{i}")
        os.makedirs(d, exist_ok=True)
        fn(rank, world, ports[i], *case, d)


def _free_ports(n):
    """n rendezvous ports the kernel reports as free right now (bind to port 0), all held open until every one is chosen
    so that they are distinct"""
    import socket
    socks = []
    for _ in range(n):
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", 0))
        socks.append(s)
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def _spawn_cases(fn, cases, tmp_path, port_base=None):
    ports = _free_ports(len(cases))
    mp.spawn(_run_cases, args=(2, ports, fn.__name__, cases, str(tmp_path)), nprocs=2, join=True)
    return [(case, tmp_path / f"case{i}") for i, case in enumerate(cases)]


def _worker(rank, world, port, n, dim, k, flags, cap, seed, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as O
    from pq_vector_b200.sharded import ShardedTopk
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 4, (n, dim)).astype(np.float32) if seed % 2 else rng.random((n, dim), dtype=np.float32)
    q = np.zeros(dim, np.float32) if seed % 2 else rng.random(dim, dtype=np.float32)
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)

    def scan(query, k_, flags_, pos_base):
        d = O.distances(data[lo:hi], query, 1 if flags_ & 1 else 0)
        return (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (np.arange(lo, hi, dtype=np.uint64))

    st = ShardedTopk(scan, lo, "cpu", cap=cap)
    rows, dd = st.search(q, k, flags)
    er, ed = O.topk_rerank(q, data, None, k, 1 if flags & 1 else 0, bool(flags & 2))
    ok = rows.tolist() == er.tolist() and dd.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process(tmp_path):
    cases = [(1000, 8, 10, 2, 4096, 2), (3000, 4, 100, 1, 64, 3), (5, 3, 10, 2, 16, 4), (2000, 6, 50, 3, 4096, 5)]
    for (n,dim,k,flags,cap,seed), d in _spawn_cases(_worker, cases, tmp_path, 29500):
        for r in range(2):
            assert np.load(d / f"ok{r}.npy")[0] == 1, (n, dim, k, flags, cap, seed)


def _batch_worker(rank, world, port, n, dim, nq, k, flags, seed, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as O
    from pq_vector_b200.sharded import ShardedBatchTopk
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    # odd seeds: small-integer grid -> many exact ties (the replay route); even seeds: continuous data (the merge route)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if seed % 2 else rng.random((n, dim), dtype=np.float32)
    queries = (rng.integers(0, 3, (nq, dim)).astype(np.float32) if seed % 2 else rng.random((nq, dim), dtype=np.float32))
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    order = 1 if flags & 1 else 0

    def keys_of(query):
        d = O.distances(data[lo:hi], query, order)
        return (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (np.arange(lo, hi, dtype=np.uint64))

    def batch(qs, k_, flags_, pos_base):  # the oracle standing in for pqv_l2_topk_batch_keys
        keys = np.full((qs.shape[0], k_ + 1), np.iinfo(np.uint64).max, dtype=np.uint64)
        cnt = np.zeros(qs.shape[0], dtype=np.uint32)
        for i, q in enumerate(qs):
            kk = np.sort(keys_of(q))[:k_ + 1]
            keys[i, :kk.size] = kk
            cnt[i] = kk.size
        if seed == 6:
            cnt[1] = 0xFFFFFFFF  # a slice that could not decide query 1 -> replay route
        return keys, cnt

    # even seeds also exercise the tie_fn route (candidates taken from the batched pass' leftovers)
    tie_fn = (lambda qi, q: keys_of(q)) if seed in (3, 7) else None
    sb = ShardedBatchTopk(batch, lambda q, k_, f_, pb: keys_of(q), lo, "cpu", tie_fn=tie_fn)
    rows, dd, cnt = sb.search(queries, k, flags)
    ok = True
    for i in range(nq):
        er, ed = O.topk_rerank(queries[i], data, None, k, order, bool(flags & 2))
        ok &= cnt[i] == er.size and rows[i, :cnt[i]].tolist() == er.tolist()
        ok &= dd[i, :cnt[i]].view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok), sb.last_replayed]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batched_exchange_matches_single_process(tmp_path):
    cases = [(2000, 8, 6, 10, 2, 2), (1500, 4, 5, 20, 1, 3), (7, 3, 4, 10, 2, 4), (1200, 6, 4, 5, 3, 6), (900, 5, 3, 3, 0, 7)]
    for (n,dim,nq,k,flags,seed), d in _spawn_cases(_batch_worker, cases, tmp_path, 31500):
        got = [np.load(d / f"ok{r}.npy") for r in range(2)]
        assert all(g[0] == 1 for g in got), (n, dim, nq, k, flags, seed)
        assert got[0][1] == got[1][1]                 # both ranks took the same replay decisions
        if seed % 2:
            assert got[0][1] > 0                      # the tie-heavy cases really exercised the replay route


def _ivf_worker(rank, world, port, n, dim, C, seed, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as O
    from pq_vector_b200.sharded import ShardedIvfBuild
    from pq_vector_b200.api import ivf_sample_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    data = rng.random((n, dim), dtype=np.float32)
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)

    def train(sample, c, max_iters, seed_):     # stand-in for Context.kmeans_train: any deterministic function of the sample
        return np.ascontiguousarray(sample[:c] * np.float32(0.5) + np.float32(0.25))

    sb = ShardedIvfBuild(lambda ids: data[lo:hi][ids], train, lambda cent: O.assign(data[lo:hi], cent, workers=2),
                         hi - lo, lo, n, dim, "cpu")
    blob = sb.build(C, 5, seed)
    # the single-process answer: same sample rule, same train, oracle assignment + lists over the whole table
    sample_ids, c = ivf_sample_rows(n, C, seed)
    cent = train(data[sample_ids], c, 5, seed)
    a = O.assign(data, cent, workers=2)
    offsets, ids = O.inverted_lists(a, c)
    want = O.index_to_bytes(dim, cent, offsets, ids)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(blob == want)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ivf_build_matches_single_process(tmp_path):
    cases = [(4000, 6, 9, 1), (333, 4, None, 2), (50, 3, 50, 3)]
    for (n,dim,C,seed), d in _spawn_cases(_ivf_worker, cases, tmp_path, 33500):
        for r in range(2):
            assert np.load(d / f"ok{r}.npy")[0] == 1, (n, dim, C, seed)


def _adist_worker(rank, world, port, n, dim, k, seed, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as O
    from pq_vector_b200.sharded import ShardedArrayDistanceTopk
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    # odd seeds: small-integer grid -> many equal f64 distances across the two slices (ties by global row)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if seed % 2 else rng.random((n, dim), dtype=np.float32)
    if seed == 8:
        data[3, 0] = np.inf
        data[n - 2, 0] = np.inf
    q = rng.random(dim) if seed % 2 == 0 else np.zeros(dim)
    if seed == 8:
        q[0] = np.inf            # rows 3 and n-2: NaN distance (sorted last); every other row: +inf
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)

    def local(query, k_):        # the oracle standing in for Dataset.array_distance_topk on this rank's slice
        return O.array_distance_topk(data[lo:hi], query, k_)

    st = ShardedArrayDistanceTopk(local, lo, "cpu")
    rows, dd = st.search(q, k)
    er, ed = O.array_distance_topk(data, q, k)
    ok = rows.tolist() == er.tolist() and dd.view(np.uint64).tolist() == ed.view(np.uint64).tolist()
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_array_distance_topk(tmp_path):
    cases = [(1000, 8, 10, 2), (3000, 4, 100, 3), (5, 3, 10, 4), (600, 6, 600, 8)]
    for (n,dim,k,seed), d in _spawn_cases(_adist_worker, cases, tmp_path, 37500):
        for r in range(2):
            assert np.load(d / f"ok{r}.npy")[0] == 1, (n, dim, k, seed)


def _ivf_search_worker(rank, world, port, n, dim, C, k, nprobe, flags, seed, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as O
    from pq_vector_b200.sharded import ShardedIvfSearch, shard_counts, shard_index
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if seed % 2 else rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy()
    offsets, ids = O.inverted_lists(O.assign(data, cent), C)
    q = data[5].copy() if seed % 2 else rng.random(dim, dtype=np.float32)
    per = (n + world - 1) // world
    bounds = [min(n, s * per) for s in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    l_off, l_ids = shard_index(offsets, ids, lo, hi)
    order = 1 if flags & 1 else 0

    def cand(query, k_, nprobe_, flags_):       # the oracle standing in for IvfIndex.search_candidates on this rank's slice
        probe = O.find_closest_centroids(query, cent, nprobe_)
        rows = O.candidate_rows(query, cent, l_off, l_ids, nprobe_)
        d = O.distances(data[lo:hi][rows], query, order) if rows.size else np.empty(0, np.float32)
        keys = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(rows.size, dtype=np.uint64)
        return keys, rows, probe

    st = ShardedIvfSearch(cand, shard_counts(offsets, ids, bounds), rank, lo, "cpu", cap=64 if seed == 5 else 4096)
    rows, dd = st.search(q, k, nprobe, flags)
    er, ed = O.topk_rerank_gather(q, data, O.candidate_rows(q, cent, offsets, ids, nprobe), k, order, bool(flags & 2))
    ok = rows.tolist() == er.tolist() and dd.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ivf_search(tmp_path):
    cases = [(2000, 8, 16, 10, 4, 2, 2), (3000, 4, 9, 100, 3, 1, 3), (900, 6, 30, 50, 30, 3, 5), (40, 3, 7, 10, 2, 2, 6)]
    for (n,dim,C,k,nprobe,flags,seed), d in _spawn_cases(_ivf_search_worker, cases, tmp_path, 39500):
        for r in range(2):
            assert np.load(d / f"ok{r}.npy")[0] == 1, (n, dim, C, k, nprobe, flags, seed)


def _ivf_batch_worker(rank, world, port, n, dim, C, nq, k, nprobe, flags, seed, out_dir):
    sys.path.insert(0, ROOT)
    import oracle as O
    from pq_vector_b200.sharded import ShardedBatchIvfSearch, ShardedIvfSearch, shard_counts, shard_index
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 3, (n, dim)).astype(np.float32) if seed % 2 else rng.random((n, dim), dtype=np.float32)
    cent = data[rng.choice(n, C, replace=False)].copy()
    offsets, ids = O.inverted_lists(O.assign(data, cent), C)
    queries = data[rng.integers(0, n, nq)].copy() if seed % 2 else rng.random((nq, dim), dtype=np.float32)
    per = (n + world - 1) // world
    bounds = [min(n, s * per) for s in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    l_off, l_ids = shard_index(offsets, ids, lo, hi)
    order = 1 if flags & 1 else 0

    def cand(query, k_, nprobe_, flags_):
        probe = O.find_closest_centroids(query, cent, nprobe_)
        rows = O.candidate_rows(query, cent, l_off, l_ids, nprobe_)
        d = O.distances(data[lo:hi][rows], query, order) if rows.size else np.empty(0, np.float32)
        return (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(rows.size, dtype=np.uint64), rows, probe

    def batch(qs, k_, nprobe_, flags_, pos_base):   # the oracle standing in for IvfIndex.search_batch_keys
        keys = np.full((qs.shape[0], k_ + 1), np.iinfo(np.uint64).max, dtype=np.uint64)
        cnt = np.zeros(qs.shape[0], dtype=np.uint32)
        for i, q in enumerate(qs):
            rows = O.candidate_rows(q, cent, l_off, l_ids, nprobe_)
            d = O.distances(data[lo:hi][rows], q, order) if rows.size else np.empty(0, np.float32)
            kk = np.sort((d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (rows.astype(np.uint64) + np.uint64(pos_base)))[:k_ + 1]
            keys[i, :kk.size] = kk
            cnt[i] = kk.size
        if seed == 6:
            cnt[2] = 0xFFFFFFFF          # a slice that could not decide query 2 -> single-query route
        return keys, cnt

    single = ShardedIvfSearch(cand, shard_counts(offsets, ids, bounds), rank, lo, "cpu")
    sb = ShardedBatchIvfSearch(batch, single, lo, "cpu")
    rows, dd, cnt = sb.search(queries, k, nprobe, flags)
    ok = True
    for i, q in enumerate(queries):
        er, ed = O.topk_rerank_gather(q, data, O.candidate_rows(q, cent, offsets, ids, nprobe), k, order, bool(flags & 2))
        ok &= cnt[i] == er.size and rows[i, :cnt[i]].tolist() == er.tolist()
        ok &= dd[i, :cnt[i]].view(np.uint32).tolist() == ed.view(np.uint32).tolist()
    replayed_ok = sb.last_replayed > 0 if seed % 2 or seed == 6 else True      # grid data / undecided slice: the replay route ran
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok and replayed_ok)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batched_ivf_search(tmp_path):
    cases = [(2000, 8, 16, 12, 10, 4, 2, 2), (1500, 4, 9, 9, 20, 3, 1, 3), (900, 6, 30, 8, 50, 30, 3, 6)]
    for (n,dim,C,nq,k,nprobe,flags,seed), d in _spawn_cases(_ivf_batch_worker, cases, tmp_path, 35500):
        for r in range(2):
            assert np.load(d / f"ok{r}.npy")[0] == 1, (n, dim, C, nq, k, nprobe, flags, seed)