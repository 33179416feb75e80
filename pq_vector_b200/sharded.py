"""One process per GPU: row-range shards + ONE all-gather of per-rank heap-entrant candidates.

Each rank scans its own slice (pqv_l2_topk_candidates), the ranks exchange a few KB of candidate keys with a
single torch.distributed all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests), and every rank
replays the reference's BinaryHeap over the union (pqv_replay_candidates) -> identical, bit-exact results on
all ranks.  The global row id of local row r on this rank is pos_base + r (the same prefix sum the reference
computes for row groups in src/df_vector/access.rs:128-144)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .api import ivf_sample_rows, merge_batch_keys, replay_candidates

DEFAULT_CAP = 4096  # candidate keys per rank carried by the single all-gather (32 KiB)


class ShardedTopk:
    def __init__(self, scan_fn, pos_base: int, device: "torch.device | str" = "cpu", group=None,
                 cap: int = DEFAULT_CAP):
        """scan_fn(query, k, flags, pos_base) -> np.uint64 candidate keys of this rank's slice
        (Dataset.l2_topk_candidates bound to the rank's resident dataset)."""
        self.scan_fn = scan_fn
        self.pos_base = int(pos_base)
        self.device = torch.device(device)
        self.group = group
        self.cap = cap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._send = torch.zeros(cap + 1, dtype=torch.int64, device=self.device)
        self._recv = torch.zeros(self.world * (cap + 1), dtype=torch.int64, device=self.device)
        self.last_gather_bytes = 0
        self._p2p = None

    def enable_p2p(self, ctx, dataset):
        """Switch the per-query exchange from the NCCL all-gather to NVLink peer writes (pqv_peer.cuh): the ranks swap
        the CUDA IPC handles of their exchange buffers once (one all-gather of 64 bytes per rank); from then on a search
        is scan -> filter -> publish into every peer's buffer -> wait for the peers' flags, with no collective call."""
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        mine = ctx.peer_exchange_create(self.world, rank, self.cap)
        send = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        if self.world > 1:
            recv = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            handles = bytes(recv.cpu().numpy().tobytes())
        else:
            handles = mine
        ctx.peer_exchange_open(handles)
        if self.world > 1:
            dist.barrier(group=self.group)   # nobody publishes before every rank has mapped every buffer
        self._p2p = dataset

    def _exchange(self, keys: np.ndarray, cap: int) -> "tuple[np.ndarray, bool]":
        if cap != self._send.numel() - 1:
            self._send = torch.zeros(cap + 1, dtype=torch.int64, device=self.device)
            self._recv = torch.zeros(self.world * (cap + 1), dtype=torch.int64, device=self.device)
        n = int(keys.size)
        host = np.zeros(cap + 1, dtype=np.int64)
        host[0] = n
        m = min(n, cap)
        host[1:1 + m] = keys[:m].view(np.int64)
        self._send.copy_(torch.from_numpy(host), non_blocking=False)
        if self.world > 1:
            dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
            got = self._recv.cpu().numpy().reshape(self.world, cap + 1)
        else:
            got = self._send.cpu().numpy()[None, :]
        self.last_gather_bytes = got.nbytes
        counts = got[:, 0]
        if int(counts.max()) > cap:
            return counts, False
        union = np.concatenate([got[r, 1:1 + int(counts[r])] for r in range(got.shape[0])]).view(np.uint64)
        return union, True

    def search(self, query, k: int, flags: int):
        if self._p2p is not None:
            got = self._p2p.l2_topk_p2p(query, k, flags, self.pos_base)   # scan + exchange + replay in one native call
            self.last_gather_bytes = 0
            if got is not None:
                return got
            # a rank had more candidates than a slot holds -- every rank saw it: all take the collective path below
        keys = np.ascontiguousarray(self.scan_fn(query, k, flags, self.pos_base), dtype=np.uint64)
        if self.world == 1:  # nothing to exchange
            self.last_gather_bytes = 0
            return replay_candidates(keys, k, flags)
        out, ok = self._exchange(keys, self.cap)
        if not ok:  # some rank had more candidates than the default payload: one more round, sized to fit
            out, ok = self._exchange(keys, int(out.max()))
            assert ok
        return replay_candidates(out, k, flags)


class ShardedBatchTopk:
    """Batched queries over row-range shards (BASELINE config C5): every rank answers the whole batch over its own
    slice in one tensor-core pass (pqv_l2_topk_batch_keys: k + 1 exact keys per query), ONE all-gather moves
    world x nq x (k + 2) x 8 B, every rank merges (pqv_merge_batch_keys).  Queries whose answer hinges on the
    reference heap's layout (exact ties) are re-run through the single-query candidate exchange of ShardedTopk, so
    every query's result is bit-identical to its own reference loop over the whole table."""

    def __init__(self, batch_fn, scan_fn, pos_base: int, device: "torch.device | str" = "cpu", group=None, tie_fn=None):
        """batch_fn(queries, k, flags, pos_base) -> (keys [nq, k+1] u64, counts [nq] u32)  (Dataset.l2_topk_batch_keys)
        scan_fn: as ShardedTopk (Dataset.l2_topk_candidates): full-scan candidates of one query.
        tie_fn(q_index, query) -> candidate keys of a flagged query from what the batched pass left on the device
        (Dataset.l2_topk_batch_tie_candidates); without it the flagged queries use scan_fn."""
        self.batch_fn = batch_fn
        self.scan_fn = scan_fn
        self.tie_fn = tie_fn
        self.pos_base = int(pos_base)
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.single = ShardedTopk(scan_fn, pos_base, device, group)
        self.last_gather_bytes = 0
        self.last_replayed = 0
        self.last_phase_ms = {}
        self._p2p = None

    def enable_p2p(self, ctx, dataset, nq: int, k: int):
        """Answer batches with ONE native call per rank (pqv_l2_topk_batch_p2p): the key lists travel over NVLink peer memory
        instead of an all-gather, merge and tie replays stay in the library.  The exchange slots are sized for nq x (k + 2)
        words; a batch that does not fit (or that a slice declines) takes the collective path below on every rank."""
        self.single.cap = max(self.single.cap, nq * (k + 2))
        self.single.enable_p2p(ctx, dataset)
        self._p2p = dataset

    def _all_gather(self, arr: np.ndarray) -> np.ndarray:
        """[world, len(arr)] int64, same on every rank"""
        send = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self.device)
        if self.world == 1:
            return send.cpu().numpy()[None, :]
        recv = torch.empty(self.world * send.numel(), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        self.last_gather_bytes += recv.numel() * 8
        return recv.cpu().numpy().reshape(self.world, send.numel())

    def search(self, queries, k: int, flags: int):
        import time
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        self.last_gather_bytes = 0
        t0 = time.perf_counter()
        if self._p2p is not None:
            got = self._p2p.l2_topk_batch_p2p(queries, k, flags, self.pos_base)
            if got is not None:
                rows, dd, cnt, self.last_replayed = got
                self.last_phase_ms = {"native_call": (time.perf_counter() - t0) * 1e3}
                return rows, dd, cnt
        keys, counts = self.batch_fn(queries, k, flags, self.pos_base)
        t1 = time.perf_counter()
        # one payload per rank: [nq, k+1] keys followed by the nq counts (widened to 64 bit)
        got = self._all_gather(np.concatenate([keys.reshape(-1).view(np.int64), counts.astype(np.int64)]))
        t2 = time.perf_counter()
        all_keys = got[:, :nq * (k + 1)].view(np.uint64).reshape(self.world, nq, k + 1)
        all_counts = got[:, nq * (k + 1):].astype(np.uint32)
        rows, dd, cnt, need = merge_batch_keys(all_keys, all_counts, k, flags)
        t3 = time.perf_counter()
        # deterministic on identical gathered data -> every rank sees the same flagged queries, in the same order
        ties = np.nonzero(need)[0]
        self.last_replayed = int(ties.size)
        # where a call's wall time went on this rank (ms): the per-rank pass, the all-gather, the merge; "ties" is added below
        self.last_phase_ms = {"pass": (t1 - t0) * 1e3, "all_gather": (t2 - t1) * 1e3, "merge": (t3 - t2) * 1e3, "ties": 0.0}
        if ties.size == 0:
            return rows, dd, cnt
        # candidates of ALL flagged queries travel together: one all-gather of their lengths, one of the padded keys
        declined = bool((all_counts == 0xFFFFFFFF).any())   # a slice without a batched pass has nothing left on the device
        use_tie = self.tie_fn is not None and not declined
        cand = [np.ascontiguousarray(self.tie_fn(int(q), queries[q]) if use_tie else
                                     self.scan_fn(queries[q], k, flags, self.pos_base), dtype=np.uint64) for q in ties]
        lens = self._all_gather(np.array([c.size for c in cand], dtype=np.int64))        # [world, n_ties]
        cap = max(int(lens.max()), 1)
        padded = np.zeros((ties.size, cap), dtype=np.int64)
        for i, c in enumerate(cand):
            padded[i, :c.size] = c.view(np.int64)
        allc = self._all_gather(padded.reshape(-1)).reshape(self.world, ties.size, cap)
        for i, q in enumerate(ties):
            union = np.concatenate([allc[r, i, :int(lens[r, i])] for r in range(self.world)]).view(np.uint64)
            r_, d_ = replay_candidates(union, k, flags)
            cnt[q] = r_.size
            rows[q, :r_.size] = r_
            dd[q, :r_.size] = d_
        self.last_phase_ms["ties"] = (time.perf_counter() - t3) * 1e3
        return rows, dd, cnt


def index_to_bytes(centroids: np.ndarray, offsets: np.ndarray, ids: np.ndarray) -> bytes:
    """IvfIndex::to_bytes (src/ivf/index.rs:65-83): u32 dim, u32 C, f32[C*dim], then per cluster u32 len + u32[len]."""
    c, dim = centroids.shape
    parts = [np.array([dim, c], dtype="<u4").tobytes(), np.ascontiguousarray(centroids, dtype="<f4").tobytes()]
    ids = np.ascontiguousarray(ids, dtype="<u4")
    for j in range(c):
        lo, hi = int(offsets[j]), int(offsets[j + 1])
        parts.append(np.array([hi - lo], dtype="<u4").tobytes())
        parts.append(ids[lo:hi].tobytes())
    return b"".join(parts)


class ShardedIvfBuild:
    """build_ivf_index (src/ivf/index.rs:152-214) over a table whose rows are sharded across the ranks:
    the training sample (<= 100 k rows, drawn from the whole table by the same rule and stream as pqv_ivf_build) is
    gathered to rank 0, k-means runs there on one GPU, the centroids (C x dim f32, 3 MB at C = 1024, dim = 768) are
    broadcast, every rank assigns ITS rows with pqv_kmeans_assign, and the per-rank assignments are gathered so that each
    rank can lay the lists out in ascending global row id -- the result is byte-identical to the single-GPU build of
    the concatenated table.

    read_rows(local_ids) -> [m, dim] f32 rows of this rank's slice      (Dataset.read per run of ids)
    train(sample [S, dim], C, max_iters, seed) -> [C, dim] centroids     (Context.kmeans_train over a scratch dataset)
    assign(centroids) -> u32[n_local] assignment of this rank's rows     (Context.kmeans_assign on the resident slice)"""

    def __init__(self, read_rows, train, assign, n_local: int, pos_base: int, n_global: int, dim: int,
                 device: "torch.device | str" = "cpu", group=None):
        self.read_rows, self.train, self.assign = read_rows, train, assign
        self.n_local, self.pos_base, self.n_global, self.dim = int(n_local), int(pos_base), int(n_global), int(dim)
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def _gather_var(self, arr: np.ndarray, dtype) -> "list[np.ndarray]":
        """all ranks' 1-D arrays (different lengths), as a list in rank order"""
        arr = np.ascontiguousarray(arr, dtype=dtype)
        if self.world == 1:
            return [arr]
        n = torch.tensor([arr.size], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(x.item()) for x in sizes]
        cap = max(max(sizes), 1)
        send = torch.zeros(cap, dtype=torch.from_numpy(arr[:0]).dtype, device=self.device)
        send[:arr.size] = torch.from_numpy(arr).to(self.device)
        recv = torch.empty(self.world * cap, dtype=send.dtype, device=self.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        got = recv.cpu().numpy().reshape(self.world, cap)
        return [got[r, :sizes[r]].astype(dtype, copy=False) for r in range(self.world)]

    def build(self, n_clusters=None, max_iters: int = 20, seed: int = 42) -> bytes:
        sample_ids, C = ivf_sample_rows(self.n_global, n_clusters, seed)        # same on every rank
        mine = (sample_ids >= self.pos_base) & (sample_ids < self.pos_base + self.n_local)
        where = np.nonzero(mine)[0]                                             # positions inside the sample
        rows = self.read_rows(sample_ids[mine].astype(np.int64) - self.pos_base) if where.size else \
            np.empty((0, self.dim), np.float32)
        parts_pos = self._gather_var(where.astype(np.int64), np.int64)
        parts_rows = self._gather_var(np.ascontiguousarray(rows, np.float32).reshape(-1).view(np.int32), np.int32)
        centroids = np.empty((C, self.dim), dtype=np.float32)
        if self.rank == 0:
            sample = np.empty((sample_ids.size, self.dim), dtype=np.float32)
            for pos, flat in zip(parts_pos, parts_rows):
                if pos.size:
                    sample[pos] = flat.view(np.float32).reshape(-1, self.dim)
            centroids[:] = self.train(sample, C, max_iters, seed)
        if self.world > 1:
            t = torch.from_numpy(centroids).to(self.device)
            dist.broadcast(t, src=0, group=self.group)
            centroids = t.cpu().numpy()
        local = np.ascontiguousarray(self.assign(centroids), dtype=np.uint32)
        assign = np.concatenate(self._gather_var(local.view(np.int32), np.int32)).view(np.uint32)   # rank order = row order
        order = np.argsort(assign, kind="stable").astype(np.uint32)             # per cluster, ascending row id
        offsets = np.zeros(C + 1, dtype=np.uint64)
        np.cumsum(np.bincount(assign, minlength=C), out=offsets[1:])
        return index_to_bytes(centroids, offsets, order)


class ShardedArrayDistanceTopk:
    """The un-indexed `array_distance` arm over row-range shards (SURVEY row a10 at config C5's 8-GPU shape): every rank
    takes the exact f64 top-k of its own slice (Dataset.array_distance_topk), ONE all-gather moves world x (1 + 2k) x 8 B,
    and every rank merges by (f64 total order with NaN last, global row) -- the k smallest of the whole table are among
    the slices' k smallest, and the key is unique per row, so all ranks produce the same list."""

    def __init__(self, local_fn, pos_base: int, device: "torch.device | str" = "cpu", group=None):
        """local_fn(query_f64, k) -> (row_idx u32 local to the slice, distance f64), ascending."""
        self.local_fn = local_fn
        self.pos_base = int(pos_base)
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.last_gather_bytes = 0

    @staticmethod
    def _ordered(d: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(d, dtype=np.float64).view(np.uint64).copy()
        b[np.isnan(d)] = np.uint64(0x7FF8000000000000)
        neg = (b >> np.uint64(63)).astype(bool)
        return np.where(neg, ~b, b | np.uint64(0x8000000000000000))

    def search(self, query, k: int):
        rows, d = self.local_fn(np.ascontiguousarray(query, dtype=np.float64), k)
        n = int(rows.size)
        send = np.zeros(1 + 2 * k, dtype=np.int64)
        send[0] = n
        send[1:1 + n] = rows.astype(np.int64) + self.pos_base
        send[1 + k:1 + k + n] = np.ascontiguousarray(d, dtype=np.float64).view(np.int64)
        if self.world > 1:
            t_send = torch.from_numpy(send).to(self.device)
            t_recv = torch.empty(self.world * send.size, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(t_recv, t_send, group=self.group)
            got = t_recv.cpu().numpy().reshape(self.world, send.size)
            self.last_gather_bytes = got.nbytes
        else:
            got = send[None, :]
            self.last_gather_bytes = 0
        all_rows = np.concatenate([got[r, 1:1 + int(got[r, 0])] for r in range(got.shape[0])]).astype(np.uint32)
        all_d = np.concatenate([got[r, 1 + k:1 + k + int(got[r, 0])] for r in range(got.shape[0])]).view(np.float64)
        order = np.lexsort((all_rows, self._ordered(all_d)))[:k]
        return all_rows[order], all_d[order]


# ---- IVF search over row-range shards ----------------------------------------------------------
def shard_counts(offsets: np.ndarray, ids: np.ndarray, bounds) -> np.ndarray:
    """counts[c, s] = rows of inverted list c that fall into shard s = [bounds[s], bounds[s+1]) (lists are ascending,
    src/ivf/index.rs:202-206, so every list is cut by binary search)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    C, world = offsets.size - 1, len(bounds) - 1
    counts = np.zeros((C, world), dtype=np.int64)
    b = np.asarray(bounds, dtype=np.int64)
    for c in range(C):
        cut = np.searchsorted(ids[offsets[c]:offsets[c + 1]], b, side="left")
        counts[c] = np.diff(cut)
    return counts


def shard_index(offsets: np.ndarray, ids: np.ndarray, lo: int, hi: int):
    """The index restricted to rows [lo, hi): every list cut to the range, ids made local (minus lo).  Same cluster
    numbering, so the centroid ranking of a query is identical on every rank."""
    offsets = np.asarray(offsets, dtype=np.int64)
    keep = (ids >= lo) & (ids < hi)
    local_ids = (ids[keep].astype(np.int64) - lo).astype(np.uint32)
    cs = np.concatenate([[0], np.cumsum(keep.astype(np.int64))])
    per_list = cs[offsets[1:]] - cs[offsets[:-1]]
    local_offsets = np.concatenate([[0], np.cumsum(per_list)]).astype(np.uint64)
    return local_offsets, local_ids


class ShardedIvfSearch:
    """TopkBuilder::search (src/ivf/search.rs:83-142) with the table's rows sharded over ranks and the index replicated:
    rank s scans only the candidates inside its row range (IvfIndex.search_candidates over its slice and the index cut to
    the slice), ONE all-gather moves the heap-entrant keys, every rank replays the reference heap over the union.
    The global candidate sequence (index.rs:57-63) is, probed list by probed list, rank 0's rows, rank 1's rows, ...
    (ascending lists, contiguous slices), so a local candidate position translates to its global position with the
    per-list per-rank counts alone."""

    def __init__(self, cand_fn, counts: np.ndarray, rank: int, lo: int, device: "torch.device | str" = "cpu", group=None,
                 cap: int = DEFAULT_CAP):
        """cand_fn(query, k, nprobe, flags) -> (keys u64, local row ids u32, probed clusters in rank order)."""
        self.cand_fn = cand_fn
        self.counts = np.asarray(counts, dtype=np.int64)
        self.rank, self.lo = int(rank), int(lo)
        self.device = torch.device(device)
        self.group = group
        self.cap = cap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.last_gather_bytes = 0

    def translate(self, keys: np.ndarray, probe: np.ndarray) -> "tuple[np.ndarray, int]":
        """local candidate positions -> global candidate positions; returns (keys with global positions, total candidates)"""
        cnt = self.counts[np.asarray(probe, dtype=np.int64)]                 # [np, world]
        lens = cnt.sum(axis=1)
        base = np.concatenate([[0], np.cumsum(lens)[:-1]]) if lens.size else np.zeros(0, np.int64)
        within = cnt[:, :self.rank].sum(axis=1)
        lp = np.concatenate([[0], np.cumsum(cnt[:, self.rank])])
        pos = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        r = np.searchsorted(lp, pos, side="right") - 1
        gpos = base[r] + within[r] + (pos - lp[r])
        return (keys & np.uint64(0xFFFFFFFF00000000)) | gpos.astype(np.uint64), int(lens.sum())

    def _exchange(self, keys: np.ndarray, rows: np.ndarray, cap: int):
        n = int(keys.size)
        send = np.zeros(1 + 2 * cap, dtype=np.int64)
        send[0] = n
        m = min(n, cap)
        send[1:1 + m] = keys[:m].view(np.int64)
        send[1 + cap:1 + cap + m] = rows[:m]
        if self.world > 1:
            t_send = torch.from_numpy(send).to(self.device)
            t_recv = torch.empty(self.world * send.size, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(t_recv, t_send, group=self.group)
            got = t_recv.cpu().numpy().reshape(self.world, send.size)
        else:
            got = send[None, :]
        self.last_gather_bytes = got.nbytes if self.world > 1 else 0
        counts = got[:, 0]
        if int(counts.max()) > cap:
            return None, None, int(counts.max())
        k_all = np.concatenate([got[s, 1:1 + int(counts[s])] for s in range(got.shape[0])]).view(np.uint64)
        r_all = np.concatenate([got[s, 1 + cap:1 + cap + int(counts[s])] for s in range(got.shape[0])]).astype(np.uint32)
        return k_all, r_all, 0

    def search(self, query, k: int, nprobe: int, flags: int):
        keys, rows, probe = self.cand_fn(query, k, nprobe, flags)
        gkeys, n_total = self.translate(np.ascontiguousarray(keys, dtype=np.uint64), probe)
        grows = rows.astype(np.int64) + self.lo
        k_all, r_all, need = self._exchange(gkeys, grows, self.cap)
        if k_all is None:  # a rank had more entrants than the default payload: one more round, sized to fit
            k_all, r_all, need = self._exchange(gkeys, grows, need)
            assert k_all is not None
        if n_total == 0 or k_all.size == 0:
            return np.empty(0, np.uint32), np.empty(0, np.float32)
        row_of = np.zeros(n_total, dtype=np.uint32)          # candidate position -> row id, filled for the entrants only
        row_of[(k_all & np.uint64(0xFFFFFFFF)).astype(np.int64)] = r_all
        return replay_candidates(k_all, k, flags, row_ids=row_of)


class ShardedBatchIvfSearch:
    """Batched IVF searches over row-range shards (config C5 with the index registered): every rank answers the whole batch
    over its slice in one masked tensor-core pass (IvfIndex.search_batch_keys: k + 1 exact keys per query among the probed
    rows of the slice), ONE all-gather, pqv_merge_batch_keys on every rank; the queries it flags (exact ties, undecided
    slices) go through ShardedIvfSearch one by one.  Every query's result equals its own search over the whole table."""

    def __init__(self, batch_fn, single: ShardedIvfSearch, pos_base: int, device: "torch.device | str" = "cpu", group=None):
        """batch_fn(queries, k, nprobe, flags, pos_base) -> (keys [nq, k+1] u64, counts [nq] u32)"""
        self.batch_fn = batch_fn
        self.single = single
        self.pos_base = int(pos_base)
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.last_gather_bytes = 0
        self.last_replayed = 0

    def search(self, queries, k: int, nprobe: int, flags: int):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        keys, counts = self.batch_fn(queries, k, nprobe, flags, self.pos_base)
        send = torch.from_numpy(np.concatenate([keys.reshape(-1).view(np.int64), counts.astype(np.int64)])).to(self.device)
        if self.world > 1:
            recv = torch.empty(self.world * send.numel(), dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            got = recv.cpu().numpy().reshape(self.world, send.numel())
            self.last_gather_bytes = got.nbytes
        else:
            got = send.cpu().numpy()[None, :]
            self.last_gather_bytes = 0
        all_keys = got[:, :nq * (k + 1)].view(np.uint64).reshape(self.world, nq, k + 1)
        all_counts = got[:, nq * (k + 1):].astype(np.uint32)
        rows, dd, cnt, need = merge_batch_keys(all_keys, all_counts, k, flags)
        ties = np.nonzero(need)[0]        # identical on every rank: the merge is deterministic on identical gathered data
        self.last_replayed = int(ties.size)
        for q in ties:
            r_, d_ = self.single.search(queries[q], k, nprobe, flags)
            cnt[q] = r_.size
            rows[q, :r_.size] = r_
            dd[q, :r_.size] = d_
        return rows, dd, cnt
