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

from .api import merge_batch_keys, replay_candidates

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

    def __init__(self, batch_fn, scan_fn, pos_base: int, device: "torch.device | str" = "cpu", group=None):
        """batch_fn(queries, k, flags, pos_base) -> (keys [nq, k+1] u64, counts [nq] u32)  (Dataset.l2_topk_batch_keys)
        scan_fn: as ShardedTopk (Dataset.l2_topk_candidates), used for the tie queries."""
        self.batch_fn = batch_fn
        self.pos_base = int(pos_base)
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.single = ShardedTopk(scan_fn, pos_base, device, group)
        self.last_gather_bytes = 0
        self.last_replayed = 0

    def search(self, queries, k: int, flags: int):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        keys, counts = self.batch_fn(queries, k, flags, self.pos_base)
        # one payload per rank: [nq, k+1] keys followed by the nq counts (widened to 64 bit)
        payload = np.concatenate([keys.reshape(-1).view(np.int64), counts.astype(np.int64)])
        send = torch.from_numpy(payload).to(self.device)
        if self.world > 1:
            recv = torch.empty(self.world * payload.size, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            got = recv.cpu().numpy().reshape(self.world, payload.size)
        else:
            got = send.cpu().numpy()[None, :]
        self.last_gather_bytes = got.nbytes
        all_keys = got[:, :nq * (k + 1)].view(np.uint64).reshape(self.world, nq, k + 1)
        all_counts = got[:, nq * (k + 1):].astype(np.uint32)
        rows, dd, cnt, need = merge_batch_keys(all_keys, all_counts, k, flags)
        # deterministic on identical gathered data -> every rank enters the same collective replays, in the same order
        self.last_replayed = int(need.sum())
        for q in np.nonzero(need)[0]:
            r, d = self.single.search(queries[q], k, flags)
            cnt[q] = r.size
            rows[q, :r.size] = r
            dd[q, :r.size] = d
        return rows, dd, cnt
