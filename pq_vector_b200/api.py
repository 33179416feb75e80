"""Thin numpy-facing wrappers over the C ABI.  Every distance is computed by libpqv.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_lib = N.lib


class PqvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pqv error {code}: {msg}")
        self.code = code
        self.message = msg


def _check(rc: int):
    if rc != N.PQV_OK:
        raise PqvError(rc, (_lib.pqv_last_error() or b"").decode())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class Context:
    """Owns devices, streams and scratch (pqv_ctx)."""

    def __init__(self, devices=None):
        self._h = N.ctxp()
        if devices is None:
            _check(_lib.pqv_init(C.byref(self._h), None, 0))
        else:
            arr = (C.c_int * len(devices))(*devices)
            _check(_lib.pqv_init(C.byref(self._h), arr, len(devices)))

    def close(self):
        if self._h:
            _lib.pqv_destroy(self._h)
            self._h = N.ctxp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return _lib.pqv_device_count(self._h)

    # ---- datasets
    def dataset(self, dim: int, n_rows_hint: int = 0) -> "Dataset":
        h = C.c_uint64()
        _check(_lib.pqv_dataset_create(self._h, dim, n_rows_hint, C.byref(h)))
        return Dataset(self, h.value, dim)

    def dataset_from(self, rows) -> "Dataset":
        rows = _f32(rows)
        ds = self.dataset(rows.shape[1], rows.shape[0])
        ds.append(rows)
        return ds

    # ---- k-means pieces
    def kmeans_assign(self, rows_or_dataset, centroids, n=None, want_sizes=False):
        centroids = _f32(centroids)
        c, dim = centroids.shape
        if isinstance(rows_or_dataset, Dataset):
            handle, rows = rows_or_dataset.handle, None
            n = rows_or_dataset.rows if n is None else n
        else:
            rows = _f32(rows_or_dataset)
            handle, n = 0, rows.shape[0]
        out = np.empty(n, dtype=np.uint32)
        sizes = np.empty(c, dtype=np.uint64) if want_sizes else None
        _check(_lib.pqv_kmeans_assign(self._h, handle, _ptr(rows, C.c_float), n, dim, _ptr(centroids, C.c_float), c,
                                      _ptr(out, C.c_uint32), _ptr(sizes, C.c_uint64)))
        return (out, sizes) if want_sizes else out

    def min_dist_update(self, rows_or_dataset, row_sel, centroid, min_dist=None):
        """init (min_dist None) -> returns new array; else updates min_dist in place (index.rs:344-370)."""
        centroid = _f32(centroid)
        dim = centroid.size
        if isinstance(rows_or_dataset, Dataset):
            handle, rows, n_rows = rows_or_dataset.handle, None, rows_or_dataset.rows
        else:
            rows = _f32(rows_or_dataset)
            handle, n_rows = 0, rows.shape[0]
        sel = None if row_sel is None else np.ascontiguousarray(row_sel, dtype=np.uint64)
        n_sel = n_rows if sel is None else sel.size
        init = min_dist is None
        if init:
            min_dist = np.empty(n_sel, dtype=np.float32)
        assert min_dist.dtype == np.float32 and min_dist.flags.c_contiguous and min_dist.size == n_sel
        _check(_lib.pqv_min_dist_update(self._h, handle, _ptr(rows, C.c_float), _ptr(sel, C.c_uint64), n_sel, dim,
                                        _ptr(centroid, C.c_float), int(init), _ptr(min_dist, C.c_float)))
        return min_dist

    def centroid_rank(self, centroids, queries, nprobe):
        centroids = _f32(centroids)
        queries = np.atleast_2d(_f32(queries))
        c, dim = centroids.shape
        if queries.shape[1] != dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {dim}, got {queries.shape[1]}")
        nq = queries.shape[0]
        out = np.empty((nq, min(max(nprobe, 1), c)), dtype=np.uint32)
        eff = C.c_uint32()
        _check(_lib.pqv_centroid_rank(self._h, _ptr(centroids, C.c_float), c, dim, _ptr(queries, C.c_float), nq,
                                      nprobe, _ptr(out, C.c_uint32), C.byref(eff)))
        return out[:, :eff.value]

    # ---- IVF index
    def ivf_build(self, dataset: "Dataset", n_clusters=None, max_iters: int = 20, seed: int = 42,
                  sum_workers: int = 0) -> "IvfIndex":
        """build_ivf_index (src/ivf/index.rs:152-214); defaults as IndexBuilder (src/ivf/parquet.rs:32-40)."""
        if n_clusters is not None and n_clusters <= 0:
            raise PqvError(N.PQV_EINVAL, "n_clusters must be > 0")
        h = C.c_uint64()
        _check(_lib.pqv_ivf_build(self._h, dataset.handle, n_clusters or 0, max_iters, seed, sum_workers, C.byref(h)))
        return IvfIndex(self, h.value)

    def kmeans_train(self, dataset: "Dataset", n_clusters: int, max_iters: int = 20, seed: int = 42,
                     sum_workers: int = 0) -> np.ndarray:
        """k_means (src/ivf/index.rs:323-457) over all rows of `dataset`; returns the C x dim centroids."""
        out = np.empty((n_clusters, dataset.dim), dtype=np.float32)
        it = C.c_uint32()
        _check(_lib.pqv_kmeans_train(self._h, dataset.handle, n_clusters, max_iters, seed, sum_workers,
                                     _ptr(out, C.c_float), C.byref(it)))
        self.last_train_iters = it.value
        return out

    def peer_exchange_create(self, world: int, rank: int, cap_keys: int = 4096) -> bytes:
        """Allocate this rank's NVLink exchange buffer; returns its 64-byte CUDA IPC handle (pqv_peer_exchange_create)."""
        h = (C.c_uint8 * 64)()
        _check(_lib.pqv_peer_exchange_create(self._h, world, rank, cap_keys, h))
        return bytes(h)

    def peer_exchange_open(self, handles: bytes):
        """Map the peers' buffers: `handles` = the world ranks' 64-byte handles concatenated in rank order."""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        _check(_lib.pqv_peer_exchange_open(self._h, buf))

    def ivf_from_bytes(self, blob: bytes) -> "IvfIndex":
        buf = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob if blob else b"\0")
        h = C.c_uint64()
        _check(_lib.pqv_ivf_from_bytes(self._h, buf, len(blob), C.byref(h)))
        return IvfIndex(self, h.value)

    def topk_stream(self, query, k, flags=N.PQV_SUM_SEQ) -> "TopkStream":
        return TopkStream(self, query, k, flags)

    def coalesce_config(self, max_batch: int = 1024, window_us: int = 0):
        _check(_lib.pqv_coalesce_config(self._h, max_batch, window_us))

    def coalesce_stats(self) -> dict:
        q, b, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(_lib.pqv_coalesce_stats(self._h, C.byref(q), C.byref(b), C.byref(m)))
        return {"queries": q.value, "batches": b.value, "max_batch": m.value}

    def last_batch_timing(self) -> dict:
        """What the batched pass of the last l2_topk_batch / pqv_l2_topk call did (queries == 0: not used)."""
        t = N.PqvBatchTiming()
        _check(_lib.pqv_last_batch_timing(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in N.PqvBatchTiming._fields_ if f != "reserved"}

    def last_assign_timing(self) -> dict:
        """What the last kmeans_assign / bench_assign did: path (1 = tcgen05 filter), ambiguous/overflow rows, ms."""
        t = N.PqvAssignTiming()
        _check(_lib.pqv_last_assign_timing(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in N.PqvAssignTiming._fields_ if f != "reserved"}

    def bench_assign(self, dataset: "Dataset", centroids, iters: int = 3, n=None, want_assign=False):
        """Device-resident assignment sweeps (inputs and outputs stay in HBM); mean CUDA-event times."""
        centroids = _f32(centroids)
        c, dim = centroids.shape
        if dim != dataset.dim:
            raise PqvError(N.PQV_EINVAL, f"dimension mismatch: dataset has {dataset.dim}, centroids have {dim}")
        n = dataset.rows if n is None else n
        out = np.empty(n, dtype=np.uint32) if want_assign else None
        t = N.PqvAssignTiming()
        _check(_lib.pqv_bench_assign(self._h, dataset.handle, n, _ptr(centroids, C.c_float), c, iters, C.byref(t),
                                     _ptr(out, C.c_uint32)))
        d = {f: getattr(t, f) for f, _ in N.PqvAssignTiming._fields_ if f != "reserved"}
        return (d, out) if want_assign else d

    def last_timing(self) -> dict:
        t = N.PqvTiming()
        _check(_lib.pqv_last_timing(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in N.PqvTiming._fields_ if f != "reserved"}


class Dataset:
    """HBM-resident dense N x dim f32 embedding block (pqv_dataset_*)."""

    def __init__(self, ctx: Context, handle: int, dim: int):
        self.ctx, self.handle, self.dim = ctx, handle, dim

    def append(self, rows):
        rows = _f32(rows)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, "Embedding data length must be a multiple of dimension")
        _check(_lib.pqv_dataset_append(self.ctx._h, self.handle, _ptr(rows, C.c_float), rows.shape[0]))

    def fill_synthetic(self, n_rows: int, seed: int, stream_first_row: int = 0):
        _check(_lib.pqv_dataset_fill_synthetic(self.ctx._h, self.handle, n_rows, seed, stream_first_row))

    @property
    def rows(self) -> int:
        n = C.c_uint64()
        _check(_lib.pqv_dataset_rows(self.ctx._h, self.handle, C.byref(n), None))
        return n.value

    def read(self, first_row: int, n_rows: int) -> np.ndarray:
        out = np.empty((n_rows, self.dim), dtype=np.float32)
        _check(_lib.pqv_dataset_read(self.ctx._h, self.handle, first_row, n_rows, _ptr(out, C.c_float)))
        return out

    def read_rows(self, row_ids) -> np.ndarray:
        ids = np.ascontiguousarray(row_ids, dtype=np.uint32).reshape(-1)
        out = np.empty((ids.size, self.dim), dtype=np.float32)
        _check(_lib.pqv_dataset_read_rows(self.ctx._h, self.handle, _ptr(ids, C.c_uint32), ids.size, _ptr(out, C.c_float)))
        return out

    def drop(self):
        if self.handle:
            _check(_lib.pqv_dataset_drop(self.ctx._h, self.handle))
            self.handle = 0

    def l2_topk(self, queries, k: int, flags: int = N.PQV_SQRT):
        """Brute-force top-k (src/ivf/search.rs:112-141 with every row a candidate).
        Returns (row_idx [nq,k] u32, dist [nq,k] f32, count [nq]); 1-D query -> 1-D results trimmed to count."""
        q = _f32(queries)
        single = q.ndim == 1
        q2 = np.atleast_2d(q)
        if q2.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q2.shape[1]}")
        nq = q2.shape[0]
        kk = max(k, 1)
        rows = np.zeros((nq, kk), dtype=np.uint32)
        dist = np.zeros((nq, kk), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        _check(_lib.pqv_l2_topk(self.ctx._h, self.handle, _ptr(q2, C.c_float), nq, k, flags, _ptr(rows, C.c_uint32),
                                _ptr(dist, C.c_float), _ptr(cnt, C.c_uint32)))
        if single:
            return rows[0, :cnt[0]].copy(), dist[0, :cnt[0]].copy()
        return rows, dist, cnt

    def l2_topk_coalesced(self, query, k: int, flags: int = N.PQV_SQRT):
        """Single-query top-k through the coalescing front door (pqv_l2_topk_coalesced): safe to call from many threads
        (ctypes drops the GIL); calls that arrive while a pass is running are answered by one batched pass."""
        q = _f32(query)
        if q.ndim != 1 or q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        kk = max(k, 1)
        rows = np.zeros(kk, dtype=np.uint32)
        dist = np.zeros(kk, dtype=np.float32)
        cnt = C.c_uint32()
        _check(_lib.pqv_l2_topk_coalesced(self.ctx._h, self.handle, _ptr(q, C.c_float), k, flags, _ptr(rows, C.c_uint32),
                                          _ptr(dist, C.c_float), C.byref(cnt)))
        return rows[:cnt.value].copy(), dist[:cnt.value].copy()

    def array_distance(self, query, metric: int = N.PQV_METRIC_L2) -> np.ndarray:
        """The DataFusion built-in `array_distance(column, literal)` as a Float64 column (pqv_array_distance)."""
        q = np.ascontiguousarray(query, dtype=np.float64).ravel()
        out = np.empty(self.rows, dtype=np.float64)
        _check(_lib.pqv_array_distance(self.ctx._h, self.handle, _ptr(q, C.c_double), q.size, metric, _ptr(out, C.c_double)))
        return out

    def array_distance_topk(self, query, k: int, metric: int = N.PQV_METRIC_L2, row_mask=None):
        """`[WHERE ..] ORDER BY array_distance(column, literal) LIMIT k` without an index: (row_idx u32, distance f64),
        ascending; row_mask (bool per row, optional) = the WHERE clause evaluated over the table."""
        q = np.ascontiguousarray(query, dtype=np.float64).ravel()
        kk = max(k, 1)
        rows = np.zeros(kk, dtype=np.uint32)
        dist = np.zeros(kk, dtype=np.float64)
        cnt = C.c_uint32()
        bits = None
        if row_mask is not None:
            m = np.ascontiguousarray(row_mask, dtype=bool)
            if m.size != self.rows:
                raise PqvError(N.PQV_EINVAL, f"row_mask has {m.size} entries, the table {self.rows} rows")
            bits = np.packbits(m, bitorder="little")
        _check(_lib.pqv_array_distance_topk_filtered(self.ctx._h, self.handle, _ptr(q, C.c_double), q.size, metric, k,
                                                     _ptr(bits, C.c_uint8), _ptr(rows, C.c_uint32), _ptr(dist, C.c_double),
                                                     C.byref(cnt)))
        return rows[:cnt.value].copy(), dist[:cnt.value].copy()

    def l2_topk_gather(self, query, row_ids, k: int, flags: int = N.PQV_SQRT):
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        ids = np.ascontiguousarray(row_ids, dtype=np.uint32)
        kk = max(k, 1)
        rows = np.zeros(kk, dtype=np.uint32)
        dist = np.zeros(kk, dtype=np.float32)
        cnt = C.c_uint32()
        _check(_lib.pqv_l2_topk_gather(self.ctx._h, self.handle, _ptr(q, C.c_float), _ptr(ids, C.c_uint32), ids.size,
                                       k, flags, _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), C.byref(cnt)))
        return rows[:cnt.value].copy(), dist[:cnt.value].copy()

    def l2_topk_candidates(self, query, k: int, flags: int = N.PQV_SQRT, pos_base: int = 0, cap: int = 4096):
        """Heap-entrant candidate keys of this (rank-local) slice; see pqv_l2_topk_candidates."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        while True:
            keys = np.empty(cap, dtype=np.uint64)
            cnt = C.c_uint64()
            rc = _lib.pqv_l2_topk_candidates(self.ctx._h, self.handle, _ptr(q, C.c_float), k, flags, pos_base,
                                             _ptr(keys, C.c_uint64), cap, C.byref(cnt))
            if rc == N.PQV_ELIMIT and cnt.value > cap:
                cap = int(cnt.value)
                continue
            _check(rc)
            return keys[:cnt.value]

    def l2_topk_candidates_p2p(self, query, k: int, flags: int = N.PQV_SQRT, pos_base: int = 0, cap_total: int = 1 << 16):
        """Scan this rank's slice and exchange the candidates with the peers over NVLink (pqv_l2_topk_candidates_p2p):
        returns the union of all ranks' candidate keys, or None when a rank overflowed its slot (all ranks see that)."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        keys = np.empty(cap_total, dtype=np.uint64)
        cnt, ovf = C.c_uint64(), C.c_uint32()
        _check(_lib.pqv_l2_topk_candidates_p2p(self.ctx._h, self.handle, _ptr(q, C.c_float), k, flags, pos_base,
                                               _ptr(keys, C.c_uint64), cap_total, C.byref(cnt), C.byref(ovf)))
        return None if ovf.value else keys[:cnt.value]

    def l2_topk_p2p(self, query, k: int, flags: int = N.PQV_SQRT, pos_base: int = 0):
        """This rank's whole sharded search in one call (pqv_l2_topk_p2p): scan, NVLink candidate exchange, heap replay.
        Returns (row_idx, dist) -- identical on every rank -- or None when a rank overflowed its slot."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        rows = np.zeros(max(k, 1), dtype=np.uint32)
        dist = np.zeros(max(k, 1), dtype=np.float32)
        cnt, ovf = C.c_uint32(), C.c_uint32()
        _check(_lib.pqv_l2_topk_p2p(self.ctx._h, self.handle, _ptr(q, C.c_float), k, flags, pos_base, _ptr(rows, C.c_uint32),
                                    _ptr(dist, C.c_float), C.byref(cnt), C.byref(ovf)))
        return None if ovf.value else (rows[:cnt.value], dist[:cnt.value])

    def l2_topk_batch_p2p(self, queries, k: int, flags: int = N.PQV_SQRT, pos_base: int = 0):
        """This rank's whole sharded batch in one call (pqv_l2_topk_batch_p2p): tensor-core pass over the slice, key lists
        exchanged over NVLink peer memory, host merge, tie queries replayed.  Returns (row_idx [nq,k], dist [nq,k], count [nq],
        replayed) -- identical on every rank -- or None when the exchange slots are too small / a slice declined the batch."""
        q = np.atleast_2d(_f32(queries))
        if q.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.shape[1]}")
        nq, kk = q.shape[0], max(k, 1)
        rows = np.zeros((nq, kk), dtype=np.uint32)
        dist = np.zeros((nq, kk), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        rep, ovf = C.c_uint32(), C.c_uint32()
        _check(_lib.pqv_l2_topk_batch_p2p(self.ctx._h, self.handle, _ptr(q, C.c_float), nq, k, flags, pos_base,
                                          _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), _ptr(cnt, C.c_uint32), C.byref(rep),
                                          C.byref(ovf)))
        return None if ovf.value else (rows, dist, cnt, rep.value)

    def l2_topk_batch_keys(self, queries, k: int, flags: int = N.PQV_SQRT, pos_base: int = 0):
        """Per-rank half of a sharded batched search: (keys [nq, k+1] u64, counts [nq] u32), see pqv.h."""
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.shape[1]}")
        nq = q.shape[0]
        keys = np.full((nq, k + 1), np.iinfo(np.uint64).max, dtype=np.uint64)
        cnt = np.zeros(nq, dtype=np.uint32)
        _check(_lib.pqv_l2_topk_batch_keys(self.ctx._h, self.handle, _ptr(q, C.c_float), nq, k, flags, pos_base,
                                           _ptr(keys, C.c_uint64), _ptr(cnt, C.c_uint32)))
        return keys, cnt

    def l2_topk_batch_tie_candidates(self, q_index: int, query, cap: int = 16384):
        """Candidate keys of query `q_index` of the batch last passed to l2_topk_batch_keys (see pqv.h)."""
        q = _f32(query)
        while True:
            keys = np.empty(cap, dtype=np.uint64)
            cnt = C.c_uint64()
            rc = _lib.pqv_l2_topk_batch_tie_candidates(self.ctx._h, self.handle, q_index, _ptr(q, C.c_float),
                                                       _ptr(keys, C.c_uint64), cap, C.byref(cnt))
            if rc == N.PQV_ELIMIT and cnt.value > cap:
                cap = int(cnt.value)
                continue
            _check(rc)
            return keys[:cnt.value]

    def bench_scan(self, query, k: int, flags: int, iters: int) -> float:
        q = _f32(query)
        ms = C.c_double()
        _check(_lib.pqv_bench_scan(self.ctx._h, self.handle, _ptr(q, C.c_float), k, flags, iters, C.byref(ms)))
        return ms.value


class IvfIndex:
    """IVF index (centroids + inverted lists), host + HBM resident (pqv_ivf_*)."""

    def __init__(self, ctx: Context, handle: int):
        self.ctx, self.handle = ctx, handle
        d, c, n = C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(_lib.pqv_ivf_info(ctx._h, handle, C.byref(d), C.byref(c), C.byref(n)))
        self.dim, self.n_clusters, self.n_ids = d.value, c.value, n.value

    def to_bytes(self) -> bytes:
        need = C.c_uint64()
        _check(_lib.pqv_ivf_to_bytes(self.ctx._h, self.handle, None, 0, C.byref(need)))
        buf = (C.c_uint8 * need.value)()
        _check(_lib.pqv_ivf_to_bytes(self.ctx._h, self.handle, buf, need.value, C.byref(need)))
        return bytes(buf)

    def centroids(self) -> np.ndarray:
        """C x dim f32 centroid table, parsed from the blob header (src/ivf/index.rs:65-83: u32 dim, u32 C, f32[C*dim])."""
        b = self.to_bytes()
        dim, c = np.frombuffer(b, dtype="<u4", count=2)
        return np.frombuffer(b, dtype="<f4", count=int(dim) * int(c), offset=8).reshape(int(c), int(dim)).copy()

    def build_stats(self) -> dict:
        it = C.c_uint32()
        ms = (C.c_double * 4)()
        _check(_lib.pqv_ivf_build_stats(self.ctx._h, self.handle, C.byref(it), ms))
        return {"lloyd_iters": it.value, "init_ms": ms[0], "lloyd_ms": ms[1], "final_assign_ms": ms[2],
                "total_ms": ms[3]}

    def candidate_rows(self, query, nprobe: int) -> np.ndarray:
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        out = np.empty(max(self.n_ids, 1), dtype=np.uint32)
        n = C.c_uint64()
        _check(_lib.pqv_ivf_candidate_rows(self.ctx._h, self.handle, _ptr(q, C.c_float), nprobe,
                                           _ptr(out, C.c_uint32), out.size, C.byref(n)))
        return out[:n.value].copy()

    def search(self, dataset: "Dataset", query, k: int, nprobe: int, flags: int = N.PQV_SQRT):
        """TopkBuilder::search (src/ivf/search.rs:76-142) over a resident table."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        rows = np.zeros(max(k, 1), dtype=np.uint32)
        dist = np.zeros(max(k, 1), dtype=np.float32)
        cnt = C.c_uint32()
        _check(_lib.pqv_ivf_search(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), k, nprobe, flags,
                                   _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), C.byref(cnt)))
        return rows[:cnt.value].copy(), dist[:cnt.value].copy()

    def search_coalesced(self, dataset: "Dataset", query, k: int, nprobe: int, flags: int = N.PQV_SQRT):
        """TopkBuilder::search through the coalescing front door (pqv_ivf_search_coalesced): call from many threads."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        rows = np.zeros(max(k, 1), dtype=np.uint32)
        dist = np.zeros(max(k, 1), dtype=np.float32)
        cnt = C.c_uint32()
        _check(_lib.pqv_ivf_search_coalesced(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), k, nprobe, flags,
                                             _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), C.byref(cnt)))
        return rows[:cnt.value].copy(), dist[:cnt.value].copy()

    def search_batch(self, dataset: "Dataset", queries, k: int, nprobe: int, flags: int = N.PQV_SQRT):
        """nq independent searches in one pass over the table (pqv_ivf_search_batch).
        Returns (row_idx [nq,k] u32, dist [nq,k] f32, count [nq])."""
        q = np.atleast_2d(_f32(queries))
        if q.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.shape[1]}")
        nq, kk = q.shape[0], max(k, 1)
        rows = np.zeros((nq, kk), dtype=np.uint32)
        dist = np.zeros((nq, kk), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        _check(_lib.pqv_ivf_search_batch(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), nq, k, nprobe, flags,
                                         _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), _ptr(cnt, C.c_uint32)))
        return rows, dist, cnt

    def search_batch_keys(self, dataset: "Dataset", queries, k: int, nprobe: int, flags: int = N.PQV_SQRT, pos_base: int = 0):
        """Per-rank half of a sharded batched IVF search (pqv_ivf_search_batch_keys): (keys [nq, k+1] u64, counts [nq] u32)."""
        q = np.atleast_2d(_f32(queries))
        if q.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.shape[1]}")
        nq = q.shape[0]
        keys = np.full((nq, k + 1), np.iinfo(np.uint64).max, dtype=np.uint64)
        cnt = np.zeros(nq, dtype=np.uint32)
        _check(_lib.pqv_ivf_search_batch_keys(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), nq, k, nprobe, flags,
                                              pos_base, _ptr(keys, C.c_uint64), _ptr(cnt, C.c_uint32)))
        return keys, cnt

    def search_candidates(self, dataset: "Dataset", query, k: int, nprobe: int, flags: int = N.PQV_SQRT, cap: int = 1 << 16):
        """Per-rank half of a sharded IVF search (pqv_ivf_search_candidates): (keys u64 with positions in this rank's
        candidate sequence, local row ids u32, probed clusters in rank order)."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        while True:
            keys = np.empty(cap, dtype=np.uint64)
            rows = np.empty(cap, dtype=np.uint32)
            probe = np.empty(max(self.n_clusters, 1), dtype=np.uint32)
            n, npe = C.c_uint64(), C.c_uint32()
            rc = _lib.pqv_ivf_search_candidates(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), k, nprobe, flags,
                                                _ptr(keys, C.c_uint64), _ptr(rows, C.c_uint32), cap, C.byref(n),
                                                _ptr(probe, C.c_uint32), C.byref(npe))
            if rc == N.PQV_ELIMIT and n.value > cap:
                cap = int(n.value)
                continue
            _check(rc)
            return keys[:n.value].copy(), rows[:n.value].copy(), probe[:npe.value].copy()

    def vector_topk(self, dataset: "Dataset", query, k: int, nprobe: int, flags: int = N.PQV_SUM_SEQ,
                    max_candidates=None, row_mask=None):
        """VectorTopKExec over a resident indexed table (pqv_vector_topk_indexed): candidates capped in rank order,
        visited in row order, rows with row_mask[r] == False dropped before scoring.
        Returns (row_idx, squared distance, candidate_rows, rows_scored)."""
        q = _f32(query)
        if q.size != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.size}")
        bits = None
        if row_mask is not None:
            m = np.ascontiguousarray(row_mask, dtype=bool)
            if m.size != dataset.rows:
                raise PqvError(N.PQV_EINVAL, f"row_mask has {m.size} entries, the table {dataset.rows} rows")
            bits = np.packbits(m, bitorder="little")
        rows = np.zeros(max(k, 1), dtype=np.uint32)
        dist = np.zeros(max(k, 1), dtype=np.float32)
        cnt, cand, scored = C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(_lib.pqv_vector_topk_indexed(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), k, nprobe, flags,
                                            (0xFFFFFFFFFFFFFFFF if max_candidates is None else int(max_candidates)), _ptr(bits, C.c_uint8), _ptr(rows, C.c_uint32),
                                            _ptr(dist, C.c_float), C.byref(cnt), C.byref(cand), C.byref(scored)))
        return rows[:cnt.value].copy(), dist[:cnt.value].copy(), cand.value, scored.value

    def vector_topk_batch(self, dataset: "Dataset", queries, k: int, nprobe: int, flags: int = N.PQV_SUM_SEQ, row_mask=None):
        """nq VectorTopKExec executions sharing one filter in one pass (pqv_vector_topk_indexed_batch).
        Returns (row_idx [nq,k] u32, dist [nq,k] f32, count [nq])."""
        q = np.atleast_2d(_f32(queries))
        if q.shape[1] != self.dim:
            raise PqvError(N.PQV_EINVAL, f"Query dimension mismatch: expected {self.dim}, got {q.shape[1]}")
        bits = None
        if row_mask is not None:
            m = np.ascontiguousarray(row_mask, dtype=bool)
            if m.size != dataset.rows:
                raise PqvError(N.PQV_EINVAL, f"row_mask has {m.size} entries, the table {dataset.rows} rows")
            bits = np.packbits(m, bitorder="little")
        nq, kk = q.shape[0], max(k, 1)
        rows = np.zeros((nq, kk), dtype=np.uint32)
        dist = np.zeros((nq, kk), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        _check(_lib.pqv_vector_topk_indexed_batch(self.ctx._h, dataset.handle, self.handle, _ptr(q, C.c_float), nq, k, nprobe, flags,
                                                  _ptr(bits, C.c_uint8), _ptr(rows, C.c_uint32), _ptr(dist, C.c_float),
                                                  _ptr(cnt, C.c_uint32)))
        return rows, dist, cnt

    def drop(self):
        if self.handle:
            _check(_lib.pqv_ivf_drop(self.ctx._h, self.handle))
            self.handle = 0


def ivf_sample_rows(n_rows: int, n_clusters=None, seed: int = 42):
    """(global row ids of the training sample pqv_ivf_build would draw for a table of n_rows, cluster count C)."""
    n, c = C.c_uint64(), C.c_uint32()
    _check(_lib.pqv_ivf_sample_rows(n_rows, n_clusters or 0, seed, None, 0, C.byref(n), C.byref(c)))
    out = np.empty(n.value, dtype=np.uint32)
    _check(_lib.pqv_ivf_sample_rows(n_rows, n_clusters or 0, seed, _ptr(out, C.c_uint32), out.size, C.byref(n), C.byref(c)))
    return out, c.value


def replay_candidates(keys, k: int, flags: int = N.PQV_SQRT, row_ids=None):
    """Replay the reference's bounded BinaryHeap over candidate keys (union over ranks)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    ids = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.uint32)
    rows = np.zeros(max(k, 1), dtype=np.uint32)
    dist = np.zeros(max(k, 1), dtype=np.float32)
    cnt = C.c_uint32()
    _check(_lib.pqv_replay_candidates(_ptr(keys, C.c_uint64), keys.size, _ptr(ids, C.c_uint32), k, flags,
                                      _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), C.byref(cnt)))
    return rows[:cnt.value].copy(), dist[:cnt.value].copy()


def merge_batch_keys(keys, counts, k: int, flags: int = N.PQV_SQRT):
    """Merge the ranks' per-query key lists (keys [world, nq, k+1] u64, counts [world, nq] u32; pqv_merge_batch_keys).
    Returns (rows [nq, k], dist [nq, k], count [nq], needs_replay [nq] bool)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    world, nq = counts.shape
    assert keys.shape == (world, nq, k + 1)
    rows = np.zeros((nq, k), dtype=np.uint32)
    dist = np.zeros((nq, k), dtype=np.float32)
    cnt = np.zeros(nq, dtype=np.uint32)
    need = np.zeros(nq, dtype=np.uint8)
    _check(_lib.pqv_merge_batch_keys(_ptr(keys, C.c_uint64), _ptr(counts, C.c_uint32), world, nq, k, flags,
                                     _ptr(rows, C.c_uint32), _ptr(dist, C.c_float), _ptr(cnt, C.c_uint32),
                                     _ptr(need, C.c_uint8)))
    return rows, dist, cnt, need.astype(bool)


class TopkStream:
    """VectorTopKExec::topk_from_batches replacement (src/df_vector/exec.rs:257-277)."""

    def __init__(self, ctx: Context, query, k: int, flags: int):
        q = _f32(query)
        self.ctx, self.k, self.dim = ctx, k, q.size
        h = C.c_uint64()
        _check(_lib.pqv_topk_stream_begin(ctx._h, q.size, _ptr(q, C.c_float), k, flags, C.byref(h)))
        self.handle = h.value

    def push(self, values):
        v = np.asarray(values)
        if v.dtype == np.float64:
            v = np.ascontiguousarray(v)
            fn, ct = _lib.pqv_topk_stream_push_f64, C.c_double
        else:
            v = _f32(v)
            fn, ct = _lib.pqv_topk_stream_push, C.c_float
        if v.size % self.dim:
            raise PqvError(N.PQV_EINVAL, "Embedding data length must be a multiple of dimension")
        _check(fn(self.ctx._h, self.handle, _ptr(v, ct), v.size // self.dim))

    def finish(self):
        rows = np.zeros(self.k, dtype=np.uint32)
        dist = np.zeros(self.k, dtype=np.float32)
        cnt = C.c_uint32()
        _check(_lib.pqv_topk_stream_finish(self.ctx._h, self.handle, _ptr(rows, C.c_uint32), _ptr(dist, C.c_float),
                                           C.byref(cnt)))
        self.handle = 0
        return rows[:cnt.value].copy(), dist[:cnt.value].copy()
