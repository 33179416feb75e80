"""ctypes binding of oracle/libpqv_oracle.so (CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package (pq_vector_b200) never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpqv_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pqv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, u32p, u64p, u8p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint8), C.POINTER(C.c_double))
        sig = {
            "pqo_squared_l2_unroll4": (C.c_float, [f32p, f32p, C.c_size_t]),
            "pqo_squared_l2_seq": (C.c_float, [f32p, f32p, C.c_size_t]),
            "pqo_squared_l2_seq_f64": (C.c_float, [f64p, f32p, C.c_size_t]),
            "pqo_distances": (None, [f32p, C.c_uint64, C.c_uint32, f32p, C.c_int, f32p]),
            "pqo_heap_topk": (C.c_size_t, [f32p, u32p, C.c_uint64, C.c_size_t, C.c_int, u32p, f32p]),
            "pqo_topk_rerank": (C.c_size_t, [f32p, f32p, u32p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int,
                                             C.c_int, u32p, f32p]),
            "pqo_topk_rerank_gather": (C.c_size_t, [f32p, f32p, u32p, C.c_uint64, C.c_uint32, C.c_size_t,
                                                    C.c_int, C.c_int, u32p, f32p]),
            "pqo_nearest_centroid": (C.c_uint32, [f32p, f32p, C.c_uint32, C.c_uint32]),
            "pqo_assign": (None, [f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint32, u32p, C.c_int]),
            "pqo_inverted_lists": (None, [u32p, C.c_uint64, C.c_uint32, u64p, u32p]),
            "pqo_find_closest_centroids": (C.c_uint32, [f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p]),
            "pqo_candidate_rows": (C.c_uint64, [f32p, f32p, C.c_uint32, C.c_uint32, u64p, u32p, C.c_uint32, u32p]),
            "pqo_min_dist_init": (None, [f32p, u64p, C.c_uint64, C.c_uint32, f32p, f32p]),
            "pqo_min_dist_update": (C.c_float, [f32p, u64p, C.c_uint64, C.c_uint32, f32p, f32p, C.c_int]),
            "pqo_kmeanspp_pick": (C.c_uint64, [f32p, C.c_uint64, C.c_float]),
            "pqo_lloyd_assign": (C.c_uint64, [f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint32, u32p, u64p, C.c_int]),
            "pqo_centroid_update": (None, [f32p, C.c_uint64, C.c_uint32, u32p, u64p, C.c_uint32, f32p]),
            "pqo_build_sizes": (C.c_int, [C.c_uint64, C.c_uint64, u64p]),
            "pqo_index_blob_size": (C.c_uint64, [C.c_uint32, C.c_uint32, u64p]),
            "pqo_index_to_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32, f32p, u64p, u32p, u8p]),
            "pqo_index_from_bytes": (C.c_int, [u8p, C.c_uint64, u32p, u32p, u64p, f32p, u64p, u32p]),
            "pqo_synth_fill": (None, [f32p, C.c_uint64, C.c_uint64, C.c_uint64]),
            "pqo_array_distance": (C.c_double, [f32p, f64p, C.c_size_t]),
            "pqo_cosine_distance": (C.c_double, [f32p, f64p, C.c_size_t]),
            "pqo_array_distance_column": (None, [f32p, C.c_uint64, C.c_uint32, f64p, C.c_int, f64p]),
            "pqo_array_distance_topk": (C.c_size_t, [f32p, C.c_uint64, C.c_uint32, f64p, C.c_int, C.c_size_t, u32p, f64p]),
            "pqo_scan_topk_mt": (C.c_size_t, [f32p, C.c_uint64, C.c_uint32, f32p, C.c_size_t, C.c_int, C.c_int,
                                              u32p, f32p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint32)


# ---------------------------------------------------------------- distances
def squared_l2_unroll4(a, b) -> np.float32:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    return np.float32(lib().pqo_squared_l2_unroll4(_p(a, C.c_float), _p(b, C.c_float), a.size))


def squared_l2_seq(values, query) -> np.float32:
    v, q = _f32(values), _f32(query)
    return np.float32(lib().pqo_squared_l2_seq(_p(v, C.c_float), _p(q, C.c_float), v.size))


def squared_l2_seq_f64(values, query) -> np.float32:
    v = np.ascontiguousarray(values, dtype=np.float64)
    q = _f32(query)
    return np.float32(lib().pqo_squared_l2_seq_f64(_p(v, C.c_double), _p(q, C.c_float), v.size))


def distances(rows, query, order=0) -> np.ndarray:
    rows, query = _f32(rows), _f32(query)
    n, dim = rows.shape
    out = np.empty(n, dtype=np.float32)
    lib().pqo_distances(_p(rows, C.c_float), n, dim, _p(query, C.c_float), order, _p(out, C.c_float))
    return out


# ---------------------------------------------------------------- top-k
def heap_topk(dist, row_ids, k, do_sqrt):
    dist = _f32(dist)
    row_ids = _u32(row_ids)
    out_r = np.empty(max(k, 1), dtype=np.uint32)
    out_d = np.empty(max(k, 1), dtype=np.float32)
    n = lib().pqo_heap_topk(_p(dist, C.c_float), _p(row_ids, C.c_uint32), dist.size, k, int(do_sqrt),
                            _p(out_r, C.c_uint32), _p(out_d, C.c_float))
    return out_r[:n].copy(), out_d[:n].copy()


def topk_rerank(query, vectors, row_ids, k, order=0, do_sqrt=True):
    """src/ivf/search.rs:112-141 (order 0, sqrt) / src/df_vector/exec.rs:257-277 (order 1, no sqrt)."""
    query, vectors = _f32(query), _f32(vectors)
    n, dim = vectors.shape
    row_ids = _u32(row_ids)
    out_r = np.empty(max(k, 1), dtype=np.uint32)
    out_d = np.empty(max(k, 1), dtype=np.float32)
    m = lib().pqo_topk_rerank(_p(query, C.c_float), _p(vectors, C.c_float), _p(row_ids, C.c_uint32), n, dim, k,
                              order, int(do_sqrt), _p(out_r, C.c_uint32), _p(out_d, C.c_float))
    return out_r[:m].copy(), out_d[:m].copy()


def topk_rerank_gather(query, table, row_ids, k, order=0, do_sqrt=True):
    query, table = _f32(query), _f32(table)
    row_ids = _u32(row_ids)
    dim = table.shape[1]
    out_r = np.empty(max(k, 1), dtype=np.uint32)
    out_d = np.empty(max(k, 1), dtype=np.float32)
    m = lib().pqo_topk_rerank_gather(_p(query, C.c_float), _p(table, C.c_float), _p(row_ids, C.c_uint32),
                                     row_ids.size, dim, k, order, int(do_sqrt), _p(out_r, C.c_uint32),
                                     _p(out_d, C.c_float))
    return out_r[:m].copy(), out_d[:m].copy()


def scan_topk_mt(rows, query, k, order=0, workers=1):
    rows, query = _f32(rows), _f32(query)
    n, dim = rows.shape
    out_r = np.empty(max(k, 1), dtype=np.uint32)
    out_d = np.empty(max(k, 1), dtype=np.float32)
    m = lib().pqo_scan_topk_mt(_p(rows, C.c_float), n, dim, _p(query, C.c_float), k, order, workers,
                               _p(out_r, C.c_uint32), _p(out_d, C.c_float))
    return out_r[:m].copy(), out_d[:m].copy()


# ---------------------------------------------------------------- IVF
def nearest_centroid(vec, centroids) -> int:
    vec, centroids = _f32(vec), _f32(centroids)
    c, dim = centroids.shape
    return int(lib().pqo_nearest_centroid(_p(vec, C.c_float), _p(centroids, C.c_float), c, dim))


def assign(data, centroids, workers=1) -> np.ndarray:
    data, centroids = _f32(data), _f32(centroids)
    n, dim = data.shape
    out = np.empty(n, dtype=np.uint32)
    lib().pqo_assign(_p(data, C.c_float), n, dim, _p(centroids, C.c_float), centroids.shape[0],
                     _p(out, C.c_uint32), workers)
    return out


def inverted_lists(assign_, n_clusters):
    a = _u32(assign_)
    offsets = np.empty(n_clusters + 1, dtype=np.uint64)
    ids = np.empty(a.size, dtype=np.uint32)
    lib().pqo_inverted_lists(_p(a, C.c_uint32), a.size, n_clusters, _p(offsets, C.c_uint64), _p(ids, C.c_uint32))
    return offsets, ids


def find_closest_centroids(query, centroids, nprobe) -> np.ndarray:
    query, centroids = _f32(query), _f32(centroids)
    c, dim = centroids.shape
    out = np.empty(c, dtype=np.uint32)
    n = lib().pqo_find_closest_centroids(_p(query, C.c_float), _p(centroids, C.c_float), c, dim, nprobe,
                                         _p(out, C.c_uint32))
    return out[:n].copy()


def candidate_rows(query, centroids, offsets, ids, nprobe) -> np.ndarray:
    query, centroids = _f32(query), _f32(centroids)
    c, dim = centroids.shape
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    ids = _u32(ids)
    out = np.empty(max(int(offsets[-1]), 1), dtype=np.uint32)
    n = lib().pqo_candidate_rows(_p(query, C.c_float), _p(centroids, C.c_float), c, dim, _p(offsets, C.c_uint64),
                                 _p(ids, C.c_uint32), nprobe, _p(out, C.c_uint32))
    return out[:n].copy()


def min_dist_init(data, row_sel, centroid):
    data, centroid = _f32(data), _f32(centroid)
    sel = None if row_sel is None else np.ascontiguousarray(row_sel, dtype=np.uint64)
    n = data.shape[0] if sel is None else sel.size
    out = np.empty(n, dtype=np.float32)
    lib().pqo_min_dist_init(_p(data, C.c_float), _p(sel, C.c_uint64), n, data.shape[1], _p(centroid, C.c_float),
                            _p(out, C.c_float))
    return out


def min_dist_update(data, row_sel, centroid, min_dist, workers=1):
    """In-place on min_dist; returns the reference's `total` for this worker count."""
    data, centroid = _f32(data), _f32(centroid)
    assert min_dist.dtype == np.float32 and min_dist.flags.c_contiguous
    sel = None if row_sel is None else np.ascontiguousarray(row_sel, dtype=np.uint64)
    n = data.shape[0] if sel is None else sel.size
    return np.float32(lib().pqo_min_dist_update(_p(data, C.c_float), _p(sel, C.c_uint64), n, data.shape[1],
                                                _p(centroid, C.c_float), _p(min_dist, C.c_float), workers))


def kmeanspp_pick(min_dist, threshold) -> int:
    md = _f32(min_dist)
    return int(lib().pqo_kmeanspp_pick(_p(md, C.c_float), md.size, float(threshold)))


def lloyd_assign(data, centroids, assign_inout, workers=1):
    data, centroids = _f32(data), _f32(centroids)
    assert assign_inout.dtype == np.uint32 and assign_inout.flags.c_contiguous
    n, dim = data.shape
    sizes = np.empty(centroids.shape[0], dtype=np.uint64)
    changed = lib().pqo_lloyd_assign(_p(data, C.c_float), n, dim, _p(centroids, C.c_float), centroids.shape[0],
                                     _p(assign_inout, C.c_uint32), _p(sizes, C.c_uint64), workers)
    return int(changed), sizes


def centroid_update(data, assign_, sizes, n_clusters):
    data = _f32(data)
    a = _u32(assign_)
    sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
    out = np.empty((n_clusters, data.shape[1]), dtype=np.float32)
    lib().pqo_centroid_update(_p(data, C.c_float), data.shape[0], data.shape[1], _p(a, C.c_uint32),
                              _p(sizes, C.c_uint64), n_clusters, _p(out, C.c_float))
    return out


def build_sizes(n_vectors, n_clusters=None):
    out = np.zeros(3, dtype=np.uint64)
    rc = lib().pqo_build_sizes(n_vectors, n_clusters or 0, _p(out, C.c_uint64))
    if rc == 1:
        raise ValueError("Cannot build IVF index with zero vectors")
    if rc == 2:
        raise ValueError("n_clusters cannot exceed number of vectors")
    return int(out[0]), int(out[1]), int(out[2])


# ---------------------------------------------------------------- index blob
def index_to_bytes(dim, centroids, offsets, ids) -> bytes:
    centroids = _f32(centroids)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    ids = _u32(ids)
    c = centroids.size // dim
    size = lib().pqo_index_blob_size(dim, c, _p(offsets, C.c_uint64))
    buf = np.empty(size, dtype=np.uint8)
    n = lib().pqo_index_to_bytes(dim, c, _p(centroids, C.c_float), _p(offsets, C.c_uint64), _p(ids, C.c_uint32),
                                 _p(buf, C.c_uint8))
    assert n == size
    return buf.tobytes()


def index_from_bytes(blob: bytes):
    b = np.frombuffer(blob, dtype=np.uint8)
    dim, c, nids = C.c_uint32(), C.c_uint32(), C.c_uint64()
    rc = lib().pqo_index_from_bytes(_p(b, C.c_uint8), b.size, C.byref(dim), C.byref(c), C.byref(nids), None, None,
                                    None)
    if rc == 1:
        raise ValueError("IVF index buffer too small")
    if rc:
        raise ValueError("IVF index buffer malformed (rc=%d)" % rc)
    centroids = np.empty((c.value, dim.value), dtype=np.float32)
    offsets = np.empty(c.value + 1, dtype=np.uint64)
    ids = np.empty(max(nids.value, 1), dtype=np.uint32)
    rc = lib().pqo_index_from_bytes(_p(b, C.c_uint8), b.size, C.byref(dim), C.byref(c), C.byref(nids),
                                    _p(centroids, C.c_float), _p(offsets, C.c_uint64), _p(ids, C.c_uint32))
    assert rc == 0
    return dim.value, centroids, offsets, ids[: nids.value].copy()


# ---------------------------------------------------------------- synthetic data
def synth(n_rows, dim, seed, first_row=0) -> np.ndarray:
    out = np.empty((n_rows, dim), dtype=np.float32)
    lib().pqo_synth_fill(_p(out, C.c_float), first_row * dim, n_rows * dim, seed)
    return out


# ---- the un-indexed array_distance arm (DataFusion built-in; PARITY UNPINNED, see pqv_oracle.c) -------------------
def array_distance_column(rows, query, metric=0) -> np.ndarray:
    rows = _f32(rows)
    q = np.ascontiguousarray(query, dtype=np.float64)
    out = np.empty(rows.shape[0], dtype=np.float64)
    lib().pqo_array_distance_column(_p(rows, C.c_float), rows.shape[0], rows.shape[1], _p(q, C.c_double), int(metric),
                                    _p(out, C.c_double))
    return out


def array_distance_topk(rows, query, k, metric=0):
    rows = _f32(rows)
    q = np.ascontiguousarray(query, dtype=np.float64)
    out_r = np.empty(max(k, 1), dtype=np.uint32)
    out_d = np.empty(max(k, 1), dtype=np.float64)
    m = lib().pqo_array_distance_topk(_p(rows, C.c_float), rows.shape[0], rows.shape[1], _p(q, C.c_double), int(metric), k,
                                      _p(out_r, C.c_uint32), _p(out_d, C.c_double))
    return out_r[:m].copy(), out_d[:m].copy()
