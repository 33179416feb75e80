"""ctypes loader for pq_vector_b200/csrc/libpqv.so (the C ABI of include/pqv.h).

There is no fallback: if the shared library is missing or does not load, importing raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpqv.so")

PQV_OK, PQV_EINVAL, PQV_ENODEV, PQV_ECUDA, PQV_ENOMEM, PQV_EHANDLE, PQV_ELIMIT = range(7)
PQV_SUM_UNROLL4, PQV_SUM_SEQ, PQV_SQRT, PQV_TIES_BY_POSITION, PQV_ROW_ORDER = 0, 1, 2, 4, 8
PQV_MAX_K, PQV_MAX_DIM = 1024, 16384
PQV_METRIC_L2, PQV_METRIC_COSINE = 0, 1


class PqvTiming(C.Structure):
    _fields_ = [("scan_ms", C.c_double), ("post_ms", C.c_double), ("total_ms", C.c_double),
                ("scan_bytes", C.c_uint64), ("launches", C.c_uint32), ("entrants", C.c_uint32),
                ("grid", C.c_uint32), ("reserved", C.c_uint32)]


class PqvAssignTiming(C.Structure):
    _fields_ = [("path", C.c_uint32), ("kind", C.c_uint32), ("rows", C.c_uint64),
                ("ambiguous_rows", C.c_uint64), ("overflow_rows", C.c_uint64), ("prep_ms", C.c_double),
                ("filter_ms", C.c_double), ("recheck_ms", C.c_double), ("pair_ms", C.c_double),
                ("total_ms", C.c_double), ("shadow_ms", C.c_double)]


class PqvBatchTiming(C.Structure):
    _fields_ = [("queries", C.c_uint32), ("declined", C.c_uint32), ("tie_queries", C.c_uint32), ("tie_batched", C.c_uint32),
                ("rows", C.c_uint64), ("sample_rows", C.c_uint64), ("candidates", C.c_uint64), ("prep_ms", C.c_double),
                ("sample_ms", C.c_double), ("filter_ms", C.c_double), ("rerank_ms", C.c_double), ("total_ms", C.c_double)]


f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
ctxp = C.c_void_p

# every symbol include/pqv.h declares, with its signature
SIGNATURES = {
    "pqv_init": (C.c_int, [C.POINTER(ctxp), C.POINTER(C.c_int), C.c_int]),
    "pqv_destroy": (None, [ctxp]),
    "pqv_last_error": (C.c_char_p, []),
    "pqv_version": (C.c_char_p, []),
    "pqv_device_count": (C.c_int, [ctxp]),
    "pqv_dataset_create": (C.c_int, [ctxp, C.c_uint32, C.c_uint64, u64p]),
    "pqv_dataset_append": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint64]),
    "pqv_dataset_rows": (C.c_int, [ctxp, C.c_uint64, u64p, u32p]),
    "pqv_dataset_drop": (C.c_int, [ctxp, C.c_uint64]),
    "pqv_dataset_fill_synthetic": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "pqv_dataset_read": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, C.c_uint64, f32p]),
    "pqv_dataset_read_rows": (C.c_int, [ctxp, C.c_uint64, u32p, C.c_uint64, f32p]),
    "pqv_l2_topk": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p, u32p]),
    "pqv_l2_topk_gather": (C.c_int, [ctxp, C.c_uint64, f32p, u32p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p,
                                     u32p]),
    "pqv_topk_stream_begin": (C.c_int, [ctxp, C.c_uint32, f32p, C.c_uint32, C.c_uint32, u64p]),
    "pqv_topk_stream_push": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint64]),
    "pqv_topk_stream_push_f64": (C.c_int, [ctxp, C.c_uint64, f64p, C.c_uint64]),
    "pqv_topk_stream_finish": (C.c_int, [ctxp, C.c_uint64, u32p, f32p, u32p]),
    "pqv_kmeans_assign": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint32, u32p, u64p]),
    "pqv_min_dist_update": (C.c_int, [ctxp, C.c_uint64, f32p, u64p, C.c_uint64, C.c_uint32, f32p, C.c_int, f32p]),
    "pqv_centroid_rank": (C.c_int, [ctxp, f32p, C.c_uint32, C.c_uint32, f32p, C.c_uint32, C.c_uint32, u32p, u32p]),
    "pqv_ivf_build": (C.c_int, [ctxp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, u64p]),
    "pqv_ivf_sample_rows": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint64, u32p, C.c_uint64, u64p, u32p]),
    "pqv_kmeans_train": (C.c_int, [ctxp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, f32p, u32p]),
    "pqv_ivf_build_stats": (C.c_int, [ctxp, C.c_uint64, u32p, f64p]),
    "pqv_ivf_from_bytes": (C.c_int, [ctxp, C.POINTER(C.c_uint8), C.c_uint64, u64p]),
    "pqv_ivf_to_bytes": (C.c_int, [ctxp, C.c_uint64, C.POINTER(C.c_uint8), C.c_uint64, u64p]),
    "pqv_ivf_info": (C.c_int, [ctxp, C.c_uint64, u32p, u32p, u64p]),
    "pqv_ivf_drop": (C.c_int, [ctxp, C.c_uint64]),
    "pqv_ivf_candidate_rows": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, u32p, C.c_uint64, u64p]),
    "pqv_ivf_search": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p,
                                 u32p]),
    "pqv_l2_topk_candidates": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_uint64,
                                         u64p]),
    "pqv_replay_candidates": (C.c_int, [u64p, C.c_uint64, u32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]),
    "pqv_peer_exchange_create": (C.c_int, [ctxp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8)]),
    "pqv_peer_exchange_open": (C.c_int, [ctxp, C.POINTER(C.c_uint8)]),
    "pqv_l2_topk_candidates_p2p": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_uint64,
                                             u64p, u32p]),
    "pqv_l2_topk_p2p": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p, u32p, u32p]),
    "pqv_l2_topk_batch_p2p": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p,
                                        u32p, u32p, u32p]),
    "pqv_l2_topk_batch_keys": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u64p,
                                         u32p]),
    "pqv_l2_topk_batch_tie_candidates": (C.c_int, [ctxp, C.c_uint64, C.c_uint32, f32p, u64p, C.c_uint64, u64p]),
    "pqv_merge_batch_keys": (C.c_int, [u64p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p, u32p,
                                       C.POINTER(C.c_uint8)]),
    "pqv_last_timing": (C.c_int, [ctxp, C.POINTER(PqvTiming)]),
    "pqv_last_batch_timing": (C.c_int, [ctxp, C.POINTER(PqvBatchTiming)]),
    "pqv_last_assign_timing": (C.c_int, [ctxp, C.POINTER(PqvAssignTiming)]),
    "pqv_bench_assign": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32,
                                   C.POINTER(PqvAssignTiming), u32p]),
    "pqv_bench_scan": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, f64p]),
    "pqv_ivf_search_batch": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p,
                                       f32p, u32p]),
    "pqv_ivf_search_batch_keys": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint32, u64p, u32p]),
    "pqv_ivf_search_candidates": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p,
                                            C.c_uint64, u64p, u32p, u32p]),
    "pqv_vector_topk_indexed": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                          C.POINTER(C.c_uint8), u32p, f32p, u32p, u64p, u64p]),
    "pqv_vector_topk_indexed_batch": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.POINTER(C.c_uint8), u32p, f32p, u32p]),
    "pqv_l2_topk_coalesced": (C.c_int, [ctxp, C.c_uint64, f32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]),
    "pqv_ivf_search_coalesced": (C.c_int, [ctxp, C.c_uint64, C.c_uint64, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p,
                                           u32p]),
    "pqv_coalesce_config": (C.c_int, [ctxp, C.c_uint32, C.c_uint32]),
    "pqv_coalesce_stats": (C.c_int, [ctxp, u64p, u64p, u64p]),
    "pqv_array_distance": (C.c_int, [ctxp, C.c_uint64, f64p, C.c_uint32, C.c_uint32, f64p]),
    "pqv_array_distance_topk": (C.c_int, [ctxp, C.c_uint64, f64p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f64p, u32p]),
    "pqv_array_distance_topk_filtered": (C.c_int, [ctxp, C.c_uint64, f64p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.POINTER(C.c_uint8), u32p, f64p, u32p]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C pq_vector_b200/csrc).  pq_vector_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()
