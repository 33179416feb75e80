"""pq_vector_b200: B200 (sm_100a) implementation of pq-vector's squared-L2 / top-k / IVF-assign hot path.

Python surface = thin ctypes wrappers over the C ABI (include/pqv.h) plus a host-side mirror of the
reference's builders (TopkBuilder / IndexBuilder / SearchResult) used by the parity tests."""
from ._native import (PQV_MAX_DIM, PQV_MAX_K, PQV_SQRT, PQV_SUM_SEQ, PQV_SUM_UNROLL4, PQV_TIES_BY_POSITION,
                      LIB_PATH)
from .api import Context, Dataset, IvfIndex, PqvError, TopkStream, merge_batch_keys, replay_candidates

__all__ = ["Context", "Dataset", "TopkStream", "IvfIndex", "PqvError", "replay_candidates", "merge_batch_keys", "PQV_SQRT", "PQV_SUM_SEQ", "PQV_SUM_UNROLL4",
           "PQV_TIES_BY_POSITION", "PQV_MAX_K", "PQV_MAX_DIM", "LIB_PATH"]
