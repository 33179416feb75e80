"""pq_vector_b200: B200 (sm_100a) implementation of pq-vector's squared-L2 / top-k / IVF-assign hot path.

Python surface = thin ctypes wrappers over the C ABI (include/pqv.h, `api.py`), the one-process-per-GPU protocol
(`sharded.py`) and a host-side mirror of the reference's public interface for this path (`builders.py`: IndexBuilder,
TopkBuilder, SearchResult, has_pq_vector_index, vector_topk) used by the parity tests that are written after the
reference's own."""
from ._native import (PQV_MAX_DIM, PQV_MAX_K, PQV_METRIC_COSINE, PQV_METRIC_L2, PQV_ROW_ORDER, PQV_SQRT, PQV_SUM_SEQ, PQV_SUM_UNROLL4, PQV_TIES_BY_POSITION,
                      LIB_PATH)
from .builders import (IndexBuilder, PqVectorError, SearchResult, TopkBuilder, VectorTopKOptions,  # noqa: E402
                       has_pq_vector_index, vector_topk)
from .session import SessionContext, SessionStateBuilder  # noqa: E402
from .api import Context, Dataset, IvfIndex, PqvError, TopkStream, merge_batch_keys, replay_candidates

__all__ = ["SessionStateBuilder", "SessionContext", "IndexBuilder", "TopkBuilder", "SearchResult", "VectorTopKOptions", "has_pq_vector_index", "vector_topk",
           "PqVectorError", "Context", "Dataset", "TopkStream", "IvfIndex", "PqvError", "replay_candidates", "merge_batch_keys", "PQV_SQRT", "PQV_SUM_SEQ", "PQV_SUM_UNROLL4",
           "PQV_TIES_BY_POSITION", "PQV_ROW_ORDER", "PQV_METRIC_L2", "PQV_METRIC_COSINE", "PQV_MAX_K", "PQV_MAX_DIM", "LIB_PATH"]
