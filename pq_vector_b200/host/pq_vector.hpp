// pq_vector.hpp -- the reference's public interface for the hot path, restated in C++ over Arrow C++ and the C ABI of
// include/pqv.h (libpqv.so).  pq-vector is a compiled (Rust) crate; no Rust toolchain exists in the authoring image, so this
// is the compiled host side above the boundary: the same names, argument meaning and error strings as the crate's
//   IndexBuilder            src/ivf/parquet.rs:21-102
//   TopkBuilder             src/ivf/search.rs:47-81   (SearchResult: search.rs:40-45)
//   has_pq_vector_index     src/ivf/parquet.rs:185-188
//   VectorTopKExec::topk_from_batches   src/df_vector/exec.rs:257-277   (vector_topk below)
// and the crate's Parquet-embedded index format (payload `PQ_VECTOR1 | u64 len | IvfIndex::to_bytes` at the footer key
// `pq_vector_index_offset`, column name at `pq_vector_embedding_column`; src/ivf/parquet.rs:105-112, 542-611).  Files written
// here are readable by the Rust crate and by pq_vector_b200/builders.py and vice versa.  Every distance is computed on the
// GPU through libpqv.so: without a B200 the calls throw (PQV_ENODEV), there is no CPU path.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace arrow {
class RecordBatch;
class Array;
}  // namespace arrow

namespace pq_vector {

/// The crate returns `Box<dyn Error>` with these messages (or DataFusionError in df_vector); here they are exceptions.
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

/// src/ivf/search.rs:40-45
struct SearchResult {
    uint32_t row_idx;
    float distance;
};

/// src/ivf/parquet.rs:185-188
bool has_pq_vector_index(const std::string &path);

/// src/ivf/parquet.rs:21-102.  Defaults as the crate: n_clusters = sqrt rule of build_ivf_index (index.rs:161-170),
/// max_iters = 20, seed = 42.  The table and the index stay resident in HBM after a build.
class IndexBuilder {
  public:
    IndexBuilder(std::string source, std::string embedding_column);
    IndexBuilder &n_clusters(size_t n_clusters);
    IndexBuilder &max_iters(size_t max_iters);
    IndexBuilder &seed(uint64_t seed);
    /// appends the index to `source` and rewrites its footer (data pages do not move), parquet.rs:57-69
    void build_inplace();
    /// writes `output` = the source's rows (vector column: no dictionary, one page per row size) + the index, parquet.rs:71-86
    void build_new(const std::string &output);

  private:
    std::string source_, embedding_column_;
    bool has_clusters_ = false;
    size_t n_clusters_ = 0, max_iters_ = 20;
    uint64_t seed_ = 42;
};

/// src/ivf/search.rs:47-81.  `search()` is the crate's async fn: callable from any number of threads, concurrent calls on
/// one file are answered by one batched pass (pqv_ivf_search_coalesced).
class TopkBuilder {
  public:
    TopkBuilder(std::string parquet_path, std::vector<float> query);
    TopkBuilder &k(size_t k);            // "k must be > 0"
    TopkBuilder &nprobe(size_t nprobe);  // "nprobe must be > 0"
    std::vector<SearchResult> search();  // "k must be set" / "nprobe must be set" / "Query dimension mismatch: ..."

  private:
    std::string path_;
    std::vector<float> query_;
    size_t k_ = 0, nprobe_ = 0;
};

/// VectorTopKExec::topk_from_batches (src/df_vector/exec.rs:257-277, 429-550): the <= k rows of `batches` closest to
/// `query` in the operator's arithmetic (sequential f32 sum, squared distance, f64 items narrowed first), ascending; null
/// rows and rows whose length differs from the query are skipped.  `column` is a List / LargeList / FixedSizeList of
/// Float32 / Float64.  Only the winners are materialised.  Returns an empty batch of the input schema when nothing scores;
/// `distances` (may be null) receives the squared distances of the returned rows.
std::shared_ptr<arrow::RecordBatch> vector_topk(const std::vector<std::shared_ptr<arrow::RecordBatch>> &batches,
                                                const std::string &column, const std::vector<float> &query, size_t k,
                                                std::vector<float> *distances = nullptr);

/// src/df_vector/options.rs:5-19 (defaults as the crate's `Default`)
struct VectorTopKOptions {
    size_t nprobe = 5;
    bool has_max_candidates = false;  // Option<usize>: None = no cap
    size_t max_candidates = 0;
};

/// What VectorTopKExec yields for ONE indexed file (execute_with_candidates + topk_from_batches,
/// src/df_vector/exec.rs:207-277): the winning file rows in the operator's output order, their squared distances, and the
/// two plan counters of the crate's snapshots (`candidate_rows`, `embeddings_fetched`).
struct VectorTopKRows {
    std::vector<uint32_t> rows;
    std::vector<float> distances;
    uint64_t candidate_rows = 0, embeddings_fetched = 0;
};

/// The operator over a file whose embedding column and index are resident: the index's candidates for `options.nprobe`
/// (index_exec.rs:152-163), the first `max_candidates` of them in rank order (CandidateCursor, access.rs:193-243), visited in
/// file order (the RowSelection of access.rs:107-176), rows failing the scan subtree's predicate dropped BEFORE scoring
/// (tests.rs:151-241), then the bounded heap of k in the operator's arithmetic (exec.rs:529-533).  `filter_mask`: the
/// predicate evaluated over the file's rows as an Arrow boolean buffer (bit r & 7 of byte r >> 3 = row r passes), at least
/// ceil(rows / 8) bytes; null = no FilterExec under the scan.  One host<->device round trip (pqv_vector_topk_indexed).
VectorTopKRows vector_topk_indexed(const std::string &parquet_path, const std::vector<float> &query, size_t k,
                                   const VectorTopKOptions &options = VectorTopKOptions(), const uint8_t *filter_mask = nullptr,
                                   size_t filter_mask_bytes = 0);

/// Drop every table and index this process keeps resident in HBM (keyed by path, size, mtime).
void drop_resident();

namespace detail {  // the file-format half, usable (and tested) without a GPU
/// parquet.rs:114-149, 176-183: (offset, embedding column) from the footer key-values; false if the file has no index
bool read_index_metadata(const std::string &path, uint64_t *offset, std::string *column);
/// parquet.rs:151-174, 191-208: the IvfIndex bytes embedded in the file + the embedding column name
std::pair<std::string, std::string> read_index_payload(const std::string &path);
/// parquet.rs:542-611: payload where the 8-byte footer tail was, then the old footer + the two key-values
void append_index_inplace(const std::string &path, const std::string &index_bytes, const std::string &embedding_column);
/// parquet.rs:210-303 (`read_parquet_with_embeddings`): validates the column as the crate does and hands each chunk's dense
/// f32 rows to `sink(values, n_rows)`; returns (rows, dim).  Float64 items are narrowed (parquet.rs:288-291).
std::pair<uint64_t, uint32_t> read_embeddings(const std::string &path, const std::string &embedding_column,
                                              const std::function<void(const float *, uint64_t, uint32_t)> &sink);
}  // namespace detail

}  // namespace pq_vector
