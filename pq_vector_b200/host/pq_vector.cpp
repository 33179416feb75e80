// pq_vector.cpp -- see pq_vector.hpp.  Host logic only: Parquet / Arrow handling, the embedded-index file format, argument
// validation with the crate's error strings, residency cache.  Every distance, ranking, k-means step and top-k runs in
// libpqv.so (include/pqv.h) on the GPU.
#include "pq_vector.hpp"

#include <arrow/api.h>
#include <arrow/io/api.h>
#include <parquet/arrow/reader.h>
#include <parquet/arrow/writer.h>
#include <parquet/file_reader.h>
#include <parquet/metadata.h>
#include <parquet/properties.h>
#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <tuple>

#include "../../include/pqv.h"

namespace pq_vector {
namespace {

constexpr char kMagic[] = "PQ_VECTOR1";                       // parquet.rs:105-106
constexpr size_t kMagicLen = sizeof(kMagic) - 1;
constexpr char kOffsetKey[] = "pq_vector_index_offset";       // parquet.rs:108-109
constexpr char kColumnKey[] = "pq_vector_embedding_column";   // parquet.rs:111-112
constexpr uint64_t kFooterSize = 8;                           // u32 metadata length + "PAR1"

[[noreturn]] void fail(const std::string &msg) { throw Error(msg); }

template <typename T>
T unwrap(arrow::Result<T> r) {
    if (!r.ok()) fail(r.status().ToString());
    return std::move(r).ValueUnsafe();
}
void check(const arrow::Status &st) {
    if (!st.ok()) fail(st.ToString());
}

// src/ivf/mod.rs:21-27
std::string embedding_column_name(const std::string &name) {
    if (name.find_first_not_of(" \t\n\r\f\v") == std::string::npos) fail("Embedding column name cannot be empty");
    return name;
}

// ---- the GPU context and what is resident on it ---------------------------------------------------------------------
pqv_ctx *gpu() {
    static std::mutex mu;
    static pqv_ctx *ctx = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!ctx && pqv_init(&ctx, nullptr, 0) != PQV_OK) {
        ctx = nullptr;
        fail(pqv_last_error());  // "no CUDA device available ...; libpqv has no CPU fallback"
    }
    return ctx;
}
void gpu_check(int rc) {
    if (rc != PQV_OK) fail(pqv_last_error());
}

using FileKey = std::tuple<std::string, uint64_t, int64_t>;  // path, size, mtime (ns)
FileKey file_key(const std::string &path) {
    struct stat st {};
    if (stat(path.c_str(), &st) != 0) fail("No such file or directory: " + path);
    return {path, (uint64_t)st.st_size, (int64_t)st.st_mtim.tv_sec * 1000000000 + st.st_mtim.tv_nsec};
}
struct ResidentTable {
    uint64_t handle = 0, rows = 0;
    uint32_t dim = 0;
};
struct ResidentIndex {
    uint64_t handle = 0;
    uint32_t dim = 0;
    std::string column;
};
std::mutex g_cache_mu;
using TableKey = std::pair<FileKey, std::string>;  // one resident block per (file state, vector column)
std::map<TableKey, ResidentTable> g_tables;
std::map<FileKey, ResidentIndex> g_indexes;

// a file that was rewritten leaves entries of its old (size, mtime) behind: release their HBM when the path comes back
// (g_cache_mu held)
void evict_stale(const FileKey &key) {
    // (an entry can only exist once a context does: nothing here touches the GPU on a cold cache)
    for (auto it = g_tables.begin(); it != g_tables.end();) {
        if (std::get<0>(it->first.first) == std::get<0>(key) && it->first.first != key) {
            pqv_dataset_drop(gpu(), it->second.handle);
            it = g_tables.erase(it);
        } else {
            ++it;
        }
    }
    for (auto it = g_indexes.begin(); it != g_indexes.end();) {
        if (std::get<0>(it->first) == std::get<0>(key) && it->first != key) {
            pqv_ivf_drop(gpu(), it->second.handle);
            it = g_indexes.erase(it);
        } else {
            ++it;
        }
    }
}

ResidentTable load_table(const std::string &path, const std::string &column) {
    pqv_ctx *ctx = gpu();
    ResidentTable t;
    try {
        auto shape = detail::read_embeddings(path, column, [&](const float *values, uint64_t n_rows, uint32_t dim) {
            if (!t.handle) {
                gpu_check(pqv_dataset_create(ctx, dim, n_rows, &t.handle));
                t.dim = dim;
            }
            gpu_check(pqv_dataset_append(ctx, t.handle, values, n_rows));  // the Arrow values buffer itself: no host copy
        });
        t.rows = shape.first;
    } catch (...) {
        if (t.handle) pqv_dataset_drop(ctx, t.handle);
        throw;
    }
    return t;
}

ResidentTable resident_table(const std::string &path, const std::string &column) {
    const TableKey key{file_key(path), column};
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_tables.find(key);
        if (it != g_tables.end()) return it->second;
        evict_stale(key.first);
    }
    ResidentTable t = load_table(path, column);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto ins = g_tables.emplace(key, t);
    if (!ins.second) pqv_dataset_drop(gpu(), t.handle);  // another thread loaded it meanwhile
    return ins.first->second;
}

ResidentIndex resident_index(const std::string &path) {
    const FileKey key = file_key(path);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_indexes.find(key);
        if (it != g_indexes.end()) return it->second;
        evict_stale(key);
    }
    auto payload = detail::read_index_payload(path);
    ResidentIndex ix;
    ix.column = payload.second;
    pqv_ctx *ctx = gpu();
    gpu_check(pqv_ivf_from_bytes(ctx, reinterpret_cast<const uint8_t *>(payload.first.data()), payload.first.size(), &ix.handle));
    uint32_t clusters = 0;
    uint64_t ids = 0;
    gpu_check(pqv_ivf_info(ctx, ix.handle, &ix.dim, &clusters, &ids));
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto ins = g_indexes.emplace(key, ix);
    if (!ins.second) pqv_ivf_drop(ctx, ix.handle);
    return ins.first->second;
}

std::shared_ptr<parquet::FileMetaData> file_metadata(const std::string &path) {
    auto file = unwrap(arrow::io::ReadableFile::Open(path));
    try {
        return parquet::ParquetFileReader::Open(file)->metadata();
    } catch (const std::exception &e) {  // parquet-cpp throws ParquetException
        fail(e.what());
    }
}

// (values, first element, validity) of a list-like column chunk: List / LargeList / FixedSizeList (exec.rs:494-519)
struct ListView {
    std::shared_ptr<arrow::Array> values;
    std::function<int64_t(int64_t)> offset;  // element offset of row r in `values` (r may be == length)
    const arrow::Array *list = nullptr;
};
bool list_view(const arrow::Array &a, ListView *out) {
    out->list = &a;
    switch (a.type_id()) {
        case arrow::Type::LIST: {
            auto &l = static_cast<const arrow::ListArray &>(a);
            out->values = l.values();
            out->offset = [&l](int64_t r) { return (int64_t)l.value_offset(r); };
            return true;
        }
        case arrow::Type::LARGE_LIST: {
            auto &l = static_cast<const arrow::LargeListArray &>(a);
            out->values = l.values();
            out->offset = [&l](int64_t r) { return (int64_t)l.value_offset(r); };
            return true;
        }
        case arrow::Type::FIXED_SIZE_LIST: {
            auto &l = static_cast<const arrow::FixedSizeListArray &>(a);
            out->values = l.values();
            const int64_t w = l.value_length(), base = l.offset();
            out->offset = [w, base](int64_t r) { return (base + r) * w; };
            return true;
        }
        default:
            return false;
    }
}

}  // namespace

// ---- file format --------------------------------------------------------------------------------------------------------
namespace detail {

bool read_index_metadata(const std::string &path, uint64_t *offset, std::string *column) {
    auto kv = file_metadata(path)->key_value_metadata();
    if (!kv) return false;
    const int io = kv->FindKey(kOffsetKey), ic = kv->FindKey(kColumnKey);
    if (io < 0 || ic < 0) return false;
    const std::string &off = kv->value(io);
    // Rust's `str::parse::<u64>`: optional '+', then digits only
    size_t p = (!off.empty() && off[0] == '+') ? 1 : 0;
    if (p == off.size()) fail("cannot parse integer from empty string");
    uint64_t v = 0;
    for (; p < off.size(); ++p) {
        if (off[p] < '0' || off[p] > '9') fail("invalid digit found in string");
        if (v > (UINT64_MAX - (uint64_t)(off[p] - '0')) / 10) fail("number too large to fit in target type");
        v = v * 10 + (uint64_t)(off[p] - '0');
    }
    if (offset) *offset = v;
    const std::string col = embedding_column_name(kv->value(ic));
    if (column) *column = col;
    return true;
}

std::pair<std::string, std::string> read_index_payload(const std::string &path) {
    uint64_t offset = 0;
    std::string column;
    if (!read_index_metadata(path, &offset, &column)) fail("Missing pq-vector index metadata in parquet footer");
    std::ifstream f(path, std::ios::binary);
    if (!f) fail("No such file or directory: " + path);
    f.seekg(0, std::ios::end);
    const uint64_t size = (uint64_t)f.tellg();
    std::string payload;
    if (offset < size) {
        payload.resize(size - offset);
        f.seekg((std::streamoff)offset);
        f.read(payload.data(), (std::streamsize)payload.size());
    }
    const size_t header = kMagicLen + 8;
    auto bad = [&](const char *why) {
        fail("Failed to decode pq-vector index payload at offset " + std::to_string(offset) + ": " + why);  // parquet.rs:205-207
    };
    if (payload.size() < header) bad("pq-vector index payload is truncated");
    if (memcmp(payload.data(), kMagic, kMagicLen) != 0) bad("Invalid pq-vector index magic");
    uint64_t len = 0;
    memcpy(&len, payload.data() + kMagicLen, 8);  // little endian host
    if (payload.size() - header < len) bad("pq-vector index bytes are truncated");
    return {payload.substr(header, len), column};
}

void append_index_inplace(const std::string &path, const std::string &index_bytes, const std::string &embedding_column) {
    const std::string column = embedding_column_name(embedding_column);
    std::shared_ptr<parquet::FileMetaData> md = file_metadata(path);
    struct stat st {};
    if (stat(path.c_str(), &st) != 0) fail("No such file or directory: " + path);
    const uint64_t file_len = (uint64_t)st.st_size;
    if (file_len < kFooterSize) fail("Parquet file too small to contain a footer");
    {
        std::ifstream f(path, std::ios::binary);
        char tail[8];
        f.seekg((std::streamoff)(file_len - kFooterSize));
        f.read(tail, 8);
        if (memcmp(tail + 4, "PARE", 4) == 0) fail("Encrypted parquet footers are not supported for in-place indexing");
        uint32_t meta_len = 0;
        memcpy(&meta_len, tail, 4);
        if ((uint64_t)meta_len + kFooterSize > file_len) fail("Parquet footer length exceeds file size");
    }
    const uint64_t index_offset = file_len - kFooterSize;  // the payload overwrites the old 8-byte tail; the old footer
                                                          // bytes stay behind as dead space, as in the crate
    std::vector<std::string> keys, values;
    if (auto kv = md->key_value_metadata())
        for (int64_t i = 0; i < kv->size(); ++i)
            if (kv->key(i) != kOffsetKey && kv->key(i) != kColumnKey) {
                keys.push_back(kv->key(i));
                values.push_back(kv->value(i));
            }
    keys.emplace_back(kOffsetKey);
    values.push_back(std::to_string(index_offset));
    keys.emplace_back(kColumnKey);
    values.push_back(column);
    // new footer = same schema, same row groups (column-chunk offsets unchanged: no data page moves), new key-values
    std::string footer;
    try {
        auto props = parquet::WriterProperties::Builder().version(md->version())->created_by(md->created_by())->build();
        auto fresh = parquet::FileMetaDataBuilder::Make(md->schema(), props)->Finish(arrow::key_value_metadata(keys, values));
        fresh->AppendRowGroups(*md);
        footer = fresh->SerializeToString();
    } catch (const std::exception &e) {
        fail(e.what());
    }
    std::FILE *f = std::fopen(path.c_str(), "r+b");
    if (!f) fail("cannot open " + path + " for writing");
    const uint64_t index_len = index_bytes.size();
    const uint32_t footer_len = (uint32_t)footer.size();
    bool ok = std::fseek(f, (long)index_offset, SEEK_SET) == 0;
    ok = ok && std::fwrite(kMagic, 1, kMagicLen, f) == kMagicLen;
    ok = ok && std::fwrite(&index_len, 1, 8, f) == 8;
    ok = ok && std::fwrite(index_bytes.data(), 1, index_bytes.size(), f) == index_bytes.size();
    ok = ok && std::fwrite(footer.data(), 1, footer.size(), f) == footer.size();
    ok = ok && std::fwrite(&footer_len, 1, 4, f) == 4;
    ok = ok && std::fwrite("PAR1", 1, 4, f) == 4;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) fail("failed to write the index to " + path);
}

std::pair<uint64_t, uint32_t> read_embeddings(const std::string &path, const std::string &embedding_column,
                                              const std::function<void(const float *, uint64_t, uint32_t)> &sink) {
    const std::string column = embedding_column_name(embedding_column);
    auto file = unwrap(arrow::io::ReadableFile::Open(path));
    std::unique_ptr<parquet::arrow::FileReader> reader = unwrap(parquet::arrow::OpenFile(file, arrow::default_memory_pool()));
    std::shared_ptr<arrow::Schema> schema;
    check(reader->GetSchema(&schema));
    const int field = schema->GetFieldIndex(column);
    if (field < 0) fail("Column '" + column + "' not found");
    // leaf columns of that top-level field
    std::vector<int> leaves;
    const auto *descr = reader->parquet_reader()->metadata()->schema();
    for (int i = 0; i < descr->num_columns(); ++i)
        if (descr->Column(i)->path()->ToDotVector().front() == column) leaves.push_back(i);
    std::shared_ptr<arrow::Table> table = unwrap(reader->ReadTable(leaves));
    std::shared_ptr<arrow::ChunkedArray> col = table->GetColumnByName(column);
    if (!col) fail("Column '" + column + "' not found");
    uint64_t rows = 0;
    uint32_t dim = 0;
    std::vector<float> narrowed;
    for (const auto &chunk : col->chunks()) {
        if (chunk->type_id() != arrow::Type::LIST) fail("Embedding column is not a list array");  // parquet.rs:236-239
        const auto &list = static_cast<const arrow::ListArray &>(*chunk);
        if (list.null_count() > 0) fail("Embedding column contains null rows");
        const auto &values = list.values();
        const bool f32 = values->type_id() == arrow::Type::FLOAT, f64 = values->type_id() == arrow::Type::DOUBLE;
        if (!f32 && !f64) fail("Embedding values are not float32/float64");
        if (values->null_count() > 0) fail("Embedding values contain nulls");
        for (int64_t r = 0; r < list.length(); ++r) {
            const int64_t len = list.value_length(r);
            if (len == 0) fail("Embedding row has zero length");
            if (dim == 0) dim = (uint32_t)len;
            else if ((uint32_t)len != dim) fail("Embedding vectors have inconsistent dimensions");
        }
        if (list.length() == 0) continue;
        const int64_t first = list.value_offset(0), count = list.length() * (int64_t)dim;
        if (f32) {
            sink(static_cast<const arrow::FloatArray &>(*values).raw_values() + first, (uint64_t)list.length(), dim);
        } else {
            const double *src = static_cast<const arrow::DoubleArray &>(*values).raw_values() + first;
            narrowed.resize((size_t)count);
            for (int64_t i = 0; i < count; ++i) narrowed[(size_t)i] = (float)src[i];  // parquet.rs:288-291 `as f32`
            sink(narrowed.data(), (uint64_t)list.length(), dim);
        }
        rows += (uint64_t)list.length();
    }
    if (rows == 0) fail("Embedding column has no rows");
    return {rows, dim};
}

}  // namespace detail

bool has_pq_vector_index(const std::string &path) { return detail::read_index_metadata(path, nullptr, nullptr); }

void drop_resident() {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_tables.empty() && g_indexes.empty()) return;
    pqv_ctx *ctx = gpu();
    for (auto &kv : g_tables) pqv_dataset_drop(ctx, kv.second.handle);
    for (auto &kv : g_indexes) pqv_ivf_drop(ctx, kv.second.handle);
    g_tables.clear();
    g_indexes.clear();
}

// ---- IndexBuilder -----------------------------------------------------------------------------------------------------------
IndexBuilder::IndexBuilder(std::string source, std::string embedding_column)
    : source_(std::move(source)), embedding_column_(std::move(embedding_column)) {}
IndexBuilder &IndexBuilder::n_clusters(size_t n) {
    has_clusters_ = true;
    n_clusters_ = n;
    return *this;
}
IndexBuilder &IndexBuilder::max_iters(size_t n) {
    max_iters_ = n;
    return *this;
}
IndexBuilder &IndexBuilder::seed(uint64_t s) {
    seed_ = s;
    return *this;
}

namespace {
struct Built {
    ResidentTable table;
    uint64_t index = 0;
    std::string blob, column;
};
// build_config (parquet.rs:88-102) + read_parquet_with_embeddings + build_ivf_index (index.rs:152-214, on the device)
Built build_index(const std::string &source, const std::string &embedding_column, bool has_clusters, size_t n_clusters,
                  size_t max_iters, uint64_t seed) {
    if (max_iters == 0) fail("max_iters must be > 0");
    if (has_clusters && n_clusters == 0) fail("n_clusters must be > 0");
    Built b;
    b.column = embedding_column_name(embedding_column);
    pqv_ctx *ctx = gpu();
    b.table = load_table(source, b.column);
    try {
        gpu_check(pqv_ivf_build(ctx, b.table.handle, has_clusters ? (uint32_t)n_clusters : 0u, (uint32_t)max_iters, seed, 0, &b.index));
        uint64_t len = 0;
        gpu_check(pqv_ivf_to_bytes(ctx, b.index, nullptr, 0, &len));
        b.blob.resize(len);
        gpu_check(pqv_ivf_to_bytes(ctx, b.index, reinterpret_cast<uint8_t *>(b.blob.data()), len, &len));
    } catch (...) {
        pqv_dataset_drop(ctx, b.table.handle);
        if (b.index) pqv_ivf_drop(ctx, b.index);
        throw;
    }
    return b;
}
// the freshly built table and index stay resident for the searches that follow
void keep_resident(const std::string &path, const Built &b) {
    const FileKey key = file_key(path);
    uint32_t dim = 0, clusters = 0;
    uint64_t ids = 0;
    gpu_check(pqv_ivf_info(gpu(), b.index, &dim, &clusters, &ids));
    std::lock_guard<std::mutex> lk(g_cache_mu);
    evict_stale(key);
    g_tables[TableKey{key, b.column}] = b.table;
    g_indexes[key] = ResidentIndex{b.index, dim, b.column};
}
}  // namespace

void IndexBuilder::build_inplace() {
    Built b = build_index(source_, embedding_column_, has_clusters_, n_clusters_, max_iters_, seed_);
    detail::append_index_inplace(source_, b.blob, b.column);
    keep_resident(source_, b);
}

void IndexBuilder::build_new(const std::string &output) {
    Built b = build_index(source_, embedding_column_, has_clusters_, n_clusters_, max_iters_, seed_);
    // write_parquet_with_index (parquet.rs:313-375): the source's rows, each column with the source's compression
    // (parquet.rs:379-470), the vector column without dictionary and with one row per data page
    auto file = unwrap(arrow::io::ReadableFile::Open(source_));
    std::unique_ptr<parquet::arrow::FileReader> reader = unwrap(parquet::arrow::OpenFile(file, arrow::default_memory_pool()));
    std::shared_ptr<arrow::Table> table = unwrap(reader->ReadTable());
    auto md = reader->parquet_reader()->metadata();
    parquet::WriterProperties::Builder props;
    props.data_pagesize((int64_t)b.table.dim * 4);
    if (md->num_row_groups() > 0) {
        auto rg = md->RowGroup(0);
        for (int i = 0; i < rg->num_columns(); ++i) {
            auto cc = rg->ColumnChunk(i);
            props.compression(cc->path_in_schema(), cc->compression());
        }
    }
    const auto *descr = md->schema();
    for (int i = 0; i < descr->num_columns(); ++i)
        if (descr->Column(i)->path()->ToDotVector().front() == b.column) props.disable_dictionary(descr->Column(i)->path());
    auto sink = unwrap(arrow::io::FileOutputStream::Open(output));
    check(parquet::arrow::WriteTable(*table, arrow::default_memory_pool(), sink, parquet::DEFAULT_MAX_ROW_GROUP_LENGTH, props.build()));
    check(sink->Close());
    detail::append_index_inplace(output, b.blob, b.column);
    keep_resident(output, b);
}

// ---- TopkBuilder --------------------------------------------------------------------------------------------------------------
TopkBuilder::TopkBuilder(std::string parquet_path, std::vector<float> query) : path_(std::move(parquet_path)), query_(std::move(query)) {}
TopkBuilder &TopkBuilder::k(size_t k) {
    if (k == 0) fail("k must be > 0");
    k_ = k;
    return *this;
}
TopkBuilder &TopkBuilder::nprobe(size_t nprobe) {
    if (nprobe == 0) fail("nprobe must be > 0");
    nprobe_ = nprobe;
    return *this;
}
std::vector<SearchResult> TopkBuilder::search() {
    if (k_ == 0) fail("k must be set");
    if (nprobe_ == 0) fail("nprobe must be set");
    const ResidentIndex ix = resident_index(path_);  // read_index_from_parquet, once per file state
    if (query_.size() != ix.dim)
        fail("Query dimension mismatch: expected " + std::to_string(ix.dim) + ", got " + std::to_string(query_.size()));
    const ResidentTable t = resident_table(path_, ix.column);
    if (t.dim != ix.dim)  // search.rs:224-231
        fail("Embedding dimension mismatch: expected " + std::to_string(ix.dim) + ", got " + std::to_string(t.dim));
    std::vector<uint32_t> rows(k_);
    std::vector<float> dist(k_);
    uint32_t n = 0;
    gpu_check(pqv_ivf_search_coalesced(gpu(), t.handle, ix.handle, query_.data(), (uint32_t)k_, (uint32_t)nprobe_, PQV_SQRT,
                                       rows.data(), dist.data(), &n));
    std::vector<SearchResult> out(n);
    for (uint32_t i = 0; i < n; ++i) out[i] = SearchResult{rows[i], dist[i]};
    return out;
}

// ---- VectorTopKExec over one resident, indexed file -------------------------------------------------------------------
VectorTopKRows vector_topk_indexed(const std::string &parquet_path, const std::vector<float> &query, size_t k,
                                   const VectorTopKOptions &options, const uint8_t *filter_mask, size_t filter_mask_bytes) {
    if (k == 0) fail("k must be > 0");
    if (options.nprobe == 0) fail("nprobe must be > 0");
    const ResidentIndex ix = resident_index(parquet_path);  // "Missing pq-vector index metadata ..." for an un-indexed file
    if (query.size() != ix.dim)                             // index_exec.rs:152-158
        fail("Query dimension mismatch: expected " + std::to_string(ix.dim) + ", got " + std::to_string(query.size()));
    const ResidentTable t = resident_table(parquet_path, ix.column);
    if (filter_mask && filter_mask_bytes < (t.rows + 7) / 8)
        fail("filter mask has " + std::to_string(filter_mask_bytes) + " bytes, the file " + std::to_string(t.rows) + " rows");
    VectorTopKRows out;
    out.rows.resize(k);
    out.distances.resize(k);
    uint32_t n = 0;
    gpu_check(pqv_vector_topk_indexed(gpu(), t.handle, ix.handle, query.data(), (uint32_t)k, (uint32_t)options.nprobe, PQV_SUM_SEQ,
                                      options.has_max_candidates ? (uint64_t)options.max_candidates : PQV_NO_CANDIDATE_CAP, filter_mask,
                                      out.rows.data(), out.distances.data(), &n, &out.candidate_rows, &out.embeddings_fetched));
    out.rows.resize(n);
    out.distances.resize(n);
    return out;
}

// ---- VectorTopKExec::topk_from_batches ---------------------------------------------------------------------------------
std::shared_ptr<arrow::RecordBatch> vector_topk(const std::vector<std::shared_ptr<arrow::RecordBatch>> &batches,
                                                const std::string &column, const std::vector<float> &query, size_t k,
                                                std::vector<float> *distances) {
    if (batches.empty()) fail("vector_topk needs at least one record batch (for the schema)");
    if (k == 0) fail("k must be > 0");
    const std::shared_ptr<arrow::Schema> schema = batches[0]->schema();
    const int ci = schema->GetFieldIndex(column);
    if (ci < 0) fail("Vector column '" + column + "' not found in schema");  // exec.rs:247-255
    pqv_ctx *ctx = gpu();
    uint64_t stream = 0;
    gpu_check(pqv_topk_stream_begin(ctx, (uint32_t)query.size(), query.data(), (uint32_t)k, PQV_SUM_SEQ, &stream));
    std::vector<std::pair<uint32_t, int64_t>> origin;  // pushed sequence -> (batch, row)
    std::vector<float> pack32;
    std::vector<double> pack64;
    auto finish_into = [&](std::vector<uint32_t> &idx, std::vector<float> &dist) {
        uint32_t n = 0;
        idx.resize(k);
        dist.resize(k);
        const int rc = pqv_topk_stream_finish(ctx, stream, idx.data(), dist.data(), &n);  // releases the stream either way
        stream = 0;
        gpu_check(rc);
        idx.resize(n);
        dist.resize(n);
    };
    try {
        for (size_t b = 0; b < batches.size(); ++b) {
            const arrow::Array &arr = *batches[b]->column(ci);
            ListView lv;
            if (!list_view(arr, &lv)) fail("Vector column must be list or fixed-size list");  // exec.rs:516-518
            const bool f32 = lv.values->type_id() == arrow::Type::FLOAT, f64 = lv.values->type_id() == arrow::Type::DOUBLE;
            if (!f32 && !f64) fail("Vector column must be Float32 or Float64 list");           // exec.rs:547-549
            const int64_t n = arr.length(), dim = (int64_t)query.size();
            // rows the operator scores: not null (exec.rs:496-498), length == query length (:526-528, :537-539)
            int64_t kept = 0;
            bool dense = true;
            for (int64_t r = 0; r < n; ++r) {
                const bool keep = arr.IsValid(r) && lv.offset(r + 1) - lv.offset(r) == dim;
                dense = dense && keep;
                kept += keep;
            }
            if (kept == 0) continue;
            const float *v32 = f32 ? static_cast<const arrow::FloatArray &>(*lv.values).raw_values() : nullptr;
            const double *v64 = f64 ? static_cast<const arrow::DoubleArray &>(*lv.values).raw_values() : nullptr;
            if (dense && lv.offset(n) - lv.offset(0) == n * dim) {  // the values buffer is the dense row block itself
                if (f32) gpu_check(pqv_topk_stream_push(ctx, stream, v32 + lv.offset(0), (uint64_t)n));
                else gpu_check(pqv_topk_stream_push_f64(ctx, stream, v64 + lv.offset(0), (uint64_t)n));
                for (int64_t r = 0; r < n; ++r) origin.emplace_back((uint32_t)b, r);
            } else {
                pack32.clear();
                pack64.clear();
                for (int64_t r = 0; r < n; ++r) {
                    if (!(arr.IsValid(r) && lv.offset(r + 1) - lv.offset(r) == dim)) continue;
                    if (f32) pack32.insert(pack32.end(), v32 + lv.offset(r), v32 + lv.offset(r) + dim);
                    else pack64.insert(pack64.end(), v64 + lv.offset(r), v64 + lv.offset(r) + dim);
                    origin.emplace_back((uint32_t)b, r);
                }
                if (f32) gpu_check(pqv_topk_stream_push(ctx, stream, pack32.data(), (uint64_t)kept));
                else gpu_check(pqv_topk_stream_push_f64(ctx, stream, pack64.data(), (uint64_t)kept));
            }
        }
    } catch (...) {
        std::vector<uint32_t> i0;
        std::vector<float> d0;
        try {
            finish_into(i0, d0);
        } catch (...) {
        }
        throw;
    }
    std::vector<uint32_t> idx;
    std::vector<float> dist;
    finish_into(idx, dist);
    if (distances) *distances = dist;
    if (idx.empty()) return unwrap(arrow::RecordBatch::MakeEmpty(schema));
    // only the winners are materialised (the crate builds ScalarValues for every candidate row, exec.rs:472)
    std::vector<std::shared_ptr<arrow::RecordBatch>> rows;
    rows.reserve(idx.size());
    for (uint32_t i : idx) rows.push_back(batches[origin[i].first]->Slice(origin[i].second, 1));
    return unwrap(arrow::ConcatenateRecordBatches(rows));
}

}  // namespace pq_vector
