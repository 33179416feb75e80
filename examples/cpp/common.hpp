// The role of the reference's examples/common/mod.rs for the C++ examples: environment defaults, `ensure_indexed`,
// `read_embedding_at_row`.
#pragma once
#include <arrow/api.h>
#include <arrow/io/api.h>
#include <parquet/arrow/reader.h>
#include <sys/stat.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "../../pq_vector_b200/host/pq_vector.hpp"

namespace common {

inline std::string env_or(const char *name, const std::string &fallback) {
    const char *v = std::getenv(name);
    return v && *v ? std::string(v) : fallback;
}
inline std::string source() { return env_or("PQ_VECTOR_SOURCE", "data/vldb_2025.parquet"); }
inline std::string indexed() { return env_or("PQ_VECTOR_INDEXED", "data/vldb_2025_indexed.parquet"); }

inline bool exists(const std::string &path) {
    struct stat st {};
    return stat(path.c_str(), &st) == 0;
}

// examples/common/mod.rs `ensure_indexed`: build the indexed copy unless it is already there
inline void ensure_indexed(const std::string &src, const std::string &out) {
    if (exists(out) && pq_vector::has_pq_vector_index(out)) return;
    pq_vector::IndexBuilder(src, "embedding").build_new(out);
}

// examples/common/mod.rs `read_embedding_at_row`
inline std::vector<float> read_embedding_at_row(const std::string &path, const std::string &column, size_t row) {
    std::vector<float> out;
    size_t seen = 0;
    pq_vector::detail::read_embeddings(path, column, [&](const float *values, uint64_t n_rows, uint32_t dim) {
        if (out.empty() && row >= seen && row < seen + n_rows) out.assign(values + (row - seen) * dim, values + (row - seen + 1) * dim);
        seen += n_rows;
    });
    if (out.empty()) throw pq_vector::Error("row " + std::to_string(row) + " out of range (" + std::to_string(seen) + " rows)");
    return out;
}

}  // namespace common
