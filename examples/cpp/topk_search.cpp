// Run a top-k search against an indexed Parquet file (the reference's examples/topk_search.rs through pq_vector.hpp).
// Optional env vars: PQ_VECTOR_SOURCE, PQ_VECTOR_INDEXED, PQ_VECTOR_QUERY_ROW (default 0)
#include <cstdio>
#include <iostream>

#include "common.hpp"

int main() {
    try {
        const std::string source = common::source(), indexed = common::indexed();
        const size_t query_row = std::stoull(common::env_or("PQ_VECTOR_QUERY_ROW", "0"));
        common::ensure_indexed(source, indexed);
        const std::vector<float> query = common::read_embedding_at_row(indexed, "embedding", query_row);
        const auto results = pq_vector::TopkBuilder(indexed, query).k(5).nprobe(5).search();
        std::cout << "Top 5 neighbors for row " << query_row << ":\n";
        for (size_t rank = 0; rank < results.size(); ++rank)
            std::printf("%zu. row %u distance %.4f\n", rank + 1, results[rank].row_idx, results[rank].distance);
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
