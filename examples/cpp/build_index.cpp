// Build an IVF index and write it to a new Parquet file (the reference's examples/build_index.rs through pq_vector.hpp).
// Optional env vars: PQ_VECTOR_SOURCE (default data/vldb_2025.parquet), PQ_VECTOR_INDEXED (default data/vldb_2025_indexed.parquet)
#include <iostream>

#include "common.hpp"

int main() {
    try {
        const std::string source = common::source(), indexed = common::indexed();
        std::cout << "Building IVF index from " << source << "...\n";
        pq_vector::IndexBuilder(source, "embedding").build_new(indexed);
        std::cout << "Wrote indexed parquet to " << indexed << "\n";
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
