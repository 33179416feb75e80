// pqv_host_cli -- thin command-line driver over pq_vector.hpp; tests/test_host_cpp.py runs it against the Python mirror
// (pq_vector_b200/builders.py) and the oracle.  One sub-command per entry point of the reference's API for the path.
#include <arrow/api.h>
#include <arrow/io/api.h>
#include <parquet/arrow/reader.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "pq_vector.hpp"

namespace pv = pq_vector;

static std::string slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw pv::Error("cannot read " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
static std::vector<float> read_query(const std::string &path) {
    const std::string raw = slurp(path);
    std::vector<float> q(raw.size() / 4);
    memcpy(q.data(), raw.data(), q.size() * 4);
    return q;
}
static uint32_t bits(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return b;
}
static void configure(pv::IndexBuilder &b, int argc, char **argv, int first) {
    if (argc > first && strcmp(argv[first], "-") != 0) b.n_clusters(std::stoull(argv[first]));
    if (argc > first + 1) b.max_iters(std::stoull(argv[first + 1]));
    if (argc > first + 2) b.seed(std::stoull(argv[first + 2]));
}

int main(int argc, char **argv) {
    const std::string cmd = argc > 1 ? argv[1] : "";
    try {
        if (cmd == "has-index" && argc == 3) {
            std::cout << (pv::has_pq_vector_index(argv[2]) ? 1 : 0) << "\n";
        } else if (cmd == "append-index" && argc == 5) {
            pv::detail::append_index_inplace(argv[2], slurp(argv[3]), argv[4]);
        } else if (cmd == "read-index" && argc == 4) {
            auto p = pv::detail::read_index_payload(argv[2]);
            std::ofstream(argv[3], std::ios::binary).write(p.first.data(), (std::streamsize)p.first.size());
            std::cout << p.second << "\n";
        } else if (cmd == "read-embeddings" && argc == 4) {
            uint64_t h = 0xcbf29ce484222325ull;  // FNV-1a over the f32 bytes, as benches/query.rs:498-560 keys vectors
            auto shape = pv::detail::read_embeddings(argv[2], argv[3], [&](const float *v, uint64_t n, uint32_t dim) {
                const unsigned char *p = reinterpret_cast<const unsigned char *>(v);
                for (uint64_t i = 0; i < n * dim * 4; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
            });
            std::cout << shape.first << " " << shape.second << " " << h << "\n";
        } else if (cmd == "build-inplace" && argc >= 4) {
            pv::IndexBuilder b(argv[2], argv[3]);
            configure(b, argc, argv, 4);
            b.build_inplace();
        } else if (cmd == "build-new" && argc >= 5) {
            pv::IndexBuilder b(argv[2], argv[3]);
            configure(b, argc, argv, 5);
            b.build_new(argv[4]);
        } else if (cmd == "search" && argc == 6) {
            pv::TopkBuilder t(argv[2], read_query(argv[5]));
            if (strcmp(argv[3], "-") != 0) t.k(std::stoull(argv[3]));
            if (strcmp(argv[4], "-") != 0) t.nprobe(std::stoull(argv[4]));
            for (const auto &r : t.search()) std::cout << r.row_idx << " " << bits(r.distance) << "\n";
        } else if (cmd == "vector-topk" && argc == 7) {  // <parquet> <vector column> <k> <query file> <key column (int64)>
            auto file = arrow::io::ReadableFile::Open(argv[2]).ValueOrDie();
            auto reader = parquet::arrow::OpenFile(file, arrow::default_memory_pool()).ValueOrDie();
            reader->set_batch_size(3);  // several small batches: the winners come from different ones
            std::shared_ptr<arrow::RecordBatchReader> rb = reader->GetRecordBatchReader().ValueOrDie();
            std::vector<std::shared_ptr<arrow::RecordBatch>> batches = rb->ToRecordBatches().ValueOrDie();
            std::vector<float> dist;
            auto out = pv::vector_topk(batches, argv[3], read_query(argv[5]), std::stoull(argv[4]), &dist);
            auto keys = std::static_pointer_cast<arrow::Int64Array>(out->GetColumnByName(argv[6]));
            for (int64_t i = 0; i < out->num_rows(); ++i) std::cout << keys->Value(i) << " " << bits(dist[(size_t)i]) << "\n";
        } else if (cmd == "vector-topk-indexed" && (argc == 7 || argc == 8)) {
            // <indexed parquet> <k> <nprobe> <max_candidates or -> <query file> [filter: packed bits file]
            pv::VectorTopKOptions opt;
            opt.nprobe = std::stoull(argv[4]);
            if (strcmp(argv[5], "-") != 0) {
                opt.has_max_candidates = true;
                opt.max_candidates = std::stoull(argv[5]);
            }
            std::string mask = argc == 8 ? slurp(argv[7]) : std::string();
            auto r = pv::vector_topk_indexed(argv[2], read_query(argv[6]), std::stoull(argv[3]), opt,
                                             argc == 8 ? reinterpret_cast<const uint8_t *>(mask.data()) : nullptr, mask.size());
            std::cout << r.candidate_rows << " " << r.embeddings_fetched << "\n";
            for (size_t i = 0; i < r.rows.size(); ++i) std::cout << r.rows[i] << " " << bits(r.distances[i]) << "\n";
        } else {
            std::cerr << "usage: pqv_host_cli has-index|append-index|read-index|read-embeddings|build-inplace|build-new|search|vector-topk|vector-topk-indexed ...\n";
            return 2;
        }
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    pv::drop_resident();
    return 0;
}
