// pqv_ivf_impl.cuh -- host-side mirror of the reference's IVF layer on top of the kernels (included at the
// end of pqv_capi.cu; same translation unit so it shares pqv_ctx / DeviceState).
//
//   build    : build_ivf_index + k_means                    src/ivf/index.rs:152-214, 323-457
//   blob     : IvfIndex::to_bytes / from_bytes              src/ivf/index.rs:65-128   (byte-identical format)
//   search   : TopkBuilder::topk minus the Parquet I/O      src/ivf/search.rs:83-142
//   candidates: IvfIndex::candidate_rows                    src/ivf/index.rs:57-63
//
// Every distance, argmin, centroid mean and top-k comes from the GPU kernels; the host keeps what the
// reference keeps serial: RNG draws, the f32 running sums of the k-means++ pick (index.rs:356-383, whose
// chunking depends on the worker count), list bookkeeping.  The RNG stream is NOT rand 0.8.5's
// ChaCha12 (unverifiable here): the sampled rows differ from a Rust build, every step after the
// draws is bit-identical given the same draws (tests pin sweeps/assignments/updates against the oracle).
#pragma once

namespace {

struct SplitMix64 {
    u64 s;
    explicit SplitMix64(u64 seed) : s(seed) {}
    u64 next() {
        u64 z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    u64 below(u64 n) {  // uniform in [0, n), rejection sampling
        const u64 lim = ~0ull - (~0ull % n);
        u64 v;
        do v = next();
        while (v >= lim);
        return v % n;
    }
    float unit_f32() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }  // [0,1), 24 bits
};

// `amount` distinct indices from [0, n) in random order (role of rand::seq::index::sample)
std::vector<uint32_t> sample_indices(SplitMix64 &rng, u64 n, u64 amount) {
    std::vector<uint32_t> out;
    out.reserve(amount);
    if (amount * 2 >= n) {  // partial Fisher-Yates
        std::vector<uint32_t> all(n);
        for (u64 i = 0; i < n; ++i) all[i] = (uint32_t)i;
        for (u64 i = 0; i < amount; ++i) {
            const u64 j = i + rng.below(n - i);
            std::swap(all[i], all[j]);
        }
        out.assign(all.begin(), all.begin() + amount);
    } else {  // Floyd's algorithm, then shuffle; membership in an open-addressing table (a std::map costs 25 ms at 100 k draws)
        u64 cap = 16;
        while (cap < amount * 4) cap <<= 1;
        const uint32_t EMPTY = 0xFFFFFFFFu;  // never a valid index here: amount * 2 < n <= 2^32 - 1 leaves it unused or handled below
        std::vector<uint32_t> table(cap, EMPTY);
        bool has_empty_value = false;  // the value 0xFFFFFFFF itself (only possible when n = 2^32)
        auto contains_or_insert = [&](uint32_t v) -> bool {  // true if v was already present
            if (v == EMPTY) {
                const bool was = has_empty_value;
                has_empty_value = true;
                return was;
            }
            u64 h = ((u64)v * 0x9E3779B97F4A7C15ull) >> 32;
            for (u64 i = h & (cap - 1);; i = (i + 1) & (cap - 1)) {
                if (table[i] == v) return true;
                if (table[i] == EMPTY) {
                    table[i] = v;
                    return false;
                }
            }
        };
        for (u64 j = n - amount; j < n; ++j) {
            const uint32_t t = (uint32_t)rng.below(j + 1);
            if (contains_or_insert(t)) {
                contains_or_insert((uint32_t)j);
                out.push_back((uint32_t)j);
            } else {
                out.push_back(t);
            }
        }
        for (u64 i = amount; i > 1; --i) std::swap(out[i - 1], out[rng.below(i)]);
    }
    return out;
}

struct IvfIndex {
    uint32_t dim = 0, n_clusters = 0;
    std::vector<float> centroids;
    std::vector<u64> offsets;  // n_clusters + 1
    std::vector<uint32_t> ids;  // host copy of the lists; after a device build it is fetched on first use
    u64 n_ids = 0;
    uint32_t max_id = 0;         // largest row id in any list (from_bytes: taken from the blob; device build: n - 1)
    bool host_ids = true;        // `ids` holds the lists
    bool lists_on_device = false;  // d_offsets / d_ids already hold the lists (device-built index)
    // device mirror (device 0 of the context), created on first search
    bool resident = false;
    DevBuf<float> d_centroids, d_cdist;
    DevBuf<u64> d_offsets, d_probe_prefix;
    DevBuf<uint32_t> d_ids, d_probe_cluster, d_cand;
    // batched search: cluster of every row (from the lists), valid for a table of row_cluster_rows rows
    DevBuf<uint32_t> d_row_cluster;
    u64 row_cluster_rows = 0;
    uint32_t build_iters = 0;   // Lloyd iterations the build ran
    double build_ms[4] = {0, 0, 0, 0};  // sample+init, lloyd, final assign, total
    // host maps row -> (cluster, index inside the cluster's list), for the tie replay of batched searches; built on first use
    std::vector<uint32_t> h_row_cluster, h_row_listpos;
    u64 h_row_maps_rows = 0;
};

void csr_from_assign(const uint32_t *assign, u64 n, uint32_t n_clusters, std::vector<u64> &offsets,
                     std::vector<uint32_t> &ids) {
    // src/ivf/index.rs:202-206: per-cluster row ids ascending
    offsets.assign((size_t)n_clusters + 1, 0);
    for (u64 i = 0; i < n; ++i) offsets[assign[i] + 1]++;
    for (uint32_t c = 0; c < n_clusters; ++c) offsets[c + 1] += offsets[c];
    ids.resize(n);
    std::vector<u64> cur(offsets.begin(), offsets.end() - 1);
    for (u64 i = 0; i < n; ++i) ids[cur[assign[i]]++] = (uint32_t)i;
}

// Inverted lists built on the device (csr_*_kernel, pqv_kernels.cuh).  *ok = false (nothing launched) when the cluster
// count does not fit the kernels' shared-memory histogram; the caller then builds the lists on the host.
constexpr uint32_t CSR_MAX_C = 51200;  // 200 KiB of u32 counters

int csr_device(DeviceState &D, const uint32_t *d_assign, u64 n, uint32_t C, u64 *d_offsets, uint32_t *d_ids, bool *ok) {
    *ok = false;
    static const bool force_host = [] {  // PQV_CSR=host: build the lists on the host (the path clusters > CSR_MAX_C take)
        const char *e = getenv("PQV_CSR");
        return e && !strcmp(e, "host");
    }();
    if (force_host || C > CSR_MAX_C || n == 0 || n > 0xFFFFFFFFull) return PQV_OK;
    u64 R = (n + (u64)D.sm_count * 4 - 1) / ((u64)D.sm_count * 4);
    R = std::min<u64>(std::max<u64>((R + 255) / 256 * 256, 256), 4096);
    const u64 nb_max = std::max<u64>((64ull << 20) / C, 1);  // counts matrix <= 256 MiB
    if ((n + R - 1) / R > nb_max) R = ((n + nb_max - 1) / nb_max + 255) / 256 * 256;
    const uint32_t NB = (uint32_t)((n + R - 1) / R);
    PQV_TRY(D.csr_counts.ensure((size_t)C * NB));
    PQV_TRY(D.csr_totals.ensure(C));
    const size_t smem = (size_t)C * 4;
    if (smem > 48 * 1024) {
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(pqv::csr_count_kernel), smem));
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(pqv::csr_scatter_kernel), smem));
    }
    pqv::csr_count_kernel<<<NB, 256, smem, D.stream>>>(d_assign, n, (uint32_t)R, C, NB, D.csr_counts.p);
    pqv::csr_scan_kernel<<<C, 256, 0, D.stream>>>(D.csr_counts.p, NB, D.csr_totals.p);
    pqv::csr_offsets_kernel<<<1, 1024, 0, D.stream>>>(D.csr_totals.p, C, d_offsets);
    pqv::csr_scatter_kernel<<<NB, 256, smem, D.stream>>>(d_assign, n, (uint32_t)R, C, NB, D.csr_counts.p, d_offsets, d_ids);
    CU_TRY(cudaGetLastError());
    *ok = true;
    return PQV_OK;
}

int assign_device(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, const float *d_centroids, uint32_t C,
                  uint32_t *d_out, const ShadowView *sv = nullptr) {
    return assign_dispatch(D, d_rows, n, dim, d_centroids, C, d_out, nullptr, false, sv);  // tcgen05 filter or exact SIMT, pqv_tc_host.cuh
}

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

IvfIndex *find_index(pqv_ctx *ctx, u64 h) {
    auto it = ctx->indexes.find(h);
    return it == ctx->indexes.end() ? nullptr : static_cast<IvfIndex *>(it->second);
}

// An index may only be paired with a table that holds every row its lists name: the gather kernels read
// data + id * dim without further checks, and a stale or foreign blob (file rewritten, index of another table) must fail
// as cleanly as the reference does, not with an illegal address.
int check_index_fits(const IvfIndex &ix, const Dataset &ds) {
    if (ix.n_ids > ds.n_rows)
        return fail(PQV_EINVAL, "index lists hold %llu rows, dataset has %llu", (unsigned long long)ix.n_ids, (unsigned long long)ds.n_rows);
    if (ix.n_ids && (u64)ix.max_id >= ds.n_rows)
        return fail(PQV_EINVAL, "index lists name row %u, dataset has %llu rows", ix.max_id, (unsigned long long)ds.n_rows);
    return PQV_OK;
}

int index_make_resident(DeviceState &D, IvfIndex &ix) {
    if (ix.resident) return PQV_OK;
    PQV_TRY(ix.d_centroids.ensure(ix.centroids.size()));
    PQV_TRY(ix.d_cdist.ensure(ix.n_clusters));
    PQV_TRY(ix.d_probe_cluster.ensure(ix.n_clusters));
    PQV_TRY(ix.d_probe_prefix.ensure((size_t)ix.n_clusters + 1));
    PQV_TRY(ix.d_cand.ensure(std::max<size_t>(ix.n_ids, 1)));
    CU_TRY(cudaMemcpyAsync(ix.d_centroids.p, ix.centroids.data(), ix.centroids.size() * 4, cudaMemcpyHostToDevice, D.stream));
    if (!ix.lists_on_device) {
        PQV_TRY(ix.d_offsets.ensure(ix.offsets.size()));
        PQV_TRY(ix.d_ids.ensure(std::max<size_t>(ix.n_ids, 1)));
        CU_TRY(cudaMemcpyAsync(ix.d_offsets.p, ix.offsets.data(), ix.offsets.size() * 8, cudaMemcpyHostToDevice, D.stream));
        if (ix.n_ids)
            CU_TRY(cudaMemcpyAsync(ix.d_ids.p, ix.ids.data(), ix.n_ids * 4, cudaMemcpyHostToDevice, D.stream));
        ix.lists_on_device = true;
    }
    CU_TRY(cudaStreamSynchronize(D.stream));
    ix.resident = true;
    return PQV_OK;
}

// host copy of the lists of a device-built index (blob serialisation, candidate_rows, the host-ranked search)
int index_host_ids(DeviceState &D, IvfIndex &ix) {
    if (ix.host_ids) return PQV_OK;
    ix.ids.resize(ix.n_ids);
    if (ix.n_ids) {
        CU_TRY(cudaMemcpyAsync(ix.ids.data(), ix.d_ids.p, ix.n_ids * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
    }
    ix.host_ids = true;
    return PQV_OK;
}

void index_free(IvfIndex *ix) {
    ix->d_centroids.release();
    ix->d_cdist.release();
    ix->d_offsets.release();
    ix->d_probe_prefix.release();
    ix->d_ids.release();
    ix->d_probe_cluster.release();
    ix->d_cand.release();
    ix->d_row_cluster.release();
    delete ix;
}

// find_closest_centroids (index.rs:130-149) on a resident index; host gets the ranked cluster ids
int rank_clusters(DeviceState &D, IvfIndex &ix, const float *query, uint32_t nprobe, std::vector<uint32_t> &ranked) {
    PQV_TRY(D.d_query.ensure(ix.dim));
    PQV_TRY(D.h_query.ensure(ix.dim));
    memcpy(D.h_query.p, query, (size_t)ix.dim * 4);
    CU_TRY(cudaMemcpyAsync(D.d_query.p, D.h_query.p, (size_t)ix.dim * 4, cudaMemcpyHostToDevice, D.stream));
    PQV_TRY(dist_launch(D, ix.d_centroids.p, nullptr, ix.n_clusters, ix.dim, D.d_query.p, ix.d_cdist.p, 0));
    std::vector<float> dist(ix.n_clusters);
    CU_TRY(cudaMemcpyAsync(dist.data(), ix.d_cdist.p, (size_t)ix.n_clusters * 4, cudaMemcpyDeviceToHost, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    std::vector<uint32_t> idx(ix.n_clusters);
    for (uint32_t i = 0; i < ix.n_clusters; ++i) idx[i] = i;
    merge_sort_stable(idx, [&](uint32_t a, uint32_t b) { return dist[a] < dist[b]; });
    const uint32_t np = std::min(nprobe, ix.n_clusters);
    ranked.assign(idx.begin(), idx.begin() + np);
    return PQV_OK;
}

constexpr uint32_t IVF_RANK_MAX_C = 16384;  // ivf_rank_kernel sorts pow2(C) u64 keys in shared memory (128 KiB)

bool ivf_fused_enabled() {
    static const bool on = [] {
        const char *e = getenv("PQV_IVF_FUSED");
        return !(e && (!strcmp(e, "off") || !strcmp(e, "0")));
    }();
    return on;
}

// TopkBuilder::topk (src/ivf/search.rs:83-142) with one host<->device round trip: centroid distances, ranking
// (index.rs:130-149), list expansion (index.rs:57-63), gathered scan + entrant filter and the entrants' row ids are all
// enqueued back to back; the host only replays the reference heap over the ~1e3 entrants.  *done = false (nothing
// written) when a centroid distance is NaN or the entrant buffer overflowed: the caller then takes the host-ranked path.
// RowOrder != nullptr selects VectorTopKExec's candidate handling (exec.rs:207-245): the first max_candidates candidates
// in rank order, visited in ascending row order, rows whose filter bit is clear dropped before scoring.
struct RowOrder {
    u64 max_candidates = ~0ull;      // ~0: no cap
    const uint8_t *h_mask = nullptr;  // host, ceil(n_rows / 8) bytes, bit r (LSB first) = row r passes the filter; null = all
    u64 candidate_rows = 0;           // out: probed candidates before cap and filter
    u64 rows_scored = 0;              // out
};
// EntrantsOut != nullptr: return the heap-entrant keys (bits(d) << 32 | candidate position), their row ids and the probed
// clusters in rank order instead of replaying the heap (one process per GPU: the replay happens over all ranks' entrants)
struct EntrantsOut {
    std::vector<u64> keys;
    std::vector<uint32_t> rows;
    std::vector<uint32_t> probe;
    u64 n_cand = 0;
};

int ivf_search_fused(pqv_ctx *ctx, Dataset &ds, DeviceState &D, IvfIndex &ix, const float *query, uint32_t k,
                     uint32_t nprobe, uint32_t flags, uint32_t *out_rows, float *out_dist, uint32_t *out_count,
                     bool *done, RowOrder *ro = nullptr, EntrantsOut *eo = nullptr, u64 seq_limit = ~0ull) {
    // seq_limit (rank-order form only): scan just the first seq_limit candidates of the sequence
    *done = false;
    static const bool trace = getenv("PQV_TRACE") != nullptr;
    double tt[8] = {0};
    if (trace) tt[0] = now_ms();
    const uint32_t C = ix.n_clusters, np = std::min(nprobe, C), cp2 = pow2ceil(C);
    const u64 n_bound = ro ? std::min<u64>(ix.n_ids, ro->max_candidates) : std::min<u64>(ix.n_ids, seq_limit);
    const int order = (flags & PQV_SUM_SEQ) ? 1 : 0;
    PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(pqv::ivf_rank_kernel), (size_t)IVF_RANK_MAX_C * 8));
    PQV_TRY(D.d_query.ensure(ix.dim));
    PQV_TRY(D.h_query.ensure(ix.dim));
    PQV_TRY(D.final_topk.ensure(PQV_MAX_K));
    PQV_TRY(D.ent_out.ensure((size_t)(1u << 16) + 1));
    const uint32_t cap = (uint32_t)std::min<size_t>(D.ent_out.cap - 1, 0xFFFFFFF0u);
    PQV_TRY(D.ent_rows.ensure(cap));
    PQV_TRY(D.ivf_info.ensure(4));
    PQV_TRY(D.h_ivf_info.ensure(4));
    PQV_TRY(D.h_ent_out.ensure((size_t)ENT_FIRST_CHUNK + 1));
    PQV_TRY(D.h_ent_rows.ensure(ENT_FIRST_CHUNK));
    if (trace) tt[1] = now_ms();
    memcpy(D.h_query.p, query, (size_t)ix.dim * 4);
    CU_TRY(cudaMemcpyAsync(D.d_query.p, D.h_query.p, (size_t)ix.dim * 4, cudaMemcpyHostToDevice, D.stream));
    // (the info words and the entrant counter are zeroed by ivf_rank_kernel)
    const u64 n_words = (ds.n_rows + 31) / 32;
    const uint32_t bm_blocks = (uint32_t)((n_words + pqv::BM_WORDS_PER_BLOCK - 1) / pqv::BM_WORDS_PER_BLOCK);
    if (ro) {
        PQV_TRY(D.vt_bitmap.ensure(n_words));
        PQV_TRY(D.vt_sums.ensure(bm_blocks));
        CU_TRY(cudaMemsetAsync(D.vt_bitmap.p, 0, n_words * 4, D.stream));
        if (ro->h_mask) {
            PQV_TRY(D.vt_mask.ensure(n_words));
            CU_TRY(cudaMemcpyAsync(D.vt_mask.p, ro->h_mask, (size_t)((ds.n_rows + 7) / 8), cudaMemcpyHostToDevice, D.stream));
        }
    }
    PQV_TRY(dist_launch(D, ix.d_centroids.p, nullptr, C, ix.dim, D.d_query.p, ix.d_cdist.p, 0));
    pqv::ivf_rank_kernel<<<1, 1024, (size_t)cp2 * 8, D.stream>>>(ix.d_cdist.p, C, cp2, np, ix.d_offsets.p, ix.d_probe_cluster.p,
                                                                 ix.d_probe_prefix.p, D.ivf_info.p, D.ent_out.p);
    CU_TRY(cudaGetLastError());
    if (!ro) {
        pqv::ivf_expand_kernel<<<dim3(np, 32), 256, 0, D.stream>>>(ix.d_ids.p, ix.d_offsets.p, ix.d_probe_cluster.p,
                                                                  ix.d_probe_prefix.p, ix.d_cand.p);
        if (seq_limit != ~0ull) pqv::clamp_count_kernel<<<1, 1, 0, D.stream>>>(D.ivf_info.p, seq_limit);
    } else {
        const uint32_t *d_mask = ro->h_mask ? D.vt_mask.p : nullptr;
        pqv::ivf_mark_kernel<<<dim3(np, 8), 256, 0, D.stream>>>(ix.d_ids.p, ix.d_offsets.p, ix.d_probe_cluster.p,
                                                                ix.d_probe_prefix.p, ro->max_candidates, ds.n_rows, D.vt_bitmap.p);
        pqv::bitmap_count_kernel<<<bm_blocks, 256, 0, D.stream>>>(D.vt_bitmap.p, d_mask, n_words, D.vt_sums.p);
        pqv::bitmap_scan_kernel<<<1, 1024, 0, D.stream>>>(D.vt_sums.p, bm_blocks, D.ivf_info.p);
        pqv::bitmap_compact_kernel<<<bm_blocks, 256, 0, D.stream>>>(D.vt_bitmap.p, d_mask, n_words, D.vt_sums.p, ix.d_cand.p);
    }
    CU_TRY(cudaGetLastError());
    if (trace) tt[2] = now_ms();
    ScanGeom g;
    PQV_TRY(enqueue_scan(ctx, D, ds.shards[0].d_data, ix.d_cand.p, n_bound, ds.dim, D.d_query.p, k, order, 0u, nullptr,
                         D.final_topk.p, D.ent_out.p, cap, true, &g, D.ivf_info.p));
    // the kernel leaves count, first keys, their row ids and the info words in page-locked host memory itself
    const uint32_t first = std::min<uint32_t>(ENT_FIRST_CHUNK, cap);
    pqv::ent_rows_kernel<<<32, 256, 0, D.stream>>>(D.ent_out.p, cap, ix.d_cand.p, D.ent_rows.p, D.h_ent_out.p, D.h_ent_rows.p, first,
                                                   D.ivf_info.p, D.h_ivf_info.p);
    CU_TRY(cudaGetLastError());
    if (eo) {
        eo->probe.resize(np);
        CU_TRY(cudaMemcpyAsync(eo->probe.data(), ix.d_probe_cluster.p, (size_t)np * 4, cudaMemcpyDeviceToHost, D.stream));
    }
    if (trace) tt[3] = now_ms();
    CU_TRY(cudaStreamSynchronize(D.stream));
    if (trace) tt[4] = now_ms();
    const u64 n_cand = D.h_ivf_info.p[0], count = D.h_ent_out.p[0];
    if ((uint32_t)D.h_ivf_info.p[1] != 0 || count > cap) return PQV_OK;
    if (ro) {
        ro->candidate_rows = D.h_ivf_info.p[2];
        ro->rows_scored = n_cand;
    }
    std::vector<u64> entrants(count);
    std::vector<uint32_t> erows(count);
    const u64 got = std::min<u64>(count, first);
    memcpy(entrants.data(), D.h_ent_out.p + 1, got * 8);
    memcpy(erows.data(), D.h_ent_rows.p, got * 4);
    if (count > got) {
        CU_TRY(cudaMemcpyAsync(entrants.data() + got, D.ent_out.p + 1 + got, (count - got) * 8, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaMemcpyAsync(erows.data() + got, D.ent_rows.p + got, (count - got) * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
    }
    // a NaN distance among the entrants: only the reference loop over every candidate answers that (topk_one's fallback)
    if (!(flags & PQV_TIES_BY_POSITION) && any_nan_key(entrants.data(), entrants.size())) return PQV_OK;  // *done stays false
    if (eo) {
        eo->keys.swap(entrants);
        eo->rows.swap(erows);
        eo->n_cand = n_cand;
        *done = true;
        return PQV_OK;
    }
    size_t cnt = 0;
    const bool fast = !(flags & PQV_TIES_BY_POSITION) &&
                      topk_without_replay(entrants.data(), count, k, flags, [&](uint32_t i, uint32_t) { return erows[i]; },
                                          out_rows, out_dist, &cnt);
    std::vector<u64> by_pos;  // (position << 32 | entrant index), sorted by position
    if (!fast) {
        by_pos.resize(count);
        for (u64 i = 0; i < count; ++i) by_pos[i] = ((u64)key_pos(entrants[i]) << 32) | i;
        radix_sort_field(by_pos, 32);
    }
    pqv_timing tm{};
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, D.ev[0], D.ev[1]);
    cudaEventElapsedTime(&b, D.ev[1], D.ev[2]);
    tm.scan_ms = a;
    tm.post_ms = b;
    tm.total_ms = a + b;
    tm.scan_bytes = n_cand * (u64)ds.dim * 4;
    tm.launches = ro ? 12 : 8;
    tm.grid = g.grid;
    tm.entrants = (uint32_t)count;
    if (fast) {
        // answered without the heap replay
    } else if (flags & PQV_TIES_BY_POSITION) {
        // the entrant set contains the k smallest keys (every key below the final threshold entered)
        std::vector<u64> pairs(count);
        for (u64 i = 0; i < count; ++i) pairs[i] = (by_pos[i] & 0xFFFFFFFF00000000ull) | erows[(uint32_t)by_pos[i]];
        RowMap row_of;
        row_of.pairs = pairs.data();
        row_of.n_pairs = pairs.size();
        cnt = emit_by_position(entrants, row_of, k, flags, out_rows, out_dist);
    } else {
        cnt = replay_ordered(
            count,
            [&](size_t i) {
                const uint32_t e = (uint32_t)by_pos[i];
                return ReplayItem{key_dist(entrants[e]), erows[e]};
            },
            k, flags, out_rows, out_dist);
    }
    *out_count = (uint32_t)cnt;
    if (trace) {
        tt[5] = now_ms();
        fprintf(stderr, "[pqv trace] ivf_search: setup %.1f us, enqueue-rank %.1f, enqueue-scan %.1f, sync wait %.1f, replay %.1f\n",
                (tt[1] - tt[0]) * 1e3, (tt[2] - tt[1]) * 1e3, (tt[3] - tt[2]) * 1e3, (tt[4] - tt[3]) * 1e3, (tt[5] - tt[4]) * 1e3);
    }
    ctx->last = tm;
    *done = true;
    return PQV_OK;
}

// function-local device / pinned scratch: released on every exit path
template <typename T>
struct ScopedDev : DevBuf<T> {
    ScopedDev() = default;
    ScopedDev(const ScopedDev &) = delete;
    ScopedDev &operator=(const ScopedDev &) = delete;
    ~ScopedDev() { this->release(); }
};
template <typename T>
struct ScopedPin : PinBuf<T> {
    ScopedPin() = default;
    ScopedPin(const ScopedPin &) = delete;
    ScopedPin &operator=(const ScopedPin &) = delete;
    ~ScopedPin() { this->release(); }
};

// k_means (src/ivf/index.rs:323-457) over the ns rows at d_sample (device): k-means++ over an init set, Lloyd until nothing
// changes or max_iters; the centroids are left in D.d_centroids (C x dim).  Shared by pqv_ivf_build and pqv_kmeans_train.
int kmeans_train_device(DeviceState &D, const float *d_sample, u64 ns, uint32_t dim, uint32_t C, uint32_t max_iters, u64 seed,
                        uint32_t sum_workers, uint32_t *iters_out, double *t_init_out) {
    // ---- k_means (index.rs:323-457)
    SplitMix64 rng(seed);  // the reference re-seeds StdRng with the same seed inside k_means (index.rs:327)
    const u64 init_n = std::max<u64>(std::min<u64>(ns, 50000), C);  // index.rs:332
    std::vector<uint32_t> init_idx;
    if (init_n == ns) {
        init_idx.resize(ns);
        for (u64 i = 0; i < ns; ++i) init_idx[i] = (uint32_t)i;
    } else {
        init_idx = sample_indices(rng, ns, init_n);
    }
    PQV_TRY(D.d_centroids.ensure((size_t)C * dim));
    CU_TRY(cudaMemsetAsync(D.d_centroids.p, 0, (size_t)C * dim * 4, D.stream));  // vec![0.0; C*dim] (index.rs:330)
    ScopedDev<uint32_t> d_init;
    ScopedDev<float> d_md;
    PQV_TRY(d_init.ensure(init_n));
    PQV_TRY(d_md.ensure(init_n));
    auto free_tmp = [&]() {
        d_init.release();
        d_md.release();
    };
    cudaError_t ce = cudaMemcpyAsync(d_init.p, init_idx.data(), init_n * 4, cudaMemcpyHostToDevice, D.stream);
    const u64 first_choice = rng.below(init_n);  // index.rs:340
    auto set_centroid = [&](uint32_t i, uint32_t sample_row) {
        return cudaMemcpyAsync(D.d_centroids.p + (size_t)i * dim, d_sample + (size_t)sample_row * dim, (size_t)dim * 4,
                               cudaMemcpyDeviceToDevice, D.stream);
    };
    if (ce == cudaSuccess) ce = set_centroid(0, init_idx[first_choice]);
    if (ce != cudaSuccess) {
        free_tmp();
        return fail(PQV_ECUDA, "k-means init failed: %s", cudaGetErrorString(ce));
    }
    int rc = dist_launch(D, d_sample, d_init.p, init_n, dim, D.d_centroids.p, d_md.p, 0);  // index.rs:344-352
    ScopedPin<float> h_md;  // pinned: the sweep result comes back 1023 times (host loop only, allocated there)
    float *md = nullptr;
    unsigned hw = std::thread::hardware_concurrency();
    const u64 workers = std::max<u64>(1, std::min<u64>(sum_workers ? sum_workers : (hw ? hw : 1), init_n));  // index.rs:259-265
    const u64 chunk = (init_n + workers - 1) / workers;
    const u64 n_chunks = (init_n + chunk - 1) / chunk;
    std::vector<float> local(n_chunks);
    static const bool trace_pp = getenv("PQV_TRACE") != nullptr;
    double tp[4] = {0, 0, 0, 0}, t_a = 0, t_b = 0, t_c = 0;
    // The picks on the device (pqv_kmeanspp.cuh): sweep and pick kernels alternate on the stream, the random stream lives in
    // device memory, nothing comes back to the host until the last centroid is set.  PQV_KMEANSPP=host keeps the host loop
    // below (also taken by init sets that do not fit one CTA's shared memory: more than 50 000 clusters).
    static const bool pp_host = getenv("PQV_KMEANSPP") && !strcmp(getenv("PQV_KMEANSPP"), "host");
    bool on_device = false;
    if (!rc && !pp_host && init_n <= pqv::kpp::MAX_ROWS && C > 1) {
        ScopedDev<unsigned long long> d_rng;
        ScopedDev<uint32_t> d_dbg;
        rc = d_rng.ensure(1);
        if (!rc && trace_pp) rc = d_dbg.ensure(10 * (size_t)C);
        const size_t pick_smem = (size_t)((init_n + 3) & ~(u64)3) * 4;
        auto pick_kern = pqv::kpp::kmeanspp_pick_kernel<4>;
        if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void *>(pick_kern), pick_smem);
        const unsigned long long state = rng.s;
        if (!rc) {
            ce = cudaMemcpyAsync(d_rng.p, &state, 8, cudaMemcpyHostToDevice, D.stream);
            if (ce != cudaSuccess) rc = fail(PQV_ECUDA, "k-means++ random state upload failed: %s", cudaGetErrorString(ce));
        }
        if (trace_pp) t_a = now_ms();
        // the init set is swept C - 1 times: keep as much of it in L2 as stays there (measured: ~55 % of the L2 size; PQV_SWEEP_KEEP
        // = percent of the set overrides, PQV_SWEEP_KEEP=off takes the general sweep kernel)
        SweepTune tune;
        tune.repeated = true;
        {
            int l2_bytes = 0;
            cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, D.dev);
            const u64 n_groups = (init_n + 31) / 32, group_bytes = (u64)32 * dim * 4;
            tune.keep_groups = std::min<u64>(n_groups, (u64)(0.55 * (double)l2_bytes) / group_bytes);
            if (const char *e = getenv("PQV_SWEEP_KEEP")) {
                if (!strcmp(e, "off")) tune.repeated = false;
                else tune.keep_groups = std::min<u64>(n_groups, (u64)(atof(e) * 0.01 * (double)n_groups));
            }
        }
        for (uint32_t i = 1; i < C && !rc; ++i) {
            rc = dist_launch(D, d_sample, d_init.p, init_n, dim, D.d_centroids.p + (size_t)(i - 1) * dim, d_md.p, 1, nullptr, &tune);
            if (rc) break;
            pick_kern<<<1, pqv::kpp::THREADS, pick_smem, D.stream>>>(
                d_md.p, (uint32_t)init_n, (uint32_t)chunk, (uint32_t)n_chunks, d_rng.p, d_init.p, d_sample, dim,
                D.d_centroids.p + (size_t)i * dim, trace_pp ? d_dbg.p + 10 * (size_t)i : nullptr, 1.0f);
            ce = cudaGetLastError();
            if (ce != cudaSuccess) rc = fail(PQV_ECUDA, "k-means++ pick launch failed: %s", cudaGetErrorString(ce));
        }
        if (!rc) {
            ce = cudaStreamSynchronize(D.stream);
            if (ce != cudaSuccess) rc = fail(PQV_ECUDA, "k-means++ on the device failed: %s", cudaGetErrorString(ce));
        }
        if (tune.repeated) cudaCtxResetPersistingL2Cache();  // the init set's evict_last lines have no claim on the L2 any more
        if (trace_pp) {
            fprintf(stderr, "[pqv trace] k-means++ on the device (%u picks over %llu rows, %llu chunk sums): %.1f ms\n", C - 1,
                    (unsigned long long)init_n, (unsigned long long)n_chunks, now_ms() - t_a);
            std::vector<uint32_t> dbg(10 * (size_t)C, 0);
            if (!rc && cudaMemcpy(dbg.data(), d_dbg.p, dbg.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
                double ph[6] = {0, 0, 0, 0, 0, 0}, iters = 0, pos = 0, chain = 0, pro = 0;
                for (uint32_t i = 1; i < C; ++i) {
                    for (int j = 1; j <= 5; ++j) ph[j] += dbg[10 * (size_t)i + j];
                    iters += dbg[10 * (size_t)i + 6];
                    chain += dbg[10 * (size_t)i + 7];
                    pro += dbg[10 * (size_t)i + 8];
                    pos += dbg[10 * (size_t)i] == 0xFFFFFFFFu ? 0 : dbg[10 * (size_t)i];
                }
                const double inv = 1.0 / (C - 1);
                fprintf(stderr, "[pqv trace] pick kernel, mean cycles: load %.0f, chunk sums %.0f (first chain %.0f, prologue sums %.0f), "
                                "draw + prologue %.0f, walk %.0f (%.1f blocks, mean pick position %.0f), row copy %.0f\n", ph[1] * inv,
                        (ph[2] - ph[1]) * inv, (chain - ph[1]) * inv, (pro - ph[1]) * inv, (ph[3] - ph[2]) * inv,
                        (ph[4] - ph[3]) * inv, iters * inv, pos * inv, (ph[5] - ph[4]) * inv);
            }
            d_dbg.release();
        }
        d_rng.release();
        on_device = true;
    }
    if (!rc && !on_device && C > 1) {
        rc = h_md.ensure(init_n);
        md = h_md.p;
    }
    for (uint32_t i = 1; i < C && !rc && !on_device; ++i) {
        if (trace_pp) t_a = now_ms();
        // the sweep writes its result to d_md and, in the same pass, to the page-locked host buffer md
        rc = dist_launch(D, d_sample, d_init.p, init_n, dim, D.d_centroids.p + (size_t)(i - 1) * dim, d_md.p, 1, md);
        if (rc) break;
        ce = cudaStreamSynchronize(D.stream);
        if (trace_pp) t_b = now_ms();
        if (ce != cudaSuccess) {
            rc = fail(PQV_ECUDA, "k-means++ sweep failed: %s", cudaGetErrorString(ce));
            break;
        }
        // index.rs:356-370: one serial f32 sum per worker chunk, then the chunk sums added in chunk order.  The chunk
        // chains are independent, so they are advanced side by side (same bits, ~n_chunks x the instruction-level
        // parallelism of walking them one after the other).
        std::fill(local.begin(), local.end(), 0.0f);
        const u64 full_chunks = init_n / chunk;  // chunks with all `chunk` elements
        for (u64 o = 0; o < chunk; ++o) {
            const float *col = md + o;
            for (u64 c = 0; c < full_chunks; ++c) local[c] += col[c * chunk];
        }
        for (u64 s = full_chunks * chunk; s < init_n; ++s) local[full_chunks] += md[s];
        float total = 0.0f;
        for (u64 c = 0; c < n_chunks; ++c) total += local[c];
        if (trace_pp) t_c = now_ms();
        if (total > 0.0f) {  // index.rs:372-383
            const float threshold = rng.unit_f32() * total;
            float cumsum = 0.0f;
            for (u64 s = 0; s < init_n; ++s) {
                cumsum += md[s];
                if (cumsum >= threshold) {
                    ce = set_centroid(i, init_idx[s]);
                    break;
                }
            }
        } else {  // index.rs:384-389
            ce = set_centroid(i, init_idx[rng.below(init_n)]);
        }
        if (ce != cudaSuccess) rc = fail(PQV_ECUDA, "k-means++ pick failed: %s", cudaGetErrorString(ce));
        if (trace_pp) {
            const double t_d = now_ms();
            tp[0] += t_b - t_a;
            tp[1] += t_c - t_b;
            tp[2] += t_d - t_c;
        }
    }
    if (trace_pp && !on_device)
        fprintf(stderr, "[pqv trace] k-means++ (%u picks over %llu rows): sweep+readback %.1f ms, chunk sums %.1f ms, cumsum pick %.1f ms\n",
                C - 1, (unsigned long long)init_n, tp[0], tp[1], tp[2]);
    h_md.release();
    free_tmp();
    if (rc) return rc;
    if (t_init_out) *t_init_out = now_ms();

    // ---- Lloyd (index.rs:392-454).  Assignments, the changed count, the member lists and the centroid update all stay
    // on the device; one 8-byte read per iteration decides whether to stop.
    ScopedDev<uint32_t> d_asg[2], d_mids;
    ScopedDev<u64> d_moff, d_changed;
    ScopedPin<u64> h_changed;
    auto free_lloyd = [&]() {
        d_asg[0].release();
        d_asg[1].release();
        d_mids.release();
        d_moff.release();
        d_changed.release();
        h_changed.release();
    };
    rc = d_asg[0].ensure(ns);
    if (!rc) rc = d_asg[1].ensure(ns);
    if (!rc) rc = d_mids.ensure(ns);
    if (!rc) rc = d_moff.ensure((size_t)C + 1);
    if (!rc) rc = d_changed.ensure(1);
    if (!rc) rc = h_changed.ensure(1);
    if (!rc) {
        ce = cudaMemsetAsync(d_asg[0].p, 0, ns * 4, D.stream);  // vec![0usize; n] (index.rs:392)
        if (ce != cudaSuccess) rc = fail(PQV_ECUDA, "Lloyd init failed: %s", cudaGetErrorString(ce));
    }
    // the 20 sweeps read the same sample: its 16-bit shadow (operand of the tensor-core filter) is built once
    ShadowView tsv;
    bool have_tsv = false;
    if (!rc && shadow_layout_ok(dim, d_sample) && assign_path_for(d_sample, ns, dim, D.d_centroids.p, C) == ASSIGN_TC) {
        rc = temp_shadow(D, d_sample, ns, dim, &tsv);
        have_tsv = !rc;
    }
    int cur = 0;  // d_asg[cur] = assignments of the previous iteration
    std::vector<uint32_t> h_assign;  // host fallback only (cluster count beyond the device list builder)
    std::vector<u64> moff;
    std::vector<uint32_t> mids;
    for (uint32_t iter = 0; iter < max_iters && !rc; ++iter) {
        if (iters_out) *iters_out = iter + 1;
        uint32_t *prev = d_asg[cur].p, *next = d_asg[cur ^ 1].p;
        rc = assign_device(D, d_sample, ns, dim, D.d_centroids.p, C, next, have_tsv ? &tsv : nullptr);
        if (rc) break;
        ce = cudaMemsetAsync(d_changed.p, 0, 8, D.stream);
        if (ce == cudaSuccess) {
            pqv::count_changed_kernel<<<D.sm_count * 2, 256, 0, D.stream>>>(prev, next, ns,
                                                                           reinterpret_cast<unsigned long long *>(d_changed.p));
            ce = cudaGetLastError();
        }
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_changed.p, d_changed.p, 8, cudaMemcpyDeviceToHost, D.stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(D.stream);
        if (ce != cudaSuccess) {
            rc = fail(PQV_ECUDA, "Lloyd assignment failed: %s", cudaGetErrorString(ce));
            break;
        }
        cur ^= 1;
        if (h_changed.p[0] == 0) break;  // index.rs:432-434
        bool on_device = false;
        rc = csr_device(D, next, ns, C, d_moff.p, d_mids.p, &on_device);
        if (rc) break;
        if (!on_device) {
            h_assign.resize(ns);
            ce = cudaMemcpyAsync(h_assign.data(), next, ns * 4, cudaMemcpyDeviceToHost, D.stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(D.stream);
            if (ce == cudaSuccess) {
                csr_from_assign(h_assign.data(), ns, C, moff, mids);
                ce = cudaMemcpyAsync(d_mids.p, mids.data(), ns * 4, cudaMemcpyHostToDevice, D.stream);
            }
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_moff.p, moff.data(), ((size_t)C + 1) * 8, cudaMemcpyHostToDevice, D.stream);
            if (ce != cudaSuccess) {
                rc = fail(PQV_ECUDA, "Lloyd update upload failed: %s", cudaGetErrorString(ce));
                break;
            }
        }
        pqv::centroid_update_kernel<<<C, 256, 0, D.stream>>>(d_sample, dim, d_mids.p, d_moff.p, D.d_centroids.p);
        ce = cudaGetLastError();
        if (ce == cudaSuccess && !on_device) ce = cudaStreamSynchronize(D.stream);  // mids/moff are reused next iteration
        if (ce != cudaSuccess) {
            rc = fail(PQV_ECUDA, "centroid update failed: %s", cudaGetErrorString(ce));
            break;
        }
    }
    if (!rc) {
        ce = cudaStreamSynchronize(D.stream);
        if (ce != cudaSuccess) rc = fail(PQV_ECUDA, "Lloyd failed: %s", cudaGetErrorString(ce));
    }
    free_lloyd();
    if (rc) return rc;

    return PQV_OK;
}

}  // namespace

extern "C" {

int pqv_ivf_build(pqv_ctx *ctx, uint64_t handle, uint32_t n_clusters_or_0, uint32_t max_iters, uint64_t seed,
                  uint32_t sum_workers, uint64_t *out_index) {
    if (!ctx || !out_index) return fail(PQV_EINVAL, "null argument");
    if (max_iters == 0) return fail(PQV_EINVAL, "max_iters must be > 0");  // src/ivf/parquet.rs:89-91
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    const bool multi = ds->shards.size() > 1;  // table spread over several devices of this context
    const u64 n = ds->n_rows;
    const uint32_t dim = ds->dim;
    if (n == 0) return fail(PQV_EINVAL, "Cannot build IVF index with zero vectors");  // index.rs:157-159
    const u64 C64 = n_clusters_or_0 ? n_clusters_or_0 : (u64)std::ceil(std::sqrt((double)n));  // index.rs:161-167
    if (C64 > n) return fail(PQV_EINVAL, "n_clusters cannot exceed number of vectors");          // index.rs:168-170
    const uint32_t C = (uint32_t)C64;
    u64 sample_size = std::max<u64>(n / 20, 1);                                                   // index.rs:172-174
    sample_size = std::min<u64>(sample_size, 100000);
    sample_size = std::min<u64>(std::max<u64>(sample_size, C), n);

    Shard &sh = ds->shards[0];
    DeviceState &D = ctx->devs[sh.di];
    DevGuard guard(D.dev);
    const double t_begin = now_ms();
    IvfIndex *ix = new IvfIndex();
    ix->dim = dim;
    ix->n_clusters = C;
    auto bail = [&](int rc) {
        index_free(ix);
        return rc;
    };
#define IVF_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__) return bail(rc__);  \
    } while (0)
#define IVF_CU(expr)                                                                                              \
    do {                                                                                                          \
        cudaError_t e__ = (expr);                                                                                 \
        if (e__ != cudaSuccess) return bail(fail(PQV_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)));    \
    } while (0)

    // ---- training sample (index.rs:182-187, 222-242)
    const float *d_sample = sh.d_data;
    const u64 ns = sample_size;
    if (multi) {
        // the sample (or, for a tiny table, every row) is collected on the first device; training runs there
        std::vector<uint32_t> sidx;
        if (sample_size != n) {
            SplitMix64 rng_s(seed);
            sidx = sample_indices(rng_s, n, sample_size);
        } else {
            sidx.resize(n);
            for (u64 i = 0; i < n; ++i) sidx[i] = (uint32_t)i;
        }
        IVF_TRY(D.d_tmp_rows.ensure((size_t)ns * dim));
        IVF_TRY(gather_rows_multi(ctx, *ds, sidx.data(), ns, D, D.d_tmp_rows.p));
        d_sample = D.d_tmp_rows.p;
    } else if (sample_size != n) {
        SplitMix64 rng_s(seed);
        std::vector<uint32_t> sidx = sample_indices(rng_s, n, sample_size);
        IVF_TRY(D.d_row_ids.ensure(ns));
        IVF_TRY(D.d_tmp_rows.ensure((size_t)ns * dim));
        IVF_CU(cudaMemcpyAsync(D.d_row_ids.p, sidx.data(), ns * 4, cudaMemcpyHostToDevice, D.stream));
        pqv::gather_rows_kernel<<<D.sm_count * 8, 256, 0, D.stream>>>(sh.d_data, D.d_row_ids.p, ns, dim, D.d_tmp_rows.p);
        IVF_CU(cudaGetLastError());
        IVF_CU(cudaStreamSynchronize(D.stream));
        d_sample = D.d_tmp_rows.p;
    }

    // ---- k_means (index.rs:323-457)
    double t_init = t_begin;
    IVF_TRY(kmeans_train_device(D, d_sample, ns, dim, C, max_iters, seed, sum_workers, &ix->build_iters, &t_init));
    const double t_lloyd = now_ms();

    // ---- final assignment of all N rows (index.rs:189-206); the inverted lists are built on the device and stay
    // there as the resident index, the host copy of the row ids is fetched when somebody asks for it (index_host_ids)
    ix->centroids.resize((size_t)C * dim);
    IVF_CU(cudaMemcpyAsync(ix->centroids.data(), D.d_centroids.p, (size_t)C * dim * 4, cudaMemcpyDeviceToHost, D.stream));
    static const bool trace = getenv("PQV_TRACE") != nullptr;
    double tf[4] = {now_ms(), 0, 0, 0};
    IVF_TRY(D.d_assign.ensure(n));
    if (multi) {
        // every shard assigns its own rows against its own copy of the centroids (all sweeps enqueued before the first wait);
        // the assignments meet on the first device, where the lists are built
        IVF_CU(cudaStreamSynchronize(D.stream));  // centroids final
        for (Shard &s2 : ds->shards) {
            if (!s2.n_rows) continue;
            DeviceState &D2 = ctx->devs[s2.di];
            DevGuard g2(D2.dev);
            if (&D2 != &D) {
                IVF_TRY(D2.d_centroids.ensure((size_t)C * dim));
                IVF_CU(cudaMemcpyPeerAsync(D2.d_centroids.p, D2.dev, D.d_centroids.p, D.dev, (size_t)C * dim * 4, D2.stream));
                IVF_TRY(D2.d_assign.ensure(s2.n_rows));
            }
            ShadowView fsv;
            bool have_fsv = false, built = false;
            IVF_TRY(sweep_shadow(D2, ds, s2.d_data, s2.n_rows, &fsv, &have_fsv, &built, &s2));
            uint32_t *dst = (&D2 == &D) ? D.d_assign.p + s2.first_row : D2.d_assign.p;
            IVF_TRY(assign_device(D2, s2.d_data, s2.n_rows, dim, D2.d_centroids.p, C, dst, have_fsv ? &fsv : nullptr));
            if (&D2 != &D)
                IVF_CU(cudaMemcpyPeerAsync(D.d_assign.p + s2.first_row, D.dev, D2.d_assign.p, D2.dev, s2.n_rows * 4, D2.stream));
        }
        for (Shard &s2 : ds->shards) {
            DeviceState &D2 = ctx->devs[s2.di];
            DevGuard g2(D2.dev);
            IVF_CU(cudaStreamSynchronize(D2.stream));
        }
    } else {
        // the table's own shadow: stays with the shard for later sweeps and batched searches
        ShadowView fsv;
        bool have_fsv = false, built = false;
        IVF_TRY(sweep_shadow(D, ds, sh.d_data, n, &fsv, &have_fsv, &built));
        IVF_TRY(assign_device(D, sh.d_data, n, dim, D.d_centroids.p, C, D.d_assign.p, have_fsv ? &fsv : nullptr));
    }
    if (trace) {
        cudaStreamSynchronize(D.stream);
        tf[1] = now_ms();
    }
    IVF_TRY(ix->d_offsets.ensure((size_t)C + 1));
    IVF_TRY(ix->d_ids.ensure(n));
    bool lists_built = false;
    IVF_TRY(csr_device(D, D.d_assign.p, n, C, ix->d_offsets.p, ix->d_ids.p, &lists_built));
    if (trace) {
        cudaStreamSynchronize(D.stream);
        tf[2] = now_ms();
        fprintf(stderr, "[pqv trace] ivf_build final: assign %.2f ms (device sweep %.2f), lists %.2f ms\n", tf[1] - tf[0],
                ctx->last_assign.total_ms, tf[2] - tf[1]);
    }
    ix->n_ids = n;
    ix->max_id = (uint32_t)(n - 1);
    if (lists_built) {
        ix->offsets.resize((size_t)C + 1);
        IVF_CU(cudaMemcpyAsync(ix->offsets.data(), ix->d_offsets.p, ((size_t)C + 1) * 8, cudaMemcpyDeviceToHost, D.stream));
        IVF_CU(cudaStreamSynchronize(D.stream));
        ix->lists_on_device = true;
        ix->host_ids = false;
    } else {
        std::vector<uint32_t> full(n);
        IVF_CU(cudaMemcpyAsync(full.data(), D.d_assign.p, n * 4, cudaMemcpyDeviceToHost, D.stream));
        IVF_CU(cudaStreamSynchronize(D.stream));
        csr_from_assign(full.data(), n, C, ix->offsets, ix->ids);
    }
    const double t_end = now_ms();
    ix->build_ms[0] = t_init - t_begin;
    ix->build_ms[1] = t_lloyd - t_init;
    ix->build_ms[2] = t_end - t_lloyd;
    ix->build_ms[3] = t_end - t_begin;
#undef IVF_TRY
#undef IVF_CU
    const u64 h = ctx->next_handle++;
    ctx->indexes[h] = ix;
    *out_index = h;
    return PQV_OK;
}

int pqv_kmeans_train(pqv_ctx *ctx, uint64_t handle, uint32_t n_clusters, uint32_t max_iters, uint64_t seed,
                     uint32_t sum_workers, float *out_centroids, uint32_t *out_iters) {
    if (!ctx || !out_centroids) return fail(PQV_EINVAL, "null argument");
    if (max_iters == 0) return fail(PQV_EINVAL, "max_iters must be > 0");
    if (n_clusters == 0) return fail(PQV_EINVAL, "Cluster count must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (ds->n_rows == 0) return fail(PQV_EINVAL, "Cannot build IVF index with zero vectors");
    if (n_clusters > ds->n_rows) return fail(PQV_EINVAL, "n_clusters cannot exceed number of vectors");
    Shard &sh = ds->shards[0];
    DeviceState &D = ctx->devs[sh.di];
    DevGuard guard(D.dev);
    uint32_t iters = 0;
    const float *d_rows = sh.d_data;
    if (ds->shards.size() > 1) {
        // a table spread over several devices: k_means sees ALL rows in table order -- they are collected on the first device
        // (this entry point is the reference's k_means over a training set, which pqv_ivf_build keeps to <= 100 000 rows)
        if (ds->n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row ids are u32");
        std::vector<uint32_t> all(ds->n_rows);
        for (u64 i = 0; i < ds->n_rows; ++i) all[i] = (uint32_t)i;
        PQV_TRY(D.d_tmp_rows.ensure((size_t)ds->n_rows * ds->dim));
        PQV_TRY(gather_rows_multi(ctx, *ds, all.data(), ds->n_rows, D, D.d_tmp_rows.p));
        d_rows = D.d_tmp_rows.p;
    }
    PQV_TRY(kmeans_train_device(D, d_rows, ds->n_rows, ds->dim, n_clusters, max_iters, seed, sum_workers, &iters, nullptr));
    CU_TRY(cudaMemcpyAsync(out_centroids, D.d_centroids.p, (size_t)n_clusters * ds->dim * 4, cudaMemcpyDeviceToHost, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    if (out_iters) *out_iters = iters;
    return PQV_OK;
}

int pqv_ivf_sample_rows(uint64_t n_rows, uint32_t n_clusters_or_0, uint64_t seed, uint32_t *out_rows, uint64_t cap,
                        uint64_t *out_n, uint32_t *out_clusters) {
    if (!out_n || !out_clusters) return fail(PQV_EINVAL, "null argument");
    if (n_rows == 0) return fail(PQV_EINVAL, "Cannot build IVF index with zero vectors");  // index.rs:157-159
    if (n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row ids are u32");
    const u64 C64 = n_clusters_or_0 ? n_clusters_or_0 : (u64)std::ceil(std::sqrt((double)n_rows));  // index.rs:161-167
    if (C64 > n_rows) return fail(PQV_EINVAL, "n_clusters cannot exceed number of vectors");
    u64 sample_size = std::max<u64>(n_rows / 20, 1);  // index.rs:172-174
    sample_size = std::min<u64>(sample_size, 100000);
    sample_size = std::min<u64>(std::max<u64>(sample_size, C64), n_rows);
    *out_n = sample_size;
    *out_clusters = (uint32_t)C64;
    if (sample_size > cap || !out_rows) return out_rows ? fail(PQV_ELIMIT, "%llu sample rows do not fit the caller's buffer", (unsigned long long)sample_size) : PQV_OK;
    if (sample_size == n_rows) {  // index.rs:182-183: the whole table, in order
        for (u64 i = 0; i < n_rows; ++i) out_rows[i] = (uint32_t)i;
    } else {
        SplitMix64 rng(seed);
        const std::vector<uint32_t> idx = sample_indices(rng, n_rows, sample_size);
        memcpy(out_rows, idx.data(), sample_size * 4);
    }
    return PQV_OK;
}

int pqv_ivf_build_stats(pqv_ctx *ctx, uint64_t index, uint32_t *out_lloyd_iters, double *out_ms4) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    if (out_lloyd_iters) *out_lloyd_iters = ix->build_iters;
    if (out_ms4) memcpy(out_ms4, ix->build_ms, sizeof ix->build_ms);
    return PQV_OK;
}

int pqv_ivf_from_bytes(pqv_ctx *ctx, const uint8_t *bytes, uint64_t len, uint64_t *out_index) {
    if (!ctx || !out_index || (!bytes && len)) return fail(PQV_EINVAL, "null argument");
    if (len < 8) return fail(PQV_EINVAL, "IVF index buffer too small");  // index.rs:88-90
    uint32_t dim, C;
    memcpy(&dim, bytes, 4);
    memcpy(&C, bytes + 4, 4);
    if (dim == 0) return fail(PQV_EINVAL, "Embedding dimension must be > 0");
    if (C == 0) return fail(PQV_EINVAL, "Cluster count must be > 0");
    u64 off = 8;
    const u64 cbytes = (u64)dim * C * 4;
    if (off + cbytes > len) return fail(PQV_EINVAL, "IVF index buffer truncated in centroids");
    IvfIndex *ix = new IvfIndex();
    ix->dim = dim;
    ix->n_clusters = C;
    ix->centroids.resize((size_t)dim * C);
    memcpy(ix->centroids.data(), bytes + off, cbytes);
    off += cbytes;
    ix->offsets.assign((size_t)C + 1, 0);
    for (uint32_t c = 0; c < C; ++c) {
        uint32_t l;
        if (off + 4 > len) {
            delete ix;
            return fail(PQV_EINVAL, "IVF index buffer truncated in list %u", c);
        }
        memcpy(&l, bytes + off, 4);
        off += 4;
        if (off + (u64)l * 4 > len) {
            delete ix;
            return fail(PQV_EINVAL, "IVF index buffer truncated in list %u", c);
        }
        const size_t base = ix->ids.size();
        ix->ids.resize(base + l);
        if (l) memcpy(ix->ids.data() + base, bytes + off, (size_t)l * 4);
        off += (u64)l * 4;
        ix->offsets[c + 1] = ix->offsets[c] + l;
    }
    ix->n_ids = ix->ids.size();
    for (uint32_t id : ix->ids) ix->max_id = std::max(ix->max_id, id);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const u64 h = ctx->next_handle++;
    ctx->indexes[h] = ix;
    *out_index = h;
    return PQV_OK;
}

int pqv_ivf_to_bytes(pqv_ctx *ctx, uint64_t index, uint8_t *out, uint64_t cap, uint64_t *out_len) {
    if (!ctx || !out_len) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    const u64 need = 8 + (u64)ix->centroids.size() * 4 + (u64)ix->n_clusters * 4 + ix->n_ids * 4;
    *out_len = need;
    if (!out || cap < need) return out ? fail(PQV_ELIMIT, "blob needs %llu bytes", (unsigned long long)need) : PQV_OK;
    {
        DeviceState &D = ctx->devs[0];
        DevGuard guard(D.dev);
        PQV_TRY(index_host_ids(D, *ix));
    }
    uint8_t *p = out;  // index.rs:65-83
    memcpy(p, &ix->dim, 4);
    p += 4;
    memcpy(p, &ix->n_clusters, 4);
    p += 4;
    memcpy(p, ix->centroids.data(), ix->centroids.size() * 4);
    p += ix->centroids.size() * 4;
    for (uint32_t c = 0; c < ix->n_clusters; ++c) {
        const uint32_t l = (uint32_t)(ix->offsets[c + 1] - ix->offsets[c]);
        memcpy(p, &l, 4);
        p += 4;
        if (l) memcpy(p, ix->ids.data() + ix->offsets[c], (size_t)l * 4);
        p += (size_t)l * 4;
    }
    return PQV_OK;
}

int pqv_ivf_info(pqv_ctx *ctx, uint64_t index, uint32_t *out_dim, uint32_t *out_clusters, uint64_t *out_ids) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    if (out_dim) *out_dim = ix->dim;
    if (out_clusters) *out_clusters = ix->n_clusters;
    if (out_ids) *out_ids = ix->n_ids;
    return PQV_OK;
}

int pqv_ivf_drop(pqv_ctx *ctx, uint64_t index) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    DevGuard guard(ctx->devs[0].dev);
    cudaStreamSynchronize(ctx->devs[0].stream);
    index_free(ix);
    ctx->indexes.erase(index);
    return PQV_OK;
}

int pqv_ivf_candidate_rows(pqv_ctx *ctx, uint64_t index, const float *query, uint32_t nprobe, uint32_t *out_rows,
                           uint64_t cap, uint64_t *out_n) {
    if (!ctx || !query || !out_n) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    PQV_TRY(index_make_resident(D, *ix));
    PQV_TRY(index_host_ids(D, *ix));
    std::vector<uint32_t> ranked;
    PQV_TRY(rank_clusters(D, *ix, query, nprobe, ranked));
    u64 total = 0;
    for (uint32_t c : ranked) total += ix->offsets[c + 1] - ix->offsets[c];
    *out_n = total;
    if (total > cap || (total && !out_rows)) return fail(PQV_ELIMIT, "%llu candidate rows do not fit the caller's buffer", (unsigned long long)total);
    u64 o = 0;
    for (uint32_t c : ranked) {  // index.rs:57-63
        const u64 l = ix->offsets[c + 1] - ix->offsets[c];
        if (l) memcpy(out_rows + o, ix->ids.data() + ix->offsets[c], l * 4);
        o += l;
    }
    return PQV_OK;
}

// one query of TopkBuilder::search over a resident table + index (the body of pqv_ivf_search; ctx->mu held)
static int ivf_search_one(pqv_ctx *ctx, Dataset *ds, DeviceState &D, IvfIndex *ix, const float *query, uint32_t k,
                          uint32_t nprobe, uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    PQV_TRY(index_make_resident(D, *ix));
    const bool multi = ds->shards.size() > 1;  // table spread over several devices: ranking here, candidates split by owner (topk_one)
    if (!multi && ivf_fused_enabled() && ix->n_clusters <= IVF_RANK_MAX_C && ix->n_ids && k <= PQV_MAX_K) {  // (larger k: topk_one's full replay)
        bool done = false;
        PQV_TRY(ivf_search_fused(ctx, *ds, D, *ix, query, k, nprobe, flags, out_row_idx, out_dist, out_count, &done));
        if (done) return PQV_OK;  // otherwise: NaN distance or entrant overflow -> host-ranked path below
    }
    PQV_TRY(index_host_ids(D, *ix));
    std::vector<uint32_t> ranked;
    PQV_TRY(rank_clusters(D, *ix, query, nprobe, ranked));
    if (multi) {
        std::vector<uint32_t> rows;
        for (uint32_t c : ranked) rows.insert(rows.end(), ix->ids.begin() + ix->offsets[c], ix->ids.begin() + ix->offsets[c + 1]);
        *out_count = 0;
        if (rows.empty()) return PQV_OK;
        return topk_one(ctx, *ds, query, rows.data(), rows.size(), k, flags, out_row_idx, out_dist, out_count);
    }
    const uint32_t np = (uint32_t)ranked.size();
    std::vector<u64> prefix((size_t)np + 1, 0);
    for (uint32_t r = 0; r < np; ++r) prefix[r + 1] = prefix[r] + (ix->offsets[ranked[r] + 1] - ix->offsets[ranked[r]]);
    const u64 n_cand = prefix[np];
    if (n_cand == 0) {
        *out_count = 0;
        return PQV_OK;
    }
    CU_TRY(cudaMemcpyAsync(ix->d_probe_cluster.p, ranked.data(), (size_t)np * 4, cudaMemcpyHostToDevice, D.stream));
    CU_TRY(cudaMemcpyAsync(ix->d_probe_prefix.p, prefix.data(), ((size_t)np + 1) * 8, cudaMemcpyHostToDevice, D.stream));
    pqv::ivf_expand_kernel<<<np, 256, 0, D.stream>>>(ix->d_ids.p, ix->d_offsets.p, ix->d_probe_cluster.p,
                                                     ix->d_probe_prefix.p, ix->d_cand.p);
    CU_TRY(cudaGetLastError());
    // candidate position -> row id on the host, from the host copy of the lists
    const std::function<uint32_t(uint32_t)> row_fn = [&](uint32_t pos) -> uint32_t {
        const uint32_t r = (uint32_t)(std::upper_bound(prefix.begin(), prefix.end(), (u64)pos) - prefix.begin()) - 1;
        return ix->ids[ix->offsets[ranked[r]] + (pos - prefix[r])];
    };
    return topk_one(ctx, *ds, query, nullptr, n_cand, k, flags, out_row_idx, out_dist, out_count, nullptr, 0,
                    ix->d_cand.p, &row_fn);
}

int pqv_ivf_search(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k, uint32_t nprobe,
                   uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!ctx || !query || !out_row_idx || !out_dist || !out_count) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");  // src/ivf/search.rs:72
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    PQV_TRY(check_topk_args(k, ds->dim, flags, true));
    if (ix->dim != ds->dim) return fail(PQV_EINVAL, "Query dimension mismatch: expected %u, got %u", ix->dim, ds->dim);
    PQV_TRY(check_index_fits(*ix, *ds));
    DeviceState &D = ctx->devs[ds->shards[0].di];
    DevGuard guard(D.dev);
    return ivf_search_one(ctx, ds, D, ix, query, k, nprobe, flags, out_row_idx, out_dist, out_count);
}

// the masked batched pass shared by pqv_ivf_search_batch (final results) and pqv_ivf_search_batch_keys (raw k + 1 keys per
// query for the sharded merge).  handled[q] = 1 where the pass decided query q.  ctx->mu held, device selected.
static int ivf_batch_masked(pqv_ctx *ctx, Dataset *ds, DeviceState &D, IvfIndex *ix, const float *queries, uint32_t n_queries,
                            uint32_t k, uint32_t nprobe, uint32_t flags, uint32_t *out_row_idx, float *out_dist,
                            uint32_t *out_count, std::vector<uint8_t> &handled, u64 *raw_keys, uint32_t *raw_count,
                            uint32_t pos_base, const uint8_t *h_row_mask = nullptr, BatchTieOut *tie_out = nullptr) {
    const uint32_t C = ix->n_clusters, dim = ds->dim, np = std::min(nprobe, C);
    handled.assign(n_queries, 0);
    ctx->last_batch = pqv_batch_timing{};
    ctx->batch_state.valid = false;
    uint32_t cp2 = 32;
    while (cp2 < C) cp2 <<= 1;
    const bool can_batch = n_queries && ix->n_ids && batch_path_applies(*ds, ds->shards[0].d_data, n_queries, k) &&
                           !(flags & PQV_TIES_BY_POSITION) && (size_t)cp2 * 8 <= 128 * 1024 && (size_t)n_queries * np < (1ull << 31);
    if (can_batch) {
        namespace T = pqv::tc;
        const uint32_t nq = n_queries;
        const uint32_t nq_pad = (nq + T::BN - 1) / T::BN * T::BN, qwords = nq_pad / 32;
        // row -> cluster, once per (index, table size)
        if (ix->row_cluster_rows != ds->n_rows) {
            PQV_TRY(ix->d_row_cluster.ensure(ds->n_rows));
            CU_TRY(cudaMemsetAsync(ix->d_row_cluster.p, 0xFF, (size_t)ds->n_rows * 4, D.stream));
            pqv::row_cluster_kernel<<<dim3(C, 8), 256, 0, D.stream>>>(ix->d_ids.p, ix->d_offsets.p, ds->n_rows, ix->d_row_cluster.p);
            CU_TRY(cudaGetLastError());
            ix->row_cluster_rows = ds->n_rows;
        }
        // all rankings in one launch
        const bool vec4 = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(ix->d_centroids.p) & 15) == 0);
        const size_t smem = (size_t)((dim + 3u) & ~3u) * 4 + (size_t)SCAN_WARPS * pqv::TileCfg<0, true>::TILE_FLOATS * 4;
        auto *k_vec = pqv::l2_dist_batch_kernel<true, SCAN_WARPS>;
        auto *k_sca = pqv::l2_dist_batch_kernel<false, SCAN_WARPS>;
        PQV_TRY(ensure_dyn_smem(vec4 ? reinterpret_cast<const void *>(k_vec) : reinterpret_cast<const void *>(k_sca), smem));
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(pqv::rank_batch_kernel), (size_t)cp2 * 8));
        PQV_TRY(D.d_tmp_rows.ensure((size_t)nq * dim));
        PQV_TRY(D.d_dist.ensure((size_t)nq * C));
        PQV_TRY(D.d_assign.ensure((size_t)nq * np));
        PQV_TRY(D.d_row_ids.ensure(nq));
        PQV_TRY(D.vt_mask.ensure((size_t)C * qwords));
        std::vector<uint32_t> nan_flags(nq);
        const u64 NG = ((u64)C + 31) / 32;
        const uint32_t gy_max = 16384;
        CU_TRY(cudaMemcpyAsync(D.d_tmp_rows.p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, D.stream));
        for (uint32_t q0 = 0; q0 < nq; q0 += gy_max) {
            const uint32_t cnt = std::min(gy_max, nq - q0);
            const dim3 grid((uint32_t)((NG + SCAN_WARPS - 1) / SCAN_WARPS), cnt);
            if (vec4) k_vec<<<grid, SCAN_WARPS * 32, smem, D.stream>>>(ix->d_centroids.p, C, dim, D.d_tmp_rows.p + (size_t)q0 * dim, D.d_dist.p + (size_t)q0 * C);
            else k_sca<<<grid, SCAN_WARPS * 32, smem, D.stream>>>(ix->d_centroids.p, C, dim, D.d_tmp_rows.p + (size_t)q0 * dim, D.d_dist.p + (size_t)q0 * C);
        }
        pqv::rank_batch_kernel<<<nq, 1024, (size_t)cp2 * 8, D.stream>>>(D.d_dist.p, C, cp2, np, D.d_assign.p, D.d_row_ids.p);
        CU_TRY(cudaMemsetAsync(D.vt_mask.p, 0, (size_t)C * qwords * 4, D.stream));
        pqv::probe_build_kernel<<<(uint32_t)(((u64)nq * np + 255) / 256), 256, 0, D.stream>>>(D.d_assign.p, nq, np, D.d_row_ids.p, qwords, D.vt_mask.p);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(nan_flags.data(), D.d_row_ids.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, D.stream));
        const uint32_t *d_row_mask = nullptr;
        if (h_row_mask) {  // the scan subtree's filter, shared by every query of the batch
            const u64 n_words = (ds->n_rows + 31) / 32;
            PQV_TRY(D.vt_bitmap.ensure(n_words));
            CU_TRY(cudaMemsetAsync(D.vt_bitmap.p + (n_words - 1), 0, 4, D.stream));
            CU_TRY(cudaMemcpyAsync(D.vt_bitmap.p, h_row_mask, (size_t)((ds->n_rows + 7) / 8), cudaMemcpyHostToDevice, D.stream));
            d_row_mask = D.vt_bitmap.p;
        }
        BatchMask bm{ix->d_row_cluster.p, D.vt_mask.p, qwords, d_row_mask};
        PQV_TRY(batch_topk(ctx, D, *ds, ds->n_rows, dim, queries, nq, k, flags, out_row_idx, out_dist, out_count, handled, raw_keys,
                           raw_count, pos_base, &bm, nullptr, nullptr, tie_out));
        ctx->batch_state.valid = false;  // the candidate segments left on the device are masked: not for the dense tie API
        CU_TRY(cudaStreamSynchronize(D.stream));  // nan_flags is in (batch_topk may have returned before its own sync)
        for (uint32_t q = 0; q < nq; ++q)
            if (nan_flags[q]) handled[q] = 0;
        if (tie_out) {  // a NaN ranking has no usable candidate sequence on the device
            size_t o = 0;
            for (size_t t = 0; t < tie_out->queries.size(); ++t)
                if (!nan_flags[tie_out->queries[t]]) {
                    tie_out->queries[o] = tie_out->queries[t];
                    tie_out->cnt[o] = tie_out->cnt[t];
                    tie_out->T[o] = tie_out->T[t];
                    ++o;
                }
            tie_out->queries.resize(o);
            tie_out->cnt.resize(o);
            tie_out->T.resize(o);
        }
    }
    return PQV_OK;
}

// Tie queries of a batched IVF search (rank-order form), replayed from a SHORT prefix of their candidate sequences.
// The reference heap (src/ivf/search.rs:115-127) walks the query's sequence -- probed lists in rank order -- and admits a
// row iff it beats the current k-th smallest.  Let T = qT (theta_select_kernel): every row with d <= T is among the query's
// candidates of the batched pass, and at least k of them exist.  Let P be the sequence position of the k-th candidate
// (in sequence order) with d <= T.  From P on the heap's threshold is <= T, so whatever it admits behind P is a candidate the
// pass already holds with its exact distance; only the rows at positions <= P can enter with d > T, and those are scanned
// exactly (ivf_search_fused limited to P + 1 candidates: ~2 % of the sequence at k = 100).  The union, in sequence order,
// replays to the reference's answer, layout ties included.  Queries this cannot place (candidates outside the lists' maps,
// fewer than k candidates under T, a declined fused scan) stay unhandled for the full single-query pipeline.
static int index_host_row_maps(DeviceState &D, IvfIndex &ix, u64 n_rows) {
    if (ix.h_row_maps_rows == n_rows && !ix.h_row_cluster.empty()) return PQV_OK;
    PQV_TRY(index_host_ids(D, ix));
    ix.h_row_cluster.assign(n_rows, 0xFFFFFFFFu);
    ix.h_row_listpos.assign(n_rows, 0u);
    for (uint32_t c = 0; c < ix.n_clusters; ++c)
        for (u64 i = ix.offsets[c]; i < ix.offsets[c + 1]; ++i) {
            const uint32_t r = ix.ids[i];
            if (r < n_rows) {
                ix.h_row_cluster[r] = c;
                ix.h_row_listpos[r] = (uint32_t)(i - ix.offsets[c]);
            }
        }
    ix.h_row_maps_rows = n_rows;
    return PQV_OK;
}

static int ivf_ties_short_prefix(pqv_ctx *ctx, Dataset *ds, DeviceState &D, IvfIndex *ix, const float *queries, uint32_t nq,
                                 uint32_t k, uint32_t nprobe, uint32_t flags, const BatchTieOut &ties, uint32_t *out_row_idx,
                                 float *out_dist, uint32_t *out_count, std::vector<uint8_t> &handled, uint32_t *n_resolved) {
    *n_resolved = 0;
    const size_t nt = ties.queries.size();
    if (!nt || !ivf_fused_enabled() || ix->n_clusters > IVF_RANK_MAX_C) return PQV_OK;
    const uint32_t C = ix->n_clusters, np = std::min(nprobe, C), dim = ds->dim;
    PQV_TRY(index_host_row_maps(D, *ix, ds->n_rows));
    // the pass' rankings (D.d_assign: [nq][np] cluster ids in rank order) and the tie queries' candidate segments
    std::vector<uint32_t> ranked((size_t)nq * np);
    CU_TRY(cudaMemcpyAsync(ranked.data(), D.d_assign.p, ranked.size() * 4, cudaMemcpyDeviceToHost, D.stream));
    std::vector<size_t> seg_off(nt + 1, 0);
    for (size_t t = 0; t < nt; ++t) seg_off[t + 1] = seg_off[t] + ties.cnt[t];
    std::vector<u64> segs(seg_off[nt]);
    for (size_t t = 0; t < nt; ++t)
        if (ties.cnt[t])
            CU_TRY(cudaMemcpyAsync(segs.data() + seg_off[t], D.tb_seg.p + (size_t)ties.queries[t] * ties.cap_q, (size_t)ties.cnt[t] * 8,
                                   cudaMemcpyDeviceToHost, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    struct Item {
        uint32_t pos, row;
        float d;
    };
    struct Tie {
        std::vector<Item> items;  // the pass' candidates with their sequence positions, ascending
        u64 P = ~0ull;            // sequence position of the k-th candidate with d <= T (~0: cannot be placed)
        EntrantsOut eo;
        bool scanned = false;
    };
    std::vector<Tie> work(nt);
    const size_t nth = std::max<size_t>(1, std::min<size_t>(tie_threads(), (nt + 1) / 2));
    auto parallel = [&](auto &&fn) {  // fn(t) for every tie query, independent of one another
        if (nth <= 1) {
            for (size_t t = 0; t < nt; ++t) fn(t);
            return;
        }
        std::vector<std::thread> th;
        for (size_t w = 1; w < nth; ++w)
            th.emplace_back([&, w] {
                for (size_t t = w; t < nt; t += nth) fn(t);
            });
        for (size_t t = 0; t < nt; t += nth) fn(t);
        for (auto &x : th) x.join();
    };
    // phase A (host threads): place every candidate in its query's sequence, find P
    parallel([&](size_t t) {
        Tie &W = work[t];
        const uint32_t q = ties.queries[t];
        const uint32_t *rk = ranked.data() + (size_t)q * np;
        std::vector<int32_t> rank_of(C, -1);
        std::vector<u64> prefix((size_t)np + 1);
        prefix[0] = 0;
        for (uint32_t r = 0; r < np; ++r) {
            rank_of[rk[r]] = (int32_t)r;
            prefix[r + 1] = prefix[r] + (ix->offsets[rk[r] + 1] - ix->offsets[rk[r]]);
        }
        const float T = ties.T[t];
        if (prefix[np] > 0xFFFFFFFFull || !(T >= 0.f)) return;
        const size_t b0 = seg_off[t], b1 = seg_off[t + 1];
        W.items.reserve(b1 - b0);
        for (size_t i = b0; i < b1; ++i) {
            if (i + 16 < b1) {  // the two maps are tens of MB: hide the misses
                const uint32_t ahead = key_pos(segs[i + 16]);
                if (ahead < ds->n_rows) {
                    __builtin_prefetch(&ix->h_row_cluster[ahead]);
                    __builtin_prefetch(&ix->h_row_listpos[ahead]);
                }
            }
            const uint32_t row = key_pos(segs[i]);
            const uint32_t c = row < ds->n_rows ? ix->h_row_cluster[row] : 0xFFFFFFFFu;
            if (c == 0xFFFFFFFFu || rank_of[c] < 0) {
                W.items.clear();
                return;
            }
            W.items.push_back(Item{(uint32_t)(prefix[rank_of[c]] + ix->h_row_listpos[row]), row, key_dist(segs[i])});
        }
        std::sort(W.items.begin(), W.items.end(), [](const Item &x, const Item &y) { return x.pos < y.pos; });
        uint32_t under = 0;
        for (const Item &it : W.items)
            if (it.d <= T && ++under == k) {
                W.P = it.pos;
                break;
            }
    });
    // phase B (one stream): exact scans of the sequence prefixes [0, P]
    for (size_t t = 0; t < nt; ++t) {
        Tie &W = work[t];
        if (W.P == ~0ull) continue;
        bool done = false;
        uint32_t dummy_rows[1], dummy_cnt = 0;
        float dummy_dist[1];
        PQV_TRY(ivf_search_fused(ctx, *ds, D, *ix, queries + (size_t)ties.queries[t] * dim, k, nprobe, flags, dummy_rows, dummy_dist,
                                 &dummy_cnt, &done, nullptr, &W.eo, W.P + 1));
        W.scanned = done;
    }
    // phase C (host threads): union in sequence order, reference loop
    parallel([&](size_t t) {
        Tie &W = work[t];
        if (!W.scanned) return;
        const uint32_t q = ties.queries[t];
        std::vector<Item> merged;
        merged.reserve(W.eo.keys.size() + W.items.size());
        for (size_t i = 0; i < W.eo.keys.size(); ++i) merged.push_back(Item{key_pos(W.eo.keys[i]), W.eo.rows[i], key_dist(W.eo.keys[i])});
        for (const Item &it : W.items)
            if (it.pos > W.P) merged.push_back(it);
        std::sort(merged.begin(), merged.end(), [](const Item &x, const Item &y) { return x.pos < y.pos; });
        out_count[q] = (uint32_t)replay_ordered(
            merged.size(), [&](size_t i) { return ReplayItem{merged[i].d, merged[i].row}; }, k, flags, out_row_idx + (size_t)q * k,
            out_dist + (size_t)q * k);
        handled[q] = 1;
    });
    for (size_t t = 0; t < nt; ++t) *n_resolved += work[t].scanned ? 1u : 0u;
    return PQV_OK;
}

// Batched IVF search: nq independent TopkBuilder::search calls (src/ivf/search.rs:83-142; PQV_ROW_ORDER: the candidate
// handling of VectorTopKExec, src/df_vector/exec.rs:207-277, without cap and filter) over one resident table + index,
// answered by ONE tensor-core pass over the table restricted, per query, to the rows of the
// clusters that query probes: all centroid rankings in one launch (l2_dist_batch_kernel + rank_batch_kernel), the probe
// sets as a [cluster][query] bit matrix, the row -> cluster map from the lists, and the mask applied in the filter's
// epilogues (pqv_tc.cuh: one word per 32 queries).  A query whose k + 1 best candidates hold an exact tie, a NaN centroid
// distance, or a batch the filter declines goes through the single-query pipeline -- every result equals its own call.
int pqv_ivf_search_batch(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *queries, uint32_t n_queries, uint32_t k,
                         uint32_t nprobe, uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    if (n_queries && (!queries || !out_row_idx || !out_dist || !out_count)) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");
    const bool row_order = (flags & PQV_ROW_ORDER) != 0;
    flags &= ~PQV_ROW_ORDER;
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (ix->dim != ds->dim) return fail(PQV_EINVAL, "Query dimension mismatch: expected %u, got %u", ix->dim, ds->dim);
    PQV_TRY(check_index_fits(*ix, *ds));
    DeviceState &D = ctx->devs[ds->shards[0].di];
    DevGuard guard(D.dev);
    PQV_TRY(index_make_resident(D, *ix));
    const uint32_t C = ix->n_clusters, dim = ds->dim;
    const bool multi = ds->shards.size() > 1;  // several devices: every query through the split single-query pipeline
    std::vector<uint8_t> handled(n_queries, 0), part;
    for (uint32_t q0 = 0; q0 < n_queries && !multi; q0 += BATCH_MAX_QUERIES) {  // the pass' scratch grows with the batch
        const uint32_t nq = std::min(BATCH_MAX_QUERIES, n_queries - q0);
        BatchTieOut tie_out;
        static const bool short_prefix = !(getenv("PQV_IVF_TIE_PREFIX") && !strcmp(getenv("PQV_IVF_TIE_PREFIX"), "off"));
        const bool want_ties = !row_order && short_prefix;  // the row-order form walks another sequence (ascending rows)
        PQV_TRY(ivf_batch_masked(ctx, ds, D, ix, queries + (size_t)q0 * dim, nq, k, nprobe, flags, out_row_idx + (size_t)q0 * k,
                                 out_dist + (size_t)q0 * k, out_count + q0, part, nullptr, nullptr, 0, nullptr,
                                 want_ties ? &tie_out : nullptr));
        if (want_ties && !tie_out.queries.empty()) {
            const pqv_batch_timing keep = ctx->last_batch;  // the prefix scans below overwrite the single-query timing only
            uint32_t resolved = 0;
            PQV_TRY(ivf_ties_short_prefix(ctx, ds, D, ix, queries + (size_t)q0 * dim, nq, k, nprobe, flags, tie_out,
                                          out_row_idx + (size_t)q0 * k, out_dist + (size_t)q0 * k, out_count + q0, part, &resolved));
            ctx->last_batch = keep;
            ctx->last_batch.tie_queries += (uint32_t)tie_out.queries.size();
            ctx->last_batch.tie_batched += resolved;
        }
        for (uint32_t i = 0; i < nq; ++i) handled[q0 + i] = part[i];
    }
    if (getenv("PQV_TRACE")) {
        uint32_t left = 0;
        for (uint32_t q = 0; q < n_queries; ++q) left += handled[q] ? 0u : 1u;
        fprintf(stderr, "[pqv trace] ivf_search_batch: %u queries, %u tie queries (%u replayed from a short prefix), %u left for the "
                        "single-query pipeline\n", n_queries, ctx->last_batch.tie_queries, ctx->last_batch.tie_batched, left);
    }
    for (uint32_t q = 0; q < n_queries; ++q) {
        if (handled[q]) continue;
        const float *qv = queries + (size_t)q * dim;
        uint32_t *orow = out_row_idx + (size_t)q * k;
        float *odist = out_dist + (size_t)q * k;
        if (!row_order) {
            PQV_TRY(ivf_search_one(ctx, ds, D, ix, qv, k, nprobe, flags, orow, odist, out_count + q));
        } else {
            RowOrder ro;
            bool done = false;
            out_count[q] = 0;
            if (!multi && ivf_fused_enabled() && C <= IVF_RANK_MAX_C && ix->n_ids)
                PQV_TRY(ivf_search_fused(ctx, *ds, D, *ix, qv, k, nprobe, flags, orow, odist, out_count + q, &done, &ro));
            if (!done) {  // host-ranked selection in row order (as pqv_vector_topk_indexed)
                PQV_TRY(index_host_ids(D, *ix));
                std::vector<uint32_t> ranked, rows;
                PQV_TRY(rank_clusters(D, *ix, qv, nprobe, ranked));
                for (uint32_t c : ranked) rows.insert(rows.end(), ix->ids.begin() + ix->offsets[c], ix->ids.begin() + ix->offsets[c + 1]);
                std::sort(rows.begin(), rows.end());
                if (!rows.empty()) PQV_TRY(topk_one(ctx, *ds, qv, rows.data(), rows.size(), k, flags, orow, odist, out_count + q));
            }
        }
    }
    return PQV_OK;
}


// per-rank half of a sharded batched IVF search: the k + 1 smallest exact keys (bits(d) << 32 | pos_base + local row) of
// every query among the probed rows of THIS rank's slice (index = the lists cut to the slice); out_count[q] = 0xFFFFFFFF
// where the slice could not decide the query.  Merge with pqv_merge_batch_keys; flagged queries: pqv_ivf_search_candidates.
int pqv_ivf_search_batch_keys(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *queries, uint32_t n_queries, uint32_t k,
                              uint32_t nprobe, uint32_t flags, uint32_t pos_base, uint64_t *out_keys, uint32_t *out_count) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    if (n_queries && (!queries || !out_keys || !out_count)) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (flags & PQV_TIES_BY_POSITION) return fail(PQV_EINVAL, "batch keys are only defined for the reference tie order");
    if (ix->dim != ds->dim) return fail(PQV_EINVAL, "Query dimension mismatch: expected %u, got %u", ix->dim, ds->dim);
    if (ds->shards.size() != 1) return fail(PQV_EINVAL, "pqv_ivf_search_batch_keys needs a single-device dataset");
    PQV_TRY(check_index_fits(*ix, *ds));
    if ((u64)pos_base + ds->n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "global row ids are u32");
    if (n_queries > BATCH_MAX_QUERIES)
        return fail(PQV_ELIMIT, "at most %u queries per pqv_ivf_search_batch_keys call (got %u)", BATCH_MAX_QUERIES, n_queries);
    for (uint32_t q = 0; q < n_queries; ++q) out_count[q] = 0xFFFFFFFFu;
    if (n_queries && ix->n_ids == 0) {  // nothing of this slice is in any list
        for (uint32_t q = 0; q < n_queries; ++q) out_count[q] = 0;
        return PQV_OK;
    }
    DeviceState &D = ctx->devs[ds->shards[0].di];
    DevGuard guard(D.dev);
    PQV_TRY(index_make_resident(D, *ix));
    std::vector<uint8_t> handled;
    std::vector<uint32_t> counts(n_queries, 0xFFFFFFFFu);
    PQV_TRY(ivf_batch_masked(ctx, ds, D, ix, queries, n_queries, k, nprobe, flags, nullptr, nullptr, nullptr, handled,
                             reinterpret_cast<u64 *>(out_keys), counts.data(), pos_base));
    for (uint32_t q = 0; q < n_queries; ++q) out_count[q] = handled[q] ? counts[q] : 0xFFFFFFFFu;
    return PQV_OK;
}


int pqv_ivf_search_candidates(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k, uint32_t nprobe,
                              uint32_t flags, uint64_t *out_keys, uint32_t *out_rows, uint64_t cap, uint64_t *out_count,
                              uint32_t *out_probe, uint32_t *out_nprobe_eff) {
    if (!ctx || !query || !out_count || !out_probe || !out_nprobe_eff || (cap && (!out_keys || !out_rows))) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (flags & PQV_TIES_BY_POSITION) return fail(PQV_EINVAL, "candidates are only defined for the reference tie order");
    if (ix->dim != ds->dim) return fail(PQV_EINVAL, "Query dimension mismatch: expected %u, got %u", ix->dim, ds->dim);
    if (ds->shards.size() != 1) return fail(PQV_EINVAL, "pqv_ivf_search_candidates needs a single-device dataset");
    PQV_TRY(check_index_fits(*ix, *ds));
    DeviceState &D = ctx->devs[ds->shards[0].di];
    DevGuard guard(D.dev);
    PQV_TRY(index_make_resident(D, *ix));
    EntrantsOut eo;
    bool done = false;
    if (ivf_fused_enabled() && ix->n_clusters <= IVF_RANK_MAX_C && ix->n_ids) {
        uint32_t dummy_rows[1], dummy_cnt = 0;
        float dummy_dist[1];
        PQV_TRY(ivf_search_fused(ctx, *ds, D, *ix, query, k, nprobe, flags, dummy_rows, dummy_dist, &dummy_cnt, &done, nullptr, &eo));
    }
    if (!done) {  // host-ranked path: NaN centroid distance, entrant overflow, empty or very wide index
        PQV_TRY(index_host_ids(D, *ix));
        PQV_TRY(rank_clusters(D, *ix, query, nprobe, eo.probe));
        std::vector<uint32_t> rows;
        for (uint32_t c : eo.probe) rows.insert(rows.end(), ix->ids.begin() + ix->offsets[c], ix->ids.begin() + ix->offsets[c + 1]);
        eo.keys.clear();
        eo.rows.clear();
        if (!rows.empty()) {
            for (uint32_t r : rows)
                if (r >= ds->n_rows) return fail(PQV_EINVAL, "row id %u is out of range (%llu rows)", r, (unsigned long long)ds->n_rows);
            PQV_TRY(topk_one(ctx, *ds, query, rows.data(), rows.size(), k, flags, nullptr, nullptr, nullptr, &eo.keys, 0));
            eo.rows.resize(eo.keys.size());
            for (size_t i = 0; i < eo.keys.size(); ++i) eo.rows[i] = rows[key_pos(eo.keys[i])];
        }
    }
    *out_nprobe_eff = (uint32_t)eo.probe.size();
    memcpy(out_probe, eo.probe.data(), eo.probe.size() * 4);
    *out_count = eo.keys.size();
    if (eo.keys.size() > cap) return fail(PQV_ELIMIT, "%zu candidate keys do not fit the caller's buffer of %llu", eo.keys.size(), (unsigned long long)cap);
    if (!eo.keys.empty()) {
        memcpy(out_keys, eo.keys.data(), eo.keys.size() * 8);
        memcpy(out_rows, eo.rows.data(), eo.rows.size() * 4);
    }
    return PQV_OK;
}

// Batched VectorTopKExec over a resident indexed table: n_queries executions of the operator (exec.rs:207-277) that share
// one scan subtree filter (row_mask, may be NULL), no candidate cap.  One masked tensor-core pass (probe sets AND filter
// bitmap inside the epilogues); undecided queries go through pqv_vector_topk_indexed's single-query pipeline.
int pqv_vector_topk_indexed_batch(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *queries, uint32_t n_queries,
                                  uint32_t k, uint32_t nprobe, uint32_t flags, const uint8_t *row_mask, uint32_t *out_row_idx,
                                  float *out_dist, uint32_t *out_count) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    if (n_queries && (!queries || !out_row_idx || !out_dist || !out_count)) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (ix->dim != ds->dim) return fail(PQV_EINVAL, "Query dimension mismatch: expected %u, got %u", ix->dim, ds->dim);
    PQV_TRY(check_index_fits(*ix, *ds));
    DeviceState &D = ctx->devs[ds->shards[0].di];
    DevGuard guard(D.dev);
    PQV_TRY(index_make_resident(D, *ix));
    const uint32_t dim = ds->dim;
    const bool multi = ds->shards.size() > 1;
    std::vector<uint8_t> handled(n_queries, 0), part;
    for (uint32_t q0 = 0; q0 < n_queries && !multi; q0 += BATCH_MAX_QUERIES) {
        const uint32_t nq = std::min(BATCH_MAX_QUERIES, n_queries - q0);
        PQV_TRY(ivf_batch_masked(ctx, ds, D, ix, queries + (size_t)q0 * dim, nq, k, nprobe, flags, out_row_idx + (size_t)q0 * k,
                                 out_dist + (size_t)q0 * k, out_count + q0, part, nullptr, nullptr, 0, row_mask));
        for (uint32_t i = 0; i < nq; ++i) handled[q0 + i] = part[i];
    }
    for (uint32_t q = 0; q < n_queries; ++q) {
        if (handled[q]) continue;
        const float *qv = queries + (size_t)q * dim;
        uint32_t *orow = out_row_idx + (size_t)q * k;
        float *odist = out_dist + (size_t)q * k;
        RowOrder ro;
        ro.h_mask = row_mask;
        bool done = false;
        out_count[q] = 0;
        if (!multi && ivf_fused_enabled() && ix->n_clusters <= IVF_RANK_MAX_C && ix->n_ids)
            PQV_TRY(ivf_search_fused(ctx, *ds, D, *ix, qv, k, nprobe, flags, orow, odist, out_count + q, &done, &ro));
        if (!done) {
            PQV_TRY(index_host_ids(D, *ix));
            std::vector<uint32_t> ranked, rows;
            PQV_TRY(rank_clusters(D, *ix, qv, nprobe, ranked));
            for (uint32_t c : ranked) rows.insert(rows.end(), ix->ids.begin() + ix->offsets[c], ix->ids.begin() + ix->offsets[c + 1]);
            std::sort(rows.begin(), rows.end());
            if (row_mask) {
                size_t o = 0;
                for (uint32_t r : rows)
                    if (row_mask[r >> 3] & (1u << (r & 7))) rows[o++] = r;
                rows.resize(o);
            }
            if (!rows.empty()) PQV_TRY(topk_one(ctx, *ds, qv, rows.data(), rows.size(), k, flags, orow, odist, out_count + q));
        }
    }
    return PQV_OK;
}

int pqv_vector_topk_indexed(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k, uint32_t nprobe,
                            uint32_t flags, uint64_t max_candidates, const uint8_t *row_mask, uint32_t *out_row_idx,
                            float *out_dist, uint32_t *out_count, uint64_t *out_candidate_rows, uint64_t *out_rows_scored) {
    if (!ctx || !query || !out_row_idx || !out_dist || !out_count) return fail(PQV_EINVAL, "null argument");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    IvfIndex *ix = find_index(ctx, index);
    if (!ix) return fail(PQV_EHANDLE, "unknown index handle %llu", (unsigned long long)index);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (ix->dim != ds->dim) return fail(PQV_EINVAL, "Query dimension mismatch: expected %u, got %u", ix->dim, ds->dim);  // index_exec.rs:152-158
    const bool multi = ds->shards.size() > 1;  // several devices: host-side selection, candidates split by owner (topk_one)
    PQV_TRY(check_index_fits(*ix, *ds));
    DeviceState &D = ctx->devs[ds->shards[0].di];
    DevGuard guard(D.dev);
    PQV_TRY(index_make_resident(D, *ix));
    RowOrder ro;
    ro.max_candidates = max_candidates;   // PQV_NO_CANDIDATE_CAP = ~0: None = no cap (options.rs:10-11)
    ro.h_mask = row_mask;
    *out_count = 0;
    if (max_candidates == 0) {
        // Some(0): target_candidates = 0 (exec.rs:222-223) -- the index scan still reports its candidates, nothing is
        // fetched or scored, the operator emits no row
        std::vector<uint32_t> ranked0;
        PQV_TRY(rank_clusters(D, *ix, query, nprobe, ranked0));
        u64 total0 = 0;
        for (uint32_t c : ranked0) total0 += ix->offsets[c + 1] - ix->offsets[c];
        if (out_candidate_rows) *out_candidate_rows = total0;
        if (out_rows_scored) *out_rows_scored = 0;
        return PQV_OK;
    }
    if (!multi && ivf_fused_enabled() && ix->n_clusters <= IVF_RANK_MAX_C && ix->n_ids && k <= PQV_MAX_K) {  // (larger k: topk_one's full replay)
        bool done = false;
        PQV_TRY(ivf_search_fused(ctx, *ds, D, *ix, query, k, nprobe, flags, out_row_idx, out_dist, out_count, &done, &ro));
        if (done) {
            if (out_candidate_rows) *out_candidate_rows = ro.candidate_rows;
            if (out_rows_scored) *out_rows_scored = ro.rows_scored;
            return PQV_OK;
        }
    }
    // host-ranked path (NaN centroid distance, entrant overflow, very large C): same selection on the host
    PQV_TRY(index_host_ids(D, *ix));
    std::vector<uint32_t> ranked;
    PQV_TRY(rank_clusters(D, *ix, query, nprobe, ranked));
    std::vector<uint32_t> rows;
    u64 total = 0;
    for (uint32_t c : ranked) {
        const u64 b = ix->offsets[c], e = ix->offsets[c + 1];
        for (u64 i = b; i < e; ++i, ++total)
            if (total < ro.max_candidates) rows.push_back(ix->ids[i]);
    }
    std::sort(rows.begin(), rows.end());
    if (row_mask) {
        size_t o = 0;
        for (uint32_t r : rows)
            if (row_mask[r >> 3] & (1u << (r & 7))) rows[o++] = r;
        rows.resize(o);
    }
    if (out_candidate_rows) *out_candidate_rows = total;
    if (out_rows_scored) *out_rows_scored = rows.size();
    if (rows.empty()) return PQV_OK;
    for (uint32_t r : rows)
        if (r >= ds->n_rows) return fail(PQV_EINVAL, "row id %u is out of range (%llu rows)", r, (unsigned long long)ds->n_rows);
    return topk_one(ctx, *ds, query, rows.data(), rows.size(), k, flags, out_row_idx, out_dist, out_count);
}

}  // extern "C"

static void pqv_free_all_indexes(pqv_ctx *ctx) {
    if (ctx->indexes.empty()) return;
    DevGuard guard(ctx->devs[0].dev);
    for (auto &kv : ctx->indexes) index_free(static_cast<IvfIndex *>(kv.second));
    ctx->indexes.clear();
}
