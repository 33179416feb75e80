// pqv_tc_host.cuh -- host launch sequence of the tcgen05 assignment filter (included by pqv_capi.cu after
// DeviceState / pqv_ctx are defined).  assign_dispatch() is the one entry used by pqv_kmeans_assign, the IVF
// build (Lloyd + final assignment) and pqv_bench_assign.
#pragma once

#include "pqv_tc.cuh"

namespace {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    });
    return fn;
}

// row-major [rows][dim] table (f32, or its fp16 shadow) viewed as a 2-D tensor {dim (inner), rows}; box = one 128-byte
// stage row (32 f32 / 64 f16 columns) x box_rows rows, 128 B swizzle; out-of-bounds elements (row tail, dim tail) are
// filled with zeros and still count towards the transaction bytes
int make_row_tmap(CUtensorMap *tm, const void *d_ptr, u64 rows, uint32_t dim, uint32_t box_rows, int kind) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return fail(PQV_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const bool f16 = kind == pqv::tc::KIND_F16;
    cuuint64_t gdim[2] = {dim, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)pqv::tc::bk_elems(kind), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(d_ptr), gdim,
                     gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PQV_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu dim=%u kind=%d)", (int)r,
                                       (unsigned long long)rows, dim, kind);
    return PQV_OK;
}

// ------------------------------------------------------------------------------------------------
// 16-bit operand shadow (pqv_half.cuh)
// ------------------------------------------------------------------------------------------------
struct ShadowView {
    const __half *h = nullptr;
    const float2 *stats = nullptr;
    const float *mu = nullptr;
    const pqv::half16::Globals *g = nullptr;
};

// PQV_TC_KIND=tf32 keeps the tensor-core filters on the f32 rows (kind::tf32); default: the fp16 shadow wherever the layout
// allows it (dim % 8 == 0: 16-byte aligned rows of halves for TMA)
bool half_kind_enabled() {  // read on every call: tests and benchmarks switch kinds inside one process
    const char *e = getenv("PQV_TC_KIND");
    return !(e && !strcmp(e, "tf32"));
}
bool shadow_layout_ok(uint32_t dim, const void *d_rows) {
    return half_kind_enabled() && (dim % 8 == 0) && dim >= 64 && ((reinterpret_cast<uintptr_t>(d_rows) & 15) == 0);
}

constexpr u64 SHADOW_MEAN_ROWS = 65536;  // rows the data mean is taken over

// zeroes the shadow's globals and derives the data mean and the operand scale from the first rows of the table
int shadow_begin(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, float *mu, pqv::half16::Globals *g) {
    const u64 ns = std::min<u64>(n, SHADOW_MEAN_ROWS);
    CU_TRY(cudaMemsetAsync(g, 0, sizeof(pqv::half16::Globals), D.stream));
    PQV_TRY(D.ts_mean_part.ensure((size_t)pqv::half16::MEAN_SLICES * dim));
    pqv::half16::column_mean_partial_kernel<<<dim3((dim + 127) / 128, pqv::half16::MEAN_SLICES), 128, 0, D.stream>>>(
        d_rows, ns, dim, D.ts_mean_part.p);
    pqv::half16::column_mean_kernel<<<(dim + 127) / 128, 128, 0, D.stream>>>(D.ts_mean_part.p, ns, dim, mu);
    // absmax of the sample lands (as bits) in the `reserved-for-scale` word, then becomes the power-of-two scale in place
    uint32_t *word = reinterpret_cast<uint32_t *>(&g->scale);
    pqv::half16::absmax_kernel<<<(uint32_t)D.sm_count * 4, 256, 0, D.stream>>>(d_rows, ns * dim, word);
    pqv::half16::scale_from_absmax_kernel<<<1, 32, 0, D.stream>>>(word, &g->scale);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

int shadow_launch(DeviceState &D, const float *d_rows, u64 first, u64 n, uint32_t dim, const float *mu, __half *h, float2 *stats,
                  pqv::half16::Globals *g) {
    if (n == 0) return PQV_OK;
    pqv::half16::shadow_rows_kernel<<<(uint32_t)D.sm_count * 8, 256, 0, D.stream>>>(d_rows, first, n, dim, mu, h, stats, g);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

// the shadow of a resident shard, created / extended on demand (enqueued on D.stream); *ms_out += its build time when one ran
int shard_shadow(DeviceState &D, const Dataset &ds, Shard &sh, ShadowView *out, bool *built = nullptr) {
    HalfShadow &S = sh.shadow;
    const u64 n = sh.n_rows;
    const uint32_t dim = ds.dim;
    if (built) *built = false;
    if (S.cap < n || !S.h) {
        CU_TRY(cudaStreamSynchronize(D.stream));
        S.drop();
        const u64 cap = std::max<u64>(n, sh.cap_rows);
        cudaError_t e = cudaMalloc((void **)&S.h, (size_t)cap * dim * sizeof(__half));
        if (e == cudaSuccess) e = cudaMalloc((void **)&S.stats, (size_t)cap * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc((void **)&S.mu, (size_t)dim * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void **)&S.g, sizeof(pqv::half16::Globals));
        if (e != cudaSuccess) {
            S.drop();
            cudaGetLastError();
            return fail(PQV_ENOMEM, "16-bit shadow of %llu x %u rows (%.2f GB): %s", (unsigned long long)cap, dim,
                        cap * (double)dim * 2 / 1e9, cudaGetErrorString(e));
        }
        S.cap = cap;
        S.rows = 0;
    }
    if (S.rows == 0 && n) PQV_TRY(shadow_begin(D, sh.d_data, n, dim, S.mu, S.g));
    if (S.rows < n) {
        CU_TRY(cudaEventRecord(D.ev_shadow[0], D.stream));
        PQV_TRY(shadow_launch(D, sh.d_data, S.rows, n - S.rows, dim, S.mu, S.h, S.stats, S.g));
        CU_TRY(cudaEventRecord(D.ev_shadow[1], D.stream));
        S.rows = n;
        if (built) *built = true;
    }
    out->h = S.h;
    out->stats = S.stats;
    out->mu = S.mu;
    out->g = S.g;
    return PQV_OK;
}

// shadow of n rows at d_rows that are not a resident shard (scratch of the device state; valid until the next call)
int temp_shadow(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, ShadowView *out) {
    PQV_TRY(D.ts_half.ensure((size_t)n * dim));
    PQV_TRY(D.ts_stats.ensure(n));
    PQV_TRY(D.ts_mu.ensure(dim));
    PQV_TRY(D.ts_g.ensure(1));
    PQV_TRY(shadow_begin(D, d_rows, n, dim, D.ts_mu.p, D.ts_g.p));
    CU_TRY(cudaEventRecord(D.ev_shadow[0], D.stream));
    PQV_TRY(shadow_launch(D, d_rows, 0, n, dim, D.ts_mu.p, D.ts_half.p, D.ts_stats.p, D.ts_g.p));
    CU_TRY(cudaEventRecord(D.ev_shadow[1], D.stream));
    out->h = D.ts_half.p;
    out->stats = D.ts_stats.p;
    out->mu = D.ts_mu.p;
    out->g = D.ts_g.p;
    return PQV_OK;
}

enum AssignPath { ASSIGN_SIMT = 0, ASSIGN_TC = 1 };

// PQV_TC_PAIR=off selects the single-CTA tcgen05 kernel (cta_group::1); default: CTA pairs (cta_group::2)
bool tc_pair_enabled() {
    static const bool on = [] {
        const char *e = getenv("PQV_TC_PAIR");
        return !(e && (!strcmp(e, "off") || !strcmp(e, "0")));
    }();
    return on;
}

// launch geometry of tc_rows_x_table(_pair)_kernel over num_mb 128-row tiles: CTAs, and 128-row tiles per CTA
struct TcGrid {
    bool pair;
    uint32_t grid, tiles_per_cta;
};
TcGrid tc_grid_for(const DeviceState &D, uint32_t num_mb) {
    TcGrid t;
    t.pair = tc_pair_enabled() && D.sm_count >= 2;
    if (t.pair) {
        const uint32_t num_mb2 = (num_mb + 1) / 2, pairs = std::min<uint32_t>(num_mb2, (uint32_t)D.sm_count / 2);
        t.grid = 2 * pairs;
        t.tiles_per_cta = (num_mb2 + pairs - 1) / pairs;
    } else {
        t.grid = std::min<uint32_t>(num_mb, (uint32_t)D.sm_count);
        t.tiles_per_cta = (num_mb + t.grid - 1) / t.grid;
    }
    return t;
}

// PQV_ASSIGN=simt|tc forces a path (tc still requires the layout preconditions); default: tc whenever it applies
int assign_path_for(const float *d_rows, u64 n, uint32_t dim, const float *d_cent, uint32_t C) {
    const bool layout_ok = (dim % 4 == 0) && dim >= (uint32_t)pqv::tc::BK && ((reinterpret_cast<uintptr_t>(d_rows) & 15) == 0) &&
                           ((reinterpret_cast<uintptr_t>(d_cent) & 15) == 0) && n < 0xFFFFFFFFull && tmap_encoder() != nullptr;
    const char *e = getenv("PQV_ASSIGN");
    if (e && !strcmp(e, "simt")) return ASSIGN_SIMT;
    if (e && !strcmp(e, "tc")) return layout_ok ? ASSIGN_TC : ASSIGN_SIMT;
    // below these sizes the exact SIMT kernel finishes in microseconds and the filter's fixed launches dominate
    return (layout_ok && n >= 2048 && C >= 16) ? ASSIGN_TC : ASSIGN_SIMT;
}

int assign_simt(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, const float *d_cent, uint32_t C, uint32_t *d_out) {
    const uint32_t grid = (uint32_t)((n + pqv::AS_BM - 1) / pqv::AS_BM);
    const bool vec4 = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_rows) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(d_cent) & 15) == 0);
    if (vec4) pqv::kmeans_assign_kernel<true><<<grid, 256, 0, D.stream>>>(d_rows, n, dim, d_cent, C, d_out);
    else pqv::kmeans_assign_kernel<false><<<grid, 256, 0, D.stream>>>(d_rows, n, dim, d_cent, C, d_out);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

// one launcher for every (epilogue, operand kind, single / pair) instance: raises the kernel's dynamic shared-memory limit
// once, then launches
template <class Epi, int KIND, bool PAIR>
int tc_launch_inst(uint32_t grid, cudaStream_t st, const CUtensorMap &tmA, const CUtensorMap &tmB, const pqv::tc::GemmShape &shape,
                   const typename Epi::Params &p) {
    namespace T = pqv::tc;
    constexpr size_t smem = PAIR ? T::smem_bytes_pair(Epi::STAGES_PAIR, Epi::EXCH_BYTES) : T::smem_bytes_single(Epi::STAGES_SINGLE, Epi::EXCH_BYTES);
    static_assert(smem <= 227 * 1024, "tensor-core kernel exceeds the shared memory of an SM");
    if (PAIR) PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(T::tc_rows_x_table_pair_kernel<Epi, KIND>), smem));
    else PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(T::tc_rows_x_table_kernel<Epi, KIND>), smem));
    if (PAIR) T::tc_rows_x_table_pair_kernel<Epi, KIND><<<grid, T::tc_threads(Epi::SPLIT), smem, st>>>(tmA, tmB, shape, p);
    else T::tc_rows_x_table_kernel<Epi, KIND><<<grid, T::tc_threads(Epi::SPLIT), smem, st>>>(tmA, tmB, shape, p);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}
template <class Epi>
int tc_launch(bool pair, int kind, uint32_t grid, cudaStream_t st, const CUtensorMap &tmA, const CUtensorMap &tmB,
              const pqv::tc::GemmShape &shape, const typename Epi::Params &p) {
    namespace T = pqv::tc;
    if (kind == T::KIND_F16)
        return pair ? tc_launch_inst<Epi, T::KIND_F16, true>(grid, st, tmA, tmB, shape, p)
                    : tc_launch_inst<Epi, T::KIND_F16, false>(grid, st, tmA, tmB, shape, p);
    return pair ? tc_launch_inst<Epi, T::KIND_TF32, true>(grid, st, tmA, tmB, shape, p)
                : tc_launch_inst<Epi, T::KIND_TF32, false>(grid, st, tmA, tmB, shape, p);
}

// sv != nullptr: the rows' 16-bit shadow (kind::f16 filter, statistics from the shadow pass); nullptr: f32 rows under
// kind::tf32 with the per-sweep row_stats_kernel pass
int assign_tc(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, const float *d_cent, uint32_t C, uint32_t *d_out,
              bool time_it, const ShadowView *sv) {
    namespace T = pqv::tc;
    const int kind = sv ? T::KIND_F16 : T::KIND_TF32;
    const uint32_t bke = (uint32_t)T::bk_elems(kind);
    const uint32_t num_mb = (uint32_t)((n + T::BM - 1) / T::BM);
    const uint32_t num_nb = (C + T::BN - 1) / T::BN;
    const uint32_t num_kb = (dim + bke - 1) / bke;
    const uint32_t cn_len = num_nb * T::BN;
    const u64 pair_cap64 = std::min<u64>(4 * n + 1024, 0xFFFFFFF0ull);
    PQV_TRY(D.tc_bp.ensure((size_t)C * dim));  // f32-sized: holds the tf32-rounded table or its fp16 form
    PQV_TRY(D.tc_mu.ensure(dim));
    PQV_TRY(D.tc_cn.ensure(2 * (size_t)cn_len));
    PQV_TRY(D.tc_wc.ensure(cn_len / 32));
    if (!sv) PQV_TRY(D.tc_stats.ensure(n));
    PQV_TRY(D.tc_u32.ensure(16));
    PQV_TRY(D.tc_perm.ensure(C));
    PQV_TRY(D.tc_okeys.ensure(C));
    PQV_TRY(D.tc_amb_rows.ensure(n));
    PQV_TRY(D.tc_pairs.ensure(pair_cap64));
    PQV_TRY(D.tc_best.ensure(n));
    PQV_TRY(D.tc_ovf_rows.ensure(n));
    uint32_t *bounds = D.tc_u32.p, *counts = D.tc_u32.p + 8;

    const TcGrid tg = tc_grid_for(D, num_mb);
    CUtensorMap tmA, tmB;
    PQV_TRY(make_row_tmap(&tmA, sv ? (const void *)sv->h : (const void *)d_rows, n, dim, T::BM, kind));
    PQV_TRY(make_row_tmap(&tmB, D.tc_bp.p, C, dim, tg.pair ? T::BN / 2 : T::BN, kind));

    if (time_it) CU_TRY(cudaEventRecord(D.ev[0], D.stream));
    CU_TRY(cudaMemsetAsync(D.tc_u32.p, 0, 16 * sizeof(uint32_t), D.stream));
    CU_TRY(cudaMemsetAsync(D.tc_wc.p, 0, (size_t)(cn_len / 32) * sizeof(uint32_t), D.stream));
    PQV_TRY(D.ts_mean_part.ensure((size_t)pqv::half16::MEAN_SLICES * dim));
    T::centroid_mean_partial_kernel<<<dim3((dim + 127) / 128, T::CMEAN_SLICES), 128, 0, D.stream>>>(d_cent, C, dim, D.ts_mean_part.p);
    T::centroid_mean_kernel<<<(dim + 127) / 128, 128, 0, D.stream>>>(D.ts_mean_part.p, C, dim, sv ? sv->mu : nullptr, D.tc_mu.p, bounds);
    // operand scale of the centred table: bounds[5] = max |c - mu| (fp16 only), then the power of two derived from it (1 for tf32)
    if (sv) T::centroid_absmax_kernel<<<(uint32_t)D.sm_count, 256, 0, D.stream>>>(d_cent, C, dim, D.tc_mu.p, bounds + 5);
    pqv::half16::scale_from_absmax_kernel<<<1, 32, 0, D.stream>>>(bounds + 5, reinterpret_cast<float *>(bounds + 5));
    // column order of the filter: ascending |c - mu|^2 (pqv_tc.cuh, centroid_order_kernel)
    const uint32_t *perm = nullptr;
    static const bool order_off = getenv("PQV_ASSIGN_ORDER") && !strcmp(getenv("PQV_ASSIGN_ORDER"), "off");
    if (C <= T::ORDER_MAX_C && C > 1 && !order_off) {
        uint32_t cp2 = 2;
        while (cp2 < C) cp2 <<= 1;
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(T::centroid_order_kernel), (size_t)T::ORDER_MAX_C * 8));
        T::centroid_spread_kernel<<<C, 128, 0, D.stream>>>(d_cent, C, dim, D.tc_mu.p, D.tc_okeys.p);
        T::centroid_order_kernel<<<1, 1024, (size_t)cp2 * 8, D.stream>>>(D.tc_okeys.p, C, cp2, D.tc_perm.p);
        perm = D.tc_perm.p;
    }
    if (sv) {
        T::centroid_prep_kernel<T::KIND_F16><<<cn_len, 128, 0, D.stream>>>(d_cent, C, dim, D.tc_mu.p, D.tc_bp.p, D.tc_cn.p,
                                                                          D.tc_cn.p + cn_len, D.tc_wc.p, cn_len, bounds, sv->g, perm);
    } else {
        T::centroid_prep_kernel<T::KIND_TF32><<<cn_len, 128, 0, D.stream>>>(d_cent, C, dim, D.tc_mu.p, D.tc_bp.p, D.tc_cn.p,
                                                                           D.tc_cn.p + cn_len, D.tc_wc.p, cn_len, bounds, nullptr, perm);
        T::row_stats_kernel<<<(uint32_t)D.sm_count * 8, 256, 0, D.stream>>>(d_rows, n, dim, D.tc_mu.p, D.tc_stats.p);
    }
    CU_TRY(cudaGetLastError());

    T::AssignTcParams p;
    p.stats = sv ? sv->stats : D.tc_stats.p;
    p.cn = D.tc_cn.p;
    p.wv = D.tc_cn.p + cn_len;
    p.wc = reinterpret_cast<const float *>(D.tc_wc.p);
    p.bounds = bounds;
    p.perm = perm;
    p.scale_a = sv ? &sv->g->scale : nullptr;
    p.assign = d_out;
    p.counts = counts;
    p.amb_rows = D.tc_amb_rows.p;
    p.pairs = D.tc_pairs.p;
    p.best = D.tc_best.p;
    p.ovf_rows = D.tc_ovf_rows.p;
    p.n = n;
    p.pair_cap = (uint32_t)pair_cap64;
    p.dim = dim;
    p.C = C;
    const T::GemmShape shape{num_mb, num_nb, num_kb};
    if (time_it) CU_TRY(cudaEventRecord(D.ev[1], D.stream));
    PQV_TRY(tc_launch<T::AssignEpi>(tg.pair, kind, tg.grid, D.stream, tmA, tmB, shape, p));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[2], D.stream));
    // exact f32 chains: one per (row, candidate) pair of the ambiguous rows; the full scan for the overflow rows
    T::pair_exact_kernel<<<(uint32_t)D.sm_count * 4, T::PAIR_WARPS * 32, 0, D.stream>>>(d_rows, dim, d_cent, counts, p.pair_cap,
                                                                                     D.tc_pairs.p, D.tc_best.p);
    if (time_it) CU_TRY(cudaEventRecord(D.ev[4], D.stream));
    {
        const uint32_t slice_len = pqv::AS_BN;  // one centroid tile per CTA: a handful of rows is spread over C / 64 CTAs
        dim3 grid_ovf((uint32_t)std::min<u64>((n + pqv::AS_BM - 1) / pqv::AS_BM, (u64)D.sm_count * 2), (C + slice_len - 1) / slice_len);
        pqv::kmeans_assign_kernel<true, true><<<grid_ovf, 256, 0, D.stream>>>(d_rows, 0, dim, d_cent, C, nullptr, D.tc_ovf_rows.p,
                                                                              counts + 1, D.tc_best.p, slice_len);
        // ... which leaves lists of up to FEW_ROWS_MAX rows (the usual handful) to the warp-per-(row, 32 centroids) kernel
        const size_t per_warp = ((size_t)((dim + 3u) & ~3u) + pqv::TileCfg<0, true>::TILE_FLOATS) * 4;
        const uint32_t fw = (uint32_t)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / per_warp));
        auto few = pqv::few_rows_assign_kernel<true>;
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(few), per_warp * fw));
        few<<<(uint32_t)D.sm_count * 2, fw * 32, per_warp * fw, D.stream>>>(d_rows, dim, d_cent, C, D.tc_ovf_rows.p, counts + 1,
                                                                             D.tc_best.p);
    }
    T::best_finalize_kernel<<<(uint32_t)D.sm_count * 2, 256, 0, D.stream>>>(counts, D.tc_amb_rows.p, D.tc_ovf_rows.p, D.tc_best.p,
                                                                            d_out);
    CU_TRY(cudaGetLastError());
    if (time_it) CU_TRY(cudaEventRecord(D.ev[3], D.stream));
    return PQV_OK;
}

// device offset of the (ambiguous, overflow) row counters inside D.tc_u32 after an assign_tc launch
constexpr size_t TC_COUNTS_OFFSET = 8;

// Enqueues the assignment of n device rows on D.stream (no synchronisation).  *path_out reports the path taken.
// sv: the 16-bit shadow of exactly these rows if the caller has one (a resident shard's, or one it built for rows it
// sweeps repeatedly); otherwise a scratch shadow is built here when the layout allows it (one extra pass over the rows).
int assign_dispatch(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, const float *d_cent, uint32_t C, uint32_t *d_out,
                    int *path_out = nullptr, bool time_it = false, const ShadowView *sv = nullptr, int *kind_out = nullptr,
                    bool *shadow_built = nullptr) {
    const int path = assign_path_for(d_rows, n, dim, d_cent, C);
    if (path_out) *path_out = path;
    if (kind_out) *kind_out = 0;
    if (shadow_built) *shadow_built = false;
    if (path == ASSIGN_TC) {
        ShadowView tmp;
        if (!sv && shadow_layout_ok(dim, d_rows)) {
            PQV_TRY(temp_shadow(D, d_rows, n, dim, &tmp));
            sv = &tmp;
            if (shadow_built) *shadow_built = true;
        }
        if (sv && !shadow_layout_ok(dim, d_rows)) sv = nullptr;
        if (kind_out) *kind_out = sv ? 1 : 0;
        return assign_tc(D, d_rows, n, dim, d_cent, C, d_out, time_it, sv);
    }
    if (time_it) CU_TRY(cudaEventRecord(D.ev[0], D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[1], D.stream));
    PQV_TRY(assign_simt(D, d_rows, n, dim, d_cent, C, d_out));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[2], D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[4], D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[3], D.stream));
    return PQV_OK;
}

// Reads the events recorded by assign_dispatch(time_it = true) into ctx->last_assign; the stream must be idle.
void record_assign_timing(pqv_ctx *ctx, DeviceState &D, int path, u64 rows, const uint32_t counts[2], bool first_piece,
                          int kind = 0, bool shadow_built = false) {
    float prep = 0.f, filt = 0.f, post = 0.f;
    cudaEventElapsedTime(&prep, D.ev[0], D.ev[1]);
    cudaEventElapsedTime(&filt, D.ev[1], D.ev[2]);
    cudaEventElapsedTime(&post, D.ev[2], D.ev[3]);
    float pair_ms = 0.f, shadow_ms = 0.f;
    cudaEventElapsedTime(&pair_ms, D.ev[2], D.ev[4]);
    if (shadow_built) cudaEventElapsedTime(&shadow_ms, D.ev_shadow[0], D.ev_shadow[1]);
    pqv_assign_timing &t = ctx->last_assign;
    if (first_piece) t = pqv_assign_timing{};
    t.path = (uint32_t)path;
    t.kind = (uint32_t)kind;
    t.rows += rows;
    t.ambiguous_rows += counts[0];
    t.overflow_rows += counts[1];
    t.prep_ms += prep;
    t.filter_ms += filt;
    t.recheck_ms += post;
    t.pair_ms += pair_ms;
    t.shadow_ms += shadow_ms;
    t.total_ms += (double)prep + filt + post;
}

// ------------------------------------------------------------------------------------------------
// batched brute-force top-k (pqv_tc.cuh, "Batched brute-force top-k")
// ------------------------------------------------------------------------------------------------
constexpr uint32_t BATCH_MIN_QUERIES = 4;          // below this the single-query scans are cheaper than the prep
constexpr uint32_t BATCH_MAX_QUERIES = 4096;       // queries per pass: the sample-bound matrix U is nq x 512 k floats (8.6 GB here)
constexpr u64 BATCH_TOTAL_CAND = 64ull << 20;      // candidate (row, query) records kept per batch (8 B each)

// PQV_BATCH=off disables the path (every query then takes the single-query scan)
struct BatchMask {
    const uint32_t *row_cluster;  // device, [n]
    const uint32_t *probe_T;      // device, [C][qwords]
    uint32_t qwords;              // ceil(nq / BN) * BN / 32
    const uint32_t *row_mask;     // device, [ceil(n / 32)] or null
};

// masked (IVF) batches: what the caller needs to replay its tie queries from a short prefix of their candidate sequences
struct BatchTieOut {
    std::vector<uint32_t> queries;  // tie queries of the pass (index into the batch)
    std::vector<uint32_t> cnt;      // candidates of each in its device segment (D.tb_seg + q * cap_q)
    std::vector<float> T;           // qT of each (theta_select_kernel)
    uint32_t cap_q = 0;
};

bool batch_path_applies(const Dataset &ds, const float *d_rows, uint32_t nq, uint32_t k, bool any_shards = false) {
    const char *e = getenv("PQV_BATCH");
    if (e && !strcmp(e, "off")) return false;
    return (any_shards || ds.shards.size() == 1) && nq >= BATCH_MIN_QUERIES && (ds.dim % 4 == 0) && ds.dim >= (uint32_t)pqv::tc::BK &&
           ((reinterpret_cast<uintptr_t>(d_rows) & 15) == 0) && ds.n_rows < 0xFFFFFFFFull && ds.n_rows >= 1 &&
           k + 1 <= (uint32_t)pqv::tc::SEL_MAX && tmap_encoder() != nullptr;
}

// Tie queries of a dense batch, resolved together (pqv_tie.cuh): one pass over the sample prefix [0, S) gives the exact
// distances of every prefix row against every tie query, one CTA per query turns its row of that matrix into a superset of
// the rows the reference heap admits inside the prefix, and the host replays the reference loop over those keys + the
// query's candidates behind the prefix (still on the device from the batched pass).  Queries whose entrant region
// overflowed (adversarial row order: most rows enter the heap) are returned in `slow` for the one-by-one path.
// PQV_TIE_BATCH=off sends every tie query through the one-by-one path instead.
constexpr uint32_t TIE_CHUNK_Q = 128;   // tie queries per pass: the distance matrix is TIE_CHUNK_Q x S floats (268 MB)
constexpr uint32_t TIE_ENT_CAP = 8192;  // entrant keys per query (expected: 2048 + ~k ln(S / 2048))

bool tie_batch_enabled() {
    const char *e = getenv("PQV_TIE_BATCH");
    return !(e && !strcmp(e, "off"));
}

// host threads for the independent heap replays of a tie batch (PQV_TIE_THREADS, default min(16, hardware threads):
// 50 replays take 8.9 ms on one thread, 1.9 ms on 8, 1.0 ms on 13)
size_t tie_threads() {
    if (const char *e = getenv("PQV_TIE_THREADS")) return (size_t)std::max(1, atoi(e));
    const unsigned hw = std::thread::hardware_concurrency();
    return std::min<size_t>(16, hw ? hw : 1);
}

int resolve_ties_together(DeviceState &D, const float *d_rows, u64 S, uint32_t dim, int order, uint32_t k, uint32_t flags,
                          const std::vector<uint32_t> &ties, const std::vector<uint32_t> &h_cnt, uint32_t cap_q,
                          uint32_t *out_rows, float *out_dist, uint32_t *out_count, std::vector<uint8_t> &handled,
                          std::vector<uint32_t> &slow) {
    namespace TI = pqv::tie;
    const u64 ldo = (S + 63) & ~63ull;
    const uint32_t kcap = std::max<uint32_t>(32, pow2ceil(k));
    const size_t region = (size_t)TIE_ENT_CAP + 1;
    cudaStream_t st = D.stream;
    static const bool trace = getenv("PQV_TRACE") != nullptr;
    for (size_t c0 = 0; c0 < ties.size(); c0 += TIE_CHUNK_Q) {
        const uint32_t tq = (uint32_t)std::min<size_t>(TIE_CHUNK_Q, ties.size() - c0);
        PQV_TRY(D.tie_dmat.ensure((size_t)tq * ldo));
        PQV_TRY(D.tie_qsel.ensure(tq));
        PQV_TRY(D.tie_ent.ensure((size_t)tq * region));
        PQV_TRY(D.h_tie_ent.ensure((size_t)tq * region));
        PQV_TRY(D.h_tie_qsel.ensure(tq));
        size_t seg_total = 0;
        for (uint32_t j = 0; j < tq; ++j) {
            D.h_tie_qsel.p[j] = ties[c0 + j];
            seg_total += std::min<uint32_t>(h_cnt[ties[c0 + j]], cap_q);
        }
        PQV_TRY(D.h_tie_seg.ensure(std::max<size_t>(seg_total, 1)));
        const double t_begin = trace ? trace_now_ms() : 0.0;
        CU_TRY(cudaMemcpyAsync(D.tie_qsel.p, D.h_tie_qsel.p, (size_t)tq * 4, cudaMemcpyHostToDevice, st));
        if (trace) CU_TRY(cudaEventRecord(D.ev[0], st));
        const dim3 grid((uint32_t)((S + TI::TN - 1) / TI::TN), (tq + TI::TM - 1) / TI::TM);
        if (order == 0)
            TI::prefix_dist_matrix_kernel<0><<<grid, 256, 0, st>>>(d_rows, S, dim, D.tb_Q.p, D.tie_qsel.p, tq, D.tie_dmat.p, ldo);
        else
            TI::prefix_dist_matrix_kernel<1><<<grid, 256, 0, st>>>(d_rows, S, dim, D.tb_Q.p, D.tie_qsel.p, tq, D.tie_dmat.p, ldo);
        CU_TRY(cudaGetLastError());
        if (trace) CU_TRY(cudaEventRecord(D.ev[1], st));
        TI::prefix_entrants_kernel<<<tq, 1024, 0, st>>>(D.tie_dmat.p, ldo, (uint32_t)S, k, kcap, D.tie_ent.p, TIE_ENT_CAP);
        CU_TRY(cudaGetLastError());
        if (trace) CU_TRY(cudaEventRecord(D.ev[2], st));
        CU_TRY(cudaMemcpyAsync(D.h_tie_ent.p, D.tie_ent.p, (size_t)tq * region * 8, cudaMemcpyDeviceToHost, st));
        size_t off = 0;
        for (uint32_t j = 0; j < tq; ++j) {
            const uint32_t q = ties[c0 + j], cq = std::min<uint32_t>(h_cnt[q], cap_q);
            if (cq)
                CU_TRY(cudaMemcpyAsync(D.h_tie_seg.p + off, D.tb_seg.p + (size_t)q * cap_q, (size_t)cq * 8, cudaMemcpyDeviceToHost, st));
            off += cq;
        }
        CU_TRY(cudaStreamSynchronize(st));
        const double t_dev = trace ? trace_now_ms() : 0.0;
        // the replays are independent: spread them over a few host threads (each ~0.1 ms: sort by position + heap replay)
        std::vector<size_t> seg_off(tq);
        std::vector<uint32_t> todo;
        off = 0;
        for (uint32_t j = 0; j < tq; ++j) {
            const uint32_t q = ties[c0 + j];
            seg_off[j] = off;
            off += std::min<uint32_t>(h_cnt[q], cap_q);
            if (D.h_tie_ent.p[(size_t)j * region] > TIE_ENT_CAP) slow.push_back(q);
            else todo.push_back(j);
        }
        auto replay_range = [&](size_t lo, size_t hi) {
            std::vector<u64> ent;
            RowMap identity;
            for (size_t t = lo; t < hi; ++t) {
                const uint32_t j = todo[t], q = ties[c0 + j], cq = std::min<uint32_t>(h_cnt[q], cap_q);
                const u64 *reg = D.h_tie_ent.p + (size_t)j * region;
                const u64 *seg = D.h_tie_seg.p + seg_off[j];
                ent.assign(reg + 1, reg + 1 + reg[0]);
                for (uint32_t i = 0; i < cq; ++i)
                    if (key_pos(seg[i]) >= S) ent.push_back(seg[i]);
                out_count[q] = (uint32_t)replay_reference_heap(ent, identity, k, flags, out_rows + (size_t)q * k, out_dist + (size_t)q * k,
                                                                /*known_tie=*/true);
                handled[q] = 1;
            }
        };
        const size_t nt = std::min<size_t>(tie_threads(), (todo.size() + 3) / 4);  // at least 4 replays per thread
        if (nt <= 1) {
            replay_range(0, todo.size());
        } else {
            std::vector<std::thread> th;
            const size_t per = (todo.size() + nt - 1) / nt;
            for (size_t t = 1; t < nt; ++t) th.emplace_back(replay_range, std::min(t * per, todo.size()), std::min((t + 1) * per, todo.size()));
            replay_range(0, std::min(per, todo.size()));
            for (auto &x : th) x.join();
        }
        if (trace) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, D.ev[0], D.ev[1]);
            cudaEventElapsedTime(&b, D.ev[1], D.ev[2]);
            fprintf(stderr, "[pqv trace] tie batch: %u queries, prefix %llu rows: matrix %.3f ms, entrants %.3f ms, enqueue->sync %.3f ms, "
                            "replay %.3f ms on %zu thread(s), %zu overflowed\n",
                    tq, (unsigned long long)S, a, b, t_dev - t_begin, trace_now_ms() - t_dev, std::max<size_t>(nt, 1), tq - todo.size());
        }
    }
    return PQV_OK;
}

// Answers every query it can decide exactly; handled[q] = 0 marks the queries the caller must run through the
// single-query path (ties whose order depends on the reference heap's layout, or a declined batch).
int batch_topk(pqv_ctx *ctx, DeviceState &D, Dataset &ds, u64 n, uint32_t dim, const float *queries, uint32_t nq,
               uint32_t k, uint32_t flags, uint32_t *out_rows, float *out_dist, uint32_t *out_count,
               std::vector<uint8_t> &handled, u64 *raw_keys = nullptr, uint32_t *raw_count = nullptr, uint32_t pos_base = 0,
               const BatchMask *bmask = nullptr, Shard *shard_in = nullptr, pqv_batch_timing *bt_io = nullptr,
               BatchTieOut *tie_out = nullptr) {
    // shard_in + bt_io (raw mode only): one shard of a table spread over several devices, driven from its own host thread
    // -- nothing of the context is written (timing goes to *bt_io, no tie state is kept)
    // bmask != null (batched IVF search): only (row, query) pairs whose row lies in a cluster the query probes count;
    // tie queries are left unhandled (their order follows the IVF candidate sequence, which the caller replays)
    // raw mode (raw_keys != null, the per-rank half of a sharded search): per query the k + 1 smallest exact keys of
    // this slice go to raw_keys[q*(k+1) ..] with pos_base added to the positions, raw_count[q] = how many are valid, or
    // 0xFFFFFFFF when this slice could not decide the query (the caller falls back to the single-query exchange)
    namespace T = pqv::tc;
    Shard &sh = shard_in ? *shard_in : ds.shards[0];
    const float *d_rows = sh.d_data;
    handled.assign(nq, 0);
    pqv_batch_timing &bt = bt_io ? *bt_io : ctx->last_batch;
    bt = pqv_batch_timing{};
    bt.queries = nq;
    bt.rows = n;
    // operand kind: the shard's fp16 shadow when the layout allows it (created on first use), else the f32 rows under tf32
    ShadowView sv;
    int kind = T::KIND_TF32;
    if (shadow_layout_ok(dim, d_rows) && n == sh.n_rows) {
        const int rc = shard_shadow(D, ds, sh, &sv);
        if (rc == PQV_OK) kind = T::KIND_F16;
        else if (rc != PQV_ENOMEM) return rc;  // no room for the shadow: the tf32 form needs none
    }
    // non-finite or huge queries: the reference's NaN/inf heap behaviour is reproduced by the single-query path only
    const float qlim = 1e15f;
    for (size_t i = 0; i < (size_t)nq * dim; ++i)
        if (!(fabsf(queries[i]) < qlim)) {
            bt.declined = 1;
            return PQV_OK;
        }
    const int order = (flags & PQV_SUM_SEQ) ? 1 : 0;
    const bool by_pos = (flags & PQV_TIES_BY_POSITION) != 0;
    const uint32_t nq_pad = (nq + T::BN - 1) / T::BN * T::BN;
    const uint32_t bke = (uint32_t)T::bk_elems(kind);
    const uint32_t num_nb = nq_pad / T::BN, num_kb = (dim + bke - 1) / bke;
    const uint32_t num_mb = (uint32_t)((n + T::BM - 1) / T::BM);
    // sample prefix that fixes theta_q: a query keeps ~k n / S candidates for the exact re-rank, the sample pass costs ~S / n of
    // the filter pass.  Small k affords a shorter prefix (measured at 6.25 M x 1024 x 768, k = 10: n / 16 -> 8.56 ms per batch,
    // n / 24 -> 8.40, n / 32 -> 8.20, n / 48 -> 8.46); PQV_BATCH_SAMPLE_DIV overrides
    static const u64 div_env = getenv("PQV_BATCH_SAMPLE_DIV") ? (u64)std::max(1, atoi(getenv("PQV_BATCH_SAMPLE_DIV"))) : 0;
    const u64 sample_div = div_env ? div_env : (k <= 16 ? 32 : (k <= 48 ? 24 : 16));
    const u64 S = std::min<u64>(n, std::min<u64>(std::max<u64>(n / sample_div, 65536), 524288));
    const uint32_t num_mb_s = (uint32_t)((S + T::BM - 1) / T::BM);
    const uint32_t ldU = num_mb_s * T::BM;
    const TcGrid tg = tc_grid_for(D, num_mb), tg_s = tc_grid_for(D, num_mb_s);
    const uint32_t grid = tg.grid, grid_s = tg_s.grid;
    const u64 rows_per_cta = (u64)tg.tiles_per_cta * T::BM;
    const uint32_t region_cap = (uint32_t)std::min<u64>(rows_per_cta * nq, std::max<u64>(BATCH_TOTAL_CAND / grid, 1024));
    const uint32_t cap_q = (uint32_t)std::min<u64>(n, std::max<u64>(BATCH_TOTAL_CAND / nq, 4096));

    PQV_TRY(D.tb_Q.ensure((size_t)nq * dim));
    PQV_TRY(D.tb_Qp.ensure((size_t)nq * dim));
    PQV_TRY(D.tb_qf.ensure(4 * (size_t)nq_pad));
    PQV_TRY(D.tc_wc.ensure(nq_pad / 32));
    PQV_TRY(D.tb_u32.ensure(4 + (size_t)nq + grid));
    PQV_TRY(D.tb_U.ensure((size_t)nq_pad * ldU));
    PQV_TRY(D.tb_cand.ensure((size_t)grid * region_cap));
    PQV_TRY(D.tb_seg.ensure((size_t)nq * cap_q));
    const uint32_t kout = raw_keys ? k + 1 : k;
    PQV_TRY(D.tb_keys.ensure((size_t)nq * kout));
    PQV_TRY(D.tc_mu.ensure(dim));
    if (kind == T::KIND_TF32 && sh.norms_cap < n) {
        sh.drop_norms();
        cudaError_t e = cudaMalloc((void **)&sh.d_norms, (size_t)n * sizeof(float2));
        if (e != cudaSuccess) return fail(PQV_ENOMEM, "row-norm cache of %llu rows: %s", (unsigned long long)n, cudaGetErrorString(e));
        sh.norms_cap = n;
    }
    PQV_TRY(D.h_batch_keys.ensure((size_t)nq * kout + nq + 2));
    float *qw = D.tb_qf.p, *q2 = D.tb_qf.p + nq_pad, *qtheta = D.tb_qf.p + 2 * (size_t)nq_pad, *qT = D.tb_qf.p + 3 * (size_t)nq_pad;
    uint32_t *qbounds = D.tb_u32.p, *dflags = D.tb_u32.p + 1, *cntq = D.tb_u32.p + 4, *region_count = D.tb_u32.p + 4 + nq;
    PQV_TRY(D.tb_info.ensure(nq));

    CUtensorMap tmAs, tmA, tmB;
    const void *a_ptr = kind == T::KIND_F16 ? (const void *)sv.h : (const void *)d_rows;
    PQV_TRY(make_row_tmap(&tmAs, a_ptr, S, dim, T::BM, kind));
    PQV_TRY(make_row_tmap(&tmA, a_ptr, n, dim, T::BM, kind));
    PQV_TRY(make_row_tmap(&tmB, D.tb_Qp.p, nq, dim, tg.pair ? T::BN / 2 : T::BN, kind));

    cudaStream_t st = D.stream;
    CU_TRY(cudaEventRecord(D.ev[0], st));
    CU_TRY(cudaMemsetAsync(D.tb_u32.p, 0, (4 + (size_t)nq + grid) * sizeof(uint32_t), st));
    CU_TRY(cudaMemsetAsync(D.tc_wc.p, 0, (size_t)(nq_pad / 32) * sizeof(uint32_t), st));
    CU_TRY(cudaMemcpyAsync(D.tb_Q.p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st));
    // operand scale of the queries: qbounds[2] = max |q| (fp16 only), then the power of two derived from it (1 for tf32)
    if (kind == T::KIND_F16) pqv::half16::absmax_kernel<<<(uint32_t)D.sm_count, 256, 0, st>>>(D.tb_Q.p, (u64)nq * dim, qbounds + 2);
    pqv::half16::scale_from_absmax_kernel<<<1, 32, 0, st>>>(qbounds + 2, reinterpret_cast<float *>(qbounds + 2));
    if (kind == T::KIND_F16) {
        T::query_prep_kernel<T::KIND_F16><<<nq_pad, 128, 0, st>>>(D.tb_Q.p, nq, dim, D.tb_Qp.p, qw, q2, D.tc_wc.p, nq_pad, qbounds, sv.g);
    } else {
        T::query_prep_kernel<T::KIND_TF32><<<nq_pad, 128, 0, st>>>(D.tb_Q.p, nq, dim, D.tb_Qp.p, qw, q2, D.tc_wc.p, nq_pad, qbounds, nullptr);
        if (sh.norms_rows != n) {  // |x|^2 per row: once per dataset state, not per batch
            CU_TRY(cudaMemsetAsync(D.tc_mu.p, 0, (size_t)dim * sizeof(float), st));
            T::row_stats_kernel<<<(uint32_t)D.sm_count * 8, 256, 0, st>>>(d_rows, n, dim, D.tc_mu.p, sh.d_norms);
            sh.norms_rows = n;
        }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(D.ev[1], st));

    T::BatchParams p;
    p.stats = kind == T::KIND_F16 ? sv.stats : sh.d_norms;
    p.qw = qw;
    p.qtheta = qtheta;
    p.qwc = reinterpret_cast<const float *>(D.tc_wc.p);
    p.qbounds = qbounds;
    p.scale_a = kind == T::KIND_F16 ? &sv.g->scale : nullptr;
    p.U = D.tb_U.p;
    p.ldU = ldU;
    p.cand = D.tb_cand.p;
    p.region_cap = region_cap;
    p.region_count = region_count;
    p.flags = dflags;
    p.dim = dim;
    p.row_cluster = bmask ? bmask->row_cluster : nullptr;
    p.probe_T = bmask ? bmask->probe_T : nullptr;
    p.qwords = bmask ? bmask->qwords : 0;
    p.row_mask = bmask ? bmask->row_mask : nullptr;
    // phase A: upper bounds over the first S rows -> theta_q
    p.n = S;
    PQV_TRY(tc_launch<T::BatchEpi<T::BATCH_SAMPLE>>(tg_s.pair, kind, grid_s, st, tmAs, tmB, T::GemmShape{num_mb_s, num_nb, num_kb}, p));
    const float delta = (float)(order == 1 ? dim + 8 : dim / 4 + 12) * 5.9604645e-08f;
    {
        const uint32_t M = k <= 128 ? 2048u : 16384u;  // chunk minima kept per query (>= 16 k)
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(T::theta_select_kernel), (size_t)16384 * 4));
        T::theta_select_kernel<<<nq_pad, 256, M * sizeof(float), st>>>(D.tb_U.p, ldU, (uint32_t)S, k, nq, q2, delta, M, qtheta, qT);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(D.ev[2], st));
    // phase B: candidates over all rows
    p.n = n;
    PQV_TRY(tc_launch<T::BatchEpi<T::BATCH_FILTER>>(tg.pair, kind, grid, st, tmA, tmB, T::GemmShape{num_mb, num_nb, num_kb}, p));
    CU_TRY(cudaEventRecord(D.ev[3], st));
    // exact distances of the candidates, per-query selection
    if (order == 0)
        T::pair_dist_kernel<0><<<dim3(8, grid), T::PairCfg<0>::WARPS * 32, 0, st>>>(d_rows, dim, D.tb_Q.p, D.tb_cand.p, region_cap,
                                                                                   region_count, D.tb_seg.p, cap_q, cntq);
    else
        T::pair_dist_kernel<1><<<dim3(32, grid), T::PairCfg<1>::WARPS * 32, 0, st>>>(d_rows, dim, D.tb_Q.p, D.tb_cand.p, region_cap,
                                                                                    region_count, D.tb_seg.p, cap_q, cntq);
    T::topk_select_kernel<<<nq, 256, 0, st>>>(D.tb_seg.p, cap_q, cntq, k, (flags & PQV_SQRT) ? 1 : 0, D.tb_keys.p, D.tb_info.p,
                                              kout);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(D.ev[4], st));
    u64 *h_keys = D.h_batch_keys.p;
    uint32_t *h_info = reinterpret_cast<uint32_t *>(h_keys + (size_t)nq * kout);
    uint32_t *h_flags = h_info + nq;  // + region counts are not needed on the host
    CU_TRY(cudaMemcpyAsync(h_keys, D.tb_keys.p, (size_t)nq * kout * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h_info, D.tb_info.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h_flags, dflags, 4, cudaMemcpyDeviceToHost, st));
    std::vector<uint32_t> h_cnt(nq);
    CU_TRY(cudaMemcpyAsync(h_cnt.data(), cntq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    std::vector<float> h_qT;
    if (tie_out) {
        h_qT.resize(nq);
        CU_TRY(cudaMemcpyAsync(h_qT.data(), qT, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
    float ms[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], D.ev[i], D.ev[i + 1]);
    bt.prep_ms = ms[0];
    bt.sample_ms = ms[1];
    bt.filter_ms = ms[2];
    bt.rerank_ms = ms[3];
    bt.total_ms = (double)ms[0] + ms[1] + ms[2] + ms[3];
    bt.sample_rows = S;
    for (uint32_t q = 0; q < nq; ++q) bt.candidates += h_cnt[q];
    if (*h_flags) {  // non-finite rows or a full candidate region: nothing of this batch is trusted
        bt.declined = 1;
        return PQV_OK;
    }
    if (raw_keys) {
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t info = h_info[q];
            if (info & T::SEL_OVERFLOW) {
                raw_count[q] = 0xFFFFFFFFu;
                continue;
            }
            const uint32_t cnt = info & 0xFFFFu;
            for (uint32_t i = 0; i < cnt; ++i) raw_keys[(size_t)q * kout + i] = h_keys[(size_t)q * kout + i] + pos_base;
            raw_count[q] = cnt;
            handled[q] = 1;
        }
        if (shard_in) return PQV_OK;
        pqv_ctx::BatchState &bs = ctx->batch_state;  // the per-query candidate segments stay on the device for tie queries
        bs.valid = true;
        bs.S = S;
        bs.nq = nq;
        bs.k = k;
        bs.flags = flags;
        bs.cap_q = cap_q;
        bs.pos_base = pos_base;
        bs.dev_index = (int)(&D - ctx->devs.data());
        bs.seg_count = h_cnt;
        return PQV_OK;
    }
    std::vector<uint32_t> ties;
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t info = h_info[q];
        if (info & T::SEL_OVERFLOW) continue;  // candidate list incomplete: the caller runs the full single-query scan
        if ((info & T::SEL_TIE) && !by_pos) {
            if (!bmask) {
                ties.push_back(q);
            } else if (tie_out && h_cnt[q] <= cap_q) {
                tie_out->queries.push_back(q);
                tie_out->cnt.push_back(h_cnt[q]);
                tie_out->T.push_back(h_qT[q]);
                tie_out->cap_q = cap_q;
            }
            continue;
        }
        const uint32_t cnt = info & 0xFFFFu;
        for (uint32_t i = 0; i < cnt; ++i) {
            const u64 key = h_keys[(size_t)q * k + i];
            const float d = key_dist(key);
            out_rows[(size_t)q * k + i] = key_pos(key);
            out_dist[(size_t)q * k + i] = (flags & PQV_SQRT) ? sqrtf(d) : d;
        }
        out_count[q] = cnt;
        handled[q] = 1;
    }
    // Tie queries: the order (or the kept set) hinges on the layout of the reference's BinaryHeap, so its push sequence is
    // replayed.  Rows the heap ever admits at positions >= S are all among the query's candidates
    // (theta_q bounds the k-th smallest distance of the first S rows, the heap's threshold from position S on), and the
    // admissions inside [0, S) come from the exact single-query scan of that prefix alone.
    bt.tie_queries = (uint32_t)ties.size();
    std::vector<u64> ent, seg_keys;
    RowMap identity;
    std::vector<uint32_t> slow;
    if (tie_batch_enabled() && !ties.empty())
        PQV_TRY(resolve_ties_together(D, d_rows, S, dim, order, k, flags, ties, h_cnt, cap_q, out_rows, out_dist, out_count, handled,
                                      slow));
    else
        slow.swap(ties);
    bt.tie_batched = bt.tie_queries - (uint32_t)slow.size();
    for (uint32_t q : slow) {
        ent.clear();
        uint32_t dummy_cnt = 0;
        PQV_TRY(topk_one(ctx, ds, queries + (size_t)q * dim, nullptr, 0, k, flags, nullptr, nullptr, &dummy_cnt, &ent, 0, nullptr,
                         nullptr, S));
        const uint32_t cq = std::min<uint32_t>(h_cnt[q], cap_q);
        seg_keys.resize(cq);
        CU_TRY(cudaMemcpy(seg_keys.data(), D.tb_seg.p + (size_t)q * cap_q, (size_t)cq * 8, cudaMemcpyDeviceToHost));
        for (u64 key : seg_keys)
            if (key_pos(key) >= S) ent.push_back(key);
        out_count[q] = (uint32_t)replay_reference_heap(ent, identity, k, flags, out_rows + (size_t)q * k, out_dist + (size_t)q * k);
        handled[q] = 1;
    }
    return PQV_OK;
}

}  // namespace
