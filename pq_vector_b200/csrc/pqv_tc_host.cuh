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

// row-major [rows][dim] f32 table viewed as a 2-D tensor {dim (inner), rows}; box = 32 columns x box_rows rows, 128 B swizzle;
// out-of-bounds elements (row tail, dim % 32 tail) are filled with zeros and still count towards the transaction bytes
int make_row_tmap(CUtensorMap *tm, const float *d_ptr, u64 rows, uint32_t dim, uint32_t box_rows) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return fail(PQV_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[2] = {dim, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * 4};
    cuuint32_t box[2] = {(cuuint32_t)pqv::tc::BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d_ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PQV_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu dim=%u)", (int)r,
                                       (unsigned long long)rows, dim);
    return PQV_OK;
}

enum AssignPath { ASSIGN_SIMT = 0, ASSIGN_TC = 1 };

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

int assign_tc(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, const float *d_cent, uint32_t C, uint32_t *d_out,
              bool time_it) {
    namespace T = pqv::tc;
    const uint32_t num_mb = (uint32_t)((n + T::BM - 1) / T::BM);
    const uint32_t num_nb = (C + T::BN - 1) / T::BN;
    const uint32_t num_kb = (dim + T::BK - 1) / T::BK;
    const uint32_t cn_len = num_nb * T::BN;
    const u64 pair_cap64 = std::min<u64>(4 * n + 1024, 0xFFFFFFF0ull);
    PQV_TRY(D.tc_bp.ensure((size_t)C * dim));
    PQV_TRY(D.tc_mu.ensure(dim));
    PQV_TRY(D.tc_cn.ensure(2 * (size_t)cn_len));
    PQV_TRY(D.tc_stats.ensure(n));
    PQV_TRY(D.tc_u32.ensure(8));
    PQV_TRY(D.tc_amb_rows.ensure(n));
    PQV_TRY(D.tc_pairs.ensure(pair_cap64));
    PQV_TRY(D.tc_best.ensure(n));
    PQV_TRY(D.tc_ovf_rows.ensure(n));
    uint32_t *bounds = D.tc_u32.p, *counts = D.tc_u32.p + 4;

    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(T::assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES);
    });
    CU_TRY(attr_err);

    CUtensorMap tmA, tmB;
    PQV_TRY(make_row_tmap(&tmA, d_rows, n, dim, T::BM));
    PQV_TRY(make_row_tmap(&tmB, D.tc_bp.p, C, dim, T::BN));

    if (time_it) CU_TRY(cudaEventRecord(D.ev[0], D.stream));
    CU_TRY(cudaMemsetAsync(D.tc_u32.p, 0, 8 * sizeof(uint32_t), D.stream));
    T::centroid_mean_kernel<<<(dim + 127) / 128, 128, 0, D.stream>>>(d_cent, C, dim, D.tc_mu.p, bounds);
    T::centroid_prep_kernel<<<cn_len, 128, 0, D.stream>>>(d_cent, C, dim, D.tc_mu.p, D.tc_bp.p, D.tc_cn.p, D.tc_cn.p + cn_len,
                                                              cn_len, bounds);
    T::row_stats_kernel<<<(uint32_t)D.sm_count * 8, 256, 0, D.stream>>>(d_rows, n, dim, D.tc_mu.p, D.tc_stats.p);
    CU_TRY(cudaGetLastError());

    T::AssignTcParams p;
    p.stats = D.tc_stats.p;
    p.cn = D.tc_cn.p;
    p.wv = D.tc_cn.p + cn_len;
    p.bounds = bounds;
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
    p.num_mb = num_mb;
    p.num_nb = num_nb;
    p.num_kb = num_kb;
    const uint32_t grid = std::min<uint32_t>(num_mb, (uint32_t)D.sm_count);
    if (time_it) CU_TRY(cudaEventRecord(D.ev[1], D.stream));
    T::assign_tc_kernel<<<grid, T::THREADS, T::SMEM_BYTES, D.stream>>>(tmA, tmB, p);
    CU_TRY(cudaGetLastError());
    if (time_it) CU_TRY(cudaEventRecord(D.ev[2], D.stream));
    // exact f32 chains: one per (row, candidate) pair of the ambiguous rows; the full scan for the overflow rows
    T::pair_exact_kernel<<<(uint32_t)D.sm_count * 4, T::PAIR_WARPS * 32, 0, D.stream>>>(d_rows, dim, d_cent, counts, p.pair_cap,
                                                                                     D.tc_pairs.p, D.tc_best.p);
    if (time_it) CU_TRY(cudaEventRecord(D.ev[4], D.stream));
    {
        const uint32_t slice_len = 2 * pqv::AS_BN;
        dim3 grid_ovf((uint32_t)((n + pqv::AS_BM - 1) / pqv::AS_BM), (C + slice_len - 1) / slice_len);
        pqv::kmeans_assign_kernel<true, true><<<grid_ovf, 256, 0, D.stream>>>(d_rows, 0, dim, d_cent, C, nullptr, D.tc_ovf_rows.p,
                                                                              counts + 1, D.tc_best.p, slice_len);
    }
    T::best_finalize_kernel<<<(uint32_t)D.sm_count * 2, 256, 0, D.stream>>>(counts, D.tc_amb_rows.p, D.tc_ovf_rows.p, D.tc_best.p,
                                                                            d_out);
    CU_TRY(cudaGetLastError());
    if (time_it) CU_TRY(cudaEventRecord(D.ev[3], D.stream));
    return PQV_OK;
}

// Enqueues the assignment of n device rows on D.stream (no synchronisation).  *path_out reports the path taken.
int assign_dispatch(DeviceState &D, const float *d_rows, u64 n, uint32_t dim, const float *d_cent, uint32_t C, uint32_t *d_out,
                    int *path_out = nullptr, bool time_it = false) {
    const int path = assign_path_for(d_rows, n, dim, d_cent, C);
    if (path_out) *path_out = path;
    if (path == ASSIGN_TC) return assign_tc(D, d_rows, n, dim, d_cent, C, d_out, time_it);
    if (time_it) CU_TRY(cudaEventRecord(D.ev[0], D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[1], D.stream));
    PQV_TRY(assign_simt(D, d_rows, n, dim, d_cent, C, d_out));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[2], D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[4], D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[3], D.stream));
    return PQV_OK;
}

// Reads the events recorded by assign_dispatch(time_it = true) into ctx->last_assign; the stream must be idle.
void record_assign_timing(pqv_ctx *ctx, DeviceState &D, int path, u64 rows, const uint32_t counts[2], bool first_piece) {
    float prep = 0.f, filt = 0.f, post = 0.f;
    cudaEventElapsedTime(&prep, D.ev[0], D.ev[1]);
    cudaEventElapsedTime(&filt, D.ev[1], D.ev[2]);
    cudaEventElapsedTime(&post, D.ev[2], D.ev[3]);
    float pair_ms = 0.f;
    cudaEventElapsedTime(&pair_ms, D.ev[2], D.ev[4]);
    pqv_assign_timing &t = ctx->last_assign;
    if (first_piece) t = pqv_assign_timing{};
    t.path = (uint32_t)path;
    t.rows += rows;
    t.ambiguous_rows += counts[0];
    t.overflow_rows += counts[1];
    t.prep_ms += prep;
    t.filter_ms += filt;
    t.recheck_ms += post;
    t.pair_ms += pair_ms;
    t.total_ms += (double)prep + filt + post;
}

}  // namespace
