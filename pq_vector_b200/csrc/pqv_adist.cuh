// pqv_adist.cuh -- the un-indexed `array_distance` arm.
//
// `ORDER BY array_distance(col, [literal]) LIMIT k` over a file WITHOUT an embedded index is not rewritten by
// pq-vector's optimizer rule (src/df_vector/physical.rs:198-214 only fires when the index is present), so the stock
// DataFusion plan runs: the built-in UDF `array_distance` (crate datafusion-functions-nested 52.1.0, Cargo.lock:1041-1042;
// NOT under /root/reference) evaluated per row, followed by SortExec(TopK).  Call sites that reach it:
// benches/query.rs:79-81 (the ground-truth arm of the recall printout), examples/datafusion_sql.rs:54-55.
//
// Upstream algorithm restated from its published source (no copy available here -- PARITY UNPINNED):
//   both lists are cast to Float64; sum = fold over elements, in order, of (v1 - v2)^2 in f64; result = sqrt(sum).
// The per-row sum is again a serial chain, so the kernel has the same shape as group_distance (pqv_kernels.cuh): a
// warp owns 32 rows, all lanes load each row coalesced and compute the independent terms (f32 -> f64 widen,
// subtract, square -- never contracted), the terms are transposed through a padded shared-memory tile, and lane r
// runs row r's serial f64 chain.  HBM is read once; FP64 work is 4 operations per element (B200: 64 FP64 lanes per
// SM and clock => 1.7 ms per 10 M x 768, below the 4.7 ms the bytes take).
//
// Cosine distance (BASELINE north_star names it; the reference has none) is additive and shares the kernel:
//   1 - dot(x, q) / (sqrt(|x|^2) * sqrt(|q|^2)), the three sums folded sequentially in f64.
//
// Top-k of the Float64 column = what SortExec(TopK) keeps: the k smallest by f64 total order (NaN last), and -- where
// the stock operator's order among equal keys is unspecified -- ties by ascending row.  Exact radix select over the
// 96-bit composite key (ordered distance bits, row): 12 histogram passes over the 8-byte column, no host round trip.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqv_kernels.cuh"

namespace pqv {

constexpr int ADIST_L2 = 0;      // sqrt(sum (x - q)^2)         (DataFusion array_distance)
constexpr int ADIST_COSINE = 1;  // 1 - x.q / (|x| |q|)          (additive; no reference semantics)

constexpr int ADIST_BLK = 64;      // columns per block: a warp reads 256 contiguous bytes of each row per request
constexpr int ADIST_TSTRIDE = 68;  // floats per tile row: 64 + 4; 272 B = 16 x 17 -> LDS.128 by row is conflict-free
constexpr int ADIST_WARPS = 8;
constexpr int ADIST_TILE_BYTES = 32 * ADIST_TSTRIDE * 4;  // 8704 B per warp
constexpr uint32_t ADIST_QSMEM_MAX_DIM = 4096;            // f64 query staged in shared memory up to here (32 KB)

// One warp owns 32 consecutive rows.  Per 64-column block all lanes load the rows coalesced (16 rows x 2 halves kept in
// registers), park the RAW f32 values in the warp's padded shared-memory tile, and lane r then walks row r: widen to
// f64, subtract, square (independent, they run ahead) and the serial f64 add chain in element order.  The loads of the
// NEXT block (or of the next group's first block) are issued before the chain starts, so every warp has 8 KB in flight
// while it computes.  Earlier form (f64 terms transposed instead of raw f32): twice the tile, 178 registers, 8 warps
// per SM -> 3.6 TB/s at best (profiles/r01_adist_*.md).
//   VEC2  : dim even and 8-byte aligned rows -> 64-bit loads; otherwise two scalar loads per lane and row.
//   QSMEM : f64 query staged in shared memory (dim <= ADIST_QSMEM_MAX_DIM); otherwise read through L1 (uniform address).
template <int METRIC, bool VEC2, bool QSMEM>
__global__ void __launch_bounds__(ADIST_WARPS * 32)
array_distance_kernel(const float *__restrict__ data, const u64 n, const uint32_t dim, const double *__restrict__ query,
                      const double q_norm2, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int HB = 16;  // rows per register half
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t dim_pad = (dim + ADIST_BLK - 1) & ~(uint32_t)(ADIST_BLK - 1);
    double *s_q = reinterpret_cast<double *>(smem_raw);
    float *tile = reinterpret_cast<float *>(smem_raw + (QSMEM ? (size_t)dim_pad * 8 : 0)) + (size_t)warp * 32 * ADIST_TSTRIDE;
    const float *trow = tile + lane * ADIST_TSTRIDE;
    if (QSMEM) {
        for (uint32_t i = tid; i < dim_pad; i += ADIST_WARPS * 32) s_q[i] = i < dim ? query[i] : 0.0;
        __syncthreads();
    }
    const u64 NG = (n + 31) >> 5;
    const u64 gstride = (u64)gridDim.x * ADIST_WARPS;
    const uint32_t ncb = dim_pad / ADIST_BLK;
    u64 g = (u64)blockIdx.x * ADIST_WARPS + warp;
    if (g >= NG) return;

    float2 va[HB], vb[HB];
    auto issue_half = [&](const u64 gg, const uint32_t cb, const int r0, float2(&v)[HB]) {
        // gg >= NG (nothing left for this warp): every load is predicated off
        const uint32_t c0 = cb * ADIST_BLK + (VEC2 ? 2 * lane : lane);
        const bool live = gg < NG;
        const bool in0 = live && c0 < dim;
        const bool in1 = live && (VEC2 ? in0 : c0 + 32 < dim);
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            u64 row = gg * 32 + (r0 + j);
            row = row < n ? row : n - 1;
            const float *rp = data + row * dim + c0;
            v[j] = make_float2(0.f, 0.f);
            if (VEC2) {
                if (in0) v[j] = ld_stream_v2(rp);
            } else {
                if (in0) v[j].x = ld_stream_f32(rp);
                if (in1) v[j].y = ld_stream_f32(rp + 32);
            }
        }
    };
    auto park_half = [&](const int r0, const float2(&v)[HB]) {
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            float *tr = tile + (r0 + j) * ADIST_TSTRIDE;
            if (VEC2) {
                *reinterpret_cast<float2 *>(tr + 2 * lane) = v[j];
            } else {
                tr[lane] = v[j].x;
                tr[32 + lane] = v[j].y;
            }
        }
    };
    double acc0, acc1;  // L2: acc0 = sum; cosine: acc0 = dot, acc1 = |x|^2
    auto step = [&](const float xf, const double q) {
        const double x = (double)xf;
        if (METRIC == ADIST_COSINE) {
            acc0 = __dadd_rn(acc0, __dmul_rn(x, q));
            acc1 = __dadd_rn(acc1, __dmul_rn(x, x));
        } else {
            const double d = __dsub_rn(x, q);
            acc0 = __dadd_rn(acc0, __dmul_rn(d, d));
        }
    };

    issue_half(g, 0, 0, va);
    issue_half(g, 0, HB, vb);
    for (;;) {
        acc0 = 0.0;
        acc1 = 0.0;
#pragma unroll 1
        for (uint32_t cb = 0; cb < ncb; ++cb) {
            park_half(0, va);
            park_half(HB, vb);
            __syncwarp();
            const bool last = cb + 1 == ncb;
            const u64 ng = last ? g + gstride : g;
            const uint32_t ncb_next = last ? 0u : cb + 1;
            issue_half(ng, ncb_next, 0, va);
            issue_half(ng, ncb_next, HB, vb);
            // serial chain in element order, lane = row
            const uint32_t c0 = cb * ADIST_BLK;
            const uint32_t cnt = dim - c0 < (uint32_t)ADIST_BLK ? dim - c0 : (uint32_t)ADIST_BLK;
            if (cnt == (uint32_t)ADIST_BLK) {
#pragma unroll
                for (int jj = 0; jj < ADIST_BLK / 4; ++jj) {
                    const float4 x = *reinterpret_cast<const float4 *>(trow + 4 * jj);
                    double2 qa, qb;
                    if (QSMEM) {
                        qa = *reinterpret_cast<const double2 *>(s_q + c0 + 4 * jj);
                        qb = *reinterpret_cast<const double2 *>(s_q + c0 + 4 * jj + 2);
                    } else {
                        qa = make_double2(__ldg(query + c0 + 4 * jj), __ldg(query + c0 + 4 * jj + 1));
                        qb = make_double2(__ldg(query + c0 + 4 * jj + 2), __ldg(query + c0 + 4 * jj + 3));
                    }
                    step(x.x, qa.x);
                    step(x.y, qa.y);
                    step(x.z, qb.x);
                    step(x.w, qb.y);
                }
            } else {
                for (uint32_t e = 0; e < cnt; ++e) step(trow[e], QSMEM ? s_q[c0 + e] : __ldg(query + c0 + e));
            }
            __syncwarp();
        }
        const u64 row = g * 32 + lane;
        if (row < n) {
            double r;
            if (METRIC == ADIST_COSINE) r = __dsub_rn(1.0, __ddiv_rn(acc0, __dmul_rn(__dsqrt_rn(acc1), __dsqrt_rn(q_norm2))));
            else r = __dsqrt_rn(acc0);
            out[row] = r;
        }
        g += gstride;
        if (g >= NG) break;
    }
}

// ------------------------------------------------------------------------------------------------
// exact top-k of an f64 column: radix select over (ordered bits, row)
// ------------------------------------------------------------------------------------------------
// f64 total order with every NaN last (DataFusion sorts NaN above +inf; the sign of a NaN is platform noise)
__device__ __forceinline__ u64 f64_ordered_bits(const double d) {
    u64 b = (u64)__double_as_longlong(d);
    if ((b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) b = 0x7FF8000000000000ull;
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct SelState {
    u64 prefix;           // ordered-bits digits fixed so far (passes 0..7, high byte first)
    uint32_t prefix_row;  // row digits fixed so far (passes 8..11)
    uint32_t k_rem;       // rank still to resolve inside the current prefix class (1-based)
    uint32_t out_count;
    uint32_t pad;
    uint32_t hist[256];
};

__global__ void sel_init_kernel(SelState *st, const uint32_t k) {
    const uint32_t t = threadIdx.x;
    if (t == 0) {
        st->prefix = 0;
        st->prefix_row = 0;
        st->k_rem = k;
        st->out_count = 0;
    }
    if (t < 256) st->hist[t] = 0;
}

// pass 0..7: digit = byte (7 - pass) of the ordered bits among keys whose higher bytes equal the prefix;
// pass 8..11: digit = byte (11 - pass) of the row among keys with bits == prefix and higher row bytes equal.
// row_mask (may be null): bit r set = row r takes part (the scan subtree's filter of a stock-plan query)
__device__ __forceinline__ bool sel_row_live(const uint32_t *__restrict__ row_mask, const u64 i) {
    return !row_mask || ((row_mask[i >> 5] >> (i & 31)) & 1u);
}

__global__ void __launch_bounds__(256) sel_hist_kernel(const double *__restrict__ col, const u64 n, const int pass, SelState *st,
                                                       const uint32_t *__restrict__ row_mask) {
    __shared__ uint32_t s_hist[256];
    const uint32_t tid = threadIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    const u64 prefix = st->prefix;
    const uint32_t prow = st->prefix_row;
    const u64 n_round = (n + 31) & ~31ull;  // whole warps stay converged for match_any
    for (u64 i = (u64)blockIdx.x * 256 + tid; i < n_round; i += (u64)gridDim.x * 256) {
        bool take = false;
        uint32_t digit = 0;
        if (i < n && sel_row_live(row_mask, i)) {
            const u64 u = f64_ordered_bits(col[i]);
            if (pass < 8) {
                const int sh = 8 * (7 - pass);
                take = pass == 0 || (u >> (sh + 8)) == (prefix >> (sh + 8));
                digit = (uint32_t)(u >> sh) & 255u;
            } else {
                const int sh = 8 * (11 - pass);
                const uint32_t r = (uint32_t)i;
                take = u == prefix && (pass == 8 || (r >> (sh + 8)) == (prow >> (sh + 8)));
                digit = (r >> sh) & 255u;
            }
        }
        // warp-aggregated: distances of one table share their high bytes, so most lanes hit one bin
        const uint32_t key = take ? digit : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (take && (uint32_t)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&s_hist[digit], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (s_hist[tid]) atomicAdd(&st->hist[tid], s_hist[tid]);
}

__global__ void sel_pick_kernel(const int pass, SelState *st) {
    __shared__ uint32_t s_h[256];
    const uint32_t t = threadIdx.x;
    s_h[t] = st->hist[t];
    st->hist[t] = 0;
    __syncthreads();
    if (t == 0) {
        uint32_t k_rem = st->k_rem, cum = 0, b = 0;
        for (; b < 255; ++b) {
            if (cum + s_h[b] >= k_rem) break;
            cum += s_h[b];
        }
        st->k_rem = k_rem - cum;
        if (pass < 8) st->prefix |= (u64)b << (8 * (7 - pass));
        else st->prefix_row |= b << (8 * (11 - pass));
    }
}

// every key <= (prefix, prefix_row): exactly min(k, n) of them, in arbitrary order (the host sorts k items)
__global__ void __launch_bounds__(256) sel_collect_kernel(const double *__restrict__ col, const u64 n, SelState *st,
                                                          const uint32_t cap, double *__restrict__ out_dist,
                                                          uint32_t *__restrict__ out_row,
                                                          const uint32_t *__restrict__ row_mask) {
    const u64 prefix = st->prefix;
    const uint32_t prow = st->prefix_row;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        if (!sel_row_live(row_mask, i)) continue;
        const double d = col[i];
        const u64 u = f64_ordered_bits(d);
        if (u < prefix || (u == prefix && (uint32_t)i <= prow)) {
            const uint32_t slot = atomicAdd(&st->out_count, 1u);
            if (slot < cap) {
                out_dist[slot] = d;
                out_row[slot] = (uint32_t)i;
            }
        }
    }
}

}  // namespace pqv
