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

constexpr int ADIST_TSTRIDE = 66;  // doubles per tile row: 64 + 2; 528 B = 16 x 33 -> LDS.128 conflict-free
constexpr int ADIST_WARPS = 4;
constexpr int ADIST_TILE_BYTES = 32 * ADIST_TSTRIDE * 8;

__device__ __forceinline__ float2 ld_stream_v2(const float *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// MODE 0: L2, two columns per lane (dim even, 8-byte aligned rows)      tile pair = (t(col), t(col+1))
// MODE 1: L2, one column per lane (any dim / alignment)                 tile pair = (t(col), unused)
// MODE 2: cosine, one column per lane                                   tile pair = (x*q, x*x)
template <int MODE, int RB>
__global__ void __launch_bounds__(ADIST_WARPS * 32)
array_distance_kernel(const float *__restrict__ data, const u64 n, const uint32_t dim, const double *__restrict__ query,
                      const double q_norm2, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t CPL = (MODE == 0) ? 2u : 1u;  // columns per lane per block
    constexpr uint32_t BLK = 32u * CPL;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tile = reinterpret_cast<double *>(smem_raw) + (size_t)warp * 32 * ADIST_TSTRIDE;
    const double *trow = tile + lane * ADIST_TSTRIDE;
    const u64 NG = (n + 31) >> 5;
    const uint32_t ncb = (dim + BLK - 1) / BLK;

    for (u64 g = (u64)blockIdx.x * ADIST_WARPS + warp; g < NG; g += (u64)gridDim.x * ADIST_WARPS) {
        const u64 g_first = g * 32;
        double acc0 = 0.0, acc1 = 0.0;  // L2: acc0 = sum; cosine: acc0 = dot, acc1 = |x|^2
        for (uint32_t cb = 0; cb < ncb; ++cb) {
            const uint32_t col = cb * BLK + lane * CPL;
            const bool inb = col < dim;
            double q0 = 0.0, q1 = 0.0;
            if (inb) {
                q0 = __ldg(query + col);
                if (MODE == 0) q1 = __ldg(query + col + 1);
            }
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += RB) {
                float2 v[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    u64 row = g_first + (r0 + j);
                    row = row < n ? row : n - 1;
                    const float *rp = data + row * dim + col;
                    v[j] = make_float2(0.f, 0.f);
                    if (inb) {
                        if (MODE == 0) v[j] = ld_stream_v2(rp);
                        else v[j].x = ld_stream_f32(rp);
                    }
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    double2 t;
                    if (MODE == 2) {
                        const double x = (double)v[j].x;
                        t.x = __dmul_rn(x, q0);
                        t.y = __dmul_rn(x, x);
                    } else {
                        const double d0 = __dsub_rn((double)v[j].x, q0);
                        t.x = __dmul_rn(d0, d0);
                        const double d1 = __dsub_rn((double)v[j].y, q1);
                        t.y = (MODE == 0) ? __dmul_rn(d1, d1) : 0.0;
                    }
                    *reinterpret_cast<double2 *>(tile + (r0 + j) * ADIST_TSTRIDE + 2 * lane) = t;
                }
            }
            __syncwarp();
            // serial chain, lane = row; cnt = lanes of this block that hold real columns
            const uint32_t left = dim - cb * BLK;
            const uint32_t cnt = left >= BLK ? 32u : (left + CPL - 1) / CPL;
            if (cnt == 32u) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double2 t = *reinterpret_cast<const double2 *>(trow + 2 * j);
                    acc0 = __dadd_rn(acc0, t.x);
                    if (MODE == 0) acc0 = __dadd_rn(acc0, t.y);
                    if (MODE == 2) acc1 = __dadd_rn(acc1, t.y);
                }
            } else {
                for (uint32_t j = 0; j < cnt; ++j) {
                    const double2 t = *reinterpret_cast<const double2 *>(trow + 2 * j);
                    acc0 = __dadd_rn(acc0, t.x);
                    if (MODE == 0) acc0 = __dadd_rn(acc0, t.y);  // MODE 0 has dim even: both columns are real
                    if (MODE == 2) acc1 = __dadd_rn(acc1, t.y);
                }
            }
            __syncwarp();
        }
        const u64 row = g_first + lane;
        if (row < n) {
            double r;
            if (MODE == 2) r = __dsub_rn(1.0, __ddiv_rn(acc0, __dmul_rn(__dsqrt_rn(acc1), __dsqrt_rn(q_norm2))));
            else r = __dsqrt_rn(acc0);
            out[row] = r;
        }
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
__global__ void __launch_bounds__(256) sel_hist_kernel(const double *__restrict__ col, const u64 n, const int pass, SelState *st) {
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
        if (i < n) {
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
                                                          uint32_t *__restrict__ out_row) {
    const u64 prefix = st->prefix;
    const uint32_t prow = st->prefix_row;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
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
