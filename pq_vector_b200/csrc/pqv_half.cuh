// pqv_half.cuh -- the 16-bit operand shadow of a resident table (sm_100a).
//
// The tensor-core FILTERS of this library (assignment sweeps src/ivf/index.rs:189-206, 395-430; batched top-k over
// src/ivf/search.rs:112-141 / src/df_vector/exec.rs:257-277) never decide a result from the tensor-core value alone: every
// emitted index / distance comes from the exact serial f32 chain over the ORIGINAL f32 rows.  So the operand the tensor
// cores read may be narrower than f32 as long as its error is accounted for -- and ncu shows the f32/tf32 filter bound by
// the L2 -> shared-memory operand fill (10.4 TB/s for 32 KB per 128 x 256 x 32 block and SM), not by the tensor pipe.  A
// 2-byte operand halves that fill and doubles the MMA rate (kind::f16).
//
// fp16, not bf16: the filter's ambiguity window is proportional to the operand rounding error.  fp16 keeps an 11-bit
// significand (round-to-nearest 2^-11, better than the 2^-10 truncation the tensor core applies to f32 under kind::tf32),
// bf16 an 8-bit one (2^-8: an 8x wider window, i.e. 8x the exact re-checks) at the same tensor rate.  What fp16 lacks is
// range, so every operand is multiplied by a power of two first (exact in f32, undone exactly in the epilogue's FFMA
// constant): the scale puts the largest magnitude of a sample of the table at 2^12, which leaves a factor 16 of headroom
// above and 26 binades below.  A row with a scaled value outside +-65504 is simply marked non-finite here and takes the
// exact path (as rows with inf/NaN always did).  Values below the fp16 normal range are flushed to zero ON PURPOSE so that
// no assumption about subnormal handling in the tensor core is needed; the residual norm below is measured after the flush.
//
// Per row the shadow pass leaves (|x|^2, x.mu) in f32 -- what row_stats_kernel computed per sweep before, now once per
// table -- and folds the row's relative residual |x - half(x)| / |x| into one table-wide maximum kappa.  Error model used
// by the consumers (pqv_tc.cuh), with a >= |x|, B^ = half(B'), bn = |B^|, rn = |B' - B^|:
//     x.B' - mma(x^, B^)  =  r_x.B^  +  x.r_b  +  accumulation error
//     |.|                <=  kappa a bn  +  a rn  +  eps_acc (1 + kappa) a bn
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqv_kernels.cuh"

namespace pqv {
namespace half16 {

struct Globals {
    uint32_t kappa_bits;      // max over finite rows of |x - x^| / |x| (f32 bits, >= 0; atomicMax)
    uint32_t x2max_bits;      // max over finite rows of |x|^2
    uint32_t nonfinite_rows;  // rows with inf / NaN / fp16 overflow (stats.x = +inf for them)
    float scale;              // power of two the rows were multiplied by before rounding to fp16
};

// max |v| over count floats as f32 bits (non-negative floats order as unsigned; NaN ranks above inf): atomicMax into *out
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ v, u64 count, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(v[i]) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
// the power of two that maps absmax to [2^11, 2^12]; 1 for an all-zero or non-finite sample
__device__ __forceinline__ float scale_for_absmax(uint32_t absmax_bits) {
    const float a = __uint_as_float(absmax_bits);
    if (!(a > 0.f) || !(a < 3.0e38f)) return 1.f;
    int e;
    frexpf(a, &e);                       // a = f 2^e, f in [0.5, 1)
    const int sh = max(-60, min(60, 12 - e));  // two such scales multiply (rows x table) and their product must stay a normal f32
    return ldexpf(1.f, sh);
}
__global__ void scale_from_absmax_kernel(const uint32_t *absmax_bits, float *scale_out) {  // may alias: converted in place
    if (threadIdx.x == 0 && blockIdx.x == 0) *scale_out = scale_for_absmax(*absmax_bits);
}

constexpr float HALF_MIN_NORMAL = 6.103515625e-05f;  // 2^-14
constexpr float HALF_MAX = 65504.f;

// round-to-nearest-even to fp16, magnitudes below the normal range flushed to +-0, overflow -> +-inf
__device__ __forceinline__ __half to_half_flushed(float v) {
    const __half h = __float2half_rn(fabsf(v) < HALF_MIN_NORMAL ? 0.f : v);
    return h;
}

// mu[col] = mean of column col over the first n_sample rows (any vector is valid for the consumers; a data mean keeps
// |c - mu| small for centroids that are means of the data).  One thread per column: coalesced across the warp.
// Two deterministic steps (a single thread per column walking 65 536 rows took 4.4 ms per shadow, i.e. per index build):
// blockIdx.y = slice of the sample rows -> part[slice][col], then the slices are added in slice order.
constexpr uint32_t MEAN_SLICES = 64;
__global__ void __launch_bounds__(128) column_mean_partial_kernel(const float *__restrict__ rows, u64 n_sample, uint32_t dim,
                                                                  float *__restrict__ part) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    const u64 per = (n_sample + MEAN_SLICES - 1) / MEAN_SLICES;
    const u64 b = (u64)blockIdx.y * per, e = b + per < n_sample ? b + per : n_sample;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    u64 r = b;
    for (; r + 3 < e; r += 4) {
        s0 += rows[(r + 0) * dim + col];
        s1 += rows[(r + 1) * dim + col];
        s2 += rows[(r + 2) * dim + col];
        s3 += rows[(r + 3) * dim + col];
    }
    for (; r < e; ++r) s0 += rows[r * dim + col];
    part[(size_t)blockIdx.y * dim + col] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(128) column_mean_kernel(const float *__restrict__ part, u64 n_sample, uint32_t dim,
                                                          float *__restrict__ mu) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    float s = 0.f;
    for (uint32_t i = 0; i < MEAN_SLICES; ++i) s += part[(size_t)i * dim + col];
    const float m = s / (float)(n_sample ? n_sample : 1);
    mu[col] = (fabsf(m) < 1e30f) ? m : 0.f;  // non-finite data: fall back to the origin (still a valid choice)
}

// rows [first, first + n) of the table: out16 = half(rows), stats = (|x|^2, x.mu), kappa / x2max / nonfinite folded into g.
// One warp per row (dim % 8 == 0): a lane converts 8 consecutive columns per step -- two 128-bit loads, one 128-bit store.
__global__ void __launch_bounds__(256) shadow_rows_kernel(const float *__restrict__ rows, u64 first, u64 n, uint32_t dim,
                                                          const float *__restrict__ mu, __half *__restrict__ out16,
                                                          float2 *__restrict__ stats, Globals *__restrict__ g) {
    const float scale = g->scale, inv_scale = 1.f / scale;  // powers of two: both products below are exact
    const uint32_t lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const uint32_t n8 = dim >> 3;
    float kap = 0.f, x2m = 0.f;
    uint32_t bad_rows = 0;
    for (u64 i = warp; i < n; i += nwarps) {
        const u64 r = first + i;
        const float *p = rows + r * dim;
        __half *o = out16 + r * dim;
        float s = 0.f, t = 0.f, e = 0.f;
        bool bad = false;
        for (uint32_t c = lane; c < n8; c += 32) {
            const float4 a = ld_stream_v4(p + 8 * c), b = ld_stream_v4(p + 8 * c + 4);
            const float4 ma = __ldg(reinterpret_cast<const float4 *>(mu + 8 * c)), mb = __ldg(reinterpret_cast<const float4 *>(mu + 8 * c + 4));
            const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            const float m[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
            __half h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                h[j] = to_half_flushed(x[j] * scale);
                const float back = __half2float(h[j]) * inv_scale;
                const float d = x[j] - back;  // exact in f32 when |x| is finite and in range (few-bit difference)
                s = __fmaf_rn(x[j], x[j], s);
                t = __fmaf_rn(x[j], m[j], t);
                e = __fmaf_rn(d, d, e);
                bad |= !(fabsf(__half2float(h[j])) <= HALF_MAX);  // inf after overflow, NaN input
            }
            uint4 pack;
            pack.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
            pack.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
            pack.z = (uint32_t)__half_as_ushort(h[4]) | ((uint32_t)__half_as_ushort(h[5]) << 16);
            pack.w = (uint32_t)__half_as_ushort(h[6]) | ((uint32_t)__half_as_ushort(h[7]) << 16);
            *reinterpret_cast<uint4 *>(o + 8 * c) = pack;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            t += __shfl_xor_sync(0xffffffffu, t, off);
            e += __shfl_xor_sync(0xffffffffu, e, off);
        }
        bad = __any_sync(0xffffffffu, bad) || !(s < 1e30f);
        if (lane == 0) {
            stats[r] = make_float2(bad ? __int_as_float(0x7f800000) : s, bad ? 0.f : t);
            if (bad) {
                ++bad_rows;
            } else {
                x2m = fmaxf(x2m, s);
                // relative residual, rounded up: f32 sums of dim non-negative terms carry <= (dim + 32) 2^-24 relative error each
                if (s > 0.f) kap = fmaxf(kap, sqrtf(e / s));
            }
        }
    }
    if (lane == 0) {
        if (kap > 0.f) atomicMax(&g->kappa_bits, __float_as_uint(kap));
        if (x2m > 0.f) atomicMax(&g->x2max_bits, __float_as_uint(x2m));
        if (bad_rows) atomicAdd(&g->nonfinite_rows, bad_rows);
    }
}

}  // namespace half16
}  // namespace pqv
