// pqv_kmeanspp.cuh -- the k-means++ pick on the device (src/ivf/index.rs:354-390).
//
// Per new centroid the reference (a) min-updates the distance of every init row against the last centroid while summing
// the updated values in one serial f32 chain PER WORKER CHUNK, adds the chunk sums in chunk order (total), (b) draws
// threshold = rand * total, (c) walks the array with ONE serial f32 chain and takes the first slot whose running sum
// reaches the threshold (none: the centroid stays zero), or a uniform random slot when total == 0.  (a)'s sweep is
// l2_dist_kernel; this kernel is the rest: one CTA, the distance array in shared memory, the random stream in device memory
// -- so the 1023 picks of a 1024-cluster build need no host round trip at all.
//
// The chunk sums are short independent chains (one thread each).  The walk (c) is up to 50 000 DEPENDENT adds; it runs
// exactly, but in parallel, block by block (tests/test_kmeanspp_pick_model.py is the CPU model of this procedure):
// inside the binade of the running sum x = m u (u = ulp(x), 2^23 <= m < 2^24) an f32 add is the integer map
//     m -> m + a + tie ((m + a) & 1),   a = round-half-down(d / u), tie = the discarded bits are exactly one half
// which depends on m only through its parity: element functions are pairs (delta | m even, delta | m odd), they compose
// associatively, and a block of 4096 elements is a prefix scan.  The first element whose result leaves the binade
// (m >= 2^24) is re-done with a real f32 add, and the scan restarts behind it with the new ulp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqv {
namespace kpp {

constexpr uint32_t THREADS = 1024;
constexpr uint32_t SAT = 1u << 25, TOP = 1u << 24;
constexpr uint32_t PROLOGUE = 2048;      // plain adds before the first block: the first binades hold a handful of elements each, and
                                         // about this many dependent adds (with their stores) fit beside the worker-chunk chains (3125 elements each at 16 workers)
constexpr uint32_t SERIAL_BURST = 256;   // plain adds when the state is zero / tiny or a block made almost no progress
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t MAX_ROWS = 50000;     // init set of the reference (index.rs:332); 200 000 B of shared memory

struct Fn {
    uint32_t e, o;  // delta for an even / odd input, saturating at SAT
};
__device__ __forceinline__ Fn fn_identity() { return Fn{0u, 0u}; }
// g first, then f
__device__ __forceinline__ Fn fn_compose(const Fn g, const Fn f) {
    const uint32_t e = g.e + ((g.e & 1u) ? f.o : f.e);
    const uint32_t o = g.o + (((1u + g.o) & 1u) ? f.o : f.e);
    return Fn{min(e, SAT), min(o, SAT)};
}
__device__ __forceinline__ Fn fn_elem(const uint32_t a, const bool tie) {
    uint32_t e = a, o = a;
    if (tie) {
        e += e & 1u;
        o += (1u + o) & 1u;
    }
    return Fn{min(e, SAT), min(o, SAT)};
}
// element bits -> (a, tie) in units of 2^(E - 150), E = biased exponent of the running sum
__device__ __forceinline__ void classify(const uint32_t db, const int E, uint32_t &a, bool &tie) {
    a = 0u;
    tie = false;
    if ((db & 0x7FFFFFFFu) == 0u) return;
    int Ed = (int)(db >> 23);
    uint32_t Md = db & 0x7FFFFFu;
    if (Ed == 0) Ed = 1;
    else Md |= 0x800000u;
    const int sh = E - Ed;
    if (sh <= 0) {
        a = SAT;  // at least 2^23 units: leaves the binade for sure
    } else if (sh < 25) {
        const uint32_t rem = Md & ((1u << sh) - 1u), half = 1u << (sh - 1);
        a = (Md >> sh) + (rem > half ? 1u : 0u);
        tie = rem == half;
    }
}

// plain serial adds over smd[begin, end) from x: eight at a time with the loads issued first and ONE threshold test per
// batch (the running sum never decreases: terms >= 0), so that the dependent chain is the adds alone.  Returns the first
// index whose running sum reaches thr (NONE: not reached), *x = the running sum behind the last element consumed.
__device__ __forceinline__ uint32_t serial_walk(const float *smd, uint32_t begin, const uint32_t end, float *x, const float thr,
                                                const bool ordered) {
    float y = *x;
    uint32_t i = begin;
    if (ordered) {
        for (; i + 8 <= end; i += 8) {
            float v[8], p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = smd[i + j];
            p[0] = __fadd_rn(y, v[0]);
#pragma unroll
            for (int j = 1; j < 8; ++j) p[j] = __fadd_rn(p[j - 1], v[j]);
            if (p[7] >= thr) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (p[j] >= thr) {
                        *x = p[j];
                        return i + (uint32_t)j;
                    }
            }
            y = p[7];
        }
    }
    for (; i < end; ++i) {
        y = __fadd_rn(y, smd[i]);
        if (y >= thr) {
            *x = y;
            return i;
        }
    }
    *x = y;
    return NONE;
}

struct SplitMix64Dev {
    unsigned long long s;
    __device__ unsigned long long next() {
        unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    __device__ unsigned long long below(unsigned long long n) {
        const unsigned long long lim = ~0ull - (~0ull % n);
        unsigned long long v;
        do v = next();
        while (v >= lim);
        return v % n;
    }
    __device__ float unit_f32() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
};

// sixteen chain adds, in element order
// PQV_KPP_ACC(s, v) = s + v, rounded once.  Written as fma(v, one, s) with `one` = 1.0f arriving as a kernel argument (so
// that it stays an FFMA): the product v * 1 is exact, hence the result is the f32 sum bit for bit, and a chain of dependent
// FFMAs advances every 4 cycles where the measured chain of FADDs took 5.8.
#define PQV_KPP_ACC(S, V) __fmaf_rn((V), one, (S))
#define PQV_KPP_ADD16(S, A0, A1, A2, A3)                                                                           \
    do {                                                                                                           \
        S = PQV_KPP_ACC(S, A0.x); S = PQV_KPP_ACC(S, A0.y); S = PQV_KPP_ACC(S, A0.z); S = PQV_KPP_ACC(S, A0.w);    \
        S = PQV_KPP_ACC(S, A1.x); S = PQV_KPP_ACC(S, A1.y); S = PQV_KPP_ACC(S, A1.z); S = PQV_KPP_ACC(S, A1.w);    \
        S = PQV_KPP_ACC(S, A2.x); S = PQV_KPP_ACC(S, A2.y); S = PQV_KPP_ACC(S, A2.z); S = PQV_KPP_ACC(S, A2.w);    \
        S = PQV_KPP_ACC(S, A3.x); S = PQV_KPP_ACC(S, A3.y); S = PQV_KPP_ACC(S, A3.z); S = PQV_KPP_ACC(S, A3.w);    \
    } while (0)

// sixteen running sums in element order, left in place of the elements
#define PQV_KPP_SCAN4(S, A)          \
    do {                             \
        A.x = S = PQV_KPP_ACC(S, A.x); \
        A.y = S = PQV_KPP_ACC(S, A.y); \
        A.z = S = PQV_KPP_ACC(S, A.z); \
        A.w = S = PQV_KPP_ACC(S, A.w); \
    } while (0)
#define PQV_KPP_SCAN16(S, A0, A1, A2, A3) \
    do {                                  \
        PQV_KPP_SCAN4(S, A0);             \
        PQV_KPP_SCAN4(S, A1);             \
        PQV_KPP_SCAN4(S, A2);             \
        PQV_KPP_SCAN4(S, A3);             \
    } while (0)

// s + smd[i] + smd[i + 1] + ... + smd[e - 1], one rounding per add, in this order (a worker chunk's chain, index.rs:358-368).
// The chain's adds are all that may sit on the critical path: plain adds up to a 16-byte boundary, then 32 elements at a time
// from two register sets that are refilled (LDS.128) behind the adds that consumed them.
__device__ __forceinline__ float chain_sum(float s, const float *smd, uint32_t i, const uint32_t e, const float one) {
    for (; i < e && (i & 3u); ++i) s = __fadd_rn(s, smd[i]);
    if (i + 32 <= e) {
        const float4 *p = reinterpret_cast<const float4 *>(smd + i);
        float4 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
        float4 b0 = p[4], b1 = p[5], b2 = p[6], b3 = p[7];
        for (i += 32, p += 8; i + 32 <= e; i += 32, p += 8) {
            PQV_KPP_ADD16(s, a0, a1, a2, a3);
            a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
            PQV_KPP_ADD16(s, b0, b1, b2, b3);
            b0 = p[4], b1 = p[5], b2 = p[6], b3 = p[7];
        }
        PQV_KPP_ADD16(s, a0, a1, a2, a3);
        PQV_KPP_ADD16(s, b0, b1, b2, b3);
    }
    for (; i < e; ++i) s = __fadd_rn(s, smd[i]);
    return s;
}

// one CTA of THREADS threads; dynamic shared memory: n floats.  EPT = elements per thread of one walk block: a block round
// costs a fixed ~3 000 cycles of barriers and shuffles, so wider blocks mean fewer rounds (the rounds a binade crossing
// forces stay: ~log2(n / PROLOGUE) of them)
template <uint32_t EPT>
__global__ void __launch_bounds__(THREADS) kmeanspp_pick_kernel(const float *__restrict__ md, const uint32_t n, const uint32_t chunk,
                                                                 const uint32_t n_chunks, unsigned long long *__restrict__ rng_state,
                                                                 const uint32_t *__restrict__ init_idx,
                                                                 const float *__restrict__ sample, const uint32_t dim,
                                                                 float *__restrict__ centroid_out, uint32_t *__restrict__ picked_out,
                                                                 const float one_arg) {
    constexpr uint32_t BLOCK = THREADS * EPT;
    static_assert(EPT % 4u == 0u, "a thread's elements are loaded as float4");
    const uint32_t n_pad = (n + 3u) & ~3u;  // the dynamic shared array is allocated up to the next multiple of four
    extern __shared__ __align__(16) float smd[];
    __shared__ float s_local[THREADS];
    __shared__ __align__(16) float s_pre[PROLOGUE];  // running sums of the first PROLOGUE elements (computed beside the chunk sums)
    __shared__ uint32_t s_wt_e[32], s_wt_o[32];
    __shared__ uint32_t s_base, s_pick, s_mode, s_weird;
    __shared__ float s_x, s_thr;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const long long t_start = picked_out ? clock64() : 0;  // picked_out != null (PQV_TRACE): [0] = pick, [1..5], [7], [8] = cycles at the end of a phase, [6] = walk rounds
    __shared__ float s_one;
    if (tid == 0) {
        s_weird = 0u;
        s_one = one_arg;
    }
    __syncthreads();
    // 1.0f for the chains' FFMAs, read back from shared memory so that it sits in a vector register of each thread (as a
    // uniform-register operand the dependent FFMAs measured no faster than FADDs)
    const float one = *reinterpret_cast<volatile float *>(&s_one);
    bool weird = false;
    const uint32_t n4 = ((reinterpret_cast<uintptr_t>(md) & 15u) == 0u) ? (n >> 2) : 0u;
    for (uint32_t i0 = 0; i0 < n4; i0 += 4u * THREADS) {  // four loads of a thread in flight (one SM pulls the whole array)
        float4 v[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t i = i0 + j * THREADS + tid;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n4) v[j] = __ldcg(reinterpret_cast<const float4 *>(md) + i);
        }
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t i = i0 + j * THREADS + tid;
            if (i < n4) reinterpret_cast<float4 *>(smd)[i] = v[j];
            // NaN, inf, negative: the integer model does not apply -> plain chains
            weird |= !(v[j].x >= 0.f && v[j].x < 3.0e38f) || !(v[j].y >= 0.f && v[j].y < 3.0e38f) ||
                     !(v[j].z >= 0.f && v[j].z < 3.0e38f) || !(v[j].w >= 0.f && v[j].w < 3.0e38f);
        }
    }
    for (uint32_t i = (n4 << 2) + tid; i < n; i += THREADS) {
        const float v = md[i];
        smd[i] = v;
        weird |= !(v >= 0.f && v < 3.0e38f);
    }
    if (weird) s_weird = 1u;
    __syncthreads();
    if (tid == 0 && picked_out) picked_out[1] = (uint32_t)(clock64() - t_start);
    // ---- (a) chunk sums (index.rs:356-370): one serial chain per worker chunk, then the chunk sums in chunk order
    float total = 0.f;
    const uint32_t pro_n = min(PROLOGUE, n);
    const bool pro_side = n_chunks <= THREADS - 32u;
    // meanwhile the walk's first PROLOGUE running sums, which need no threshold yet: a lane of the LAST warp takes them when
    // the chunk chains leave it free (the usual 16 workers sit in warp 0), so they cost nothing on the critical path
    if (pro_side && tid == THREADS - 1u && !s_weird) {
        float y = 0.f;
        uint32_t i = 0;
        if (pro_n >= 32u) {  // two register sets of 16: one is being summed and stored while the other one's loads are in flight
            const float4 *p = reinterpret_cast<const float4 *>(smd);
            float4 *q = reinterpret_cast<float4 *>(s_pre);
            float4 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
            float4 b0 = p[4], b1 = p[5], b2 = p[6], b3 = p[7];
            for (i = 32, p += 8; i + 32 <= pro_n; i += 32, p += 8, q += 8) {
                PQV_KPP_SCAN16(y, a0, a1, a2, a3);
                q[0] = a0, q[1] = a1, q[2] = a2, q[3] = a3;
                a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
                PQV_KPP_SCAN16(y, b0, b1, b2, b3);
                q[4] = b0, q[5] = b1, q[6] = b2, q[7] = b3;
                b0 = p[4], b1 = p[5], b2 = p[6], b3 = p[7];
            }
            PQV_KPP_SCAN16(y, a0, a1, a2, a3);
            q[0] = a0, q[1] = a1, q[2] = a2, q[3] = a3;
            PQV_KPP_SCAN16(y, b0, b1, b2, b3);
            q[4] = b0, q[5] = b1, q[6] = b2, q[7] = b3;
        }
        for (; i < pro_n; ++i) {
            y = __fadd_rn(y, smd[i]);
            s_pre[i] = y;
        }
        if (picked_out) picked_out[8] = (uint32_t)(clock64() - t_start);  // end of the prologue's running sums
    }
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += THREADS) {
        const uint32_t c = c0 + tid;
        if (c < n_chunks) {
            const uint32_t b = c * chunk, e = min(n, b + chunk);
            s_local[tid] = chain_sum(0.f, smd, b, e, one);
            if (picked_out && c == 0u) picked_out[7] = (uint32_t)(clock64() - t_start);  // end of the first chunk's chain
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t m = min(THREADS, n_chunks - c0);
            for (uint32_t i = 0; i < m; ++i) total = __fadd_rn(total, s_local[i]);
        }
        __syncthreads();
    }
    // ---- (b) the draw
    if (tid == 0) {
        if (picked_out) picked_out[2] = (uint32_t)(clock64() - t_start);
        SplitMix64Dev rng{*rng_state};
        s_pick = NONE;
        s_mode = 0u;
        if (total > 0.f) {  // index.rs:372-383
            s_thr = __fmul_rn(rng.unit_f32(), total);
        } else {            // index.rs:384-389
            s_pick = (uint32_t)rng.below(n);
            s_mode = 2u;    // done
        }
        *rng_state = rng.s;
        s_x = 0.f;
        s_base = 0u;
    }
    __syncthreads();
    // ---- (c) the walk
    const bool do_walk = s_mode != 2u;
    __syncthreads();  // every thread has read the mode before the prologue test below may set it (racecheck: read / write race,
                      // and a thread that saw the new value would have skipped the barriers inside the block)
    if (do_walk) {
        const float thr = s_thr;
        if (pro_side && !s_weird) {
            // the prologue's running sums are there: the first one that reaches the threshold, if any (they never decrease)
            for (uint32_t i = tid; i < pro_n; i += THREADS)
                if (s_pre[i] >= thr && (i == 0u || !(s_pre[i - 1u] >= thr))) {
                    s_pick = i;
                    s_mode = 2u;
                }
            if (tid == 0) {
                s_x = s_pre[pro_n - 1u];
                s_base = pro_n;
                if (picked_out) {
                    picked_out[3] = (uint32_t)(clock64() - t_start);
                    picked_out[6] = 0u;
                }
            }
        } else if (tid == 0) {  // plain prologue (or the whole walk for arrays the integer model does not cover)
            const uint32_t cnt = s_weird ? n : pro_n;
            float x = 0.f;
            const uint32_t hit = serial_walk(smd, 0u, cnt, &x, thr, !s_weird);
            if (hit != NONE) {
                s_pick = hit;
                s_mode = 2u;
            }
            s_x = x;
            s_base = cnt;
            if (picked_out) {
                picked_out[3] = (uint32_t)(clock64() - t_start);
                picked_out[6] = 0u;
            }
        }
        __syncthreads();
        while (s_mode != 2u && s_base < n) {
            if (tid == 0 && picked_out) picked_out[6] += 1u;
            const uint32_t base = s_base;
            const float x = s_x;
            const uint32_t xb = __float_as_uint(x);
            const int E = (int)(xb >> 23);
            if (E < 25 || E >= 254) {  // zero / tiny / huge running sum: plain adds for a while
                __syncthreads();
                if (tid == 0) {
                    float y = x;
                    const uint32_t end = min(n, base + SERIAL_BURST);
                    const uint32_t hit = serial_walk(smd, base, end, &y, thr, true);
                    if (hit != NONE) {
                        s_pick = hit;
                        s_mode = 2u;
                    }
                    s_x = y;
                    s_base = end;
                }
                __syncthreads();
                continue;
            }
            const uint32_t m0 = (xb & 0x7FFFFFu) | 0x800000u;
            const float u = __uint_as_float((uint32_t)(E - 23) << 23);            // ulp of the running sum: 2^(E - 150)
            const float half_u = __uint_as_float((uint32_t)(E - 24) << 23);
            const float inv_u = __uint_as_float((uint32_t)(277 - E) << 23);
            const float big = __uint_as_float(((uint32_t)E << 23) | 0x400000u);   // 1.5 x 2^(E - 127): same ulp, even mantissa
            const uint32_t big_bits = ((uint32_t)E << 23) | 0x400000u;
            const float dmax = __uint_as_float((uint32_t)(E - 1) << 23);          // big + d stays inside big's binade below this
            // the threshold in units of u: m u >= thr  <=>  m >= ceil(thr / u) (the scaling is exact; a quotient that
            // underflows compares true for every m >= 2^23 like thr itself, one that overflows or is NaN never does)
            const float tq = __fmul_rn(thr, inv_u);
            const uint32_t thr_m = (tq < 33554432.f) ? (uint32_t)ceilf(tq) : SAT;
            const uint32_t lim = min(TOP, thr_m);  // the first m >= lim is the block's event: pick (m < TOP) or binade exit
            // element -> (a, tie): t = big + d is d rounded to a multiple of u (ties to even) on top of big, so a is the
            // DIFFERENCE OF THE BIT PATTERNS of t and big; the residual d - (t - big) is exact, and |residual| == u / 2 marks
            // the tie; a = the round-DOWN choice of a tie, so that the parity rule of the running sum (not of big) decides it
            uint32_t a[EPT];
            bool tie[EPT], any_tie = false;
            uint32_t sum_a = 0u;
            // the block starts at the 16-byte boundary at or below base (elements before base count as zeros), so that a
            // thread's EPT consecutive elements come in as LDS.128 -- with EPT = 12 the eight lanes of a quarter warp are
            // 48 bytes apart and hit eight different 16-byte bank groups (scalar loads at stride EPT were 4- to 16-way conflicts)
            const uint32_t base_al = base & ~3u;
            const uint32_t j0 = base_al + tid * EPT;
#pragma unroll
            for (uint32_t e4 = 0; e4 < EPT; e4 += 4) {
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + e4 < n_pad) v4 = *reinterpret_cast<const float4 *>(smd + j0 + e4);
                const float dv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q) {
                    const uint32_t e = e4 + q, j = j0 + e;
                    const float d = (j >= base && j < n) ? dv[q] : 0.f;
                    const float t = __fadd_rn(big, d);
                    const float res = __fadd_rn(d, -__fadd_rn(t, -big));
                    const bool big_d = d >= dmax;  // at least a quarter of the binade: leaves it for sure (and would overflow big's)
                    tie[e] = !big_d && fabsf(res) == half_u;
                    const uint32_t ai = __float_as_uint(t) - big_bits - ((res == -half_u) ? 1u : 0u);
                    a[e] = big_d ? SAT : ai;
                    any_tie |= tie[e];
                    sum_a = min(sum_a + a[e], SAT);
                }
            }
            // inclusive scan inside the warp: plain saturating sums unless some lane of the warp holds a tie
            Fn inc;
            if (!__any_sync(0xffffffffu, any_tie)) {
                uint32_t v = sum_a;
#pragma unroll
                for (uint32_t o = 1; o < 32; o <<= 1) {
                    const uint32_t g = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v = min(v + g, SAT);
                }
                inc = Fn{v, v};
            } else {
                Fn mine = fn_identity();
#pragma unroll
                for (uint32_t e = 0; e < EPT; ++e) mine = fn_compose(mine, fn_elem(a[e], tie[e]));
                inc = mine;
#pragma unroll
                for (uint32_t o = 1; o < 32; o <<= 1) {
                    Fn g;
                    g.e = __shfl_up_sync(0xffffffffu, inc.e, o);
                    g.o = __shfl_up_sync(0xffffffffu, inc.o, o);
                    if (lane >= o) inc = fn_compose(g, inc);
                }
            }
            if (lane == 31) {
                s_wt_e[warp] = inc.e;
                s_wt_o[warp] = inc.o;
            }
            __syncthreads();
            if (warp == 0) {
                Fn w{s_wt_e[lane], s_wt_o[lane]};
#pragma unroll
                for (uint32_t o = 1; o < 32; o <<= 1) {
                    Fn g;
                    g.e = __shfl_up_sync(0xffffffffu, w.e, o);
                    g.o = __shfl_up_sync(0xffffffffu, w.o, o);
                    if (lane >= o) w = fn_compose(g, w);
                }
                s_wt_e[lane] = w.e;  // inclusive over warps
                s_wt_o[lane] = w.o;
            }
            __syncthreads();
            // The running sum never decreases, so the state behind a thread's LAST element tells whether the block's first
            // event lies at or before it: exactly ONE thread starts below lim and does not end below it.  It alone walks its
            // elements and leaves the next round's state (no atomics); without an event the last thread does.  Everybody else
            // is done after two integer compares.
            const uint32_t m_warp = warp > 0 ? m0 + ((m0 & 1u) ? s_wt_o[warp - 1] : s_wt_e[warp - 1]) : m0;  // before the warp
            const Fn incl = warp > 0 ? fn_compose(Fn{s_wt_e[warp - 1], s_wt_o[warp - 1]}, inc) : inc;
            const uint32_t m_end = m0 + ((m0 & 1u) ? incl.o : incl.e);
            uint32_t m = __shfl_up_sync(0xffffffffu, m_end, 1);
            if (lane == 0) m = m_warp;
            if (m < lim && m_end >= lim) {
                uint32_t ev_e = 0u, mn = 0u;
                bool found = false;
#pragma unroll
                for (uint32_t e = 0; e < EPT; ++e) {
                    if (!found) {
                        mn = m + a[e] + (tie[e] ? ((m + a[e]) & 1u) : 0u);
                        if (mn >= lim) {
                            found = true;
                            ev_e = e;
                        } else {
                            m = mn;
                        }
                    }
                }
                // m = the exact state before element ev_e, mn behind it: it reaches the threshold inside the binade or leaves it
                const uint32_t i_ev = j0 + ev_e;
                if (mn < TOP) {
                    s_pick = i_ev;
                    s_mode = 2u;
                } else {  // the add that leaves the binade: a real f32 add from the exact state before it
                    float y = __fadd_rn(__fmul_rn(__uint2float_rn(m), u), smd[i_ev]);
                    uint32_t i = i_ev;
                    if (y >= thr) {
                        s_pick = i;
                        s_mode = 2u;
                    } else {
                        ++i;
                        if (i_ev - base < 16u) {  // hardly any progress (alternating magnitudes): plain adds for a while
                            const uint32_t end = min(n, i + SERIAL_BURST);
                            const uint32_t hit = serial_walk(smd, i, end, &y, thr, true);
                            if (hit != NONE) {
                                s_pick = hit;
                                s_mode = 2u;
                            }
                            i = end;
                        }
                        s_x = y;
                        s_base = i;
                    }
                }
            } else if (tid == THREADS - 1u && m_end < lim) {  // no event in this block
                s_x = __fmul_rn(__uint2float_rn(m_end), u);
                s_base = base_al + BLOCK;
            }
            __syncthreads();
        }
    }
    // ---- the picked row becomes the centroid (none: it stays as the caller initialised it -- zeros, index.rs:330)
    const uint32_t pick = s_pick;
    if (tid == 0 && picked_out) {
        picked_out[0] = pick;
        picked_out[4] = (uint32_t)(clock64() - t_start);
    }
    if (pick != NONE) {
        const float *src = sample + (size_t)init_idx[pick] * dim;
        for (uint32_t j = tid; j < dim; j += THREADS) centroid_out[j] = src[j];
    }
    if (tid == 0 && picked_out) picked_out[5] = (uint32_t)(clock64() - t_start);
}

}  // namespace kpp
}  // namespace pqv
