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

constexpr uint32_t THREADS = 1024, EPT = 4, BLOCK = THREADS * EPT;
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

// one CTA of THREADS threads; dynamic shared memory: n floats
__global__ void __launch_bounds__(THREADS) kmeanspp_pick_kernel(const float *__restrict__ md, const uint32_t n, const uint32_t chunk,
                                                                 const uint32_t n_chunks, unsigned long long *__restrict__ rng_state,
                                                                 const uint32_t *__restrict__ init_idx,
                                                                 const float *__restrict__ sample, const uint32_t dim,
                                                                 float *__restrict__ centroid_out, uint32_t *__restrict__ picked_out) {
    extern __shared__ float smd[];
    __shared__ float s_local[THREADS];
    __shared__ float s_pre[PROLOGUE];  // running sums of the first PROLOGUE elements (computed beside the chunk sums)
    __shared__ uint32_t s_wt_e[32], s_wt_o[32];
    __shared__ unsigned long long s_event;  // (position << 1 | kind) << 32 | m before the element; ~0: no event in this block
    __shared__ uint32_t s_mend, s_base, s_pick, s_mode, s_weird;
    __shared__ float s_x, s_thr;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const long long t_start = picked_out ? clock64() : 0;  // picked_out != null (PQV_TRACE): [0] = pick, [1..5] = cycles per phase
    if (tid == 0) s_weird = 0u;
    __syncthreads();
    bool weird = false;
    const uint32_t n4 = ((reinterpret_cast<uintptr_t>(md) & 15u) == 0u) ? (n >> 2) : 0u;
    for (uint32_t i = tid; i < n4; i += THREADS) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(md) + i);
        reinterpret_cast<float4 *>(smd)[i] = v;
        // NaN, inf, negative: the integer model does not apply -> plain chains
        weird |= !(v.x >= 0.f && v.x < 3.0e38f) || !(v.y >= 0.f && v.y < 3.0e38f) || !(v.z >= 0.f && v.z < 3.0e38f) || !(v.w >= 0.f && v.w < 3.0e38f);
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
    // meanwhile the walk's first PROLOGUE running sums, which need no threshold yet: a lane of the LAST warp takes them when
    // the chunk chains leave it free (the usual 16 workers sit in warp 0), so they cost nothing on the critical path
    const uint32_t pro_n = min(PROLOGUE, n);
    const bool pro_side = n_chunks <= THREADS - 32u;
    if (pro_side && tid == THREADS - 1u && !s_weird) {
        float y = 0.f;
        uint32_t i = 0;
        for (; i + 8 <= pro_n; i += 8) {  // loads first, then the dependent adds, then the stores
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = smd[i + j];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                y = __fadd_rn(y, v[j]);
                v[j] = y;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s_pre[i + j] = v[j];
        }
        for (; i < pro_n; ++i) {
            y = __fadd_rn(y, smd[i]);
            s_pre[i] = y;
        }
    }
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += THREADS) {
        const uint32_t c = c0 + tid;
        if (c < n_chunks) {
            const uint32_t b = c * chunk, e = min(n, b + chunk);
            float s = 0.f;
            uint32_t i = b;
            if (i + 16 <= e) {  // sixteen elements in registers ahead of the chain: the dependent adds are all that is left on it
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = smd[i + j];
                for (i += 16; i + 16 <= e; i += 16) {
                    float w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = smd[i + j];
#pragma unroll
                    for (int j = 0; j < 16; ++j) s = __fadd_rn(s, v[j]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = w[j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) s = __fadd_rn(s, v[j]);
            }
            for (; i < e; ++i) s = __fadd_rn(s, smd[i]);
            s_local[tid] = s;
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
            const float dmax = __uint_as_float((uint32_t)(E - 1) << 23);          // big + d stays inside big's binade below this
            // element -> (a, tie) with three f32 adds: r = (big + d) - big is d rounded to a multiple of u (ties to even), the
            // residual d - r is exact, and |residual| == u / 2 marks the tie; a = the round-DOWN choice of a tie, so that the
            // parity rule of the running sum (not of big) decides it below
            uint32_t a[EPT];
            bool tie[EPT], any_tie = false;
            uint32_t sum_a = 0u;
#pragma unroll
            for (uint32_t e = 0; e < EPT; ++e) {
                const uint32_t j = base + tid * EPT + e;
                const float d = j < n ? smd[j] : 0.f;
                const float r = __fadd_rn(__fadd_rn(big, d), -big);
                const float res = __fadd_rn(d, -r);
                const bool big_d = d >= dmax;  // at least a quarter of the binade: leaves it for sure (and would overflow big's)
                tie[e] = !big_d && fabsf(res) == half_u;
                const float af = (res == -half_u) ? __fadd_rn(r, -u) : r;
                a[e] = big_d ? SAT : __float2uint_rn(__fmul_rn(af, inv_u));
                any_tie |= tie[e];
                sum_a = min(sum_a + a[e], SAT);
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
            if (tid == 0) s_event = ~0ull;
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
            Fn excl;  // everything before this thread's first element
            excl.e = __shfl_up_sync(0xffffffffu, inc.e, 1);
            excl.o = __shfl_up_sync(0xffffffffu, inc.o, 1);
            if (lane == 0) excl = fn_identity();
            if (warp > 0) excl = fn_compose(Fn{s_wt_e[warp - 1], s_wt_o[warp - 1]}, excl);
            uint32_t m = m0 + ((m0 & 1u) ? excl.o : excl.e);
            // The running sum never decreases, so exactly ONE thread sees the block's first event: the one that starts inside
            // the binade and below the threshold and does not end that way.  It alone writes s_event (no atomics).
            uint32_t my_event = NONE, my_before = 0u;
            if (m < TOP && !(__fmul_rn(__uint2float_rn(m), u) >= thr)) {
#pragma unroll
                for (uint32_t e = 0; e < EPT; ++e) {
                    if (my_event != NONE) break;
                    const uint32_t mn = m + a[e] + (tie[e] ? ((m + a[e]) & 1u) : 0u);
                    const uint32_t j = base + tid * EPT + e;
                    if (mn >= TOP) {
                        my_event = ((tid * EPT + e) << 1);
                        my_before = m;
                    } else if (j < n && __fmul_rn(__uint2float_rn(mn), u) >= thr) {
                        my_event = ((tid * EPT + e) << 1) | 1u;
                        my_before = m;
                    } else {
                        m = mn;
                    }
                }
            }
            if (my_event != NONE) s_event = ((unsigned long long)my_event << 32) | my_before;
            if (tid == THREADS - 1) s_mend = m;
            __syncthreads();
            if (tid == 0) {
                const unsigned long long ev64 = s_event;
                const uint32_t ev = (uint32_t)(ev64 >> 32), m_before = (uint32_t)ev64;
                if (ev64 == ~0ull) {
                    s_x = __fmul_rn(__uint2float_rn(s_mend), u);
                    s_base = base + BLOCK;
                } else {
                    const uint32_t pos = ev >> 1;
                    if (ev & 1u) {
                        s_pick = base + pos;
                        s_mode = 2u;
                    } else {  // the add that leaves the binade: a real f32 add from the exact state before it
                        float y = __fadd_rn(__fmul_rn(__uint2float_rn(m_before), u), smd[base + pos]);
                        uint32_t i = base + pos;
                        if (y >= thr) {
                            s_pick = i;
                            s_mode = 2u;
                        } else {
                            ++i;
                            if (pos < 16u) {  // hardly any progress (alternating magnitudes): plain adds for a while
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
                }
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
