// pqv_kernels.cuh -- sm_100a device code for the pq-vector squared-L2 / top-k / IVF-assign hot path.
//
// Design: every distance is produced with the reference's exact f32 operation
// order (no FMA: __fsub_rn/__fmul_rn/__fadd_rn), so results are bit-identical to the Rust loops:
//   PQV_SUM_UNROLL4  src/ivf/index.rs:461-480      sum += ((d0^2+d1^2)+d2^2)+d3^2, scalar tail
//   PQV_SUM_SEQ      src/df_vector/exec.rs:529-533  dist += diff*diff
// The per-row sum is a serial chain, so one warp works on a group of 32 rows: all lanes load each
// row with coalesced 128-bit loads and compute the (independent) chain terms in parallel, the
// terms are transposed through a padded shared-memory tile, and then lane r runs row r's serial
// chain in the reference order.  HBM is read exactly once; the chain costs one FADD per term.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqv {

typedef unsigned long long u64;
constexpr u64 KEY_MAX = 0xFFFFFFFFFFFFFFFFull;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_v4(const float *p) {
    // read-only, streaming: do not allocate in L1 (each byte is used once)
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// same load with an L2 eviction policy word (createpolicy): rows a kernel is going to read again in the next launch are
// kept (evict_last), the rest of an over-sized working set is let go first (evict_first)
__device__ __forceinline__ float4 ld_policy_v4(const float *p, const unsigned long long policy) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float2 ld_stream_v2(const float *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// one chain term of src/ivf/index.rs:467-471: ((d0*d0 + d1*d1) + d2*d2) + d3*d3, d = a - b
__device__ __forceinline__ float chunk4(const float4 a, const float4 b) {
    const float d0 = __fsub_rn(a.x, b.x), d1 = __fsub_rn(a.y, b.y);
    const float d2 = __fsub_rn(a.z, b.z), d3 = __fsub_rn(a.w, b.w);
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)),
                     __fmul_rn(d3, d3));
}
__device__ __forceinline__ float sq1(const float a, const float b) {
    const float d = __fsub_rn(a, b);
    return __fmul_rn(d, d);
}

// ------------------------------------------------------------------------------------------------
// group_distance: squared-L2 of the 32 rows of group g against the vector staged in s_vec.
// Returns in lane r the distance of row (g*32 + r) (rows past n are clamped to n-1; callers mask).
//   ORDER 0: unroll-4 order, a = s_vec (query), b = row     (search.rs:117  squared_l2_distance(query, vec))
//   ORDER 1: sequential,     diff = row - s_vec             (exec.rs:531    value - q)
//   VEC4   : dim % 4 == 0 and 16-byte aligned rows -> 128-bit loads; otherwise scalar loads.
//   GATHER : row index taken from row_ids[position].
// tile: per-warp shared memory, 32 rows x TSTRIDE floats (TSTRIDE = 36, or 68 for ORDER 1 + VEC4).
// ------------------------------------------------------------------------------------------------
// CBV = float4 loads per lane per row per column block (ORDER 0 + VEC4 only): a warp then reads CBV*512
// contiguous bytes of a row per block.
template <int ORDER, bool VEC4, int CBV = 1>
struct TileCfg {
    // ORDER 1 (sequential sum, every element its own chain term): 64-column blocks, 2 terms per lane, so that the tile is
    // 8.7 KB per warp and two CTAs fit an SM (with 128-column blocks = 16.9 KB per warp only one did: 3.4 TB/s)
    static constexpr int TERMS = (ORDER == 1 && VEC4) ? 2 : CBV;  // chain terms per lane per column block
    static constexpr int TSTRIDE = 32 * TERMS + 4;                // floats; /4 is odd -> LDS.128 conflict-free
    static constexpr int TILE_FLOATS = 32 * TSTRIDE;
};

// Tuned on B200 (profiles/r01_sweep_scan.jsonl): 8 rows x 2 float4 per lane in flight (1 KB contiguous per
// row per request pair) gives 7.2 TB/s; 8 x 1 gives 6.0-6.6, 16 x 1 7.1, 4 CTAs/SM x 4 x 1 7.1.
template <int ORDER, bool VEC4, bool GATHER = false>
struct ScanDefaults {
    static constexpr int CBV = (ORDER == 0 && VEC4) ? 2 : 1;
    static constexpr int RB = (GATHER && CBV == 2) ? 4 : 8;  // the gather variant spills at 8 x 2 under 128 regs
};

template <int ORDER, bool VEC4, bool GATHER, int RB = 8, int CBV = 1, bool HINT = false>
__device__ __forceinline__ float group_distance(const float *__restrict__ data,
                                                const uint32_t *__restrict__ row_ids, const u64 n,
                                                const uint32_t dim, const u64 g, const float *s_vec,
                                                float *tile, const uint32_t lane, const unsigned long long policy = 0ull) {
    constexpr int TSTRIDE = TileCfg<ORDER, VEC4, CBV>::TSTRIDE;
    static_assert(CBV == 1 || (ORDER == 0 && VEC4), "CBV > 1 only for the unroll-4 vector path");
    const u64 pos = g * 32 + lane;
    const u64 posc = pos < n ? pos : n - 1;
    const u64 my_row = GATHER ? (u64)row_ids[posc] : posc;
    // dense groups are contiguous: row r of the group starts at gbase + r*dim (clamped at the table end)
    const u64 g_first = g * 32;
    float sum = 0.0f;
    float *trow = tile + lane * TSTRIDE;

    if constexpr (VEC4 && ORDER == 1) {
        // sequential order: 64-column blocks, one 64-bit load per lane and row (a warp reads 256 contiguous bytes of a
        // row), 2 chain terms per lane; RB rows in flight
        constexpr uint32_t BLK = 64u;
        const uint32_t ncb = (dim + BLK - 1) / BLK;
        for (uint32_t cb = 0; cb < ncb; ++cb) {
            const uint32_t col0 = cb * BLK + (lane << 1);
            const bool inb = col0 < dim;  // dim % 4 == 0: both columns of the pair are in range together
            float2 q2 = make_float2(0.f, 0.f);
            if (inb) q2 = *reinterpret_cast<const float2 *>(s_vec + col0);
            // dense: rows of the group are dim floats apart from gp; the table's last group clamps to its last row
            // (32-bit offsets: 32 rows x dim <= 2^19 floats)
            const float *gp = data + g_first * dim + col0;
            const uint32_t last_rel = (uint32_t)((n - 1 - g_first) < 31 ? (n - 1 - g_first) : 31);
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 2 * RB) {
                float2 v2[2 * RB];
#pragma unroll
                for (int j = 0; j < 2 * RB; ++j) {
                    const float *rp;
                    if constexpr (GATHER) {
                        rp = data + (u64)__shfl_sync(0xffffffffu, (uint32_t)my_row, r0 + j) * dim + col0;
                    } else {
                        const uint32_t rel = (uint32_t)(r0 + j) < last_rel ? (uint32_t)(r0 + j) : last_rel;
                        rp = gp + rel * dim;
                    }
                    v2[j] = make_float2(0.f, 0.f);
                    if (inb) v2[j] = ld_stream_v2(rp);
                }
#pragma unroll
                for (int j = 0; j < 2 * RB; ++j) {
                    float2 p;
                    p.x = sq1(v2[j].x, q2.x);
                    p.y = sq1(v2[j].y, q2.y);
                    *reinterpret_cast<float2 *>(tile + (r0 + j) * TSTRIDE + (lane << 1)) = p;
                }
            }
            __syncwarp();
            const uint32_t rem = dim - cb * BLK;  // elements left in this block (multiple of 4)
            if (rem >= BLK) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 t = *reinterpret_cast<const float4 *>(trow + 4 * j);
                    sum = __fadd_rn(sum, t.x);
                    sum = __fadd_rn(sum, t.y);
                    sum = __fadd_rn(sum, t.z);
                    sum = __fadd_rn(sum, t.w);
                }
            } else {
                for (uint32_t j = 0; j < rem; ++j) sum = __fadd_rn(sum, trow[j]);
            }
            __syncwarp();
        }
    } else if constexpr (VEC4) {
        constexpr uint32_t BLK = 128u * CBV;  // columns per block
        const uint32_t ncb = (dim + BLK - 1) / BLK;
        for (uint32_t cb = 0; cb < ncb; ++cb) {
            const uint32_t col0 = cb * BLK + (lane << 2);
            float4 q4[CBV];
            bool inb[CBV];
#pragma unroll
            for (int v = 0; v < CBV; ++v) {
                inb[v] = col0 + 128u * v < dim;
                q4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb[v]) q4[v] = *reinterpret_cast<const float4 *>(s_vec + col0 + 128u * v);
            }
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += RB) {
                float4 v4[RB][CBV];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    u64 row;
                    if constexpr (GATHER) {
                        row = (u64)__shfl_sync(0xffffffffu, (uint32_t)my_row, r0 + j);
                    } else {
                        row = g_first + (r0 + j);
                        row = row < n ? row : n - 1;
                    }
                    const float *rp = data + row * dim + col0;
#pragma unroll
                    for (int v = 0; v < CBV; ++v) {
                        v4[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (inb[v]) {
                            if constexpr (HINT) v4[j][v] = ld_policy_v4(rp + 128 * v, policy);
                            else v4[j][v] = ld_stream_v4(rp + 128 * v);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    if constexpr (ORDER == 0) {
#pragma unroll
                        for (int v = 0; v < CBV; ++v) tile[(r0 + j) * TSTRIDE + 32 * v + lane] = chunk4(q4[v], v4[j][v]);
                    } else {
                        float4 p;
                        p.x = sq1(v4[j][0].x, q4[0].x);
                        p.y = sq1(v4[j][0].y, q4[0].y);
                        p.z = sq1(v4[j][0].z, q4[0].z);
                        p.w = sq1(v4[j][0].w, q4[0].w);
                        *reinterpret_cast<float4 *>(tile + (r0 + j) * TSTRIDE + (lane << 2)) = p;
                    }
                }
            }
            __syncwarp();
            // serial chain, lane = row
            const uint32_t rem = dim - cb * BLK;  // elements left in this block (multiple of 4)
            if constexpr (ORDER == 0) {
                if (rem >= BLK) {
#pragma unroll
                    for (int j = 0; j < 8 * CBV; ++j) {
                        const float4 t = *reinterpret_cast<const float4 *>(trow + 4 * j);
                        sum = __fadd_rn(sum, t.x);
                        sum = __fadd_rn(sum, t.y);
                        sum = __fadd_rn(sum, t.z);
                        sum = __fadd_rn(sum, t.w);
                    }
                } else {
                    const uint32_t nt = rem >> 2;
                    for (uint32_t j = 0; j < nt; ++j) sum = __fadd_rn(sum, trow[j]);
                }
            } else {
                if (rem >= 128u) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float4 t = *reinterpret_cast<const float4 *>(trow + 4 * j);
                        sum = __fadd_rn(sum, t.x);
                        sum = __fadd_rn(sum, t.y);
                        sum = __fadd_rn(sum, t.z);
                        sum = __fadd_rn(sum, t.w);
                    }
                } else {
                    for (uint32_t j = 0; j < rem; ++j) sum = __fadd_rn(sum, trow[j]);
                }
            }
            __syncwarp();
        }
    } else {
        // scalar-load path (dim % 4 != 0 or unaligned base).  ORDER 0: lane handles chunk j of 4
        // consecutive elements; ORDER 1: lane handles one element.
        constexpr uint32_t EPL = (ORDER == 0) ? 4u : 1u;       // elements per lane per block
        const uint32_t n_terms = (ORDER == 0) ? (dim >> 2) : dim;
        const uint32_t nblk = (n_terms + 31u) >> 5;
        for (uint32_t cb = 0; cb < nblk; ++cb) {
            const uint32_t term = (cb << 5) + lane;
            const bool inb = term < n_terms;
            const uint32_t col = term * EPL;
            float q[EPL];
#pragma unroll
            for (uint32_t e = 0; e < EPL; ++e) q[e] = inb ? s_vec[col + e] : 0.f;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                u64 row;
                if constexpr (GATHER) {
                    row = (u64)__shfl_sync(0xffffffffu, (uint32_t)my_row, r);
                } else {
                    row = g_first + r;
                    row = row < n ? row : n - 1;
                }
                float t = 0.f;
                if (inb) {
                    const float *rp = data + row * dim + col;
                    if constexpr (ORDER == 0) {
                        const float4 v = make_float4(ld_stream_f32(rp), ld_stream_f32(rp + 1),
                                                     ld_stream_f32(rp + 2), ld_stream_f32(rp + 3));
                        t = chunk4(make_float4(q[0], q[1], q[2], q[3]), v);
                    } else {
                        t = sq1(ld_stream_f32(rp), q[0]);
                    }
                }
                tile[r * TSTRIDE + lane] = t;
            }
            __syncwarp();
            const uint32_t left = n_terms - (cb << 5);
            const uint32_t nt = left < 32u ? left : 32u;
            for (uint32_t j = 0; j < nt; ++j) sum = __fadd_rn(sum, trow[j]);
            __syncwarp();
        }
        if constexpr (ORDER == 0) {
            // scalar tail of src/ivf/index.rs:474-478: each leftover element is its own chain term
            const float *rp = data + my_row * dim;
            for (uint32_t i = (dim >> 2) << 2; i < dim; ++i) sum = __fadd_rn(sum, sq1(s_vec[i], rp[i]));
        }
    }
    return sum;
}

// ------------------------------------------------------------------------------------------------
// l2_scan_topk_kernel: distance scan + per-CTA exact top-k + heap-entrant candidate emission.
//
// CTA b owns the contiguous groups [NG*b/G, NG*(b+1)/G) and walks them in position order, WARPS
// groups per lock-step iteration.  A row passes the filter iff bits(d) < tau, where tau is the
// k-th smallest distance among the rows this CTA finished in EARLIER iterations (0xFFFFFFFF while
// it holds fewer than k), optionally capped by cap_bits (a bound carried in from rows that precede
// the whole launch).  Because tau only ever summarises rows at smaller positions, it is an upper
// bound of the reference heap's root at that position, so the rows that pass are a superset of the
// rows the reference's BinaryHeap would ever admit.  Passing keys go to
//   (a) the CTA's sorted top-k list  -> cta_topk[b][0..kcap)   key = bits(d) << 32 | position
//   (b) the entrant list             -> ent[first_pos_of_cta ...], ent_count[b]
// ------------------------------------------------------------------------------------------------
struct ScanParams {
    const float *data;
    const float *query;       // device
    const uint32_t *row_ids;  // device, GATHER only
    u64 n;                    // candidates to scan
    const u64 *n_dev;         // device, may be null: overrides n at run time (candidate count produced on the device
                              // by ivf_rank_kernel; n is then only the bound the launch was sized for)
    uint32_t dim;
    uint32_t k;
    uint32_t kcap;     // pow2 >= max(k, 32)
    uint32_t sort_n;   // pow2 >= kcap + flush_at + WARPS*32
    uint32_t flush_at;
    uint32_t cap_bits;  // static cap on tau (0xFFFFFFFF = none)
    const u64 *carry;   // device, may be null: ascending top-k keys of all rows that precede this launch;
                        // its k-th distance caps tau (streaming pushes)
    uint32_t pos_base;  // added to positions in emitted keys
    u64 *cta_topk;      // [grid][kcap]
    u64 *ent;           // [n]
    uint32_t *ent_count;  // [grid]
    float *dist_out;    // may be null: distance of candidate i of this launch at dist_out[i] (streaming top-k keeps a log of
                        // every distance so that a NaN can be answered by replaying the reference loop literally)
};

// A NaN distance breaks the threshold structure of the reference heap (src/ivf/search.rs:121-122: `d < top` is false
// against a NaN root and for a NaN candidate).  The kernels therefore never decide anything about such a row: its key is
// always emitted as an entrant, and the host, on seeing one, replays the reference loop over ALL candidates.
__device__ __forceinline__ bool is_nan_bits(const uint32_t bits) { return (bits & 0x7FFFFFFFu) > 0x7F800000u; }

__device__ __forceinline__ void bitonic_sort_smem(u64 *s, const uint32_t n, const uint32_t tid,
                                                  const uint32_t nthreads) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n >> 1); t += nthreads) {
                const uint32_t i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const uint32_t j = i + stride;
                const bool up = (i & size) == 0;
                const u64 a = s[i], b = s[j];
                if ((a > b) == up) {
                    s[i] = b;
                    s[j] = a;
                }
            }
            __syncthreads();
        }
    }
}

template <int ORDER, bool VEC4, bool GATHER, int WARPS, int RB = ScanDefaults<ORDER, VEC4, GATHER>::RB,
          int CBV = ScanDefaults<ORDER, VEC4, GATHER>::CBV, int MINB = 2>
__global__ void __launch_bounds__(WARPS * 32, MINB) l2_scan_topk_kernel(const ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE_FLOATS = TileCfg<ORDER, VEC4, CBV>::TILE_FLOATS;
    constexpr uint32_t NT = WARPS * 32;
    const uint32_t dim_pad = (p.dim + 3u) & ~3u;
    float *s_query = reinterpret_cast<float *>(smem_raw);
    float *s_tiles = s_query + dim_pad;
    u64 *s_keys = reinterpret_cast<u64 *>(s_tiles + WARPS * TILE_FLOATS);
    __shared__ uint32_t s_buf_count;
    __shared__ uint32_t s_tau;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    u64 n_rows = p.n;  // dense: stays a constant-bank operand (a register copy costs spills at 128 regs)
    if constexpr (GATHER) {
        if (p.n_dev) n_rows = *p.n_dev;
    }
    const u64 NG = (n_rows + 31) >> 5;
    const u64 g_begin = NG * blockIdx.x / gridDim.x;
    const u64 g_end = NG * (blockIdx.x + 1) / gridDim.x;

    uint32_t cap_bits = p.cap_bits;
    if (p.carry) {
        const u64 kth = p.carry[p.k - 1];
        const uint32_t t = (kth == KEY_MAX) ? 0xFFFFFFFFu : (uint32_t)(kth >> 32);
        cap_bits = t < cap_bits ? t : cap_bits;
    }
    for (uint32_t i = tid; i < p.dim; i += NT) s_query[i] = p.query[i];
    for (uint32_t i = tid; i < p.sort_n; i += NT) s_keys[i] = KEY_MAX;
    if (tid == 0) {
        s_buf_count = 0;
        s_tau = cap_bits;
    }
    __syncthreads();

    float *tile = s_tiles + warp * TILE_FLOATS;
    u64 *s_buf = s_keys + p.kcap;
    const u64 ent_base = g_begin * 32;
    uint32_t ent_written = 0;

    auto flush = [&]() {
        const uint32_t n_new = s_buf_count;
        __syncthreads();
        for (uint32_t i = tid; i < n_new; i += NT) p.ent[ent_base + ent_written + i] = s_buf[i];
        ent_written += n_new;
        for (uint32_t i = p.kcap + n_new + tid; i < p.sort_n; i += NT) s_keys[i] = KEY_MAX;
        __syncthreads();
        bitonic_sort_smem(s_keys, p.sort_n, tid, NT);
        for (uint32_t i = p.k + tid; i < p.kcap; i += NT) s_keys[i] = KEY_MAX;
        if (tid == 0) {
            s_buf_count = 0;
            const u64 kth = s_keys[p.k - 1];
            const uint32_t t = (kth == KEY_MAX) ? 0xFFFFFFFFu : (uint32_t)(kth >> 32);
            s_tau = t < cap_bits ? t : cap_bits;
        }
        __syncthreads();
    };

    for (u64 g0 = g_begin; g0 < g_end; g0 += WARPS) {
        const u64 g = g0 + warp;
        const bool active = g < g_end;  // warp-uniform
        float d = 0.f;
        if (active) d = group_distance<ORDER, VEC4, GATHER, RB, CBV>(p.data, p.row_ids, n_rows, p.dim, g, s_query, tile, lane);
        const u64 pos = g * 32 + lane;
        const uint32_t bits = __float_as_uint(d);
        const bool pass = active && pos < n_rows && (bits < s_tau || is_nan_bits(bits));
        if (p.dist_out && active && pos < n_rows) p.dist_out[pos] = d;
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        bool crossed = false;
        if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_buf_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) s_buf[base + __popc(m & ((1u << lane) - 1u))] = ((u64)bits << 32) | (u64)(uint32_t)(pos + p.pos_base);
            // the count only grows inside an iteration, so the warp whose append ends at the final
            // count sees whether the flush mark was reached; OR-reduce that across the CTA.
            crossed = base + __popc(m) >= p.flush_at;
        }
        if (__syncthreads_or(crossed)) flush();  // one barrier, CTA-uniform decision
    }
    if (s_buf_count > 0) flush();

    u64 *out = p.cta_topk + (u64)blockIdx.x * p.kcap;
    for (uint32_t i = tid; i < p.kcap; i += NT) out[i] = s_keys[i];
    if (tid == 0) p.ent_count[blockIdx.x] = ent_written;
}

// ------------------------------------------------------------------------------------------------
// topk_seq_merge_kernel: CTA j walks the sorted lists [j*lists_per_cta, min(n_lists,(j+1)*lists_per_cta))
// in order, keeping the running exact top-k P (ascending by (distance, position)); blockDim.x == kcap.
//   excl_prefix_out[b] (kcap keys) = P BEFORE list b is merged (this CTA's exclusive prefix)
//   total_out[j]       (kcap keys) = P after the CTA's last list
// P starts from `carry` (kcap keys, may be null = empty).  The next lists are prefetched into registers
// (depth PF) so the sequential chain is not exposed to global-memory latency.
// Used twice (two-level scan): level 1 with lists_per_cta = MERGE_GROUP over the scan CTAs' lists
// (-> W[b], T[j]); level 2 with one CTA over the group totals T (-> GP[j], final top-k).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t MERGE_GROUP = 16;

__global__ void __launch_bounds__(1024) topk_seq_merge_kernel(const u64 *__restrict__ lists, const uint32_t n_lists,
                                                              const uint32_t lists_per_cta, const uint32_t k,
                                                              const uint32_t kcap, const u64 *__restrict__ carry,
                                                              u64 *__restrict__ excl_prefix_out,
                                                              u64 *__restrict__ total_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s = reinterpret_cast<u64 *>(smem_raw);  // 2*kcap keys: two exchange buffers for the strides that cross warps
    constexpr int PF = 4;
    const uint32_t tid = threadIdx.x;  // blockDim.x == kcap (a power of two >= 32)
    const uint32_t b0 = blockIdx.x * lists_per_cta;
    const uint32_t b1 = min(n_lists, b0 + lists_per_cta);
    // Thread t keeps P[t] in a register.  One step: the incoming list arrives REVERSED (thread t holds its (kcap-1-t)-th
    // key), so [P | incoming reversed] is bitonic and x[t] = min(P[t], incoming[kcap-1-t]) is the half-cleaner's lower half:
    // the kcap smallest keys of the union as a bitonic sequence.  It is sorted with compare-exchanges at strides kcap/2 .. 1;
    // strides >= 32 go through shared memory (alternating buffers: one barrier each), strides < 32 through shuffles.
    u64 P = carry ? carry[tid] : KEY_MAX;
    u64 nxt[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) nxt[i] = (b0 + i < b1) ? lists[(u64)(b0 + i) * kcap + (kcap - 1 - tid)] : KEY_MAX;
    uint32_t buf = 0;
    for (uint32_t b = b0; b < b1; b += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            if (b + i >= b1) break;  // uniform
            if (excl_prefix_out) excl_prefix_out[(u64)(b + i) * kcap + tid] = P;
            u64 x = min(P, nxt[i]);
            const uint32_t bn = b + i + PF;
            nxt[i] = (bn < b1) ? lists[(u64)bn * kcap + (kcap - 1 - tid)] : KEY_MAX;
            for (uint32_t stride = kcap >> 1; stride >= 32; stride >>= 1) {
                u64 *sb = s + buf * kcap;
                sb[tid] = x;
                __syncthreads();
                const u64 y = sb[tid ^ stride];
                x = (tid & stride) ? max(x, y) : min(x, y);
                buf ^= 1u;
            }
#pragma unroll
            for (uint32_t stride = 16; stride > 0; stride >>= 1) {
                const u64 y = __shfl_xor_sync(0xffffffffu, x, stride);
                x = (tid & stride) ? max(x, y) : min(x, y);
            }
            P = tid < k ? x : KEY_MAX;
        }
    }
    total_out[(u64)blockIdx.x * kcap + tid] = P;
}

// ------------------------------------------------------------------------------------------------
// entrant_filter_kernel: CTA b first derives gthr[b] = distance bits of the k-th smallest key among all
// scan CTAs < b (and the carry) = k-th smallest of GP[b / MERGE_GROUP] U W[b] (two sorted lists, merge
// path by rank), i.e. the reference heap's root when the scan reaches CTA b's first row.  It then keeps
// the entrant keys of scan-CTA b whose distance is below it and appends them (order irrelevant, the host
// sorts by position) to out[1..]; out[0] = running count.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_keys(const u64 *a, const uint32_t n, const u64 x) {
    uint32_t lo = 0, hi = n;  // first index with a[idx] >= x
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) entrant_filter_kernel(const u64 *__restrict__ ent,
                                                             const uint32_t *__restrict__ ent_count,
                                                             const u64 *__restrict__ group_prefix,
                                                             const u64 *__restrict__ within_prefix, const uint32_t k,
                                                             const uint32_t kcap, uint32_t *__restrict__ gthr,
                                                             const u64 n_arg, const u64 *__restrict__ n_dev,
                                                             u64 *__restrict__ out, const uint32_t out_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *sA = reinterpret_cast<u64 *>(smem_raw);  // GP list
    u64 *sB = sA + kcap;                           // W list
    __shared__ uint32_t s_thr;
    const uint32_t b = blockIdx.x, G = gridDim.x;
    const u64 *A = group_prefix + (u64)(b / MERGE_GROUP) * kcap;
    const u64 *B = within_prefix + (u64)b * kcap;
    for (uint32_t i = threadIdx.x; i < kcap; i += blockDim.x) {
        sA[i] = A[i];
        sB[i] = B[i];
    }
    if (threadIdx.x == 0) s_thr = 0xFFFFFFFFu;
    __syncthreads();
    // keys are unique (position field), so exactly one real key has merged rank k-1 when >= k keys exist
    for (uint32_t i = threadIdx.x; i < 2 * kcap; i += blockDim.x) {
        const bool inA = i < kcap;
        const uint32_t idx = inA ? i : i - kcap;
        const u64 key = inA ? sA[idx] : sB[idx];
        if (key != KEY_MAX) {
            const uint32_t rank = idx + lower_bound_keys(inA ? sB : sA, kcap, key);
            if (rank == k - 1) s_thr = (uint32_t)(key >> 32);
        }
    }
    __syncthreads();
    const uint32_t thr = s_thr;
    if (threadIdx.x == 0) gthr[b] = thr;

    const u64 n = n_dev ? *n_dev : n_arg;  // must be what the scan kernel used (same CTA row ranges)
    const u64 NG = (n + 31) >> 5;
    const u64 base = (NG * b / G) * 32;
    const uint32_t cnt = ent_count[b];
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t i0 = 0; i0 < cnt; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        u64 key = 0;
        bool keep = false;
        if (i < cnt) {
            key = ent[base + i];
            keep = (uint32_t)(key >> 32) < thr || is_nan_bits((uint32_t)(key >> 32));
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            u64 slot = 0;
            if (lane == 0) slot = atomicAdd(out, (u64)__popc(m));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            const u64 idx = slot + __popc(m & ((1u << lane) - 1u));
            if (keep && idx < out_cap) out[1 + idx] = key;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// l2_dist_kernel: plain distance sweep in index.rs order (a = row, b = vec as in index.rs:350/362
// `squared_l2_distance(vec, centroid)`), optional row selection, store or k-means++ min-update
// (index.rs:363-365: if dist < slot { slot = dist }).
// ------------------------------------------------------------------------------------------------
template <bool VEC4, bool GATHER, int WARPS, int RB = 8, int CBV = 1, bool HINT = false>
__global__ void __launch_bounds__(WARPS * 32, 2) l2_dist_kernel(const float *__restrict__ data,
                                                             const uint32_t *__restrict__ row_ids, const u64 n,
                                                             const uint32_t dim, const float *__restrict__ vec,
                                                             float *__restrict__ out, const int min_update,
                                                             float *__restrict__ mirror = nullptr, const u64 keep_groups = 0) {
    // mirror (may be null): page-locked host memory that receives the same final values while the kernel runs, so the
    // k-means++ loop (1023 dependent sweeps, each followed by a host-side pick) needs no separate read-back copy
    // HINT: the rows of groups < keep_groups are loaded evict_last, the others evict_first -- a caller that sweeps the same
    // rows again and again (k-means++: 1023 sweeps of a 154 MB init set) keeps the part of them that fits in L2 resident
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE_FLOATS = TileCfg<0, VEC4, VEC4 ? CBV : 1>::TILE_FLOATS;
    const uint32_t dim_pad = (dim + 3u) & ~3u;
    float *s_vec = reinterpret_cast<float *>(smem_raw);
    float *s_tiles = s_vec + dim_pad;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < dim; i += WARPS * 32) s_vec[i] = vec[i];
    __syncthreads();
    float *tile = s_tiles + warp * TILE_FLOATS;
    const u64 NG = (n + 31) >> 5;
    unsigned long long pol_keep = 0ull, pol_go = 0ull;
    if constexpr (HINT) {
        pol_keep = l2_policy_evict_last();
        pol_go = l2_policy_evict_first();
    }
    for (u64 g = (u64)blockIdx.x * WARPS + warp; g < NG; g += (u64)gridDim.x * WARPS) {
        const float d = group_distance<0, VEC4, GATHER, RB, VEC4 ? CBV : 1, HINT>(data, row_ids, n, dim, g, s_vec, tile, lane,
                                                                                  g < keep_groups ? pol_keep : pol_go);
        const u64 pos = g * 32 + lane;
        if (pos < n) {
            float keep = d;
            if (min_update) {
                const float old = out[pos];
                if (d < old) out[pos] = d;
                else keep = old;
            } else {
                out[pos] = d;
            }
            if (mirror) mirror[pos] = keep;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Batched centroid ranking (find_closest_centroids for every query of a batch, index.rs:130-149; SURVEY row a6 "becomes
// a (queries x centroids) contraction when batched").  The table is short (C x dim, L2-resident), so the contraction
// stays in the reference's exact f32 order on the SIMT path: blockIdx.y = query, the CTA stages that query and its warps
// take groups of 32 centroids.  rank_batch_kernel then sorts each query's (distance bits, cluster) keys in shared memory
// -- for non-NaN distances exactly the reference's stable ascending sort -- and writes the first np cluster ids; a NaN
// distance raises the query's flag and the host ranks that query with the reference comparator instead.
// ------------------------------------------------------------------------------------------------
template <bool VEC4, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) l2_dist_batch_kernel(const float *__restrict__ data, const u64 n,
                                                                   const uint32_t dim, const float *__restrict__ vecs,
                                                                   float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE_FLOATS = TileCfg<0, VEC4>::TILE_FLOATS;
    const uint32_t dim_pad = (dim + 3u) & ~3u;
    float *s_vec = reinterpret_cast<float *>(smem_raw);
    float *s_tiles = s_vec + dim_pad;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *vec = vecs + (size_t)blockIdx.y * dim;
    float *o = out + (size_t)blockIdx.y * n;
    for (uint32_t i = tid; i < dim; i += WARPS * 32) s_vec[i] = vec[i];
    __syncthreads();
    float *tile = s_tiles + warp * TILE_FLOATS;
    const u64 NG = (n + 31) >> 5;
    for (u64 g = (u64)blockIdx.x * WARPS + warp; g < NG; g += (u64)gridDim.x * WARPS) {
        // index.rs:138 squared_l2_distance(query, centroid): a = query (s_vec), b = row -- group_distance<0> order
        const float d = group_distance<0, VEC4, false>(data, nullptr, n, dim, g, s_vec, tile, lane);
        const u64 pos = g * 32 + lane;
        if (pos < n) o[pos] = d;
    }
}

__global__ void __launch_bounds__(1024) rank_batch_kernel(const float *__restrict__ cdist, const uint32_t C,
                                                          const uint32_t cp2, const uint32_t np,
                                                          uint32_t *__restrict__ out_ids, uint32_t *__restrict__ nan_flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s = reinterpret_cast<u64 *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const float *d_q = cdist + (size_t)blockIdx.x * C;
    bool bad = false;
    for (uint32_t i = tid; i < cp2; i += 1024) {
        u64 key = KEY_MAX;
        if (i < C) {
            const float d = d_q[i];
            bad |= (d != d);
            key = ((u64)__float_as_uint(d) << 32) | i;
        }
        s[i] = key;
    }
    const int any_bad = __syncthreads_or(bad);
    if (tid == 0) nan_flags[blockIdx.x] = any_bad ? 1u : 0u;
    bitonic_sort_smem(s, cp2, tid, 1024);
    for (uint32_t r = tid; r < np; r += 1024) out_ids[(size_t)blockIdx.x * np + r] = (uint32_t)s[r];
}

// ------------------------------------------------------------------------------------------------
// l2_dist_wide_kernel: the same sweep for SHORT tables (centroid ranking: C rows, index.rs:130-149).  With only a
// few groups of 32 rows, one warp per group leaves the machine empty and the sweep becomes a chain of memory round
// trips; here one CTA of 8 warps owns a group, warp w loads column blocks w, w+8, ... of all 32 rows and writes the
// chain terms to a [32][dim/4] shared tile, then warp 0 runs the 32 serial chains (lane = row) in the reference order.
// dim % 4 == 0, 16-byte aligned rows; ts = row stride of the tile in floats (ts % 4 == 0, ts/4 odd).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_dist_wide_kernel(const float *__restrict__ data,
                                                           const uint32_t *__restrict__ row_ids, const u64 n,
                                                           const uint32_t dim, const uint32_t ts,
                                                           const float *__restrict__ vec, float *__restrict__ out,
                                                           const int min_update) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_terms = reinterpret_cast<float *>(smem_raw);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 g_first = (u64)blockIdx.x * 32;
    const uint32_t ncb = (dim + 127u) / 128u;
    for (uint32_t cb = warp; cb < ncb; cb += 8) {
        const uint32_t col0 = cb * 128u + (lane << 2);
        if (col0 < dim) {
            const float4 q4 = *reinterpret_cast<const float4 *>(vec + col0);
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 16) {
                float4 v4[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    u64 pos = g_first + r0 + j;
                    pos = pos < n ? pos : n - 1;
                    const u64 row = row_ids ? (u64)row_ids[pos] : pos;
                    v4[j] = ld_stream_v4(data + row * dim + col0);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) s_terms[(r0 + j) * ts + (col0 >> 2)] = chunk4(q4, v4[j]);
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        const float *trow = s_terms + lane * ts;
        const uint32_t nt = dim >> 2;
        float sum = 0.0f;
        uint32_t j = 0;
        for (; j + 4 <= nt; j += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(trow + j);
            sum = __fadd_rn(sum, t.x);
            sum = __fadd_rn(sum, t.y);
            sum = __fadd_rn(sum, t.z);
            sum = __fadd_rn(sum, t.w);
        }
        for (; j < nt; ++j) sum = __fadd_rn(sum, trow[j]);
        const u64 pos = g_first + lane;
        if (pos < n) {
            if (min_update) {
                if (sum < out[pos]) out[pos] = sum;
            } else {
                out[pos] = sum;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kmeans_assign_kernel: exact f32 argmin over all centroids for a tile of 64 rows
// (src/ivf/index.rs:244-257 / 405-415).  Register-tiled SIMT: 256 threads = 16 (centroid dir) x 16
// (row dir), each thread owns 4 rows x 4 centroids; every (row, centroid) pair keeps its own serial
// chain in the reference order, operands staged through shared memory in blocks of 32 columns.
// Ties: lowest centroid index (strict '<' while scanning centroids upwards); NaN/inf never win.
// ------------------------------------------------------------------------------------------------
constexpr int AS_BM = 64, AS_BN = 64, AS_BK = 32, AS_LD = AS_BK + 4;
constexpr uint32_t FEW_ROWS_MAX = 256;  // device-side row lists up to this length take few_rows_assign_kernel

// GATHER: the tile's rows are row_ids[0 .. *n_dev) (a device-side list: the rows the tensor-core filter could not
// decide).  The grid covers the worst case in x and surplus CTAs exit at once; blockIdx.y selects a slice of slice_len
// centroids (a multiple of AS_BN) so that a handful of rows does not serialise behind one CTA's walk over the whole table.
// Each slice folds its winner into best[row] = min(bits(distance) << 32 | centroid): distances that can win are finite and
// >= 0, so the u64 order is (distance, index) -- the strict-'<' ascending scan; best[row] stays KEY_MAX if nothing wins.
template <bool VEC4, bool GATHER = false>
__global__ void __launch_bounds__(256, 2) kmeans_assign_kernel(const float *__restrict__ rows, const u64 n_arg,
                                                               const uint32_t dim,
                                                               const float *__restrict__ centroids,
                                                               const uint32_t n_clusters,
                                                               uint32_t *__restrict__ out_assign,
                                                               const uint32_t *__restrict__ row_ids = nullptr,
                                                               const uint32_t *__restrict__ n_dev = nullptr,
                                                               u64 *__restrict__ best = nullptr,
                                                               const uint32_t slice_len = 0) {
    __shared__ __align__(16) float As[2][AS_BM * AS_LD];
    __shared__ __align__(16) float Bs[2][AS_BN * AS_LD];
    const uint32_t tid = threadIdx.x;
    const uint32_t tx = tid & 15, ty = tid >> 4;
    const u64 n = GATHER ? (u64)*n_dev : n_arg;
    if (GATHER && n <= FEW_ROWS_MAX) return;            // short lists: few_rows_assign_kernel
    const uint32_t n4 = dim >> 2;                       // full 4-chunks (chain terms)
    const uint32_t nkb = (n4 + 7) >> 3;                 // blocks of 8 chunks = 32 columns
    const uint32_t tail0 = n4 << 2, ntail = dim - tail0;  // scalar tail terms (dim % 4)

    // loader mapping: 64 rows x 8 chunks = 512 float4 per tile -> 2 per thread
    auto load_tile = [&](float *dst, const float *src, const u64 first, const u64 limit, const uint32_t kb,
                         const bool via_ids = false) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const uint32_t idx = tid + it * 256;
            const uint32_t r = idx >> 3, c = idx & 7;
            const uint32_t chunk = kb * 8 + c;
            u64 row = first + r;
            row = row < limit ? row : limit - 1;
            if (GATHER && via_ids) row = row_ids[row];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chunk < n4) {
                const float *sp = src + row * dim + (chunk << 2);
                if constexpr (VEC4) v = *reinterpret_cast<const float4 *>(sp);
                else v = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
            *reinterpret_cast<float4 *>(dst + r * AS_LD + (c << 2)) = v;
        }
    };

    // dense: one row tile per CTA (the grid covers the table).  GATHER: the list length is only known on the device, so
    // a fixed grid walks the tiles (a grid sized for the worst case would launch a million CTAs that exit at once)
    for (u64 row0 = (u64)blockIdx.x * AS_BM; row0 < n; row0 += (u64)gridDim.x * AS_BM) {
    float run_d[4];
    uint32_t run_i[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        run_d[i] = __int_as_float(0x7f800000);
        run_i[i] = 0;
    }

    const uint32_t c_begin = GATHER ? blockIdx.y * slice_len : 0u;
    const uint32_t c_end = GATHER ? min(n_clusters, c_begin + slice_len) : n_clusters;
    for (uint32_t cn = c_begin; cn < c_end; cn += AS_BN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        int buf = 0;
        if (nkb > 0) {
            load_tile(As[0], rows, row0, n, 0, true);
            load_tile(Bs[0], centroids, cn, n_clusters, 0);
        }
        __syncthreads();
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            if (kb + 1 < nkb) {
                load_tile(As[buf ^ 1], rows, row0, n, kb + 1, true);
                load_tile(Bs[buf ^ 1], centroids, cn, n_clusters, kb + 1);
            }
            const float *A = As[buf] + (ty * 4) * AS_LD;
            const float *B = Bs[buf] + tx * AS_LD;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(A + i * AS_LD + (c << 2));
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4 *>(B + (16 * j) * AS_LD + (c << 2));
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], chunk4(a[i], b[j]));
            }
            __syncthreads();
            buf ^= 1;
        }
        // scalar tail (dim % 4 != 0): each leftover column is its own chain term (index.rs:474-478)
        for (uint32_t t = 0; t < ntail; ++t) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u64 row = row0 + ty * 4 + i;
                row = row < n ? row : n - 1;
                if (GATHER) row = row_ids[row];
                const float av = rows[row * dim + tail0 + t];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t c = cn + tx + 16 * j;
                    c = c < n_clusters ? c : n_clusters - 1;
                    acc[i][j] = __fadd_rn(acc[i][j], sq1(av, centroids[(u64)c * dim + tail0 + t]));
                }
            }
        }
        // argmin over this centroid tile, first-min on ties
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float bd = __int_as_float(0x7f800000);
            uint32_t bi = 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t c = cn + tx + 16 * j;
                const float d = acc[i][j];
                if (c < n_clusters && d < bd) {  // NaN fails '<'; inf fails '<' against inf
                    bd = d;
                    bi = c;
                }
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, off);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (od < bd || (od == bd && oi < bi)) {
                    bd = od;
                    bi = oi;
                }
            }
            if (bd < run_d[i]) {
                run_d[i] = bd;
                run_i[i] = bi;
            }
        }
        __syncthreads();
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u64 row = row0 + ty * 4 + i;
            if (row < n) {
                if constexpr (GATHER) {
                    if (run_d[i] < __int_as_float(0x7f800000))
                        atomicMin(reinterpret_cast<unsigned long long *>(&best[row_ids[row]]),
                                  ((u64)__float_as_uint(run_d[i]) << 32) | (u64)run_i[i]);
                } else {
                    out_assign[row] = run_i[i];
                }
            }
        }
    }
    }  // row tiles
}

// ------------------------------------------------------------------------------------------------
// few_rows_assign_kernel: the same exact argmin pieces for a SHORT device-side list of rows (the rows the tensor-core filter
// hands to the full scan: a few per million).  The 64-row tiles of kmeans_assign_kernel<.., GATHER> spend ~70 us of
// barriers and k-blocks on four rows; here one warp takes (row, 32 centroids): the row sits in the warp's shared-memory
// slot as the "query", the 32 centroids are the "rows" of group_distance<0> (squared_l2_distance(row, centroid): the
// difference's sign does not reach the square), lane j holds centroid j's exact chain, the warp's finite minimum folds into
// best[row] = min(bits(distance) << 32 | centroid) like the tiled kernel's.  Lists longer than FEW_ROWS_MAX are left to the
// tiled kernel (which re-uses its operand tiles); both kernels are launched and the device-side count picks one.
// Dynamic shared memory: warps * (dim_pad + TileCfg<0, VEC4>::TILE_FLOATS) floats.
// ------------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256) few_rows_assign_kernel(const float *__restrict__ rows, const uint32_t dim,
                                                              const float *__restrict__ centroids, const uint32_t n_clusters,
                                                              const uint32_t *__restrict__ row_ids,
                                                              const uint32_t *__restrict__ n_dev, u64 *__restrict__ best) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE_FLOATS = TileCfg<0, VEC4>::TILE_FLOATS;
    const uint32_t n = *n_dev;
    if (n == 0 || n > FEW_ROWS_MAX) return;
    const uint32_t dim_pad = (dim + 3u) & ~3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    float *s_vec = reinterpret_cast<float *>(smem_raw) + (size_t)warp * (dim_pad + TILE_FLOATS);
    float *tile = s_vec + dim_pad;
    const uint32_t nslices = (n_clusters + 31u) >> 5;
    const uint32_t items = n * nslices;
    uint32_t have_row = 0xFFFFFFFFu;
    for (uint32_t item = blockIdx.x * warps + warp; item < items; item += gridDim.x * warps) {
        const uint32_t r = item / nslices, sl = item - r * nslices;
        const uint32_t row = row_ids[r];
        if (row != have_row) {
            __syncwarp();
            for (uint32_t i = lane; i < dim; i += 32) s_vec[i] = rows[(u64)row * dim + i];
            have_row = row;
            __syncwarp();
        }
        const float d = group_distance<0, VEC4, false>(centroids, nullptr, (u64)n_clusters, dim, (u64)sl, s_vec, tile, lane);
        const uint32_t c = (sl << 5) + lane;
        u64 key = KEY_MAX;
        if (c < n_clusters && d < __int_as_float(0x7f800000)) key = ((u64)__float_as_uint(d) << 32) | (u64)c;  // NaN / inf never win
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) key = min(key, (u64)__shfl_xor_sync(0xffffffffu, key, off));
        if (lane == 0 && key != KEY_MAX) atomicMin(reinterpret_cast<unsigned long long *>(&best[row]), key);
    }
}

// ------------------------------------------------------------------------------------------------
// synthetic data: uniform [0,1) on the 24-bit grid ((u32 >> 8) * 2^-24, the distribution of
// rand's gen::<f32>() used by benches/bench_util.rs:29-41), counter-based and keyed by
// (seed, absolute element index) so any row can be regenerated anywhere (stream defined in
// DESIGN.md section 6; the test-side checker regenerates the same stream on the CPU).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t synth_u32(const u64 seed, const u64 idx) {
    u64 z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

__global__ void __launch_bounds__(256) synth_fill_kernel(float *__restrict__ out, const u64 first_elem,
                                                         const u64 n_elems, const u64 seed) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride)
        out[i] = (float)(synth_u32(seed, first_elem + i) >> 8) * (1.0f / 16777216.0f);
}

// f64 -> f32 narrowing of a pushed batch (src/df_vector/exec.rs:542 `value as f32`)
__global__ void __launch_bounds__(256) narrow_f64_kernel(const double *__restrict__ in, float *__restrict__ out,
                                                         const u64 n) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i];
}

// ------------------------------------------------------------------------------------------------
// IVF support kernels
// ------------------------------------------------------------------------------------------------
// dense copy of selected rows (k-means sample, src/ivf/index.rs:234-239): out[i] = data[ids[i]]
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ data,
                                                          const uint32_t *__restrict__ ids, const u64 n_ids,
                                                          const uint32_t dim, float *__restrict__ out) {
    const u64 total = n_ids * dim;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const u64 r = i / dim;
        const uint32_t c = (uint32_t)(i - r * dim);
        out[i] = data[(u64)ids[r] * dim + c];
    }
}

// candidate_rows (src/ivf/index.rs:57-63): concatenate the inverted lists of the probed clusters in
// rank order.  probe_cluster[r] = cluster id of rank r, probe_prefix[r] = first candidate position of
// rank r (probe_prefix[nprobe] = n_cand).  blockIdx.x = probed cluster, blockIdx.y = slice of its list.
__global__ void __launch_bounds__(256) ivf_expand_kernel(const uint32_t *__restrict__ list_ids,
                                                         const u64 *__restrict__ list_offsets,
                                                         const uint32_t *__restrict__ probe_cluster,
                                                         const u64 *__restrict__ probe_prefix,
                                                         uint32_t *__restrict__ out_rows) {
    const uint32_t r = blockIdx.x;
    const uint32_t c = probe_cluster[r];
    const u64 src = list_offsets[c], len = list_offsets[c + 1] - src, dst = probe_prefix[r];
    for (u64 i = (u64)blockIdx.y * blockDim.x + threadIdx.x; i < len; i += (u64)gridDim.y * blockDim.x)
        out_rows[dst + i] = list_ids[src + i];
}

// ------------------------------------------------------------------------------------------------
// VectorTopKExec's candidate handling on the device (src/df_vector/exec.rs:207-245, access.rs:107-176): of the probed
// lists' rows in rank order only the first `max_candidates` count (CandidateCursor over one file = a prefix), the rows
// are then visited in FILE order (RowSelection per row group) and the scan subtree's filter drops rows before they are
// scored.  On the device: the kept candidates set their bit in an N-bit map, the map is ANDed with the caller's filter
// bitmap and compacted in order -> ascending row ids, count in info[0] (what the gathered scan reads through n_dev);
// info[2] keeps the un-capped, un-filtered candidate count (the `candidate_rows` metric of VectorIndexScanExec).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ivf_mark_kernel(const uint32_t *__restrict__ list_ids,
                                                       const u64 *__restrict__ list_offsets,
                                                       const uint32_t *__restrict__ probe_cluster,
                                                       const u64 *__restrict__ probe_prefix, const u64 max_candidates,
                                                       const u64 n_rows, uint32_t *__restrict__ bitmap) {
    const uint32_t r = blockIdx.x;
    const uint32_t c = probe_cluster[r];
    const u64 src = list_offsets[c], len = list_offsets[c + 1] - src, dst = probe_prefix[r];
    for (u64 i = (u64)blockIdx.y * blockDim.x + threadIdx.x; i < len; i += (u64)gridDim.y * blockDim.x) {
        if (dst + i >= max_candidates) break;  // positions only grow along the list
        const uint32_t row = list_ids[src + i];
        if (row < n_rows) atomicOr(&bitmap[row >> 5], 1u << (row & 31));
    }
}

constexpr uint32_t BM_WORDS_PER_THREAD = 4;
constexpr uint32_t BM_WORDS_PER_BLOCK = 256 * BM_WORDS_PER_THREAD;

__global__ void __launch_bounds__(256) bitmap_count_kernel(const uint32_t *__restrict__ bitmap,
                                                           const uint32_t *__restrict__ mask, const u64 n_words,
                                                           uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t s_warp[8];
    const u64 w0 = ((u64)blockIdx.x * 256 + threadIdx.x) * BM_WORDS_PER_THREAD;
    uint32_t cnt = 0;
#pragma unroll
    for (uint32_t i = 0; i < BM_WORDS_PER_THREAD; ++i)
        if (w0 + i < n_words) cnt += __popc(bitmap[w0 + i] & (mask ? mask[w0 + i] : 0xFFFFFFFFu));
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += s_warp[i];
        block_sums[blockIdx.x] = t;
    }
}

// one CTA: block_sums -> exclusive prefix in place; info[2] = info[0] (all probed candidates), info[0] = rows kept
__global__ void __launch_bounds__(1024) bitmap_scan_kernel(uint32_t *__restrict__ block_sums, const uint32_t n_blocks,
                                                           u64 *__restrict__ info) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        const uint32_t i = base + tid;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if ((int)lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) {
        info[2] = info[0];
        info[0] = s_carry;
    }
}

__global__ void __launch_bounds__(256) bitmap_compact_kernel(const uint32_t *__restrict__ bitmap,
                                                             const uint32_t *__restrict__ mask, const u64 n_words,
                                                             const uint32_t *__restrict__ block_offsets,
                                                             uint32_t *__restrict__ out_rows) {
    __shared__ uint32_t s_warp[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 w0 = ((u64)blockIdx.x * 256 + tid) * BM_WORDS_PER_THREAD;
    uint32_t bits[BM_WORDS_PER_THREAD];
    uint32_t cnt = 0;
#pragma unroll
    for (uint32_t i = 0; i < BM_WORDS_PER_THREAD; ++i) {
        bits[i] = (w0 + i < n_words) ? (bitmap[w0 + i] & (mask ? mask[w0 + i] : 0xFFFFFFFFu)) : 0u;
        cnt += __popc(bits[i]);
    }
    uint32_t incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = block_offsets[blockIdx.x] + incl - cnt;
    for (uint32_t w = 0; w < warp; ++w) off += s_warp[w];
#pragma unroll
    for (uint32_t i = 0; i < BM_WORDS_PER_THREAD; ++i) {
        uint32_t b = bits[i];
        while (b) {
            const uint32_t j = __ffs(b) - 1;
            out_rows[off++] = (uint32_t)((w0 + i) << 5) + j;
            b &= b - 1;
        }
    }
}

// Batched IVF search (pqv_ivf_search_batch): row -> cluster map from the inverted lists, and the probe bit matrix
// probe_T[c][q / 32] from the ranked cluster ids of every query (rank_batch_kernel output, np per query).
__global__ void __launch_bounds__(256) row_cluster_kernel(const uint32_t *__restrict__ list_ids,
                                                          const u64 *__restrict__ list_offsets, const u64 n_rows,
                                                          uint32_t *__restrict__ row_cluster) {
    const uint32_t c = blockIdx.x;
    const u64 b = list_offsets[c], e = list_offsets[c + 1];
    for (u64 i = b + (u64)blockIdx.y * blockDim.x + threadIdx.x; i < e; i += (u64)gridDim.y * blockDim.x) {
        const uint32_t r = list_ids[i];
        if (r < n_rows) row_cluster[r] = c;
    }
}
__global__ void __launch_bounds__(256) probe_build_kernel(const uint32_t *__restrict__ ranked, const uint32_t nq,
                                                          const uint32_t np, const uint32_t *__restrict__ skip,
                                                          const uint32_t qwords, uint32_t *__restrict__ probe_T) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (u64)nq * np) return;
    const uint32_t q = (uint32_t)(i / np);
    if (skip[q]) return;  // NaN centroid distance: this query is ranked and answered on the host path
    const uint32_t c = ranked[i];
    atomicOr(&probe_T[(size_t)c * qwords + (q >> 5)], 1u << (q & 31));
}

// find_closest_centroids (src/ivf/index.rs:130-149) on the device: stable ascending sort of the C query-centroid
// distances, keep the first nprobe.  For distances that are not NaN (sums of squares: >= +0, or +inf) the stable
// sort under partial_cmp is the sort by (distance bits, cluster index); a NaN distance raises *nan_flag and the caller
// ranks on the host instead.  Also emits the candidate-position prefix of the probed lists and the candidate count
// (the scan kernel reads it through ScanParams::n_dev).  One CTA of 1024 threads, cp2 = pow2 >= C keys in shared memory.
__global__ void __launch_bounds__(1024) ivf_rank_kernel(const float *__restrict__ cdist, const uint32_t C,
                                                        const uint32_t cp2, const uint32_t nprobe,
                                                        const u64 *__restrict__ list_offsets,
                                                        uint32_t *__restrict__ probe_cluster,
                                                        u64 *__restrict__ probe_prefix, u64 *__restrict__ info,
                                                        u64 *__restrict__ zero_word) {
    // info[0] = candidate count, info[1] = 1 when a centroid distance is NaN, info[2], info[3] = 0; *zero_word = 0 (the
    // entrant counter of the scan that follows): the search needs no memset launches in front of this kernel
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s = reinterpret_cast<u64 *>(smem_raw);
    __shared__ u64 s_part[1024];
    const uint32_t tid = threadIdx.x;
    bool bad = false;
    for (uint32_t i = tid; i < cp2; i += 1024) {
        u64 key = KEY_MAX;
        if (i < C) {
            const float d = cdist[i];
            bad |= (d != d);
            key = ((u64)__float_as_uint(d) << 32) | i;
        }
        s[i] = key;
    }
    const int any_bad = __syncthreads_or(bad);
    if (tid == 0) {
        info[1] = any_bad ? 1u : 0u;
        info[2] = 0;
        info[3] = 0;
        if (zero_word) *zero_word = 0;
    }
    bitonic_sort_smem(s, cp2, tid, 1024);
    __syncthreads();
    const uint32_t per = (nprobe + 1023) / 1024;
    const uint32_t b = min(nprobe, tid * per), e = min(nprobe, b + per);
    u64 local = 0;
    for (uint32_t r = b; r < e; ++r) {
        const uint32_t c = (uint32_t)s[r];
        local += list_offsets[c + 1] - list_offsets[c];
    }
    s_part[tid] = local;
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the 1024 partials: 32 per lane, then across the warp
        u64 acc = 0;
        for (uint32_t i = 0; i < 32; ++i) acc += s_part[tid * 32 + i];
        u64 incl = acc;
        for (int o = 1; o < 32; o <<= 1) {
            const u64 v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)tid >= o) incl += v;
        }
        u64 run = incl - acc;
        for (uint32_t i = 0; i < 32; ++i) {
            const u64 v = s_part[tid * 32 + i];
            s_part[tid * 32 + i] = run;
            run += v;
        }
        if (tid == 31) {
            probe_prefix[nprobe] = incl;
            info[0] = incl;
        }
    }
    __syncthreads();
    u64 run = s_part[tid];
    for (uint32_t r = b; r < e; ++r) {
        const uint32_t c = (uint32_t)s[r];
        probe_cluster[r] = c;
        probe_prefix[r] = run;
        run += list_offsets[c + 1] - list_offsets[c];
    }
}

// info[0] = min(info[0], limit): a search over a prefix of the candidate sequence only (tie replay of a batched IVF search)
__global__ void clamp_count_kernel(u64 *__restrict__ info, const u64 limit) {
    if (info[0] > limit) info[0] = limit;
}

// row ids of the surviving entrant keys of a gathered scan: rows_out[i] = cand[position of key i]
// (keys = entrant_filter_kernel's out: [0] = count, keys from [1]).
// h_keys / h_rows / h_info (may be null): page-locked host memory that receives the count ([0] of h_keys), the first
// h_first keys and row ids and the four info words while the kernel runs, so that a search needs no device-to-host copy
// behind it (three small copies cost ~25 us of DMA set-up; these are a few KB of posted writes).
__global__ void __launch_bounds__(256) ent_rows_kernel(const u64 *__restrict__ ent_out, const uint32_t cap,
                                                       const uint32_t *__restrict__ cand,
                                                       uint32_t *__restrict__ rows_out, u64 *__restrict__ h_keys = nullptr,
                                                       uint32_t *__restrict__ h_rows = nullptr, const uint32_t h_first = 0,
                                                       const u64 *__restrict__ info = nullptr, u64 *__restrict__ h_info = nullptr) {
    const u64 total = ent_out[0];
    const u64 cnt = min(total, (u64)cap);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (u64)gridDim.x * blockDim.x) {
        const u64 key = ent_out[1 + i];
        const uint32_t row = cand[(uint32_t)key];
        rows_out[i] = row;
        if (h_keys && i < h_first) {
            h_keys[1 + i] = key;
            h_rows[i] = row;
        }
    }
    if (h_keys && blockIdx.x == 0) {
        if (threadIdx.x == 0) h_keys[0] = total;
        if (info && threadIdx.x < 4) h_info[threadIdx.x] = info[threadIdx.x];
    }
}

// ------------------------------------------------------------------------------------------------
// Inverted lists on the device (src/ivf/index.rs:202-206: per-cluster row ids, ascending) = a stable counting sort of
// the row ids by cluster.  Block b owns rows [b*R, (b+1)*R):
//   csr_count_kernel    counts[c*NB + b] = rows of block b assigned to cluster c              (shared-memory histogram)
//   csr_scan_kernel     per cluster: counts[c*NB + .] -> exclusive prefix over blocks, totals[c] = list length
//   csr_offsets_kernel  offsets[0..C] = exclusive prefix of totals (u64, the CSR offsets)
//   csr_scatter_kernel  ids[offsets[c] + counts[c*NB + b] + rank within the block] = row
// In the scatter the block walks its rows 256 at a time; inside a chunk the warps take their slots one after the other
// (one barrier per warp) and a warp ranks its lanes with match_any, so every list comes out in ascending row order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csr_count_kernel(const uint32_t *__restrict__ assign, const u64 n,
                                                        const uint32_t R, const uint32_t C, const uint32_t NB,
                                                        uint32_t *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);
    for (uint32_t c = threadIdx.x; c < C; c += 256) hist[c] = 0;
    __syncthreads();
    const u64 r0 = (u64)blockIdx.x * R, r1 = min(n, r0 + R);
    for (u64 r = r0 + threadIdx.x; r < r1; r += 256) atomicAdd(&hist[assign[r]], 1u);
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < C; c += 256) counts[(u64)c * NB + blockIdx.x] = hist[c];
}

__global__ void __launch_bounds__(256) csr_scan_kernel(uint32_t *__restrict__ counts, const uint32_t NB,
                                                       uint32_t *__restrict__ totals) {
    __shared__ uint32_t s_part[256];
    uint32_t *row = counts + (u64)blockIdx.x * NB;
    const uint32_t per = (NB + 255) / 256;
    const uint32_t b = min(NB, threadIdx.x * per), e = min(NB, b + per);
    uint32_t local = 0;
    for (uint32_t i = b; i < e; ++i) local += row[i];
    s_part[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t acc = 0;
        for (uint32_t i = 0; i < 8; ++i) acc += s_part[threadIdx.x * 8 + i];
        uint32_t incl = acc;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += v;
        }
        uint32_t run = incl - acc;
        for (uint32_t i = 0; i < 8; ++i) {
            const uint32_t v = s_part[threadIdx.x * 8 + i];
            s_part[threadIdx.x * 8 + i] = run;
            run += v;
        }
        if (threadIdx.x == 31) totals[blockIdx.x] = incl;
    }
    __syncthreads();
    uint32_t run = s_part[threadIdx.x];
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t v = row[i];
        row[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(1024) csr_offsets_kernel(const uint32_t *__restrict__ totals, const uint32_t C,
                                                           u64 *__restrict__ offsets) {
    __shared__ u64 s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (C + 1023) / 1024;
    const uint32_t b = min(C, tid * per), e = min(C, b + per);
    u64 local = 0;
    for (uint32_t i = b; i < e; ++i) local += totals[i];
    s_part[tid] = local;
    __syncthreads();
    if (tid < 32) {
        u64 acc = 0;
        for (uint32_t i = 0; i < 32; ++i) acc += s_part[tid * 32 + i];
        u64 incl = acc;
        for (int o = 1; o < 32; o <<= 1) {
            const u64 v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)tid >= o) incl += v;
        }
        u64 run = incl - acc;
        for (uint32_t i = 0; i < 32; ++i) {
            const u64 v = s_part[tid * 32 + i];
            s_part[tid * 32 + i] = run;
            run += v;
        }
        if (tid == 31) offsets[C] = incl;
    }
    __syncthreads();
    u64 run = s_part[tid];
    for (uint32_t i = b; i < e; ++i) {
        offsets[i] = run;
        run += totals[i];
    }
}

__global__ void __launch_bounds__(256) csr_scatter_kernel(const uint32_t *__restrict__ assign, const u64 n,
                                                          const uint32_t R, const uint32_t C, const uint32_t NB,
                                                          const uint32_t *__restrict__ counts,
                                                          const u64 *__restrict__ offsets, uint32_t *__restrict__ ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *cur = reinterpret_cast<uint32_t *>(smem_raw);  // next free slot of every list for this block
    for (uint32_t c = threadIdx.x; c < C; c += 256) cur[c] = (uint32_t)offsets[c] + counts[(u64)c * NB + blockIdx.x];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u64 r0 = (u64)blockIdx.x * R, r1 = min(n, r0 + R);
    for (u64 base = r0; base < r1; base += 256) {
        const u64 r = base + threadIdx.x;
        const bool active = r < r1;
        const uint32_t c = active ? assign[r] : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, c);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t slot = 0;
        for (uint32_t w = 0; w < 8; ++w) {
            if (warp == w && active && lane == leader) slot = atomicAdd(&cur[c], __popc(peers));
            __syncthreads();
        }
        slot = __shfl_sync(0xffffffffu, slot, leader);
        if (active) ids[slot + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)r;
    }
}

// Lloyd bookkeeping (src/ivf/index.rs:417-420): how many assignments changed since the previous iteration
__global__ void __launch_bounds__(256) count_changed_kernel(const uint32_t *__restrict__ prev,
                                                            const uint32_t *__restrict__ next, const u64 n,
                                                            unsigned long long *__restrict__ changed) {
    unsigned long long local = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) local += prev[i] != next[i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(changed, local);
}

// centroid update (src/ivf/index.rs:436-453): per cluster, column-wise serial f32 sums over the member
// rows in ascending row order (the order the reference's `for i in 0..n` visits them), then `/= size`
// when size > 0; an empty cluster becomes the origin.  One CTA per cluster, a thread per
// column: the per-(cluster, column) chain is identical to the reference's.
__global__ void __launch_bounds__(256) centroid_update_kernel(const float *__restrict__ data, const uint32_t dim,
                                                              const uint32_t *__restrict__ member_ids,
                                                              const u64 *__restrict__ member_offsets,
                                                              float *__restrict__ centroids) {
    const uint32_t j = blockIdx.x;
    const u64 b = member_offsets[j], e = member_offsets[j + 1];
    // a thread owns up to four columns (d, d + blockDim, ...: coalesced across the CTA) and walks the members once for all of
    // them: the additions form one serial chain per (cluster, column), the loads do not depend on it -- 8 members x 4
    // columns in flight (the kernel is pure load latency: one cluster's rows are scattered over the sample)
    for (uint32_t d0 = threadIdx.x; d0 < dim; d0 += 4u * blockDim.x) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        bool on[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) on[c] = d0 + (uint32_t)c * blockDim.x < dim;
        u64 m = b;
        for (; m + 8 <= e; m += 8) {
            float v[4][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float *row = data + (u64)__ldg(member_ids + m + i) * dim + d0;
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c][i] = on[c] ? __ldg(row + (size_t)c * blockDim.x) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c] = __fadd_rn(acc[c], v[c][i]);
        }
        for (; m < e; ++m) {
            const float *row = data + (u64)__ldg(member_ids + m) * dim + d0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (on[c]) acc[c] = __fadd_rn(acc[c], __ldg(row + (size_t)c * blockDim.x));
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (on[c]) {
                float a = acc[c];
                if (e > b) a = __fdiv_rn(a, (float)(e - b));
                centroids[(u64)j * dim + d0 + (size_t)c * blockDim.x] = a;
            }
    }
}

}  // namespace pqv
