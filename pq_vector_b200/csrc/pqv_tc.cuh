// pqv_tc.cuh -- tcgen05 / TMEM / TMA path for the k-means assignment sweeps (sm_100a only).
//
// Reference loops replaced: nearest_centroid + the final assignment  src/ivf/index.rs:244-257, 189-206
//                           the Lloyd assignment step                 src/ivf/index.rs:395-430
// Both are  out[i] = argmin_j squared_l2_distance(row_i, centroid_j)  with a strict '<' scan (lowest index wins
// ties, +inf/NaN never win, default 0).  The distance itself is never returned, only the u32 argmin, so the
// N x C x dim contraction can run on the tensor cores as a FILTER as long as the emitted index is provably the
// reference's:
//
//   s_j := |c_j|^2 - 2 x.(c_j - mu)          differs from |x - c_j|^2 by a per-row constant (|x|^2 - 2 x.mu)
//   ŝ_j := cn_j - 2 * tf32_mma(x, B'_j)      B'_j = tf32_rn(c_j - mu) precomputed, x truncated by the tensor core
//   |ŝ_j - s_j| <= E(row) = |x| * W + Z      W, Z from the error model below (Cauchy-Schwarz on the dropped bits)
//   d_ref_j = d_true_j (1 + theta), |theta| <= delta = (dim/4 + 12) 2^-24   (serial f32 chain, all terms >= 0)
//
// so the reference argmin lies in  { j : ŝ_j <= min_j ŝ_j + 2E + G },  G >= 2.1 delta max_j d_true_j.  Rows whose
// set has one element are final; the others (and rows with non-finite norms or a full candidate FIFO) are
// re-evaluated with the exact serial-order f32 chain (pair_exact_kernel / the sliced kmeans_assign_kernel<.., GATHER>).
//
// Kernel shape (one persistent CTA per SM, 192 threads):
//   warp 0   : TMA producer      rows tile 128 x 32 f32 + centroid tile 256 x 32 f32 per stage (SWIZZLE_128B), 4 stages
//   warp 1   : MMA issuer        tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8, 4 per stage; owns TMEM alloc
//   warps 2-5: epilogue          tcgen05.ld 32x32b.x32 of the f32 accumulator (2 x 256 TMEM columns, double buffered),
//                                running min + candidate FIFO per row, one row per thread
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqv_kernels.cuh"

namespace pqv {
namespace tc {

constexpr int BM = 128;      // rows per tile (UMMA M)
constexpr int BN = 256;      // centroids per tile (UMMA N)
constexpr int BK = 32;       // f32 columns per stage = 128 B = one SWIZZLE_128B atom row
constexpr int UMMA_K = 8;    // tf32: 32 bytes of K per instruction
constexpr int STAGES = 4;
constexpr uint32_t A_BYTES = BM * BK * 4;
constexpr uint32_t B_BYTES = BN * BK * 4;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t TMEM_COLS = 512;  // two accumulator stages of BN f32 columns
constexpr int THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr uint32_t BAR_BYTES = 8 * (2 * STAGES + 4) + 16;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int FIFO = 8;      // candidate slots per row (unordered; a slot is free once its score left the window)

constexpr uint64_t HINT_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): c_format F32 [4,6)=1, a/b_format TF32 [7,10)/[10,13)=2,
// a/b major K [15],[16]=0, n_dim [17,23)=N>>3, m_dim [24,29)=M>>4
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int32_t c0, int32_t c1, uint32_t bar,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive f32 columns: thread t of the warp gets lane (taddr.lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout) for a K-major SWIZZLE_128B tile whose rows are
// 128 B: start address >> 4 in [0,14), LBO [16,30) unused for this layout (0), SBO [32,46) = 8 rows * 128 B = 1024 B >> 4,
// version [46,48) = 1 (sm_100), base offset 0 (tiles are 1 KB aligned), layout type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------
// preparation kernels (tiny): column mean of the centroid table, centred + tf32-rounded table, norms, bounds
// ------------------------------------------------------------------------------------------------
// bounds[0] = max_j 2 (eps_mma bn_j + rn_j)   bounds[1] = max_j |c_j|^2   bounds[2] = max_j bn_j   bounds[3] = |mu|^2
// (f32 bit patterns, >= 0; zeroed by the caller before centroid_mean_kernel)
__global__ void centroid_mean_kernel(const float *__restrict__ cent, uint32_t C, uint32_t dim, float *__restrict__ mu,
                                     uint32_t *__restrict__ bounds) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    float s = 0.f;
    for (uint32_t j = 0; j < C; ++j) s += cent[(size_t)j * dim + col];
    const float m = s / (float)C;  // any vector is valid here; the mean just keeps |c - mu| small
    mu[col] = m;
    atomicAdd(reinterpret_cast<float *>(&bounds[3]), m * m);  // |mu|^2 (order-dependent rounding: only feeds an inflated bound)
}

__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(128) centroid_prep_kernel(const float *__restrict__ cent, uint32_t C, uint32_t dim,
                                                            const float *__restrict__ mu, float *__restrict__ Bp,
                                                            float *__restrict__ cn, float *__restrict__ wv, uint32_t cn_len,
                                                            uint32_t *__restrict__ bounds) {
    const uint32_t j = blockIdx.x;
    __shared__ double red[3][4];
    if (j >= C) {  // padding entries of cn: +inf never passes "ŝ <= thr"
        if (threadIdx.x == 0 && j < cn_len) {
            cn[j] = __int_as_float(0x7f800000);
            wv[j] = 0.f;
        }
        return;
    }
    double r2 = 0.0, b2 = 0.0, c2 = 0.0;
    for (uint32_t col = threadIdx.x; col < dim; col += blockDim.x) {
        const float c = cent[(size_t)j * dim + col];
        const double cp = (double)c - (double)mu[col];
        const float b = tf32_rn((float)cp);
        Bp[(size_t)j * dim + col] = b;
        const double r = cp - (double)b;
        r2 += r * r;
        b2 += (double)b * (double)b;
        c2 += (double)c * (double)c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        b2 += __shfl_xor_sync(0xffffffffu, b2, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    const uint32_t w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[0][w] = r2;
        red[1][w] = b2;
        red[2][w] = c2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        r2 = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        b2 = red[1][0] + red[1][1] + red[1][2] + red[1][3];
        c2 = red[2][0] + red[2][1] + red[2][2] + red[2][3];
        const double up = 1.0 + 1e-6;
        const double bn = sqrt(b2) * up, rn = sqrt(r2) * up + 1e-300;
        // tensor-core error model per unit |x| |B'_j|: operand A truncated to tf32 (2^-10), B' exact in tf32, products exact in
        // f32, fp32 accumulation over dim/8 instructions (2^-19 each, conservative for a non-IEEE adder tree)
        const double eps = ldexp(1.0, -10) + ((double)(dim / 8 + 4)) * ldexp(1.0, -19);
        const float wj = (float)(2.0 * (eps * bn + rn) * up);
        const float cnj = (float)c2;
        cn[j] = cnj;
        wv[j] = wj;  // |ŝ_j - s_j| <= |x| wv[j] + rounding slack
        atomicMax(&bounds[0], __float_as_uint(wj));   // non-negative floats (and NaN/inf above them) order as unsigned
        atomicMax(&bounds[1], __float_as_uint(cnj));
        atomicMax(&bounds[2], __float_as_uint((float)bn));
    }
}

// per row: (|x|^2, x.mu) in f32 (any order: they only feed the error bounds, inflated by the consumer)
__global__ void __launch_bounds__(256) row_stats_kernel(const float *__restrict__ rows, u64 n, uint32_t dim,
                                                        const float *__restrict__ mu, float2 *__restrict__ stats) {
    const uint32_t lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const uint32_t n4 = dim >> 2;
    const float4 *m4 = reinterpret_cast<const float4 *>(mu);
    for (u64 r = warp; r < n; r += nwarps) {
        const float4 *p = reinterpret_cast<const float4 *>(rows + r * dim);
        float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
        uint32_t c = lane;
        for (; c + 32 < n4; c += 64) {
            const float4 a = ld_stream_v4(reinterpret_cast<const float *>(p + c));
            const float4 b = ld_stream_v4(reinterpret_cast<const float *>(p + c + 32));
            const float4 ma = __ldg(m4 + c), mb = __ldg(m4 + c + 32);
            s0 = __fmaf_rn(a.x, a.x, s0); s0 = __fmaf_rn(a.y, a.y, s0); s0 = __fmaf_rn(a.z, a.z, s0); s0 = __fmaf_rn(a.w, a.w, s0);
            s1 = __fmaf_rn(b.x, b.x, s1); s1 = __fmaf_rn(b.y, b.y, s1); s1 = __fmaf_rn(b.z, b.z, s1); s1 = __fmaf_rn(b.w, b.w, s1);
            t0 = __fmaf_rn(a.x, ma.x, t0); t0 = __fmaf_rn(a.y, ma.y, t0); t0 = __fmaf_rn(a.z, ma.z, t0); t0 = __fmaf_rn(a.w, ma.w, t0);
            t1 = __fmaf_rn(b.x, mb.x, t1); t1 = __fmaf_rn(b.y, mb.y, t1); t1 = __fmaf_rn(b.z, mb.z, t1); t1 = __fmaf_rn(b.w, mb.w, t1);
        }
        for (; c < n4; c += 32) {
            const float4 a = ld_stream_v4(reinterpret_cast<const float *>(p + c));
            const float4 ma = __ldg(m4 + c);
            s0 = __fmaf_rn(a.x, a.x, s0); s0 = __fmaf_rn(a.y, a.y, s0); s0 = __fmaf_rn(a.z, a.z, s0); s0 = __fmaf_rn(a.w, a.w, s0);
            t0 = __fmaf_rn(a.x, ma.x, t0); t0 = __fmaf_rn(a.y, ma.y, t0); t0 = __fmaf_rn(a.z, ma.z, t0); t0 = __fmaf_rn(a.w, ma.w, t0);
        }
        float s = s0 + s1, t = t0 + t1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            t += __shfl_xor_sync(0xffffffffu, t, o);
        }
        if (lane == 0) stats[r] = make_float2(s, t);
    }
}

// ------------------------------------------------------------------------------------------------
// the tensor-core filter
// ------------------------------------------------------------------------------------------------
struct AssignTcParams {
    const float2 *stats;    // [n] (|x|^2, x.mu) per row (row_stats_kernel)
    const float *cn;        // [num_nb * BN] centroid squared norms, +inf padded
    const float *wv;        // [num_nb * BN] per-centroid error weight w_j (0 padded): |ŝ_j - s_j| <= |x| w_j + slack
    const uint32_t *bounds; // [4] see centroid_prep_kernel
    uint32_t *assign;       // [n] out: final for rows the filter decides
    uint32_t *counts;       // [0] ambiguous rows, [1] overflow rows, [2] (row, candidate) pairs; zeroed by the caller
    uint32_t *amb_rows;     // [n]
    uint2 *pairs;           // [pair_cap] (row, candidate): one exact distance each (pair_exact_kernel)
    u64 *best;              // [n] per ambiguous row: min over its pairs of bits(distance) << 32 | candidate
    uint32_t *ovf_rows;     // [n] rows for the full exact scan
    u64 n;
    uint32_t pair_cap;
    uint32_t dim, C;
    uint32_t num_mb, num_nb, num_kb;
};

__global__ void __launch_bounds__(THREADS, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AssignTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barrier i at bar0 + 8 i: full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base-address slot
#define BAR_FULL(s) (bar0 + 8u * (uint32_t)(s))
#define BAR_EMPTY(s) (bar0 + 8u * (uint32_t)(STAGES + (s)))
#define BAR_TFULL(a) (bar0 + 8u * (uint32_t)(2 * STAGES + (a)))
#define BAR_TEMPTY(a) (bar0 + 8u * (uint32_t)(2 * STAGES + 2 + (a)))
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(BAR_FULL(s), 1);
            mbar_init(BAR_EMPTY(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(BAR_TFULL(a), 1);
            mbar_init(BAR_TEMPTY(a), 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t mb = blockIdx.x; mb < p.num_mb; mb += gridDim.x)
                for (uint32_t nb = 0; nb < p.num_nb; ++nb)
                    for (uint32_t kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                        mbar_wait(BAR_EMPTY(s), ph ^ 1u);
                        mbar_expect_tx(BAR_FULL(s), STAGE_BYTES);
                        const uint32_t sa = base + s * STAGE_BYTES;
                        tma_load_2d(sa, &tmA, (int32_t)(kb * BK), (int32_t)(mb * BM), BAR_FULL(s), HINT_EVICT_NORMAL);
                        tma_load_2d(sa + A_BYTES, &tmB, (int32_t)(kb * BK), (int32_t)(nb * BN), BAR_FULL(s), HINT_EVICT_LAST);
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0, tile = 0;
            for (uint32_t mb = blockIdx.x; mb < p.num_mb; mb += gridDim.x)
                for (uint32_t nb = 0; nb < p.num_nb; ++nb, ++tile) {
                    const uint32_t as = tile & 1u, aph = (tile >> 1) & 1u;
                    mbar_wait(BAR_TEMPTY(as), aph ^ 1u);  // epilogue has drained this accumulator stage
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (uint32_t kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                        mbar_wait(BAR_FULL(s), ph);
                        tc_fence_after();
                        const uint32_t sa = base + s * STAGE_BYTES;
                        const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < BK / UMMA_K; ++k)  // +32 B of K inside the swizzle atom = +2 in the >>4 address field
                            umma_tf32(d_tmem, adesc + 2u * k, bdesc + 2u * k, IDESC_TF32, (kb | k) != 0u);
                        umma_commit(BAR_EMPTY(s));  // frees the smem stage once these MMAs have read it
                    }
                    umma_commit(BAR_TFULL(as));  // accumulator complete
                }
        }
        __syncwarp();
    } else {
        // ===== epilogue: one row per thread; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
        const uint32_t q = warp & 3u;
        const uint32_t row_in_tile = q * 32u + lane;
        const float wmax = __uint_as_float(p.bounds[0]), cnmax = __uint_as_float(p.bounds[1]);
        const float bnmax = __uint_as_float(p.bounds[2]);
        const float mun = sqrtf(__uint_as_float(p.bounds[3])) * 1.000001f;
        const float delta = (float)(p.dim / 4 + 12) * 5.9604645e-08f;   // 2^-24: reference's serial f32 chain, terms >= 0
        const float gamma = (float)(p.dim + 32) * 5.9604645e-08f;       // row_stats_kernel's f32 sums
        const float c2 = 2.2f * delta;
        const bool table_ok = (cnmax < 1e30f) && (wmax < 1e30f) && (mun < 1e30f);
        const float inf = __int_as_float(0x7f800000);
        uint32_t tile = 0;
        for (uint32_t mb = blockIdx.x; mb < p.num_mb; mb += gridDim.x) {
            const u64 row = (u64)mb * BM + row_in_tile;
            const bool valid = row < p.n;
            const float2 st = valid ? p.stats[row] : make_float2(0.f, 0.f);
            const float x2 = st.x * (1.f + gamma) + 1e-37f;
            const float a = sqrtf(x2) * 1.000001f;
            // per centroid:  L_j = ŝ_j - a w_j <= s_j <= U_j = ŝ_j + a w_j  (up to the rounding slack folded into T2).  With
            // m = min_j U_j the reference argmin satisfies  L_j <= thr(m) = m + T2 + 2.2 delta max(m + K2, 0):  m + K2 bounds
            // |x - c_j'|^2 for the j' attaining m, because s_j and d_j differ by exactly |x|^2 - 2 x.mu <= K2.
            const float na = -a;
            const float mag = cnmax + 2.f * a * bnmax + a * wmax;                        // bound on |ŝ|, |L|, |U|
            const float kmag = x2 + 2.f * fabsf(st.y);
            const float T2 = mag * 9.5367432e-07f + 1e-37f;                              // 2^-20: cn/fma/thr roundings
            const float K2 = (x2 - 2.f * st.y) + 2.02f * gamma * a * mun + (kmag + mag) * 9.5367432e-07f;
            float m = inf;
            float fs[FIFO];
            uint32_t fi[FIFO];
#pragma unroll
            for (int e = 0; e < FIFO; ++e) {
                fs[e] = inf;
                fi[e] = NONE;
            }
            bool ovf = false;
            for (uint32_t nb = 0; nb < p.num_nb; ++nb, ++tile) {
                const uint32_t as = tile & 1u, aph = (tile >> 1) & 1u;
                mbar_wait(BAR_TFULL(as), aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * BN;
                const float4 *cn4 = reinterpret_cast<const float4 *>(p.cn + (size_t)nb * BN);
                const float4 *wv4 = reinterpret_cast<const float4 *>(p.wv + (size_t)nb * BN);
#pragma unroll 1
                for (uint32_t ch = 0; ch < BN / 32; ++ch) {
                    float v[32], w[32];
                    tmem_ld32(taddr + ch * 32u, v);
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 c = __ldg(cn4 + ch * 8 + i4);
                        const float4 ww = __ldg(wv4 + ch * 8 + i4);
                        v[4 * i4 + 0] = __fmaf_rn(-2.f, v[4 * i4 + 0], c.x);
                        v[4 * i4 + 1] = __fmaf_rn(-2.f, v[4 * i4 + 1], c.y);
                        v[4 * i4 + 2] = __fmaf_rn(-2.f, v[4 * i4 + 2], c.z);
                        v[4 * i4 + 3] = __fmaf_rn(-2.f, v[4 * i4 + 3], c.w);
                        w[4 * i4 + 0] = ww.x;
                        w[4 * i4 + 1] = ww.y;
                        w[4 * i4 + 2] = ww.z;
                        w[4 * i4 + 3] = ww.w;
                    }
                    float cm = inf;
#pragma unroll
                    for (int i = 0; i < 32; ++i) cm = fminf(cm, __fmaf_rn(a, w[i], v[i]));   // min U_j
                    m = fminf(m, cm);
                    const float thr = m + T2 + c2 * fmaxf(m + K2, 0.f);
                    bool any = false;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = __fmaf_rn(na, w[i], v[i]);                                    // L_j
                        any |= (v[i] <= thr);
                    }
                    if (any) {
                        const uint32_t j0 = nb * BN + ch * 32u;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (v[i] <= thr) {
                                bool placed = false;  // take any slot whose lower bound has left the window
#pragma unroll
                                for (int e = 0; e < FIFO; ++e) {
                                    const bool take = !placed && !(fs[e] <= thr);
                                    fs[e] = take ? v[i] : fs[e];
                                    fi[e] = take ? j0 + i : fi[e];
                                    placed |= take;
                                }
                                ovf |= !placed;
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(BAR_TEMPTY(as));
            }
            // ---- finalize the row
            const float thr = m + T2 + c2 * fmaxf(m + K2, 0.f);
            uint32_t nv = 0;
            uint32_t only = NONE;
#pragma unroll
            for (int e = 0; e < FIFO; ++e) {
                const bool ok = fs[e] <= thr;
                fi[e] = ok ? fi[e] : NONE;
                only = ok ? fi[e] : only;
                nv += ok ? 1u : 0u;
            }
            const bool finite = table_ok && (x2 < 1e30f) && (m < 1e30f) && (m > -1e30f) && (kmag < 1e30f);
            bool is_ovf = valid && (ovf || !finite || nv == 0u);
            bool is_amb = valid && !is_ovf && nv > 1u;
            if (valid && !is_ovf && nv == 1u) p.assign[row] = only;
            // (row, candidate) pairs of the ambiguous rows: warp-aggregated reservation
            const uint32_t cnt = is_amb ? nv : 0u;
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t pbase = 0;
            if (total) {
                if (lane == 0) pbase = atomicAdd(&p.counts[2], total);
                pbase = __shfl_sync(0xffffffffu, pbase, 0);
            }
            const bool fits = (u64)pbase + total <= (u64)p.pair_cap;
            if (is_amb) {
                uint32_t slot = pbase + incl - cnt;
#pragma unroll
                for (int e = 0; e < FIFO; ++e)
                    if (fi[e] != NONE) {
                        // a reservation that does not fit leaves sentinels behind and the row takes the full scan
                        if (slot < p.pair_cap) p.pairs[slot] = fits ? make_uint2((uint32_t)row, fi[e]) : make_uint2(NONE, NONE);
                        ++slot;
                    }
                if (!fits) {
                    is_amb = false;
                    is_ovf = true;
                }
            }
            const uint32_t amb_mask = __ballot_sync(0xffffffffu, is_amb);
            const uint32_t ovf_mask = __ballot_sync(0xffffffffu, is_ovf);
            uint32_t amb_base = 0, ovf_base = 0;
            if (lane == 0) {
                if (amb_mask) amb_base = atomicAdd(&p.counts[0], (uint32_t)__popc(amb_mask));
                if (ovf_mask) ovf_base = atomicAdd(&p.counts[1], (uint32_t)__popc(ovf_mask));
            }
            amb_base = __shfl_sync(0xffffffffu, amb_base, 0);
            ovf_base = __shfl_sync(0xffffffffu, ovf_base, 0);
            const uint32_t below = (1u << lane) - 1u;
            if (is_amb) {
                p.amb_rows[amb_base + (uint32_t)__popc(amb_mask & below)] = (uint32_t)row;
                p.best[row] = KEY_MAX;
            }
            if (is_ovf) {
                p.ovf_rows[ovf_base + (uint32_t)__popc(ovf_mask & below)] = (uint32_t)row;
                p.best[row] = KEY_MAX;
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
#undef BAR_FULL
#undef BAR_EMPTY
#undef BAR_TFULL
#undef BAR_TEMPTY
}

// ------------------------------------------------------------------------------------------------
// exact re-evaluation in the reference order (src/ivf/index.rs:461-480); dim % 4 == 0 on this path.
// One warp per 32 (row, candidate) pairs, same scheme as group_distance: all lanes load a pair's row and centroid with
// coalesced 128-bit loads and compute the independent chain terms, the terms are transposed through a padded
// shared-memory tile, and lane p runs pair p's serial chain.  The per-row winner is an atomicMin over
// bits(distance) << 32 | candidate: squared distances are >= 0 and finite here, so the u64 order is (distance, index) --
// the reference's strict-'<' ascending scan (index.rs:251).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

constexpr int PAIR_WARPS = 8, PAIR_TS = 36;

__global__ void __launch_bounds__(PAIR_WARPS * 32) pair_exact_kernel(const float *__restrict__ rows, uint32_t dim,
                                                                     const float *__restrict__ cent,
                                                                     const uint32_t *__restrict__ counts, uint32_t pair_cap,
                                                                     const uint2 *__restrict__ pairs, u64 *__restrict__ best) {
    __shared__ __align__(16) float tiles[PAIR_WARPS][32 * PAIR_TS];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *tile = tiles[wib];
    const uint32_t npairs = min(counts[2], pair_cap);
    const uint32_t ngroups = (npairs + 31u) / 32u;
    const uint32_t ncb = (dim + 127u) / 128u;
    for (uint32_t g = blockIdx.x * PAIR_WARPS + wib; g < ngroups; g += gridDim.x * PAIR_WARPS) {
        const uint32_t idx = g * 32u + lane;
        const uint2 pr = idx < npairs ? pairs[idx] : make_uint2(NONE, NONE);
        const bool valid = pr.x != NONE;
        const uint32_t row_l = valid ? pr.x : 0u, cand_l = valid ? pr.y : 0u;
        float sum = 0.f;
        for (uint32_t cb = 0; cb < ncb; ++cb) {
            const uint32_t col = cb * 128u + (lane << 2);
            const bool inb = col < dim;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
                float4 xv[8], cv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t rr = __shfl_sync(0xffffffffu, row_l, r0 + j);
                    const uint32_t cc = __shfl_sync(0xffffffffu, cand_l, r0 + j);
                    xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    cv[j] = xv[j];
                    if (inb) {
                        xv[j] = ldg4(rows + (u64)rr * dim + col);
                        cv[j] = ldg4(cent + (size_t)cc * dim + col);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) tile[(r0 + j) * PAIR_TS + lane] = chunk4(xv[j], cv[j]);
            }
            __syncwarp();
            // columns past dim contribute +0.0 terms: x + (+0.0) == x bit for bit for the non-negative partial sums
            const float4 *tr = reinterpret_cast<const float4 *>(tile + lane * PAIR_TS);
#pragma unroll
            for (int t4 = 0; t4 < 8; ++t4) {
                const float4 t = tr[t4];
                sum = __fadd_rn(sum, t.x);
                sum = __fadd_rn(sum, t.y);
                sum = __fadd_rn(sum, t.z);
                sum = __fadd_rn(sum, t.w);
            }
            __syncwarp();
        }
        if (valid) atomicMin(reinterpret_cast<unsigned long long *>(&best[row_l]), ((u64)__float_as_uint(sum) << 32) | (u64)cand_l);
    }
}

// assign[row] = low word of best[row] for the rows of both lists (ambiguous: pair_exact_kernel; overflow: the sliced exact
// scan).  KEY_MAX means no finite distance won: the reference's default, cluster 0 (index.rs:245).
__global__ void __launch_bounds__(256) best_finalize_kernel(const uint32_t *__restrict__ counts,
                                                            const uint32_t *__restrict__ amb_rows,
                                                            const uint32_t *__restrict__ ovf_rows,
                                                            const u64 *__restrict__ best, uint32_t *__restrict__ assign) {
    const uint32_t na = counts[0], total = na + counts[1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t row = i < na ? amb_rows[i] : ovf_rows[i - na];
        const uint32_t c = (uint32_t)(best[row] & 0xFFFFFFFFull);
        assign[row] = (c == NONE) ? 0u : c;
    }
}

}  // namespace tc
}  // namespace pqv
