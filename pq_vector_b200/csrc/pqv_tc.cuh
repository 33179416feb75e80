// pqv_tc.cuh -- tcgen05 / TMEM / TMA path for the k-means assignment sweeps and the batched top-k (sm_100a only).
//
// Reference loops replaced: nearest_centroid + the final assignment  src/ivf/index.rs:244-257, 189-206
//                           the Lloyd assignment step                 src/ivf/index.rs:395-430
// Both are  out[i] = argmin_j squared_l2_distance(row_i, centroid_j)  with a strict '<' scan (lowest index wins
// ties, +inf/NaN never win, default 0).  The distance itself is never returned, only the u32 argmin, so the
// N x C x dim contraction can run on the tensor cores as a FILTER as long as the emitted index is provably the
// reference's:
//
//   s_j := |c_j|^2 - 2 x.(c_j - mu)          differs from |x - c_j|^2 by a per-row constant (|x|^2 - 2 x.mu)
//   ŝ_j := cn_j - 2 * mma(x^, B^_j)          B^_j = the rounded (c_j - mu), x^ = the operand the tensor core sees
//   |ŝ_j - s_j| <= a w_j                      a >= |x|, w_j = 2 (rn_j + (kappa + eps_acc (1 + kappa)) bn_j)   (pqv_half.cuh)
//   d_ref_j = d_true_j (1 + theta), |theta| <= delta = (dim/4 + 12) 2^-24   (serial f32 chain, all terms >= 0)
//
// so the reference argmin lies in  { j : ŝ_j - a w_j <= min_j (ŝ_j + a w_j) + G },  G >= 2.2 delta max_j d_true_j.  Rows
// whose set has one element are final; the others (and rows with non-finite norms or a full candidate FIFO) are
// re-evaluated with the exact serial-order f32 chain (pair_exact_kernel / the sliced kmeans_assign_kernel<.., GATHER>).
//
// Operand kinds (template parameter KIND of the kernels):
//   KIND_F16   the table's 16-bit shadow (pqv_half.cuh) under tcgen05 kind::f16: 64 columns per 128-byte stage row, half
//              the L2 -> shared-memory fill per flop and twice the MMA rate of tf32, kappa measured (~2^-12.3)
//   KIND_TF32  the f32 rows read directly under kind::tf32 (the tensor core truncates them: kappa = 2^-10); used when no
//              shadow exists (dim % 8 != 0, PQV_TC_KIND=tf32)
//
// Kernel shape (persistent, CTA pairs = clusters of two CTAs on the two SMs of a TPC):
//   warp 0   : TMA producer      own 128-row tile of A + half of the 256-row table tile per stage (SWIZZLE_128B), 6 stages
//   warp 1   : MMA issuer        tcgen05.mma.cta_group::2, M=256 over both SMs, N=256, 32 bytes of K per instruction
//   warps 2+ : epilogue          tcgen05.ld 32x32b.x32 of the f32 accumulator (2 x 256 TMEM columns, double buffered);
//                                SPLIT warps share each TMEM lane quarter and split a tile's columns
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqv_half.cuh"
#include "pqv_kernels.cuh"

namespace pqv {
namespace tc {

enum { KIND_TF32 = 0, KIND_F16 = 1 };
constexpr int BM = 128;      // rows per tile (UMMA M)
constexpr int BN = 256;      // centroids per tile (UMMA N)
constexpr int BK = 32;       // 4-byte slots per stage row = 128 B = one SWIZZLE_128B atom row
__host__ __device__ constexpr int bk_elems(int kind) { return kind == KIND_F16 ? 64 : 32; }  // operand columns per stage
constexpr int UMMA_STEPS = 4;  // MMAs per stage: 32 bytes of K each (8 tf32 / 16 f16 elements)
constexpr int STAGES = 4;
constexpr uint32_t A_BYTES = BM * BK * 4;
constexpr uint32_t B_BYTES = BN * BK * 4;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t TMEM_COLS = 512;  // two accumulator stages of BN f32 columns
// threads per CTA: TMA warp + MMA warp + 4 * SPLIT epilogue warps.  TMEM lane quarter q is only readable by warps with
// warp % 4 == q, so with SPLIT = 2 two warps share every row and each takes half of a tile's columns: a single epilogue warp
// per SM sub-partition is latency-bound (ncu: 0.23 IPC) and the MMAs end up waiting for the accumulator to be drained
constexpr int tc_threads(int split) { return 64 + 128 * split; }
constexpr int EPI_WARP0 = 2;
__host__ __device__ constexpr uint32_t bar_bytes(int stages) { return 8u * (2u * (uint32_t)stages + 4u) + 16u; }
// shared memory behind the barriers: the hand-over area of a split epilogue (Epi::EXCH_BYTES) and EPI_TAB_FLOATS floats of
// per-column constants the epilogue reads for every tile (centroid norms / halved query thresholds), staged once per CTA
constexpr uint32_t EPI_TAB_FLOATS = 4096;
constexpr uint32_t EPI_TAB_BYTES = EPI_TAB_FLOATS * 4;
constexpr size_t smem_bytes_single(int stages, uint32_t exch_bytes) {
    return (size_t)stages * STAGE_BYTES + bar_bytes(stages) + exch_bytes + EPI_TAB_BYTES + 1024;  // +1024: manual 1 KB alignment
}
// CTA-pair variant (tcgen05 cta_group::2): UMMA M = 256 over two SMs, each CTA stages its own 128 rows of A and HALF of
// the B tile (128 table rows), so a stage is 32 KB per CTA instead of 48 KB and six stages fit
constexpr int STAGES2 = 6;
constexpr uint32_t B2_BYTES = (BN / 2) * BK * 4;
constexpr uint32_t STAGE2_BYTES = A_BYTES + B2_BYTES;
constexpr size_t smem_bytes_pair(int stages, uint32_t exch_bytes) {
    return (size_t)stages * STAGE2_BYTES + bar_bytes(stages) + exch_bytes + EPI_TAB_BYTES + 1024;
}
constexpr uint32_t NONE = 0xFFFFFFFFu;

constexpr uint64_t HINT_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): c_format F32 [4,6)=1, a/b_format [7,10)/[10,13)
// (kind::tf32: TF32 = 2; kind::f16: F16 = 0, BF16 = 1), a/b major K [15],[16]=0, n_dim [17,23)=N>>3, m_dim [24,29)=M>>4
__host__ __device__ constexpr uint32_t idesc_for(int kind, int m) {
    return (1u << 4) | (kind == KIND_TF32 ? ((2u << 7) | (2u << 10)) : 0u) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

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
template <int KIND>
__device__ __forceinline__ void umma_single(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// ---- cta_group::2 (CTA pair) forms
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {  // same offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// destination in this CTA's shared memory, completion bytes on an mbarrier that may live in the peer CTA (`bar` is a
// shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *tm, int32_t c0, int32_t c1, uint32_t bar,
                                                 uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives (once every tcgen05.mma issued so far has completed on both SMs) on the barrier at the same offset in every CTA
// of cta_mask
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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

// the same load without the wait: several can be in flight; tmem_ld_fence() makes their registers readable.  The empty
// volatile asm per register pins every use behind the wait (volatile asms keep their order, plain arithmetic does not).
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32], float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        asm volatile("" : "+r"(r[i]));
        v[i] = __uint_as_float(r[i]);
    }
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout) for a K-major SWIZZLE_128B tile whose rows are
// 128 B: start address >> 4 in [0,14), LBO [16,30) unused for this layout (0), SBO [32,46) = 8 rows * 128 B = 1024 B >> 4,
// version [46,48) = 1 (sm_100), base offset 0 (tiles are 1 KB aligned), layout type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------
// preparation kernels (tiny): column mean of the centroid table, centred + rounded table, norms, bounds
// ------------------------------------------------------------------------------------------------
// bounds[] (f32 bit patterns, >= 0; zeroed by the caller before centroid_mean_kernel):
//   [0] max_j w_j   [1] max_j |c_j|^2   [2] max_j bn_j   [3] |mu_d|^2   [4] |mu - mu_d|^2
//   [5] max |c - mu| over the table (centroid_absmax_kernel), replaced by the table's operand scale (a power of two, f32)
// mu = column mean of the centroids (the vector the table is centred on); mu_d = the vector the per-row statistic x.mu_d
// was computed against: the same mu (mu_data == nullptr: row_stats_kernel ran for this sweep) or the table's cached data
// mean (pqv_half.cuh).  The consumer needs x.mu only inside a bound: x.mu = x.mu_d + x.(mu - mu_d), |x.(mu - mu_d)| <= |x| |mu - mu_d|.
// column sums of the table in CMEAN_SLICES row slices (blockIdx.y), added in slice order by centroid_mean_kernel: one thread
// per column walking all C rows cost 46 us per sweep, more than the centroid preparation itself
constexpr uint32_t CMEAN_SLICES = 16;
__global__ void centroid_mean_partial_kernel(const float *__restrict__ cent, uint32_t C, uint32_t dim, float *__restrict__ part) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    const uint32_t per = (C + CMEAN_SLICES - 1) / CMEAN_SLICES;
    const uint32_t b = blockIdx.y * per, e = min(C, b + per);
    float s = 0.f;
    for (uint32_t j = b; j < e; ++j) s += cent[(size_t)j * dim + col];
    part[(size_t)blockIdx.y * dim + col] = s;
}
__global__ void centroid_mean_kernel(const float *__restrict__ part, uint32_t C, uint32_t dim, const float *__restrict__ mu_data,
                                     float *__restrict__ mu, uint32_t *__restrict__ bounds) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    float s = 0.f;
    for (uint32_t i = 0; i < CMEAN_SLICES; ++i) s += part[(size_t)i * dim + col];
    const float m = s / (float)C;  // any vector is valid here; the mean just keeps |c - mu| small
    mu[col] = m;
    const float md = mu_data ? mu_data[col] : m;
    const float dl = m - md;
    // order-dependent rounding of these sums only feeds inflated bounds
    atomicAdd(reinterpret_cast<float *>(&bounds[3]), md * md);
    atomicAdd(reinterpret_cast<float *>(&bounds[4]), dl * dl);
}

// max |c_j[col] - mu[col]| over the table as f32 bits (atomicMax into *out): picks the fp16 operand scale of the table
__global__ void __launch_bounds__(256) centroid_absmax_kernel(const float *__restrict__ cent, uint32_t C, uint32_t dim,
                                                              const float *__restrict__ mu, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    const u64 count = (u64)C * dim;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(cent[i] - mu[i % dim]) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// accumulation error of the tensor core per unit |x^| |B^|: products of tf32 / f16 operands are exact in f32, the f32
// accumulation over the K dimension is charged 2^-19 per 8 products (conservative for a non-IEEE adder tree)
__device__ __forceinline__ double eps_acc(uint32_t dim) { return ((double)(dim / 8 + 4)) * ldexp(1.0, -19); }
// operand-A residual per unit |x|: the table-wide maximum measured by the shadow pass (inflated by the f32 summation error
// of its two sums), or 2^-10 for f32 rows truncated to tf32 by the tensor core
__device__ __forceinline__ double kappa_of(const half16::Globals *g, uint32_t dim) {
    if (!g) return ldexp(1.0, -10);
    return (double)__uint_as_float(g->kappa_bits) * (1.0 + 2.0 * (double)(dim + 32) * ldexp(1.0, -24) + 1e-6) + 1e-30;
}

// Column order of the table inside the filter.  A row's score s_j = |c_j - mu|^2 - 2 (x - mu).(c_j - mu) + const grows with
// |c_j - mu|^2 (clusters of few members sit far from the mean: the term ranges over 1 .. 70 for config C3 while the projection
// term has a spread of ~1), and so does the error weight w_j.  Sorting the columns by |c_j - mu|^2 puts the likely winners into
// the first chunks -- the running bound m is final after a few tiles and the remaining chunks fail the chunk pre-test for
// all 32 rows of a warp -- and makes w homogeneous inside a 4-column group / 32-column chunk.  perm[pos] = original index; the
// epilogue maps back before anything leaves the kernel, so tie-breaking (lowest ORIGINAL index, src/ivf/index.rs:251) is untouched.
__global__ void __launch_bounds__(128) centroid_spread_kernel(const float *__restrict__ cent, uint32_t C, uint32_t dim,
                                                              const float *__restrict__ mu, u64 *__restrict__ keys) {
    const uint32_t j = blockIdx.x;
    __shared__ float red[4];
    float b2 = 0.f;
    for (uint32_t col = threadIdx.x; col < dim; col += blockDim.x) {
        const float d = cent[(size_t)j * dim + col] - mu[col];
        b2 += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = b2;
    __syncthreads();
    if (threadIdx.x == 0) {
        b2 = (red[0] + red[1]) + (red[2] + red[3]);
        // non-negative f32 bits order as unsigned; NaN / inf sort behind every finite value (any order is valid)
        keys[j] = ((u64)(__float_as_uint(b2) & 0x7FFFFFFFu) << 32) | j;
    }
}
constexpr uint32_t ORDER_MAX_C = 8192;   // one CTA sorts the keys in shared memory (64 KB); wider tables keep their order
__global__ void __launch_bounds__(1024) centroid_order_kernel(const u64 *__restrict__ keys, uint32_t C, uint32_t cp2,
                                                              uint32_t *__restrict__ perm) {
    extern __shared__ u64 okeys[];
    for (uint32_t i = threadIdx.x; i < cp2; i += blockDim.x) okeys[i] = i < C ? keys[i] : ~0ull;
    __syncthreads();
    bitonic_sort_smem(okeys, cp2, threadIdx.x, blockDim.x);
    for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) perm[i] = (uint32_t)okeys[i];
}

// B^_j = rounded (c_j - mu) in the operand type of the kernel (tf32-rounded f32, or fp16 with the sub-normal range flushed),
// cn_j = |c_j|^2, w_j, and wc[j / 32] = max of w over the 32-column chunk of j
template <int KIND>
__global__ void __launch_bounds__(128) centroid_prep_kernel(const float *__restrict__ cent, uint32_t C, uint32_t dim,
                                                            const float *__restrict__ mu, void *__restrict__ Bp_,
                                                            float *__restrict__ cn, float *__restrict__ wv,
                                                            uint32_t *__restrict__ wc, uint32_t cn_len,
                                                            uint32_t *__restrict__ bounds, const half16::Globals *__restrict__ hg,
                                                            const uint32_t *__restrict__ perm) {
    const uint32_t j = blockIdx.x;   // position inside the filter; row `src` of the caller's table
    __shared__ double red[3][4];
    const float sb = KIND == KIND_F16 ? __uint_as_float(bounds[5]) : 1.f;  // power of two (scale_from_absmax_kernel)
    if (j >= C) {  // padding entries of cn: +inf never passes "ŝ <= thr"
        if (threadIdx.x == 0 && j < cn_len) {
            cn[j] = __int_as_float(0x7f800000);
            wv[j] = 0.f;
        }
        return;
    }
    double r2 = 0.0, b2 = 0.0, c2 = 0.0;
    const uint32_t src = perm ? perm[j] : j;
    for (uint32_t col = threadIdx.x; col < dim; col += blockDim.x) {
        const float c = cent[(size_t)src * dim + col];
        const double cp = (double)c - (double)mu[col];
        float b;
        if (KIND == KIND_F16) {
            const __half h = half16::to_half_flushed((float)cp * sb);
            reinterpret_cast<__half *>(Bp_)[(size_t)j * dim + col] = h;
            b = __half2float(h) / sb;  // exact: sb is a power of two
        } else {
            b = tf32_rn((float)cp);
            reinterpret_cast<float *>(Bp_)[(size_t)j * dim + col] = b;
        }
        const double r = cp - (double)b;  // includes the f32 rounding of cp and an overflow to inf (-> non-finite bounds)
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
        const double kap = kappa_of(hg, dim);
        const double eps = kap + eps_acc(dim) * (1.0 + kap);
        const float wj = (float)(2.0 * (eps * bn + rn) * up);
        const float cnj = (float)c2;
        cn[j] = cnj;
        wv[j] = wj;  // |ŝ_j - s_j| <= |x| wv[j] + rounding slack
        atomicMax(&wc[j >> 5], __float_as_uint(wj));  // non-negative floats (and NaN/inf above them) order as unsigned
        atomicMax(&bounds[0], __float_as_uint(wj));
        atomicMax(&bounds[1], __float_as_uint(cnj));
        atomicMax(&bounds[2], __float_as_uint((float)bn));
    }
}

// per row: (|x|^2, x.mu) in f32 (any order: they only feed the error bounds, inflated by the consumer).  Only the
// KIND_TF32 path runs this per sweep; with a 16-bit shadow the statistics are part of the shadow (pqv_half.cuh).
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
    const float2 *stats;    // [n] (|x|^2, x.mu_d) per row (shadow pass or row_stats_kernel)
    const float *cn;        // [num_nb * BN] centroid squared norms, +inf padded
    const float *wv;        // [num_nb * BN] per-centroid error weight w_j (0 padded): |ŝ_j - s_j| <= |x| w_j + slack
    const float *wc;        // [num_nb * BN / 32] per 32-column chunk: max_j w_j (chunk pre-test)
    const uint32_t *bounds; // [8] see centroid_mean_kernel
    const uint32_t *perm;   // [C] column of the filter -> row of the caller's table (centroid_order_kernel), nullptr = identity
    const float *scale_a;   // operand scale of the rows (half16::Globals::scale), nullptr = 1
    uint32_t *assign;       // [n] out: final for rows the filter decides
    uint32_t *counts;       // [0] ambiguous rows, [1] overflow rows, [2] (row, candidate) pairs, [4..7] overflow reasons; zeroed by the caller
    uint32_t *amb_rows;     // [n]
    uint2 *pairs;           // [pair_cap] (row, candidate): one exact distance each (pair_exact_kernel)
    u64 *best;              // [n] per ambiguous row: min over its pairs of bits(distance) << 32 | candidate
    uint32_t *ovf_rows;     // [n] rows for the full exact scan
    u64 n;
    uint32_t pair_cap;
    uint32_t dim, C;
};

// barrier i at bar0 + 8 i: full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base-address slot and a u32
// scratch counter for the epilogue
#define BAR_FULL(s) (bar0 + 8u * (uint32_t)(s))
#define BAR_EMPTY(s) (bar0 + 8u * (uint32_t)(STAGES + (s)))
#define BAR_TFULL(a) (bar0 + 8u * (uint32_t)(2 * STAGES + (a)))
#define BAR_TEMPTY(a) (bar0 + 8u * (uint32_t)(2 * STAGES + 2 + (a)))

struct GemmShape {
    uint32_t num_mb, num_nb, num_kb;  // 128-row tiles of A, 256-row tiles of B, k blocks of bk_elems(KIND) columns
};

// what an epilogue warp needs: barrier base, TMEM base, its TMEM lane quarter, the CTA's scratch counter and the CTA's
// walk over the 128-row tiles (mb = mb0, mb0 + mb_stride, ... < mb_end; in the CTA-pair kernel both CTAs of a pair take
// the same number of steps, the odd CTA's last tile may lie past the table: its rows are simply not valid)
struct EpiCtx {
    uint32_t tfull0;     // shared::cta address of this CTA's tfull[0] barrier (tfull[1] follows at +8)
    uint32_t tempty0;    // address of tempty[0] of the CTA that issues the MMAs: shared::cta (single CTA) or shared::cluster
    uint32_t remote;     // 1: tempty0 is a shared::cluster address (CTA-pair kernel)
    uint32_t tmem_base, q, lane;
    uint32_t *counter;   // shared memory, zeroed before the roles start
    uint32_t mb0, mb_stride, mb_end;
    uint32_t h;          // column slice this warp drains (0 .. SPLIT-1)
    uint32_t *exch;      // shared memory, Epi::EXCH_BYTES: hand-over between the warps of a row (split epilogue)
    float *tab;          // shared memory, EPI_TAB_FLOATS floats: per-column constants staged by the epilogue warps
    uint32_t et;         // index of this thread among the epilogue threads (0 .. 128 * SPLIT - 1)
    // wait for accumulator `tile` of this CTA; returns the TMEM address of this warp's 32 lanes x BN columns
    __device__ __forceinline__ uint32_t acquire(uint32_t tile) const {
        const uint32_t as = tile & 1u, aph = (tile >> 1) & 1u;
        mbar_wait(tfull0 + 8u * as, aph);
        tc_fence_after();
        return tmem_base + ((q * 32u) << 16) + as * BN;
    }
    // One arrival per WARP: every thread orders its tcgen05.ld before the warp-level sync, lane 0 then arrives for all 32.
    // (Per-thread `mbarrier.arrive.release.cluster` cost an ERRBAR + membar sequence in every epilogue thread: 15 % of the
    // epilogue warps' samples in ncu, and it delayed the hand-back of the accumulator stage to the MMA warp.)
    __device__ __forceinline__ void release(uint32_t tile) const {
        tc_fence_before();
        __syncwarp();
        if (lane == 0u) {
            const uint32_t bar = tempty0 + 8u * (tile & 1u);
            if (remote) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
            else mbar_arrive(bar);
        }
    }
    // all epilogue threads of the CTA (named barrier 9)
    __device__ __forceinline__ void sync_epilogue(uint32_t nthreads) const {
        asm volatile("bar.sync 9, %0;" ::"r"(nthreads) : "memory");
    }
};

// D[128 x 256 per tile] = A[rows x dim] . B[table x dim]^T on the tensor cores, persistent over the row tiles of A;
// Epi::run consumes every accumulator tile straight from TMEM (nothing of D is ever written to memory as a matrix).
template <class Epi, int KIND>
__global__ void __launch_bounds__(tc_threads(Epi::SPLIT), 1)
tc_rows_x_table_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape g,
                       const typename Epi::Params p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int STAGES = Epi::STAGES_SINGLE;   // shared-memory ring depth of this epilogue's instance (shadows the default)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
    uint32_t *const counter = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot + 8u - raw));
    constexpr int BKE = bk_elems(KIND);
    constexpr uint32_t IDESC = idesc_for(KIND, BM);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(BAR_FULL(s), 1);
            mbar_init(BAR_EMPTY(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(BAR_TFULL(a), 1);
            mbar_init(BAR_TEMPTY(a), 4 * Epi::SPLIT);   // one arrival per epilogue warp
        }
        *counter = 0u;
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
            for (uint32_t mb = blockIdx.x; mb < g.num_mb; mb += gridDim.x)
                for (uint32_t nb = 0; nb < g.num_nb; ++nb)
                    for (uint32_t kb = 0; kb < g.num_kb; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                        mbar_wait(BAR_EMPTY(s), ph ^ 1u);
                        mbar_expect_tx(BAR_FULL(s), STAGE_BYTES);
                        const uint32_t sa = base + s * STAGE_BYTES;
                        tma_load_2d(sa, &tmA, (int32_t)(kb * BKE), (int32_t)(mb * BM), BAR_FULL(s), HINT_EVICT_NORMAL);
                        tma_load_2d(sa + A_BYTES, &tmB, (int32_t)(kb * BKE), (int32_t)(nb * BN), BAR_FULL(s), HINT_EVICT_LAST);
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t it = 0, tile = 0;
            for (uint32_t mb = blockIdx.x; mb < g.num_mb; mb += gridDim.x)
                for (uint32_t nb = 0; nb < g.num_nb; ++nb, ++tile) {
                    const uint32_t as = tile & 1u, aph = (tile >> 1) & 1u;
                    mbar_wait(BAR_TEMPTY(as), aph ^ 1u);  // epilogue has drained this accumulator stage
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (uint32_t kb = 0; kb < g.num_kb; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                        mbar_wait(BAR_FULL(s), ph);
                        tc_fence_after();
                        const uint32_t sa = base + s * STAGE_BYTES;
                        const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < UMMA_STEPS; ++k)  // +32 B of K inside the swizzle atom = +2 in the >>4 address field
                            umma_single<KIND>(d_tmem, adesc + 2u * k, bdesc + 2u * k, IDESC, (kb | k) != 0u);
                        umma_commit(BAR_EMPTY(s));  // frees the smem stage once these MMAs have read it
                    }
                    umma_commit(BAR_TFULL(as));  // accumulator complete
                }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps: one row per thread; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
        uint8_t *const after_bars = smem_raw + (bar0 + bar_bytes(STAGES) - raw);
        Epi::run(EpiCtx{BAR_TFULL(0), BAR_TEMPTY(0), 0u, tmem_base, warp & 3u, lane, counter, blockIdx.x, gridDim.x, g.num_mb,
                        (warp - 2u) >> 2, reinterpret_cast<uint32_t *>(after_bars),
                        reinterpret_cast<float *>(after_bars + Epi::EXCH_BYTES), threadIdx.x - 64u},
                 g, p);
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) Epi::finish(counter, p);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// The same contraction on CTA PAIRS (tcgen05 cta_group::2, clusters of two CTAs on the two SMs of a TPC): one UMMA covers
// 256 rows x 256 table entries, CTA r of the pair owns rows [128 r, 128 r + 128) of it (its own TMEM lanes) and stages
// only ITS half of the B tile (table entries [128 r, 128 r + 128)); the tensor cores of both SMs read both halves.  Per
// 128 x 256 x 32 block an SM now ingests 32 KB instead of 48 KB -- the L2 -> SM fill, not the tensor pipe, bounds this
// kernel (profiles/r01_assign_tc_full.md).  Barrier protocol:
//   full[s]    leader only, count 1: the leader's producer arms it with the bytes of BOTH CTAs; both producers' TMA loads
//              complete on it (cp.async.bulk.tensor ... cta_group::2, barrier address mapped into the leader)
//   empty[s]   in each CTA, count 1: tcgen05.commit multicast to both CTAs when the MMAs that read stage s are done
//   tfull[a]   in each CTA, count 1: multicast commit when accumulator a is complete (each CTA drains its own TMEM)
//   tempty[a]  leader only, count 8 x SPLIT: one arrival per epilogue WARP of both CTAs (remote arrive from the peer)
template <class Epi, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc_threads(Epi::SPLIT), 1)
tc_rows_x_table_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape g,
                            const typename Epi::Params p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int STAGES2 = Epi::STAGES_PAIR;    // shared-memory ring depth of this epilogue's instance (shadows the default)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar0 = base + STAGES2 * STAGE2_BYTES;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES2 + 4);
    uint32_t *const counter = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot + 8u - raw));
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const uint32_t num_mb2 = (g.num_mb + 1u) >> 1;  // 256-row super tiles
    constexpr int BKE = bk_elems(KIND);
    constexpr uint32_t IDESC = idesc_for(KIND, 2 * BM);
#define BAR2_FULL(s) (bar0 + 8u * (uint32_t)(s))
#define BAR2_EMPTY(s) (bar0 + 8u * (uint32_t)(STAGES2 + (s)))
#define BAR2_TFULL(a) (bar0 + 8u * (uint32_t)(2 * STAGES2 + (a)))
#define BAR2_TEMPTY(a) (bar0 + 8u * (uint32_t)(2 * STAGES2 + 2 + (a)))

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(BAR2_FULL(s), 1);
            mbar_init(BAR2_EMPTY(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(BAR2_TFULL(a), 1);
            mbar_init(BAR2_TEMPTY(a), 8 * Epi::SPLIT);  // one arrival per epilogue warp of both CTAs
        }
        *counter = 0u;
        fence_barrier_init();
    }
    if (warp == 1) {  // one warp of EACH CTA of the pair takes part in the paired allocation
        tmem_alloc_pair(tmem_slot, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));
    const uint32_t lead_bar0 = mapa_shared(bar0, 0u);

    if (warp == 0) {
        // ===== TMA producer (both CTAs; each fills its own stage buffers, bytes complete on the leader's barrier) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t mb2 = pair; mb2 < num_mb2; mb2 += num_pairs)
                for (uint32_t nb = 0; nb < g.num_nb; ++nb)
                    for (uint32_t kb = 0; kb < g.num_kb; ++kb, ++it) {
                        const uint32_t s = it % STAGES2, ph = (it / STAGES2) & 1u;
                        mbar_wait(BAR2_EMPTY(s), ph ^ 1u);
                        if (rank == 0u) mbar_expect_tx(BAR2_FULL(s), 2u * STAGE2_BYTES);
                        const uint32_t sa = base + s * STAGE2_BYTES;
                        const uint32_t lead_full = lead_bar0 + 8u * s;
                        tma_load_2d_pair(sa, &tmA, (int32_t)(kb * BKE), (int32_t)((mb2 * 2u + rank) * BM), lead_full, HINT_EVICT_NORMAL);
                        tma_load_2d_pair(sa + A_BYTES, &tmB, (int32_t)(kb * BKE), (int32_t)(nb * BN + rank * (BN / 2)), lead_full,
                                         HINT_EVICT_LAST);
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives the tensor cores of both SMs =====
        if (lane == 0 && rank == 0u) {
            uint32_t it = 0, tile = 0;
            for (uint32_t mb2 = pair; mb2 < num_mb2; mb2 += num_pairs)
                for (uint32_t nb = 0; nb < g.num_nb; ++nb, ++tile) {
                    const uint32_t as = tile & 1u, aph = (tile >> 1) & 1u;
                    mbar_wait(BAR2_TEMPTY(as), aph ^ 1u);  // both CTAs' epilogues have drained this accumulator stage
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (uint32_t kb = 0; kb < g.num_kb; ++kb, ++it) {
                        const uint32_t s = it % STAGES2, ph = (it / STAGES2) & 1u;
                        mbar_wait(BAR2_FULL(s), ph);
                        tc_fence_after();
                        const uint32_t sa = base + s * STAGE2_BYTES;
                        const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < UMMA_STEPS; ++k)
                            umma_pair<KIND>(d_tmem, adesc + 2u * k, bdesc + 2u * k, IDESC, (kb | k) != 0u);
                        umma_commit_pair(BAR2_EMPTY(s), 3);  // frees stage s in both CTAs
                    }
                    umma_commit_pair(BAR2_TFULL(as), 3);  // accumulator complete in both CTAs
                }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps of both CTAs: each drains its own 128 TMEM lanes =====
        uint8_t *const after_bars = smem_raw + (bar0 + bar_bytes(STAGES2) - raw);
        Epi::run(EpiCtx{BAR2_TFULL(0), lead_bar0 + 8u * (uint32_t)(2 * STAGES2 + 2), 1u, tmem_base, warp & 3u, lane, counter,
                        pair * 2u + rank, num_pairs * 2u, num_mb2 * 2u, (warp - 2u) >> 2,
                        reinterpret_cast<uint32_t *>(after_bars), reinterpret_cast<float *>(after_bars + Epi::EXCH_BYTES),
                        threadIdx.x - 64u},
                 g, p);
    }
    // ---- teardown: nobody leaves (or frees TMEM) while the peer can still signal into this CTA
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) Epi::finish(counter, p);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
#undef BAR2_FULL
#undef BAR2_EMPTY
#undef BAR2_TFULL
#undef BAR2_TEMPTY
}

// epilogue of the k-means assignment filter (see the header of this file).
// Two warps share a row, each scanning half of every tile's columns with its own running bound m >= min_j s_j^true and its
// own candidate store in shared memory (its window is the wider one of its slice, so nothing the full-width scan keeps is
// lost); after the last tile of the row slice 1 publishes (m, store) and slice 0 finalises over both stores.
//
// Per 32-column chunk the common path is one FFMA and half a three-input minimum per column:
//     s_j = cn_j - 2 acc_j ,  smin = min_j s_j ,  awc = a * max_{j in chunk} w_j
//     m  <- min(m, smin + awc)              (an upper bound of min_j s_j^true: s_j^true <= ŝ_j + a w_j <= s_j + awc)
//     a column can only matter if  s_j - awc <= thr(m),  tested on smin for the chunk and on the 4-column minima the
//     reduction tree already holds for its eight groups.
// A warp holds 32 different rows, each with its own sequence of record minima, so SOME lane passes the chunk test in
// nearly every chunk (98.7 % measured): whatever happens then is paid by the whole warp.  The scan therefore does no
// per-column work at all: a group that passes is SAVED (its four scores as one 16-byte shared-memory store, its index
// pushed on a register stack) with predicated straight-line code, and the columns are only looked at after the row's last
// tile, against the final threshold and with each column's OWN w_j (one far-out centroid -- an empty cluster at the
// origin, src/ivf/index.rs:446-453 -- must not widen the window of its 31 neighbours) -- by then most saved groups
// (records that were overtaken) are outside the window, and every lane has the same kind of work.  The store holds
// SAVE groups per row and slice; when it is full the groups that have left the window are dropped (thr only shrinks),
// and a row whose live groups still do not fit is handed to the exact scan (overflow).
struct AssignEpi {
    typedef AssignTcParams Params;
    static constexpr int SPLIT = 2;
    static constexpr int NCH = BN / 32 / SPLIT;      // chunks per warp and tile
    static constexpr int STAGES_SINGLE = 3, STAGES_PAIR = 5;   // one stage less than the default: room for the group store
    static constexpr int SAVE = 6;                   // saved 4-column groups per row and slice
    static constexpr int GBITS = 10;                 // bits per group index on the register stack (6 x 10 in a u64): C <= 4096
    static constexpr int ETHREADS = 128 * SPLIT;     // epilogue threads per CTA
    // shared memory of the epilogue: SAVE x ETHREADS float4 entries + one uint4 per row (slice 1 -> slice 0: m, stack, count)
    static constexpr uint32_t SAVE_BYTES = SAVE * ETHREADS * 16;
    static constexpr uint32_t EXCH_BYTES = SAVE_BYTES + 128 * 16;
    static __device__ __forceinline__ void finish(uint32_t *, const Params &) {}

    static __device__ __forceinline__ float min4(const float4 v) { return fminf(fminf(v.x, v.y), fminf(v.z, v.w)); }
    // group index of entry e (0 = oldest) of a stack holding cnt entries
    static __device__ __forceinline__ uint32_t group_of(u64 stack, uint32_t cnt, uint32_t e) {
        return (uint32_t)(stack >> ((cnt - 1u - e) * GBITS)) & ((1u << GBITS) - 1u);
    }
    // drops the saved groups no column of which can still be inside the window; returns the new count
    static __device__ __forceinline__ float max4(const float4 v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
    static __device__ __noinline__ uint32_t compact(float4 *mine, uint32_t cnt, u64 *stack, float thr, float a, const float *wv) {
        uint32_t o = 0;
        u64 ns = 0;
        for (uint32_t e = 0; e < cnt; ++e) {
            const uint32_t grp = group_of(*stack, cnt, e);
            const float4 it = mine[e * ETHREADS];
            if (min4(it) - a * max4(*reinterpret_cast<const float4 *>(wv + (size_t)grp * 4u)) <= thr) {
                mine[(o++) * ETHREADS] = it;
                ns = (ns << GBITS) | grp;
            }
        }
        *stack = ns;
        return o;
    }

    static __device__ void run(const EpiCtx c, const GemmShape g, const Params &p) {
        const uint32_t q = c.q, lane = c.lane;
        const uint32_t row_in_tile = q * 32u + lane;
        // centroid norms and error weights for every tile of this CTA: staged once in shared memory when the table is short enough
        const uint32_t cn_len = g.num_nb * BN;
        const float *cnp = p.cn, *wvp = p.wv;
        const float *wgp = nullptr;   // per 4-column group: max of w (only when the table's constants fit the staging area)
        if (2u * cn_len + cn_len / 4u <= EPI_TAB_FLOATS) {
            for (uint32_t i = c.et; i < cn_len; i += ETHREADS) {
                c.tab[i] = p.cn[i];
                c.tab[cn_len + i] = p.wv[i];
            }
            for (uint32_t i = c.et; i < cn_len / 4u; i += ETHREADS)
                c.tab[2u * cn_len + i] = max4(*reinterpret_cast<const float4 *>(p.wv + (size_t)i * 4u));
            c.sync_epilogue(ETHREADS);
            cnp = c.tab;
            wvp = c.tab + cn_len;
            wgp = c.tab + 2u * cn_len;
        }
        float4 *const saves = reinterpret_cast<float4 *>(c.exch);
        float4 *const mine = saves + c.et;                                  // entry e of this thread: mine[e * ETHREADS]
        uint4 *const meta = reinterpret_cast<uint4 *>(saves + SAVE * ETHREADS) + row_in_tile;   // slice 1 -> slice 0 (same row)
        const float wmax = __uint_as_float(p.bounds[0]), cnmax = __uint_as_float(p.bounds[1]);
        const float bnmax = __uint_as_float(p.bounds[2]);
        // the tensor cores saw (sa x) and (sb B'): -2 / (sa sb) undoes both scales exactly (powers of two)
        const float neg2u = -2.f / ((p.scale_a ? *p.scale_a : 1.f) * __uint_as_float(p.bounds[5]));
        const float mun = sqrtf(__uint_as_float(p.bounds[3])) * 1.000001f;      // |mu_d|
        const float dmu = sqrtf(__uint_as_float(p.bounds[4])) * 1.000001f;      // |mu - mu_d|
        const float delta = (float)(p.dim / 4 + 12) * 5.9604645e-08f;   // 2^-24: reference's serial f32 chain, terms >= 0
        const float gamma = (float)(p.dim + 32) * 5.9604645e-08f;       // f32 sums of the row statistics
        const float c2 = 2.2f * delta;
        const bool table_ok = (cnmax < 1e30f) && (wmax < 1e30f) && (mun < 1e30f) && (dmu < 1e30f);
        const bool stack_ok = (cn_len >> 2) <= (1u << GBITS);   // every group index fits its stack field
        const float inf = __int_as_float(0x7f800000);
        uint32_t tile = 0;
        for (uint32_t mb = c.mb0; mb < c.mb_end; mb += c.mb_stride) {
            const u64 row = (u64)mb * BM + row_in_tile;
            const bool valid = row < p.n;
            const float2 st = valid ? p.stats[row] : make_float2(0.f, 0.f);
            const float x2 = st.x * (1.f + gamma) + 1e-37f;
            const float a = sqrtf(x2) * 1.000001f;
            // per column:  L_j = ŝ_j - a w_j <= s_j <= U_j = ŝ_j + a w_j  (up to the rounding slack folded into T2).  With
            // m >= min_j U_j the reference argmin satisfies  L_j <= thr(m) = m + T2 + 2.2 delta max(m + K2, 0):  m + K2 bounds
            // |x - c_j'|^2 for the j' attaining m, because s_j and d_j differ by exactly |x|^2 - 2 x.mu <= K2.
            const float mag = cnmax + 2.f * a * bnmax + 2.f * a * wmax;                  // bound on |ŝ|, |ŝ -+ a wc|
            const float kmag = x2 + 2.f * fabsf(st.y) + 2.f * a * dmu;
            const float T2 = mag * 1.9073486e-06f + 1e-37f;                              // 2^-19: cn / fma / awc / thr roundings
            const float K2 = (x2 - 2.f * st.y) + 2.f * a * dmu + 2.02f * gamma * a * mun + (kmag + mag) * 9.5367432e-07f;
            float m = inf;
            uint32_t cnt = 0;   // saved groups of this thread
            u64 stack = 0;      // their group indices (column / 4), GBITS each, newest in the low bits
            bool ovf = !stack_ok;
            for (uint32_t nb = 0; nb < g.num_nb; ++nb, ++tile) {
                const uint32_t taddr = c.acquire(tile);
                const float4 *cn4 = reinterpret_cast<const float4 *>(cnp + (size_t)nb * BN);
                const float *wcp = p.wc + (size_t)nb * (BN / 32);
#pragma unroll 1
                for (uint32_t ch = c.h * NCH; ch < (c.h + 1u) * NCH; ++ch) {
                    float v[32];
                    tmem_ld32(taddr + ch * 32u, v);
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 cc = cn4[ch * 8 + i4];
                        v[4 * i4 + 0] = __fmaf_rn(neg2u, v[4 * i4 + 0], cc.x);
                        v[4 * i4 + 1] = __fmaf_rn(neg2u, v[4 * i4 + 1], cc.y);
                        v[4 * i4 + 2] = __fmaf_rn(neg2u, v[4 * i4 + 2], cc.z);
                        v[4 * i4 + 3] = __fmaf_rn(neg2u, v[4 * i4 + 3], cc.w);
                    }
                    float sm[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) sm[i] = fminf(fminf(v[4 * i], v[4 * i + 1]), fminf(v[4 * i + 2], v[4 * i + 3]));
                    const float smin = fminf(fminf(fminf(sm[0], sm[1]), fminf(sm[2], sm[3])), fminf(fminf(sm[4], sm[5]), fminf(sm[6], sm[7])));
                    const float awc = a * __ldg(wcp + ch);
                    m = fminf(m, smin + awc);
                    float thr = m + T2 + c2 * fmaxf(m + K2, 0.f);
                    // s_j <= ts  <=>  s_j - awc <= thr (rounding of the two forms: inside T2); +inf scores (padding columns,
                    // overflowed norms) never win in the reference and never pass, even while m is still +inf
                    const float ts = fminf(thr + awc, 3.0e38f);
                    if (smin <= ts) {
                        // the groups are tested with their OWN widest w: one far-out centroid (an empty cluster at the origin)
                        // inflates awc of its chunk tenfold, and its 7 neighbour groups must not all be saved for it
                        const uint32_t g0 = (nb * BN + ch * 32u) >> 2;
                        float wg[8];
                        if (wgp) {
                            const float4 wa = *reinterpret_cast<const float4 *>(wgp + g0);
                            const float4 wb = *reinterpret_cast<const float4 *>(wgp + g0 + 4u);
                            wg[0] = wa.x, wg[1] = wa.y, wg[2] = wa.z, wg[3] = wa.w, wg[4] = wb.x, wg[5] = wb.y, wg[6] = wb.z, wg[7] = wb.w;
                        } else {
#pragma unroll
                            for (int g4 = 0; g4 < 8; ++g4) wg[g4] = max4(*reinterpret_cast<const float4 *>(wvp + (size_t)(g0 + g4) * 4u));
                        }
                        // tighten m with the groups' own bounds first (min4 + a max4(w) >= U_j of the group's best column): while
                        // the chunk of a far-out centroid holds the record, smin + awc alone would keep the window ten times
                        // too wide for every later chunk
                        float ug[8];
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) ug[g4] = __fmaf_rn(a, wg[g4], sm[g4]);
                        m = fminf(m, fminf(fminf(fminf(ug[0], ug[1]), fminf(ug[2], ug[3])), fminf(fminf(ug[4], ug[5]), fminf(ug[6], ug[7]))));
                        thr = m + T2 + c2 * fmaxf(m + K2, 0.f);
                        const float thr_c = fminf(thr, 3.0e38f);
                        uint32_t hits = 0;
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) hits |= (__fmaf_rn(-a, wg[g4], sm[g4]) <= thr_c) ? (1u << g4) : 0u;
                        if (cnt + (uint32_t)__popc(hits) > (uint32_t)SAVE) {   // rare: make room, or give the row up
                            cnt = compact(mine, cnt, &stack, thr, a, wvp);
                            if (cnt + (uint32_t)__popc(hits) > (uint32_t)SAVE) {
                                ovf = true;
                                hits = 0;
                            }
                        }
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if ((hits >> g4) & 1u) {
                                mine[cnt * ETHREADS] = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
                                stack = (stack << GBITS) | (u64)(g0 + (uint32_t)g4);
                                ++cnt;
                            }
                        }
                    }
                }
                c.release(tile);
            }
            // ---- after the row's last tile: every saved column's own bound U_j = s_j + a w_j tightens m (the column attaining
            // the minimum of U is itself inside the window, hence saved -- unless the row overflowed)
            for (uint32_t e = 0; e < cnt; ++e) {
                const float4 it = mine[e * ETHREADS];
                const float4 w4 = *reinterpret_cast<const float4 *>(wvp + (size_t)group_of(stack, cnt, e) * 4u);
                m = fminf(m, fminf(fminf(it.x + a * w4.x, it.y + a * w4.y), fminf(it.z + a * w4.z, it.w + a * w4.w)));
            }
            // ---- hand-over between the two column slices (named barrier 1 + q: the two warps that own TMEM quarter q): slice 1
            // publishes (m, stack, count, overflow); its entries are read in place by slice 0, which finalises the row
            if (c.h != 0u)
                *meta = make_uint4(__float_as_uint(m), cnt | (ovf ? 0x80000000u : 0u), (uint32_t)stack, (uint32_t)(stack >> 32));
            asm volatile("bar.sync %0, 64;" ::"r"(1u + q) : "memory");
            if (c.h == 0u) {
                const uint4 om = *meta;
                const uint32_t ocnt = om.y & 0x7FFFFFFFu;
                const u64 ostack = (u64)om.z | ((u64)om.w << 32);
                const float4 *const other = mine + 128;   // the same row's thread of slice 1
                m = fminf(m, __uint_as_float(om.x));
                ovf |= (om.y >> 31) != 0u;
                const float thr = m + T2 + c2 * fmaxf(m + K2, 0.f);
                // pass 1: which saved columns of both slices are inside the final window (bit 4 e + i of inwin)
                static_assert(2 * SAVE * 4 <= 64, "inwin holds one bit per saved column of both slices");
                uint32_t nv = 0, only = NONE;
                u64 inwin = 0;   // 2 x SAVE entries x 4 columns
                for (uint32_t e = 0; e < cnt + ocnt; ++e) {
                    const bool own = e < cnt;
                    const uint32_t grp = own ? group_of(stack, cnt, e) : group_of(ostack, ocnt, e - cnt);
                    const float4 it = own ? mine[e * ETHREADS] : other[(e - cnt) * ETHREADS];
                    const float4 w4 = *reinterpret_cast<const float4 *>(wvp + (size_t)grp * 4u);
                    const float lj[4] = {it.x - a * w4.x, it.y - a * w4.y, it.z - a * w4.z, it.w - a * w4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (lj[i] <= thr) {
                            ++nv;
                            only = grp * 4u + (uint32_t)i;
                            inwin |= 1ull << (4u * e + (uint32_t)i);
                        }
                    }
                }
                const bool finite = table_ok && (x2 < 1e30f) && (m < 1e30f) && (m > -1e30f) && (kmag < 1e30f);
                bool is_ovf = valid && (ovf || !finite || nv == 0u);
                if (is_ovf) atomicAdd(&p.counts[ovf ? 4 : (!finite ? 5 : 6)], 1u);   // diagnostics (PQV_TRACE): why the row left the filter
                bool is_amb = valid && !is_ovf && nv > 1u;
                if (valid && !is_ovf && nv == 1u) p.assign[row] = p.perm ? __ldg(p.perm + only) : only;
                // (row, candidate) pairs of the ambiguous rows: warp-aggregated reservation
                const uint32_t want = is_amb ? nv : 0u;
                uint32_t incl = want;
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
                    // pass 2: a reservation that does not fit leaves sentinels behind and the row takes the full scan
                    uint32_t slot = pbase + incl - want;
                    for (u64 bits = inwin; bits; bits &= bits - 1ull) {
                        const uint32_t b = (uint32_t)__ffsll((long long)bits) - 1u, e = b >> 2;
                        const uint32_t grp = e < cnt ? group_of(stack, cnt, e) : group_of(ostack, ocnt, e - cnt);
                        const uint32_t col = grp * 4u + (b & 3u);
                        if (slot < p.pair_cap)
                            p.pairs[slot] = fits ? make_uint2((uint32_t)row, p.perm ? __ldg(p.perm + col) : col) : make_uint2(NONE, NONE);
                        ++slot;
                    }
                    if (!fits) {
                        is_amb = false;
                        is_ovf = true;
                        atomicAdd(&p.counts[7], 1u);
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
            asm volatile("bar.sync %0, 64;" ::"r"(1u + q) : "memory");  // slice 1 may overwrite its store / meta for the next rows
        }
    }
};



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


// ================================================================================================
// Batched brute-force top-k: nq independent single-query searches (src/ivf/search.rs:112-141, src/df_vector/exec.rs:
// 257-277) answered in ONE pass over the table.  The reference has no batched entry point; every query's
// result must equal the single-query loop's.  The row tile x query tile contraction runs on the tensor cores as a filter:
//
//   s_rq := |x_r|^2 - 2 x_r.q            = d_true(r, q) - |q|^2   (per-query constant shift)
//   ŝ_rq := x2c_r - 2 tf32_mma(x_r, Q'_q)                         Q'_q = tf32_rn(q)
//   |ŝ - s| <= a_r w_q + rho_r            a_r >= |x_r|, w_q = 2 (eps_mma |Q'_q| + |q - Q'_q|), rho_r: norm/rounding slack
//
//   phase A (BATCH_SAMPLE): U_rq = ŝ + a_r w_q + rho_r >= s_rq for the first S rows; theta_q = k-th smallest U over them
//                           (+ the reference's own f32 rounding, 2.2 delta d): no row outside {s_rq <= theta_q} can be
//                           among the reference's k smallest distances, ties at the boundary included.
//   phase B (BATCH_FILTER): every (row, query) with  ŝ - a_r w_q - rho_r <= theta_q  is appended to the CTA's candidate
//                           region; pair_dist_kernel then evaluates the exact serial-order f32 distance of each candidate
//                           and topk_select_kernel keeps the k + 1 smallest (distance, row) keys per query.
// A query whose k + 1 smallest keys hold a boundary tie (d_k == d_k+1) or whose k returned values are not pairwise
// distinct is re-run through the single-query path: only there does the order depend on the reference heap's layout.
// ================================================================================================
enum { BATCH_SAMPLE = 0, BATCH_FILTER = 1 };
constexpr uint32_t FLAG_NONFINITE_ROW = 1u, FLAG_REGION_FULL = 2u;

// per query: Q^ = the rounded query in the operand type of the kernel, w_q, q2 = |q|^2 rounded up; qbounds[0] = max_q |Q^_q|;
// qwc[j / 32] = max of w over the 32-query chunk of j (zeroed by the caller)
template <int KIND>
__global__ void __launch_bounds__(128) query_prep_kernel(const float *__restrict__ Q, uint32_t nq, uint32_t dim,
                                                         void *__restrict__ Qp_, float *__restrict__ qw,
                                                         float *__restrict__ q2, uint32_t *__restrict__ qwc, uint32_t nq_pad,
                                                         uint32_t *__restrict__ qbounds, const half16::Globals *__restrict__ hg) {
    const uint32_t j = blockIdx.x;
    __shared__ double red[3][4];
    const float sq = KIND == KIND_F16 ? __uint_as_float(qbounds[2]) : 1.f;  // power of two (scale_from_absmax_kernel)
    if (j >= nq) {
        if (threadIdx.x == 0 && j < nq_pad) {
            qw[j] = 0.f;
            q2[j] = 0.f;
        }
        return;
    }
    double r2 = 0.0, b2 = 0.0, c2 = 0.0;
    for (uint32_t col = threadIdx.x; col < dim; col += blockDim.x) {
        const float c = Q[(size_t)j * dim + col];
        float b;
        if (KIND == KIND_F16) {
            const __half h = half16::to_half_flushed(c * sq);
            reinterpret_cast<__half *>(Qp_)[(size_t)j * dim + col] = h;
            b = __half2float(h) / sq;  // exact: sq is a power of two
        } else {
            b = tf32_rn(c);
            reinterpret_cast<float *>(Qp_)[(size_t)j * dim + col] = b;
        }
        const double r = (double)c - (double)b;
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
        const double kap = kappa_of(hg, dim);
        const double eps = kap + eps_acc(dim) * (1.0 + kap);  // as centroid_prep_kernel
        const float wj = (float)(2.0 * (eps * bn + rn) * up);
        qw[j] = wj;
        q2[j] = (float)(c2 * up);
        atomicMax(&qwc[j >> 5], __float_as_uint(wj));
        atomicMax(&qbounds[0], __float_as_uint((float)bn));
    }
}

struct BatchParams {
    const float2 *stats;     // [n] (.x = |x|^2 in f32, row_stats_kernel)
    const float *qw;         // [nq_pad] w_q, 0 padded
    const float *qtheta;     // [nq_pad] theta_q (-inf padded)                       BATCH_FILTER
    const float *qwc;        // [nq_pad / 32] per 32-query chunk: max_q w_q             BATCH_FILTER (chunk pre-test)
    const uint32_t *qbounds; // [0] = max_q |Q^_q| (f32 bits), [2] = operand scale of the queries (a power of two, f32)
    const float *scale_a;    // operand scale of the rows (half16::Globals::scale), nullptr = 1
    float *U;                // [nq_pad][ldU] upper bounds of s over the sample rows   BATCH_SAMPLE
    uint32_t ldU;
    uint2 *cand;             // [gridDim.x][region_cap] (row, query)                   BATCH_FILTER
    uint32_t region_cap;
    uint32_t *region_count;  // [gridDim.x]
    uint32_t *flags;         // FLAG_*
    u64 n;                   // rows covered by this launch
    uint32_t dim;
    // IVF mask (batched IVF search, all three null/0 otherwise): a (row, query) pair only counts if the row's cluster is
    // among the query's probed clusters.  probe_T[c * qwords + (q >> 5)] bit (q & 31): query q probes cluster c.
    const uint32_t *row_cluster;  // [n] cluster of each row, 0xFFFFFFFF = in no list
    const uint32_t *probe_T;      // [C][qwords]
    uint32_t qwords;              // nq_pad / 32
    const uint32_t *row_mask;     // [ceil(n / 32)] or null: bit r = row r passes the scan subtree's filter (all queries)
};

template <int MODE>
struct BatchEpi {
    typedef BatchParams Params;
    static constexpr int SPLIT = 2;  // every (row, query) pair is independent: the two warps of a row just split the columns
    static constexpr uint32_t EXCH_BYTES = 0;
    static constexpr int STAGES_SINGLE = STAGES, STAGES_PAIR = STAGES2;
    static __device__ __forceinline__ void finish(uint32_t *counter, const Params &p) {
        if (MODE == BATCH_FILTER) p.region_count[blockIdx.x] = min(*counter, p.region_cap);
    }
    static __device__ void run(const EpiCtx c, const GemmShape g, const Params &p) {
        const uint32_t lane = c.lane;
        const uint32_t row_in_tile = c.q * 32u + lane;
        // FILTER: theta_q / 2 for every query column, staged once per CTA (halving is exact; -inf padding stays -inf)
        const uint32_t nq_pad = g.num_nb * BN;
        const bool th_smem = (MODE == BATCH_FILTER) && nq_pad <= EPI_TAB_FLOATS;
        if (th_smem) {
            for (uint32_t i = c.et; i < nq_pad; i += 128u * SPLIT) c.tab[i] = 0.5f * p.qtheta[i];
            c.sync_epilogue(128u * SPLIT);
        }
        const float bnmax = __uint_as_float(p.qbounds[0]);
        // the tensor cores saw (sa x) and (sq q): u undoes both scales exactly (powers of two)
        const float u1 = 1.f / ((p.scale_a ? *p.scale_a : 1.f) * __uint_as_float(p.qbounds[2]));
        const float neg2u = -2.f * u1;
        const float gamma = (float)(p.dim + 32) * 5.9604645e-08f;  // f32 summation error of the row statistics
        uint2 *const region = (MODE == BATCH_FILTER) ? p.cand + (size_t)blockIdx.x * p.region_cap : nullptr;
        uint32_t tile = 0;
        for (uint32_t mb = c.mb0; mb < c.mb_end; mb += c.mb_stride) {
            const u64 row = (u64)mb * BM + row_in_tile;
            const bool valid = row < p.n;
            const float x2c = valid ? p.stats[row].x : 0.f;
            if (MODE == BATCH_FILTER && valid && !(x2c < 1e30f)) atomicOr(p.flags, FLAG_NONFINITE_ROW);
            const float x2hi = x2c * (1.f + 2.f * gamma) + 1e-37f;
            const float a = sqrtf(x2hi) * 1.000001f;
            const float rho = 1.5f * gamma * x2hi + (x2hi + 2.f * a * bnmax) * 9.5367432e-07f + 1e-37f;
            const float x2s = (MODE == BATCH_SAMPLE) ? x2c + rho : x2c - rho;
            const float hslack = (x2hi + 2.f * a * bnmax) * 9.5367432e-07f + 1e-37f;  // roundings of the exact test and of the pre-test
            // IVF mask: the probe words of this row's cluster (one word per 32 queries)
            const uint32_t *probe_row = nullptr;
            bool dead = false;  // filtered out for every query
            if (p.row_mask && valid) dead = !((__ldg(p.row_mask + (row >> 5)) >> (row & 31)) & 1u);
            if (p.row_cluster && valid && !dead) {
                const uint32_t cl = p.row_cluster[row];
                if (cl != 0xFFFFFFFFu) probe_row = p.probe_T + (size_t)cl * p.qwords;
            }
            for (uint32_t nb = 0; nb < g.num_nb; ++nb, ++tile) {
                const uint32_t taddr = c.acquire(tile);
                const float4 *w4 = reinterpret_cast<const float4 *>(p.qw + (size_t)nb * BN);
                const float4 *t4 = reinterpret_cast<const float4 *>(p.qtheta + (size_t)nb * BN);
#pragma unroll 1
                for (uint32_t ch = c.h * (BN / 32 / SPLIT); ch < (c.h + 1u) * (BN / 32 / SPLIT); ++ch) {
                    float v[32];
                    tmem_ld32(taddr + ch * 32u, v);
                    const uint32_t q0 = nb * BN + ch * 32u;
                    const uint32_t pm = dead ? 0u : (!p.row_cluster ? 0xFFFFFFFFu : (probe_row ? __ldg(probe_row + (q0 >> 5)) : 0u));
                    if (MODE == BATCH_SAMPLE) {
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 ww = __ldg(w4 + ch * 8 + i4);
                            const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float u = __fmaf_rn(a, wv[e], __fmaf_rn(neg2u, v[4 * i4 + e], x2s));
                                if (!((pm >> (4 * i4 + e)) & 1u)) u = 3.0e38f;  // not probed by this query: never among its k smallest
                                if (valid) p.U[(size_t)(q0 + 4 * i4 + e) * p.ldU + row] = u;  // lanes = consecutive rows: coalesced
                            }
                        }
                    } else {
                        // chunk pre-test: a pair hits iff  -2 v + x2s <= a w_q + theta_q  <=>  v + theta_q / 2 >= (x2s - a w_q) / 2;
                        // with wcq >= w_q of the chunk and the roundings of both forms inside hslack, the maximum of
                        // v + theta/2 over the chunk decides whether ANY pair of it can hit (one FADD + half a maximum per pair)
                        bool maybe = valid && !dead;
                        if (th_smem) {
                            const float4 *h4 = reinterpret_cast<const float4 *>(c.tab + q0);
                            float tm[8];
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                const float4 hh = h4[i4];
                                tm[i4] = fmaxf(fmaxf(__fmaf_rn(u1, v[4 * i4 + 0], hh.x), __fmaf_rn(u1, v[4 * i4 + 1], hh.y)),
                                               fmaxf(__fmaf_rn(u1, v[4 * i4 + 2], hh.z), __fmaf_rn(u1, v[4 * i4 + 3], hh.w)));
                            }
                            const float tmax = fmaxf(fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3])), fmaxf(fmaxf(tm[4], tm[5]), fmaxf(tm[6], tm[7])));
                            const float awq = a * __ldg(p.qwc + (q0 >> 5));
                            const float hcut = 0.5f * (x2s - awq) - hslack;
                            maybe = maybe && !(tmax < hcut);   // NaN (non-finite rows are flagged separately) falls through to the exact test
                        }
                        uint32_t mask = 0;
                        if (maybe) {
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                const float4 ww = __ldg(w4 + ch * 8 + i4);
                                const float4 tt = __ldg(t4 + ch * 8 + i4);
                                const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
                                const float tv[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const bool hit = __fmaf_rn(neg2u, v[4 * i4 + e], x2s) <= __fmaf_rn(a, wv[e], tv[e]);
                                    mask |= hit ? (1u << (4 * i4 + e)) : 0u;
                                }
                            }
                        }
                        if (!valid) mask = 0;
                        mask &= pm;
                        if (__any_sync(0xffffffffu, mask != 0u)) {
                            const uint32_t cnt = (uint32_t)__popc(mask);
                            uint32_t incl = cnt;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                                if ((int)lane >= o) incl += t;
                            }
                            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                            uint32_t base = 0;
                            if (lane == 0) base = atomicAdd(c.counter, total);
                            base = __shfl_sync(0xffffffffu, base, 0);
                            uint32_t slot = base + incl - cnt;
                            while (mask) {
                                const uint32_t i = (uint32_t)__ffs(mask) - 1u;
                                mask &= mask - 1u;
                                if (slot < p.region_cap) region[slot] = make_uint2((uint32_t)row, q0 + i);
                                else atomicOr(p.flags, FLAG_REGION_FULL);
                                ++slot;
                            }
                        }
                    }
                }
                c.release(tile);
            }
        }
    }
};

// order-preserving map of f32 bit patterns onto u32 (negative values below positive ones)
__device__ __forceinline__ uint32_t f32_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// one CTA per query: theta_q from an upper bound of the k-th smallest of U[q][0..S).  The S values are folded into M
// strided chunk minima (chunk c = elements c, c + M, ...; M >= 16 k): M distinct elements of the array, so their k-th
// smallest is >= the array's k-th smallest, and equal to it unless two of the k smallest share a chunk.  The k-th smallest
// of the M minima is then exact (MSB-first radix select in shared memory, 4 passes of 8 bits).
//   theta_q = th + (2^-20 + 2.2 delta) max(th + q2_q, 0) + rounding slack      (header of this section)
// queries >= nq (padding) get -inf so that they never produce candidates.  S < k: +inf (every row is a candidate).
__global__ void __launch_bounds__(256) theta_select_kernel(const float *__restrict__ U, uint32_t ldU, uint32_t S, uint32_t k,
                                                           uint32_t nq, const float *__restrict__ q2, float delta, uint32_t M,
                                                           float *__restrict__ qtheta, float *__restrict__ qT) {
    // qT[q] (reference-distance units): at least k rows of the sample have d_ref <= qT, and EVERY row of the table with
    // d_ref <= qT passes the filter -- d_true <= qT / (1 - delta) <= (th + |q|^2)(1 + 2.01 delta + 2^-23) stays below theta_q.
    // The IVF tie replay uses it to bound the sequence prefix that needs exact distances (pqv_ivf_impl.cuh).
    extern __shared__ float mins[];  // [M]
    const uint32_t q = blockIdx.x;
    if (q >= nq) {
        if (threadIdx.x == 0) {
            qtheta[q] = __int_as_float(0xff800000);
            qT[q] = __int_as_float(0xff800000);
        }
        return;
    }
    if (S < k) {
        if (threadIdx.x == 0) {
            qtheta[q] = __int_as_float(0x7f800000);
            qT[q] = __int_as_float(0xff800000);   // no bound from the sample: the replay takes the whole sequence
        }
        return;
    }
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining;
    const float *u = U + (size_t)q * ldU;
    const float inf = __int_as_float(0x7f800000);
    // chunk c is owned by thread c % 256: coalesced reads, no atomics
    for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) {
        float m0 = inf, m1 = inf, m2 = inf, m3 = inf;
        uint32_t i = c;
        for (; i + 3 * M < S; i += 4 * M) {
            m0 = fminf(m0, u[i]);
            m1 = fminf(m1, u[i + M]);
            m2 = fminf(m2, u[i + 2 * M]);
            m3 = fminf(m3, u[i + 3 * M]);
        }
        for (; i < S; i += M) m0 = fminf(m0, u[i]);
        mins[c] = fminf(fminf(m0, m1), fminf(m2, m3));  // +inf for an empty chunk (M > S): ranks above every real value
    }
    if (threadIdx.x == 0) {
        s_prefix = 0u;
        s_remaining = min(k, min(M, S));  // rank (1-based) of the wanted element among those matching the prefix
    }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t shift = 8u * (uint32_t)pass;
        const uint32_t himask = (pass == 3) ? 0u : (0xFFFFFFFFu << (shift + 8u));
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
            const uint32_t o = f32_ordered(mins[i]);
            if ((o & himask) == prefix) atomicAdd(&hist[(o >> shift) & 0xFFu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t rem = s_remaining, b = 0;
            for (; b < 256; ++b) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            b = b < 256 ? b : 255;
            s_prefix = prefix | (b << shift);
            s_remaining = rem;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float th = f32_unordered(s_prefix);
        const float qq = q2[q];
        const float d = fmaxf(th + qq, 0.f);
        qtheta[q] = th + (9.5367432e-07f + 2.2f * delta) * d + (fabsf(th) + qq) * 9.5367432e-07f + 1e-37f;
        qT[q] = d * (1.f + delta + 1.2e-07f);
    }
}

// Exact squared distance of every candidate (row, query) pair in the reference's operation order; the result goes to the
// query's key segment as bits(distance) << 32 | row.  One warp per 32 pairs, transposed through a padded shared tile
// exactly like group_distance / pair_exact_kernel.
//   ORDER 0: src/ivf/index.rs:461-480      sum += ((d0^2 + d1^2) + d2^2) + d3^2   (one chain term per 4 columns)
//   ORDER 1: src/df_vector/exec.rs:529-533 dist += diff * diff                    (one chain term per column)
template <int ORDER>
struct PairCfg {
    static constexpr int TERMS = ORDER == 1 ? 4 : 1;       // chain terms per lane per 128-column block
    static constexpr int TS = 32 * TERMS + 4;              // tile row stride (floats); TS/4 odd -> conflict-free LDS.128
    static constexpr int WARPS = ORDER == 1 ? 2 : 8;       // static shared memory <= 48 KB
};

template <int ORDER>
__global__ void __launch_bounds__(PairCfg<ORDER>::WARPS * 32)
pair_dist_kernel(const float *__restrict__ rows, uint32_t dim, const float *__restrict__ Q,
                 const uint2 *__restrict__ cand, uint32_t region_cap, const uint32_t *__restrict__ region_count,
                 u64 *__restrict__ seg, uint32_t cap_q, uint32_t *__restrict__ cntq) {
    constexpr int TS = PairCfg<ORDER>::TS, WARPS = PairCfg<ORDER>::WARPS, TERMS = PairCfg<ORDER>::TERMS;
    __shared__ __align__(16) float tiles[WARPS][32 * TS];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *tile = tiles[wib];
    const uint32_t npairs = region_count[blockIdx.y];
    const uint2 *pairs = cand + (size_t)blockIdx.y * region_cap;
    const uint32_t ngroups = (npairs + 31u) / 32u;
    const uint32_t ncb = (dim + 127u) / 128u;
    for (uint32_t g = blockIdx.x * WARPS + wib; g < ngroups; g += gridDim.x * WARPS) {
        const uint32_t idx = g * 32u + lane;
        const bool valid = idx < npairs;
        const uint2 pr = valid ? pairs[idx] : make_uint2(0u, 0u);
        float sum = 0.f;
        for (uint32_t cb = 0; cb < ncb; ++cb) {
            const uint32_t col = cb * 128u + (lane << 2);
            const bool inb = col < dim;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
                float4 xv[8], qv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t rr = __shfl_sync(0xffffffffu, pr.x, r0 + j);
                    const uint32_t qq = __shfl_sync(0xffffffffu, pr.y, r0 + j);
                    xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    qv[j] = xv[j];
                    if (inb) {
                        xv[j] = ldg4(rows + (u64)rr * dim + col);
                        qv[j] = ldg4(Q + (size_t)qq * dim + col);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (ORDER == 0) {
                        tile[(r0 + j) * TS + lane] = chunk4(qv[j], xv[j]);   // squared_l2_distance(query, vec), search.rs:117
                    } else {
                        float4 t;                                            // diff = value - q, exec.rs:531
                        t.x = sq1(xv[j].x, qv[j].x);
                        t.y = sq1(xv[j].y, qv[j].y);
                        t.z = sq1(xv[j].z, qv[j].z);
                        t.w = sq1(xv[j].w, qv[j].w);
                        *reinterpret_cast<float4 *>(tile + (r0 + j) * TS + (lane << 2)) = t;
                    }
                }
            }
            __syncwarp();
            // columns past dim contribute +0.0 terms: x + (+0.0) == x bit for bit for the non-negative partial sums
            const float4 *tr = reinterpret_cast<const float4 *>(tile + lane * TS);
#pragma unroll
            for (int t4 = 0; t4 < 8 * TERMS; ++t4) {
                const float4 t = tr[t4];
                sum = __fadd_rn(sum, t.x);
                sum = __fadd_rn(sum, t.y);
                sum = __fadd_rn(sum, t.z);
                sum = __fadd_rn(sum, t.w);
            }
            __syncwarp();
        }
        if (valid) {
            const uint32_t slot = atomicAdd(&cntq[pr.y], 1u);
            if (slot < cap_q) seg[(size_t)pr.y * cap_q + slot] = ((u64)__float_as_uint(sum) << 32) | (u64)pr.x;
        }
    }
}

// one CTA per query: the kk = min(k + 1, cnt) smallest keys of the query's segment (8-pass MSB-first radix select on the
// u64 keys, then a bitonic sort of the survivors in shared memory).  out_keys[q][0..kout) ascending; out_info[q] =
// count | TIE << 30 | OVERFLOW << 31 where TIE means the caller must re-run the query through the single-query path.
constexpr uint32_t SEL_TIE = 1u << 30, SEL_OVERFLOW = 1u << 31;
constexpr int SEL_MAX = 2048;  // >= PQV_MAX_K + 1, power of two

__global__ void __launch_bounds__(256) topk_select_kernel(const u64 *__restrict__ seg, uint32_t cap_q,
                                                          const uint32_t *__restrict__ cntq, uint32_t k, int apply_sqrt,
                                                          u64 *__restrict__ out_keys, uint32_t *__restrict__ out_info,
                                                          uint32_t kout) {
    const uint32_t q = blockIdx.x;
    const uint32_t raw_cnt = cntq[q];
    const uint32_t cnt = min(raw_cnt, cap_q);
    const u64 *keys = seg + (size_t)q * cap_q;
    const uint32_t kk = min(k + 1u, cnt);
    __shared__ uint32_t hist[256];
    __shared__ u64 s_prefix;
    __shared__ uint32_t s_remaining, s_fill, s_tie;
    __shared__ u64 buf[SEL_MAX];
    if (threadIdx.x == 0) {
        s_prefix = 0ull;
        s_remaining = kk;
        s_fill = 0u;
        s_tie = 0u;
    }
    __syncthreads();
    if (kk > 0) {
        for (int pass = 7; pass >= 0; --pass) {
            hist[threadIdx.x] = 0u;
            __syncthreads();
            const u64 prefix = s_prefix;
            const uint32_t shift = 8u * (uint32_t)pass;
            const u64 himask = (pass == 7) ? 0ull : (~0ull << (shift + 8u));
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
                const u64 key = keys[i];
                if ((key & himask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 0xFFu], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t rem = s_remaining, b = 0;
                for (; b < 256; ++b) {
                    if (hist[b] >= rem) break;
                    rem -= hist[b];
                }
                b = b < 256 ? b : 255;
                s_prefix = prefix | ((u64)b << shift);
                s_remaining = rem;
            }
            __syncthreads();
        }
    }
    const u64 kth = s_prefix;  // the kk-th smallest key (keys are unique: one per row)
    for (uint32_t i = threadIdx.x; i < SEL_MAX; i += blockDim.x) buf[i] = KEY_MAX;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const u64 key = keys[i];
        if (kk > 0 && key <= kth) {
            const uint32_t slot = atomicAdd(&s_fill, 1u);
            if (slot < (uint32_t)SEL_MAX) buf[slot] = key;
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= (uint32_t)SEL_MAX; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (uint32_t)SEL_MAX / 2; i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const u64 x = buf[lo], y = buf[hi];
                if ((x > y) == up) {
                    buf[lo] = y;
                    buf[hi] = x;
                }
            }
            __syncthreads();
        }
    // kout = k: the k winners; kout = k + 1 (sharded search): the boundary key too, the ranks' lists are merged later
    const uint32_t nout = min(k, cnt);
    if (kout > k && threadIdx.x == 0 && cnt > k) out_keys[(size_t)q * kout + k] = buf[k];
    for (uint32_t i = threadIdx.x; i < nout; i += blockDim.x) {
        out_keys[(size_t)q * kout + i] = buf[i];
        if (i + 1 < nout) {  // returned values must be pairwise distinct, else their order is the heap layout's
            const float d0 = __uint_as_float((uint32_t)(buf[i] >> 32)), d1 = __uint_as_float((uint32_t)(buf[i + 1] >> 32));
            const bool same = apply_sqrt ? (__fsqrt_rn(d0) == __fsqrt_rn(d1)) : (d0 == d1);
            if (same) atomicOr(&s_tie, 1u);
        }
    }
    // boundary: the k-th and (k+1)-th smallest squared distances must differ, else the kept set is the heap layout's
    if (threadIdx.x == 0 && cnt > k && (uint32_t)(buf[k - 1] >> 32) == (uint32_t)(buf[k] >> 32)) atomicOr(&s_tie, 1u);
    __syncthreads();
    if (threadIdx.x == 0)
        out_info[q] = min(kout, cnt) | (s_tie ? SEL_TIE : 0u) | ((raw_cnt > cap_q || s_fill > (uint32_t)SEL_MAX) ? SEL_OVERFLOW : 0u);
}

#undef BAR_FULL
#undef BAR_EMPTY
#undef BAR_TFULL
#undef BAR_TEMPTY

}  // namespace tc
}  // namespace pqv
