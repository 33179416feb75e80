// pqv_tie.cuh -- tie queries of the batched brute-force pass, resolved together.
//
// A query whose answer hinges on the layout of the reference's BinaryHeap (bit-equal returned distances, or a tie across
// the k boundary) needs the reference loop replayed over every row the heap ever admits.  Behind the sample prefix
// [0, S) those rows are among the query's exact-distance candidates (the batched pass keeps them); inside the prefix
// they used to come from one exact single-query scan of the prefix PER tie query (0.5 ms each: 27 of the 51 ms of a
// 1024-query top-100 batch).  Here the prefix is read ONCE for all tie queries:
//   prefix_dist_matrix_kernel  exact distances (reference summation order, bit-identical to group_distance<ORDER>) of
//                              the prefix rows against the selected queries -> a [queries][rows] f32 matrix;
//   prefix_entrants_kernel     one CTA per query walks its row of that matrix in position order with the exact running
//                              k-th smallest distance and emits every row that is below the threshold in force at the
//                              start of its 2048-row chunk (the threshold never rises, so this is a superset of the
//                              heap's admissions; the first chunk is emitted whole) as bits(d) << 32 | position.
// The host replays the reference heap over these keys + the candidates behind the prefix, as before.
#pragma once
#include "pqv_kernels.cuh"

namespace pqv {
namespace tie {

constexpr int TM = 64, TN = 64, TK = 32, TLD = TK + 4;  // queries x rows x columns per tile step
constexpr uint32_t CHUNK = 2048;                        // rows per threshold refresh in prefix_entrants_kernel
constexpr uint32_t SORT_MAX = 4096;                     // shared-memory keys: best list (<= 1024) + one chunk's candidates

// acc += one 4-column step of the reference chain
template <int ORDER>
__device__ __forceinline__ float chain4(float acc, const float4 a, const float4 b) {
    if constexpr (ORDER == 0) {
        return __fadd_rn(acc, chunk4(a, b));  // index.rs:467-471
    } else {                                  // exec.rs:529-533: dist += diff * diff, column by column
        acc = __fadd_rn(acc, sq1(a.x, b.x));
        acc = __fadd_rn(acc, sq1(a.y, b.y));
        acc = __fadd_rn(acc, sq1(a.z, b.z));
        return __fadd_rn(acc, sq1(a.w, b.w));
    }
}

// out[qi * ldo + row] = distance(queries[qsel[qi]], rows[row]) for row < n, qi < nq_sel.  dim % 4 == 0, 16-byte aligned
// rows (what the batched pass requires anyway).  256 threads = 16 (query direction) x 16 (row direction); a thread owns
// queries ty*4 + i and rows tx + 16*j, each pair its own serial chain; operands staged in blocks of 32 columns, double
// buffered.  grid = (row tiles, query tiles).
template <int ORDER>
__global__ void __launch_bounds__(256, 2) prefix_dist_matrix_kernel(const float *__restrict__ rows, const u64 n,
                                                                    const uint32_t dim,
                                                                    const float *__restrict__ queries,
                                                                    const uint32_t *__restrict__ qsel,
                                                                    const uint32_t nq_sel, float *__restrict__ out,
                                                                    const u64 ldo) {
    __shared__ __align__(16) float Qs[2][TM * TLD];
    __shared__ __align__(16) float Rs[2][TN * TLD];
    const uint32_t tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const u64 row0 = (u64)blockIdx.x * TN;
    const uint32_t q0 = blockIdx.y * TM;
    const uint32_t n4 = dim >> 2, nkb = (n4 + 7) >> 3;

    // 64 vectors x 8 four-column chunks = 512 float4 per tile -> 2 per thread
    auto load_rows = [&](float *dst, const uint32_t kb) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const uint32_t idx = tid + it * 256, r = idx >> 3, c = idx & 7, chunk = kb * 8 + c;
            u64 row = row0 + r;
            row = row < n ? row : n - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chunk < n4) v = *reinterpret_cast<const float4 *>(rows + row * dim + (chunk << 2));
            *reinterpret_cast<float4 *>(dst + r * TLD + (c << 2)) = v;
        }
    };
    auto load_queries = [&](float *dst, const uint32_t kb) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const uint32_t idx = tid + it * 256, r = idx >> 3, c = idx & 7, chunk = kb * 8 + c;
            uint32_t qi = q0 + r;
            qi = qi < nq_sel ? qi : nq_sel - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chunk < n4) v = *reinterpret_cast<const float4 *>(queries + (u64)qsel[qi] * dim + (chunk << 2));
            *reinterpret_cast<float4 *>(dst + r * TLD + (c << 2)) = v;
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int buf = 0;
    load_queries(Qs[0], 0);
    load_rows(Rs[0], 0);
    __syncthreads();
    for (uint32_t kb = 0; kb < nkb; ++kb) {
        if (kb + 1 < nkb) {
            load_queries(Qs[buf ^ 1], kb + 1);
            load_rows(Rs[buf ^ 1], kb + 1);
        }
        const float *A = Qs[buf] + (ty * 4) * TLD;
        const float *B = Rs[buf] + tx * TLD;
        // columns past dim in the last block are zero in both operands: (0 - 0)^2 = +0 and x + 0 = x for the non-negative
        // partial sums of a chain, so padding terms leave every chain's bits unchanged
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(A + i * TLD + (c << 2));
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4 *>(B + (16 * j) * TLD + (c << 2));
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = chain4<ORDER>(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t qi = q0 + ty * 4 + i;
        if (qi >= nq_sel) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u64 row = row0 + tx + 16 * j;
            if (row < n) out[(u64)qi * ldo + row] = acc[i][j];
        }
    }
}

// One CTA of 1024 threads per query: dmat row blockIdx.x holds the exact distances of rows [0, S).  Region layout
// out[blockIdx.x * (cap + 1)]: [0] = number of keys produced (may exceed cap: then the region is incomplete and the
// caller falls back), [1 ..] = keys in no particular order.  kcap = pow2ceil(max(k, 32)) <= 1024.
__global__ void __launch_bounds__(1024) prefix_entrants_kernel(const float *__restrict__ dmat, const u64 ldo,
                                                               const uint32_t S, const uint32_t k, const uint32_t kcap,
                                                               u64 *__restrict__ out, const uint32_t cap) {
    __shared__ u64 s[SORT_MAX];
    __shared__ uint32_t s_cnt;
    const uint32_t tid = threadIdx.x;
    const float *d = dmat + (u64)blockIdx.x * ldo;
    u64 *o = out + (u64)blockIdx.x * ((u64)cap + 1);

    // first chunk: while the heap fills every row is admitted -- emit all of it, then take the k smallest
    const uint32_t F = S < CHUNK ? S : CHUNK;
    for (uint32_t i = tid; i < CHUNK; i += 1024) {
        u64 key = KEY_MAX;
        if (i < F) {
            key = ((u64)__float_as_uint(d[i]) << 32) | i;
            if (i < cap) o[1 + i] = key;
        }
        s[i] = key;
    }
    __syncthreads();
    bitonic_sort_smem(s, CHUNK, tid, 1024);
    for (uint32_t i = k + tid; i < SORT_MAX; i += 1024) s[i] = KEY_MAX;  // keep s[0 .. k); candidates land at s[kcap ..]
    uint32_t total = F;
    __syncthreads();

    for (uint32_t base = CHUNK; base < S; base += CHUNK) {
        if (tid == 0) s_cnt = 0;
        const u64 kth = s[k - 1];  // k <= F here (S > CHUNK >= k), so the list is full
        const float thr = __uint_as_float((uint32_t)(kth >> 32));
        __syncthreads();
#pragma unroll
        for (uint32_t it = 0; it < CHUNK / 1024; ++it) {
            const uint32_t pos = base + it * 1024 + tid;
            if (pos < S) {
                const float dv = d[pos];
                if (dv < thr) {  // strict, as `distance < top.distance` (search.rs:121, exec.rs:476)
                    const uint32_t idx = atomicAdd(&s_cnt, 1u);
                    s[kcap + idx] = ((u64)__float_as_uint(dv) << 32) | pos;
                }
            }
        }
        __syncthreads();
        const uint32_t cnt = s_cnt;  // uniform
        if (cnt) {
            for (uint32_t i = tid; i < cnt; i += 1024)
                if ((u64)total + i < cap) o[1 + total + i] = s[kcap + i];
            total += cnt;
            uint32_t n_sort = 2 * kcap;
            while (n_sort < kcap + cnt) n_sort <<= 1;  // <= 1024 + 2048 -> 4096 = SORT_MAX
            // s[kcap + cnt .. n_sort) is KEY_MAX already (reset below after every merge)
            __syncthreads();
            bitonic_sort_smem(s, n_sort, tid, 1024);
            for (uint32_t i = k + tid; i < n_sort; i += 1024) s[i] = KEY_MAX;
        }
        __syncthreads();
    }
    if (tid == 0) o[0] = total;
}

}  // namespace tie
}  // namespace pqv
