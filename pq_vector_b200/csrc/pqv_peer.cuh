// pqv_peer.cuh -- the candidate exchange of a sharded search over NVLink peer memory (included by pqv_capi.cu).
//
// One process per GPU, every rank scans its own rows.  The only exchange of the path is the per-rank
// heap-entrant candidate list (a few KB).  With NCCL that costs a host round trip on each side of the collective plus its
// launch; here the ranks map one another's exchange buffer once (CUDA IPC) and the scan's tail kernel WRITES the rank's
// candidates straight into every peer's buffer over NVLink, publishes a sequence flag, and waits for the peers' flags:
//
//   l2_scan_topk_kernel -> merge -> entrant_filter_kernel -> peer_publish_kernel -> peer_wait_pack_kernel (writes the
//   union of all ranks' live keys into page-locked host memory)
//
// Buffer of a rank (device memory, local):   2 parities x [ world slots x (1 + cap) u64 ]  +  2 x world u64 flags.
// Slot r of parity p holds rank r's (count, keys) of the search with sequence number s, s % 2 == p; flag[p][r] == s once
// the slot is complete.  A rank can be at most one search ahead of its peers (its search s + 1 only finishes when every
// peer has published s + 1, i.e. has finished reading s), so two parities are enough.
#pragma once

namespace pqv {

// CTA d copies this rank's candidate block (count + min(count, cap) keys) into slot `rank` of peer d's buffer, makes it
// visible system-wide and then releases the flag.
// The count word also carries, in its upper half, the distance bits of this rank's final k-th smallest key (0xFFFFFFFF: the
// slice holds fewer than k rows): the reader uses it to drop the later ranks' entrants that the global heap cannot admit.
__global__ void __launch_bounds__(256) peer_publish_kernel(const u64 *__restrict__ ent_out, const uint32_t cap,
                                                           u64 *const *__restrict__ peer_base, const uint32_t rank,
                                                           const uint32_t world, const u64 seq,
                                                           const u64 *__restrict__ final_topk, const uint32_t k) {
    const uint32_t d = blockIdx.x;
    const u64 par = seq & 1ull;
    const u64 slot_words = 1ull + cap;
    u64 *slot = peer_base[d] + (par * world + rank) * slot_words;
    u64 *flag = peer_base[d] + 2ull * world * slot_words + par * world + rank;
    const u64 count = ent_out[0];
    const u64 n = count < (u64)cap ? count : (u64)cap;
    for (u64 i = threadIdx.x; i < n; i += blockDim.x) slot[1 + i] = ent_out[1 + i];
    if (threadIdx.x == 0) {
        const u64 kth = final_topk ? final_topk[k - 1] : ~0ull;
        slot[0] = (count & 0xFFFFFFFFull) | ((kth == ~0ull ? 0xFFFFFFFFull : (kth >> 32)) << 32);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
    }
}

// waits until every rank's flag of this parity carries `seq` (bounded: a lost peer must surface as an error, not a hang)
__global__ void __launch_bounds__(32) peer_wait_kernel(const u64 *__restrict__ local_base, const uint32_t cap,
                                                       const uint32_t world, const u64 seq, uint32_t *__restrict__ timed_out) {
    const u64 par = seq & 1ull;
    const u64 *flags = local_base + 2ull * world * (1ull + cap) + par * world;
    for (uint32_t r = threadIdx.x; r < world; r += 32) {
        u64 v = 0;
        uint64_t spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
            if (v == seq) break;
            __nanosleep(200);
        } while (++spins < (1ull << 27));  // ~ half a minute: ranks may be skewed by first-call allocations
        if (v != seq) atomicOr(timed_out, 1u);
    }
}

// The same wait, followed by the read-back itself: once every flag carries `seq`, the CTA packs the live keys of all slots
// (rank order) straight into page-locked HOST memory -- out[0] = timed-out | any slot over capacity << 1, out[1] = total,
// out[2 + r] = count of rank r, keys from out[2 + world] on.  A search reads back a few KB instead of the whole
// world x (1 + cap) block, and needs no separate copy after the kernel.
__global__ void __launch_bounds__(256) peer_wait_pack_kernel(const u64 *__restrict__ local_base, const uint32_t cap,
                                                             const uint32_t world, const u64 seq, u64 *__restrict__ out,
                                                             const uint32_t raw) {
    // raw != 0: a structured payload (batch key lists): every word of every slot, in place and in rank order, no filter
    __shared__ uint32_t s_timed_out;
    const u64 par = seq & 1ull;
    const u64 slot_words = 1ull + cap;
    const u64 *slots = local_base + par * world * slot_words;
    const u64 *flags = local_base + 2ull * world * slot_words + par * world;
    if (threadIdx.x == 0) s_timed_out = 0u;
    __syncthreads();
    if (threadIdx.x < 32) {
        for (uint32_t r = threadIdx.x; r < world; r += 32) {
            u64 v = 0;
            uint64_t spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
                if (v == seq) break;
                __nanosleep(200);
            } while (++spins < (1ull << 27));
            if (v != seq) atomicOr(&s_timed_out, 1u);
        }
    }
    __syncthreads();
    // Rows are numbered rank by rank, so when the global heap reaches rank r's rows it already holds the k best of ranks
    // < r: its threshold is at most T_r = min_{j<r} (rank j's final k-th distance).  An entrant of rank r with d >= T_r is
    // never admitted (admission needs d < threshold) and is dropped here; NaN keys always travel (the host decides).
    // out[2 + r] keeps the RAW count (overflow detection), out[1] = keys actually packed, in no particular order.
    __shared__ uint32_t s_thr[64], s_raw[64];
    __shared__ uint32_t s_total, s_over;
    if (threadIdx.x == 0) {
        uint32_t T = 0xFFFFFFFFu, over = 0;
        for (uint32_t r = 0; r < world; ++r) {
            const u64 head = __ldcg(slots + r * slot_words);
            const uint32_t c = (uint32_t)head;
            if (c > cap) over = 1;
            s_thr[r] = T;
            s_raw[r] = c < cap ? c : cap;
            out[2 + r] = c;
            const uint32_t kth = (uint32_t)(head >> 32);
            if (kth < T && !raw) T = kth;
        }
        s_total = 0;
        s_over = over;
    }
    __syncthreads();
    u64 *dst = out + 2 + world;
    if (raw) {
        uint32_t off = 0;
        for (uint32_t r = 0; r < world; ++r) {
            const uint32_t n = s_raw[r];
            const u64 *src = slots + r * slot_words + 1;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[off + i] = __ldcg(src + i);
            off += n;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            out[0] = (u64)s_timed_out | ((u64)s_over << 1);
            out[1] = off;
        }
        return;
    }
    for (uint32_t r = 0; r < world; ++r) {
        const uint32_t n = s_raw[r], T = s_thr[r];
        const u64 *src = slots + r * slot_words + 1;
        for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
            const uint32_t i = i0 + threadIdx.x;
            u64 key = 0;
            bool keep = false;
            if (i < n) {
                key = __ldcg(src + i);
                const uint32_t b = (uint32_t)(key >> 32);
                keep = b < T || (b & 0x7FFFFFFFu) > 0x7F800000u;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            uint32_t base = 0;
            if ((threadIdx.x & 31u) == 0u && m) base = atomicAdd(&s_total, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) dst[base + (uint32_t)__popc(m & ((1u << (threadIdx.x & 31u)) - 1u))] = key;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = (u64)s_timed_out | ((u64)s_over << 1);
        out[1] = s_total;
    }
}

}  // namespace pqv

