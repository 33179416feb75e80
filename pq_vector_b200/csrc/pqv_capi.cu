// pqv_capi.cu -- the C ABI of include/pqv.h on top of the sm_100a kernels in pqv_kernels.cuh.
//
// Host-side responsibilities: device scratch management, dataset residency (row-range shards over the
// context's devices), launch sequencing on a per-device stream, and the final exact replay of the
// reference's bounded BinaryHeap over the (small) heap-entrant superset the kernels emit
// (DESIGN.md section 4.3).  No CPU fallback exists: every distance is computed on the GPU.
#include "../../include/pqv.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <time.h>
#include <vector>

#include "pqv_kernels.cuh"
#include "pqv_half.cuh"
#include "pqv_peer.cuh"
#include "pqv_adist.cuh"
#include "pqv_tie.cuh"

using pqv::u64;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? PQV_ENOMEM : PQV_ECUDA;                 \
            return fail(code__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                              \
    } while (0)

#define PQV_TRY(expr)          \
    do {                       \
        int rc__ = (expr);     \
        if (rc__) return rc__; \
    } while (0)

// Stable ascending sort for the reference's `sort_by(|a, b| a.partial_cmp(b).unwrap_or(Equal))` sites
// (src/ivf/index.rs:143, src/ivf/search.rs:136-140, src/df_vector/exec.rs:270-274).  For NaN-free keys every stable sort
// gives the same result.  With a NaN the comparator is not an order and the outcome depends on the algorithm -- Rust's
// own changes with the std version (merge sort / driftsort), so it cannot be pinned; this is the plain top-down merge
// (split at n/2, the right element is taken only if strictly smaller) that the oracle uses, so product and checker
// agree on those inputs too.
template <typename T, typename Less>
static void merge_sort_stable_rec(T *a, T *tmp, size_t n, Less less) {
    if (n < 2) return;
    const size_t mid = n / 2;
    merge_sort_stable_rec(a, tmp, mid, less);
    merge_sort_stable_rec(a + mid, tmp, n - mid, less);
    size_t i = 0, j = mid, o = 0;
    while (i < mid && j < n) tmp[o++] = less(a[j], a[i]) ? a[j++] : a[i++];
    while (i < mid) tmp[o++] = a[i++];
    while (j < n) tmp[o++] = a[j++];
    for (size_t t = 0; t < n; ++t) a[t] = tmp[t];
}
template <typename T, typename Less>
static void merge_sort_stable(std::vector<T> &v, Less less) {
    std::vector<T> tmp(v.size());
    merge_sort_stable_rec(v.data(), tmp.data(), v.size(), less);
}

// ------------------------------------------------------------------------------------------------
// Rust std::collections::BinaryHeap<HeapItem> (max-heap on distance, NaN -> Equal), written against
// the published std source: push = sift_up(0, old_len); pop = swap last into root +
// sift_down_to_bottom(0) (+ sift_up); into_iter = backing-vector order.  Used to replay the
// reference loop src/ivf/search.rs:115-127 / src/df_vector/exec.rs:467-482 over the entrant set.
// ------------------------------------------------------------------------------------------------
namespace {

struct HeapItem {
    float distance;
    uint32_t row_idx;
};

struct RustMaxHeap {
    std::vector<HeapItem> data;
    static bool le(const HeapItem &a, const HeapItem &b) { return !(a.distance > b.distance); }
    void sift_up(size_t start, size_t pos) {
        HeapItem elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void sift_down_to_bottom(size_t pos) {
        const size_t end = data.size(), start = pos;
        HeapItem elt = data[pos];
        size_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {
            child += le(data[child], data[child + 1]) ? 1 : 0;
            data[pos] = data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            data[pos] = data[child];
            pos = child;
        }
        data[pos] = elt;
        sift_up(start, pos);
    }
    void push(HeapItem it) {
        data.push_back(it);
        sift_up(0, data.size() - 1);
    }
    void pop() {
        HeapItem last = data.back();
        data.pop_back();
        if (!data.empty()) {
            data[0] = last;
            sift_down_to_bottom(0);
        }
    }
};

// bits(d) of a NaN distance (either sign): the kernels emit such rows as entrants unconditionally (pqv_kernels.cuh)
inline bool key_is_nan(u64 key) { return ((uint32_t)(key >> 32) & 0x7FFFFFFFu) > 0x7F800000u; }
inline bool any_nan_key(const u64 *keys, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (key_is_nan(keys[i])) return true;
    return false;
}

inline float key_dist(u64 key) {
    uint32_t b = (uint32_t)(key >> 32);
    float f;
    memcpy(&f, &b, 4);
    return f;
}
inline uint32_t key_pos(u64 key) { return (uint32_t)(key & 0xFFFFFFFFull); }

// entrants: keys (bits(d) << 32 | position) in any order; a superset of every candidate the
// reference heap admits.  Replays the reference loop in position order, then (sqrt), stable sort.
// maps a candidate position to its row id: identity, a host array (gather API) or a functor (IVF lists)
struct RowMap {
    const uint32_t *ids = nullptr;
    const std::function<uint32_t(uint32_t)> *fn = nullptr;
    // sparse map for the entrants only: (position << 32 | row id), ascending (fused IVF search)
    const u64 *pairs = nullptr;
    size_t n_pairs = 0;
    uint32_t operator()(uint32_t pos) const {
        if (pairs) {
            const u64 *it = std::lower_bound(pairs, pairs + n_pairs, (u64)pos << 32);
            return (uint32_t)*it;
        }
        return fn ? (*fn)(pos) : (ids ? ids[pos] : pos);
    }
};

// stable LSD radix sort of u64 items by the 32-bit field (item >> shift) & 0xFFFFFFFF: 11-bit digits, passes whose
// digit is constant are skipped.  ~10x faster than std::sort for the ~1e3 entrant keys of a query, and the replay sits
// on the latency path of every search.
void radix_sort_field(std::vector<u64> &v, unsigned shift) {
    const size_t n = v.size();
    if (n < 2) return;
    if (n < 64) {
        std::sort(v.begin(), v.end(), [shift](u64 a, u64 b) { return (uint32_t)(a >> shift) < (uint32_t)(b >> shift); });
        return;
    }
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (u64 x : v) {
        const uint32_t f = (uint32_t)(x >> shift);
        lo &= f;
        hi |= f;
    }
    const uint32_t varying = lo ^ hi;  // bits that differ between at least two items
    std::vector<u64> tmp(n);
    u64 *src = v.data(), *dst = tmp.data();
    for (unsigned d = 0; d < 32; d += 11) {
        if (((varying >> d) & 0x7FFu) == 0) continue;
        uint32_t hist[2048] = {0};
        for (size_t i = 0; i < n; ++i) hist[((uint32_t)(src[i] >> shift) >> d) & 0x7FFu]++;
        uint32_t run = 0;
        for (uint32_t &h : hist) {
            const uint32_t c = h;
            h = run;
            run += c;
        }
        for (size_t i = 0; i < n; ++i) dst[hist[((uint32_t)(src[i] >> shift) >> d) & 0x7FFu]++] = src[i];
        std::swap(src, dst);
    }
    if (src != v.data()) memcpy(v.data(), src, n * 8);
}

// the reference loop (src/ivf/search.rs:115-127) over (distance, row id) items already in position order
struct ReplayItem {
    float d;
    uint32_t row;
};
template <typename Next>
size_t replay_ordered(size_t n, Next next, uint32_t k, uint32_t flags, uint32_t *out_rows, float *out_dist) {
    RustMaxHeap h;
    h.data.reserve((size_t)k + 1);
    for (size_t i = 0; i < n; ++i) {
        const ReplayItem e = next(i);
        HeapItem it{e.d, e.row};
        if (h.data.size() < k) {
            h.push(it);
        } else if (e.d < h.data[0].distance) {
            h.pop();
            h.push(it);
        }
    }
    std::vector<HeapItem> &r = h.data;
    if (flags & PQV_SQRT)
        for (auto &it : r) it.distance = sqrtf(it.distance);
    merge_sort_stable(r, [](const HeapItem &a, const HeapItem &b) { return a.distance < b.distance; });
    for (size_t i = 0; i < r.size(); ++i) {
        out_rows[i] = r[i].row_idx;
        out_dist[i] = r[i].distance;
    }
    return r.size();
}

// Shortcut around the heap replay.  The replay over an entrant set E keeps the k smallest distances of E; which rows
// those are is unique unless the k-th and (k+1)-th smallest distances of E are equal, and the output order (stable
// sort by the returned value, src/ivf/search.rs:134-140 / src/df_vector/exec.rs:269-274) is unique unless two
// returned values are equal.  When neither happens (and no NaN is involved) the answer is the k smallest keys in
// ascending order, whatever the heap's layout history was; otherwise return false and let the caller replay.
// row_at(i, pos) = row id of entrant i.
template <typename RowAt>
bool topk_without_replay(const u64 *keys, size_t n, uint32_t k, uint32_t flags, RowAt row_at, uint32_t *out_rows,
                         float *out_dist, size_t *out_cnt) {
    struct KeyIdx {
        u64 key;
        uint32_t idx;
        bool operator<(const KeyIdx &o) const { return key < o.key; }
    };
    static thread_local std::vector<KeyIdx> v;
    if (any_nan_key(keys, n)) return false;  // a NaN inside the reference heap changes which rows it admits: replay
    v.resize(n);
    for (size_t i = 0; i < n; ++i) v[i] = KeyIdx{keys[i], (uint32_t)i};
    const size_t take = std::min<size_t>(k, n);
    if (n > take) {
        std::nth_element(v.begin(), v.begin() + take, v.end());  // v[take] = (k+1)-th smallest, smaller ones before it
        if (take && (uint32_t)(v[take].key >> 32) == (uint32_t)(std::max_element(v.begin(), v.begin() + take)->key >> 32))
            return false;  // tie across the k boundary
    }
    std::sort(v.begin(), v.begin() + take);
    if (take && (uint32_t)(v[take - 1].key >> 32) > 0x7F800000u) return false;  // NaN (or negative-sign bits): replay
    for (size_t i = 0; i < take; ++i) {
        const float d = key_dist(v[i].key);
        out_dist[i] = (flags & PQV_SQRT) ? sqrtf(d) : d;
        if (i && out_dist[i] == out_dist[i - 1]) return false;  // equal returned values: order hinges on the heap layout
    }
    for (size_t i = 0; i < take; ++i) out_rows[i] = row_at(v[i].idx, key_pos(v[i].key));
    *out_cnt = take;
    return true;
}

// known_tie: the caller already knows the shortcut cannot decide this query (a tie query of a batch)
size_t replay_reference_heap(std::vector<u64> &entrants, const RowMap &row_of, uint32_t k, uint32_t flags,
                             uint32_t *out_rows, float *out_dist, bool known_tie = false) {
    size_t fast_cnt = 0;
    if (!known_tie &&
        topk_without_replay(entrants.data(), entrants.size(), k, flags, [&](uint32_t, uint32_t pos) { return row_of(pos); },
                            out_rows, out_dist, &fast_cnt))
        return fast_cnt;
    radix_sort_field(entrants, 0);  // by position
    return replay_ordered(
        entrants.size(),
        [&](size_t i) {
            const u64 key = entrants[i];
            return ReplayItem{key_dist(key), row_of(key_pos(key))};
        },
        k, flags, out_rows, out_dist);
}

double trace_now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

uint32_t pow2ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------
// per-device state
// ------------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    int ensure(size_t n) {
        if (n <= cap) return PQV_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e != cudaSuccess) return fail(PQV_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
        cap = want;
        return PQV_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return PQV_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost((void **)&p, n * sizeof(T));
        if (e != cudaSuccess) return fail(PQV_ENOMEM, "cudaMallocHost(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
        cap = n;
        return PQV_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr uint32_t ENT_FIRST_CHUNK = 8192;  // entrant keys fetched with the first D2H (64 KiB)

struct AppendPool;
struct DeviceState {
    int dev = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<float> d_query;
    DevBuf<u64> cta_topk, ent, final_topk, ent_out, within_prefix, group_total, group_prefix;
    DevBuf<uint32_t> ent_count, gthr, d_row_ids, d_assign;
    DevBuf<float> d_vec, d_tmp_rows, d_centroids, d_dist;
    // tcgen05 assignment filter scratch (pqv_tc_host.cuh)
    DevBuf<float> tc_bp, tc_mu, tc_cn;
    DevBuf<uint32_t> tc_wc;
    // 16-bit shadow of rows that do not belong to a resident table (host rows streamed through, the k-means training sample)
    DevBuf<__half> ts_half;
    DevBuf<float2> ts_stats;
    DevBuf<float> ts_mu, ts_mean_part;
    DevBuf<pqv::half16::Globals> ts_g;
    cudaEvent_t ev_shadow[2] = {nullptr, nullptr};
    DevBuf<float2> tc_stats;
    DevBuf<uint2> tc_pairs;
    DevBuf<u64> tc_best;
    DevBuf<uint32_t> tc_u32, tc_amb_rows, tc_ovf_rows, tc_perm;
    DevBuf<u64> tc_okeys;
    // batched top-k scratch
    DevBuf<float> tb_Q, tb_Qp, tb_qf, tb_U;
    DevBuf<uint32_t> tb_u32, tb_info;
    DevBuf<uint2> tb_cand;
    DevBuf<u64> tb_seg, tb_keys;
    PinBuf<u64> h_batch_keys;
    // tie queries of a batch resolved together (pqv_tie.cuh): prefix distance matrix, selected query ids, entrant regions
    DevBuf<float> tie_dmat;
    DevBuf<uint32_t> tie_qsel;
    DevBuf<u64> tie_ent;
    PinBuf<u64> h_tie_ent, h_tie_seg;
    PinBuf<uint32_t> h_tie_qsel;
    PinBuf<u64> h_ent_out, h_final;
    PinBuf<float> h_query;
    // fused IVF search: row ids of the entrants + (candidate count, NaN flag)
    DevBuf<uint32_t> ent_rows;
    // device-side inverted-list build (csr_*_kernel)
    DevBuf<uint32_t> csr_counts, csr_totals;
    DevBuf<u64> ivf_info;
    PinBuf<uint32_t> h_ent_rows;
    PinBuf<u64> h_ivf_info;
    // table load (pqv_dataset_append of a large pageable buffer): independent double-buffered staging lanes, one host
    // thread each (append_staged)
    struct AppendLane {
        cudaStream_t st = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        PinBuf<unsigned char> buf[2];
    };
    std::vector<AppendLane> lanes;
    AppendPool *pool = nullptr;  // persistent copy threads, created on the first large append
    // VectorTopKExec candidate handling (ivf_mark / bitmap_* kernels): candidate bitmap, filter bitmap, block sums
    DevBuf<uint32_t> vt_bitmap, vt_mask, vt_sums;
    // un-indexed array_distance arm (pqv_adist.cuh): f64 query, f64 distance column, radix-select state, k winners
    DevBuf<double> ad_query, ad_col, ad_out_dist;
    DevBuf<uint32_t> ad_out_row;
    DevBuf<pqv::SelState> ad_state;
    PinBuf<double> h_ad_dist;
    PinBuf<uint32_t> h_ad_row;
    PinBuf<pqv::SelState> h_ad_state;
};

constexpr size_t APPEND_CHUNK = 8u << 20;
constexpr size_t APPEND_STAGED_MIN = 2u << 20;   // below this one plain copy is as fast
constexpr size_t APPEND_CHUNK_MIN = 512u << 10;
constexpr size_t APPEND_MAX_LANES = 8;

// persistent worker threads (thread creation + the CUDA runtime's per-thread set-up cost ~10 ms per call otherwise)
struct AppendPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    u64 job_id = 0;
    size_t pending = 0;
    bool stop = false;
    // the current job
    DeviceState *D = nullptr;
    unsigned char *dst = nullptr;
    const unsigned char *src = nullptr;
    size_t bytes = 0, n_chunks = 0, T = 0, chunk = APPEND_CHUNK;
    cudaError_t errs[APPEND_MAX_LANES];

    void lane_run(size_t t) {
        DeviceState::AppendLane &L = D->lanes[t];
        cudaError_t e = cudaSuccess;
        size_t turn = 0;
        for (size_t c = t; c < n_chunks && e == cudaSuccess; c += T, ++turn) {
            const int b = (int)(turn & 1);
            const size_t off = c * chunk, len = std::min(chunk, bytes - off);
            e = cudaEventSynchronize(L.ev[b]);  // the DMA that last read this staging buffer is done
            if (e != cudaSuccess) break;
            memcpy(L.buf[b].p, src + off, len);
            e = cudaMemcpyAsync(dst + off, L.buf[b].p, len, cudaMemcpyHostToDevice, L.st);
            if (e == cudaSuccess) e = cudaEventRecord(L.ev[b], L.st);
        }
        const cudaError_t s2 = cudaStreamSynchronize(L.st);
        errs[t] = e != cudaSuccess ? e : s2;
    }
    void worker(size_t t, int dev) {
        cudaSetDevice(dev);
        u64 seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_job.wait(lk, [&] { return stop || job_id != seen; });
            if (stop) return;
            seen = job_id;
            if (t < T) {
                lk.unlock();
                lane_run(t);
                lk.lock();
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_job.notify_all();
        for (auto &x : th) x.join();
        th.clear();
    }
};

// 16-bit operand shadow of a shard (pqv_half.cuh): fp16 copy of the rows, per-row (|x|^2, x.mu), the data mean mu and the
// table-wide residual bound.  Created on the first tensor-core pass over the shard, extended when rows are appended.
struct HalfShadow {
    __half *h = nullptr;
    float2 *stats = nullptr;
    float *mu = nullptr;
    pqv::half16::Globals *g = nullptr;
    u64 rows = 0, cap = 0;  // rows covered / allocated
    void drop() {
        if (h) cudaFree(h);
        if (stats) cudaFree(stats);
        if (mu) cudaFree(mu);
        if (g) cudaFree(g);
        h = nullptr;
        stats = nullptr;
        mu = nullptr;
        g = nullptr;
        rows = cap = 0;
    }
};

struct Shard {
    int di = 0;  // index into ctx->devs
    float *d_data = nullptr;
    u64 cap_rows = 0, n_rows = 0, first_row = 0;
    // |x|^2 per row for the tf32 form of the batched top-k (no shadow), kept until the shard changes
    float2 *d_norms = nullptr;
    u64 norms_rows = 0, norms_cap = 0;
    HalfShadow shadow;
    void drop_norms() {
        if (d_norms) cudaFree(d_norms);
        d_norms = nullptr;
        norms_rows = norms_cap = 0;
    }
    void drop_derived() {
        drop_norms();
        shadow.drop();
    }
};

struct Dataset {
    uint32_t dim = 0;
    u64 n_rows = 0;
    std::vector<Shard> shards;
};

struct StreamState {
    uint32_t dim = 0, k = 0, flags = 0;
    u64 rows_pushed = 0;
    u64 ent_rows_bound = 0;  // rows pushed since the device entrant accumulator was last drained
    int cur = 0;
    DevBuf<float> staging[2];
    DevBuf<double> staging64;
    cudaEvent_t scan_done[2] = {nullptr, nullptr};
    cudaEvent_t copy_done = nullptr;
    DevBuf<u64> carry[2];
    int carry_cur = 0;
    DevBuf<float> d_query;
    DevBuf<u64> ent_acc;  // [0] = count, then keys
    std::vector<u64> host_entrants;
    bool any = false;
    // distance of every pushed row, in push order (4 bytes per row next to dim * 4 read): only read back when a NaN distance
    // shows up among the entrants, to replay the reference loop literally (src/df_vector/exec.rs:467-482)
    DevBuf<float> dist_log;
    bool log_ok = true;
};
constexpr u64 STREAM_LOG_MAX_ROWS = 1ull << 30;  // 4 GB of log; beyond that a NaN distance is answered from the entrants alone

static void stream_state_free(StreamState *s) {
    s->staging[0].release();
    s->staging[1].release();
    s->staging64.release();
    s->carry[0].release();
    s->carry[1].release();
    s->d_query.release();
    s->ent_acc.release();
    s->dist_log.release();
    for (auto &ev : s->scan_done)
        if (ev) cudaEventDestroy(ev);
    if (s->copy_done) cudaEventDestroy(s->copy_done);
    delete s;
}

// NVLink peer-memory candidate exchange (pqv_peer.cuh)
struct PeerExchange {
    bool ready = false;
    uint32_t world = 0, rank = 0, cap = 0;
    u64 seq = 0;
    u64 *local = nullptr;      // this rank's buffer (cudaMalloc, exported through CUDA IPC)
    std::vector<u64 *> peers;  // peers[d]: rank d's buffer mapped into this process (peers[rank] == local)
    u64 **d_peers = nullptr;   // device copy of `peers`
    uint32_t *d_timeout = nullptr;
    PinBuf<u64> h_block;       // world x (1 + cap) words + 1 (timeout flag)
    double t_enqueued = 0;     // PQV_TRACE: when the last exchange's launches were all enqueued
    size_t words() const { return 2ull * world * (1ull + cap) + 2ull * world; }
};

}  // namespace

struct pqv_ctx {
    std::vector<DeviceState> devs;
    std::map<u64, Dataset> datasets;
    std::map<u64, StreamState *> streams;
    // device buffers and events of the last finished streaming top-k, kept for the next pqv_topk_stream_begin: a
    // VectorTopKExec execution per query would otherwise pay cudaMalloc + cudaFree of ~0.5 GB of staging every time
    StreamState *stream_cache = nullptr;
    std::map<u64, void *> indexes;  // IvfIndex*, see pqv_ivf_impl.cuh
    u64 next_handle = 1;
    std::mutex mu;
    pqv_timing last{};
    pqv_assign_timing last_assign{};
    pqv_batch_timing last_batch{};
    // what pqv_l2_topk_batch_keys left on the device for pqv_l2_topk_batch_tie_candidates (valid until the next batched
    // call or any change of the dataset)
    struct BatchState {
        bool valid = false;
        u64 handle = 0, S = 0;
        uint32_t nq = 0, k = 0, flags = 0, cap_q = 0, pos_base = 0;
        int dev_index = 0;
        std::vector<uint32_t> seg_count;
    } batch_state;
    int occ_override = 0;
    int scan_variant = 0;
    int gather_variant = 0;
    int seq_gather_variant = 0;
    PeerExchange peer;
    // coalescing front door for concurrent single-query callers (pqv_l2_topk_coalesced)
    struct CoalesceReq {
        u64 handle = 0;
        u64 index = 0;       // 0: brute force (pqv_l2_topk); else the IVF index (pqv_ivf_search)
        uint32_t nprobe = 0;
        uint32_t k = 0, flags = 0;
        const float *query = nullptr;
        uint32_t *rows = nullptr;
        float *dist = nullptr;
        uint32_t *count = nullptr;
        int status = 0;
        std::string err;
        bool done = false, lead = false;
    };
    struct Coalescer {
        std::mutex m;
        std::condition_variable cv;
        std::deque<CoalesceReq *> pending;
        bool leader_active = false;
        uint32_t max_batch = 1024, window_us = 0;
        u64 n_queries = 0, n_batches = 0, max_seen = 0;
        std::map<u64, uint32_t> dims;  // dataset handle -> dim, so that enqueueing never waits on ctx->mu
    } co;
};

namespace {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE and sticky per kernel: remember the largest size set for
// (current device, kernel) and only ever raise it.  Every launch that needs more than 48 KB of dynamic shared memory goes
// through here (a context may drive several devices, and the same kernel is then launched on each of them).
static int ensure_dyn_smem(const void *kern, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> raised;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = raised[{dev, kern}];
    if (smem > cur) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PQV_ECUDA, "cudaFuncSetAttribute(%zu bytes of dynamic shared memory): %s", smem, cudaGetErrorString(e));
        cur = smem;
    }
    return PQV_OK;
}

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        // a thread that never called cudaSetDevice reports device 0 but has NO current context: runtime calls bind one
        // lazily, the driver entry points used for TMA descriptors (cuTensorMapEncodeTiled) fail with INVALID_CONTEXT.
        // Callers are arbitrary threads (tokio workers), so bind explicitly once per thread.
        static thread_local bool bound = false;
        cudaGetDevice(&prev);
        if (prev != dev || !bound) {
            cudaSetDevice(dev);
            bound = true;
        }
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------------
// scan launch
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_WARPS = 8;

struct ScanGeom {
    int variant = 0;
    uint32_t kcap, sort_n, flush_at, grid;
    size_t smem;
    bool vec4;
    bool gather = false;
    int seq_variant = 0;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is sticky per kernel: only ever raise it
template <int ORDER, bool VEC4, bool GATHER>
int scan_ensure_smem_attr(size_t smem) {
    auto kern = pqv::l2_scan_topk_kernel<ORDER, VEC4, GATHER, SCAN_WARPS>;
    return ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem);
}

template <int ORDER, bool VEC4, bool GATHER>
int scan_launch_t(const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st) {
    PQV_TRY((scan_ensure_smem_attr<ORDER, VEC4, GATHER>(smem)));
    pqv::l2_scan_topk_kernel<ORDER, VEC4, GATHER, SCAN_WARPS><<<grid, SCAN_WARPS * 32, smem, st>>>(p);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

template <int ORDER, bool VEC4, bool GATHER>
int scan_occupancy_t(size_t smem, int *occ) {
    // the occupancy query costs ~10 us: remember the answer per (variant, smem)
    static std::mutex mu;
    static std::map<size_t, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(smem);
    if (it != cache.end()) {
        *occ = it->second;
        return PQV_OK;
    }
    PQV_TRY((scan_ensure_smem_attr<ORDER, VEC4, GATHER>(smem)));
    auto kern = pqv::l2_scan_topk_kernel<ORDER, VEC4, GATHER, SCAN_WARPS>;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, SCAN_WARPS * 32, smem));
    cache[smem] = *occ;
    return PQV_OK;
}

// Tuning variants of the dense unroll-4 vector kernel (PQV_SCAN_VARIANT=n; 0 = the shipped default, which is
// the <8,2,2> configuration = entry 4 of this table).
struct ScanVariant {
    int rb, cbv, minb;
};
static const ScanVariant kVariants[] = {{8, 1, 2}, {16, 1, 2}, {4, 1, 3}, {4, 2, 2}, {8, 2, 2}, {4, 1, 4}, {2, 2, 3}, {8, 1, 1}, {16, 1, 1},
                                       {16, 2, 2}, {8, 3, 2}, {4, 3, 2}, {16, 2, 1}};

template <int RB, int CBV, int MINB>
int scan_variant_go(const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st, int *occ_out) {
    auto kern = pqv::l2_scan_topk_kernel<0, true, false, SCAN_WARPS, RB, CBV, MINB>;
    PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    if (occ_out) {
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_out, kern, SCAN_WARPS * 32, smem));
        return PQV_OK;
    }
    kern<<<grid, SCAN_WARPS * 32, smem, st>>>(p);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

static int scan_variant_dispatch(int v, const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st, int *occ_out) {
    switch (v) {
        case 1: return scan_variant_go<16, 1, 2>(p, grid, smem, st, occ_out);
        case 2: return scan_variant_go<4, 1, 3>(p, grid, smem, st, occ_out);
        case 3: return scan_variant_go<4, 2, 2>(p, grid, smem, st, occ_out);
        case 4: return scan_variant_go<8, 2, 2>(p, grid, smem, st, occ_out);
        case 5: return scan_variant_go<4, 1, 4>(p, grid, smem, st, occ_out);
        case 6: return scan_variant_go<2, 2, 3>(p, grid, smem, st, occ_out);
        case 7: return scan_variant_go<8, 1, 1>(p, grid, smem, st, occ_out);
        case 8: return scan_variant_go<16, 1, 1>(p, grid, smem, st, occ_out);
        case 9: return scan_variant_go<16, 2, 2>(p, grid, smem, st, occ_out);
        case 10: return scan_variant_go<8, 3, 2>(p, grid, smem, st, occ_out);
        case 11: return scan_variant_go<4, 3, 2>(p, grid, smem, st, occ_out);
        case 12: return scan_variant_go<16, 2, 1>(p, grid, smem, st, occ_out);
        default: return fail(PQV_EINVAL, "unknown PQV_SCAN_VARIANT %d", v);
    }
}

// Tuning variants of the GATHERED unroll-4 vector kernel (PQV_GATHER_VARIANT=n; 0 = the shipped default).  The default
// <4, 2, 2> keeps 4 KB per warp in flight (8 x 2 spills under the 128-register cap of two CTAs per SM); one CTA per SM lifts
// the cap.
static const ScanVariant kGatherVariants[] = {{4, 2, 2}, {8, 2, 1}, {16, 1, 1}, {16, 2, 1}, {8, 1, 2}, {8, 2, 2}, {8, 3, 1}};

template <int RB, int CBV, int MINB>
int gather_variant_go(const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st, int *occ_out) {
    auto kern = pqv::l2_scan_topk_kernel<0, true, true, SCAN_WARPS, RB, CBV, MINB>;
    PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    if (occ_out) {
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_out, kern, SCAN_WARPS * 32, smem));
        return PQV_OK;
    }
    kern<<<grid, SCAN_WARPS * 32, smem, st>>>(p);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

static int gather_variant_dispatch(int v, const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st, int *occ_out) {
    switch (v) {
        case 1: return gather_variant_go<8, 2, 1>(p, grid, smem, st, occ_out);
        case 2: return gather_variant_go<16, 1, 1>(p, grid, smem, st, occ_out);
        case 3: return gather_variant_go<16, 2, 1>(p, grid, smem, st, occ_out);
        case 4: return gather_variant_go<8, 1, 2>(p, grid, smem, st, occ_out);
        case 5: return gather_variant_go<8, 2, 2>(p, grid, smem, st, occ_out);
        case 6: return gather_variant_go<8, 3, 1>(p, grid, smem, st, occ_out);
        default: return fail(PQV_EINVAL, "unknown PQV_GATHER_VARIANT %d", v);
    }
}

// ... and of the gathered SEQUENTIAL-order kernel (PQV_SEQ_GATHER_VARIANT=n; 0 = the shipped <8 row pairs, two CTAs per SM>)
template <int RB, int MINB>
int seq_gather_variant_go(const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st, int *occ_out) {
    auto kern = pqv::l2_scan_topk_kernel<1, true, true, SCAN_WARPS, RB, 1, MINB>;
    PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));
    if (occ_out) {
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_out, kern, SCAN_WARPS * 32, smem));
        return PQV_OK;
    }
    kern<<<grid, SCAN_WARPS * 32, smem, st>>>(p);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}
static const ScanVariant kSeqGatherVariants[] = {{8, 1, 2}, {16, 1, 2}, {16, 1, 1}, {4, 1, 2}, {4, 1, 3}, {8, 1, 1}};
static int seq_gather_variant_dispatch(int v, const pqv::ScanParams &p, uint32_t grid, size_t smem, cudaStream_t st, int *occ_out) {
    switch (v) {
        case 1: return seq_gather_variant_go<16, 2>(p, grid, smem, st, occ_out);
        case 2: return seq_gather_variant_go<16, 1>(p, grid, smem, st, occ_out);
        case 3: return seq_gather_variant_go<4, 2>(p, grid, smem, st, occ_out);
        case 4: return seq_gather_variant_go<4, 3>(p, grid, smem, st, occ_out);
        case 5: return seq_gather_variant_go<8, 1>(p, grid, smem, st, occ_out);
        default: return fail(PQV_EINVAL, "unknown PQV_SEQ_GATHER_VARIANT %d", v);
    }
}

#define SCAN_DISPATCH(FN, order, vec4, gather, ...)                                                 \
    ((order) == 0 ? ((vec4) ? ((gather) ? FN<0, true, true>(__VA_ARGS__) : FN<0, true, false>(__VA_ARGS__))      \
                            : ((gather) ? FN<0, false, true>(__VA_ARGS__) : FN<0, false, false>(__VA_ARGS__)))   \
                  : ((vec4) ? ((gather) ? FN<1, true, true>(__VA_ARGS__) : FN<1, true, false>(__VA_ARGS__))      \
                            : ((gather) ? FN<1, false, true>(__VA_ARGS__) : FN<1, false, false>(__VA_ARGS__))))

size_t scan_smem_bytes(int order, bool vec4, uint32_t dim, uint32_t sort_n) {
    const size_t tile = (order == 1 && vec4) ? pqv::TileCfg<1, true>::TILE_FLOATS
                        : (order == 0 && vec4) ? pqv::TileCfg<0, true, pqv::ScanDefaults<0, true>::CBV>::TILE_FLOATS
                                               : pqv::TileCfg<0, false>::TILE_FLOATS;
    const size_t dim_pad = (dim + 3u) & ~3u;
    return dim_pad * 4 + (size_t)SCAN_WARPS * tile * 4 + (size_t)sort_n * 8;
}

int scan_geometry(pqv_ctx *ctx, DeviceState &D, const float *d_data, u64 n, uint32_t dim, uint32_t k, int order,
                  bool gather, ScanGeom *g) {
    g->vec4 = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_data) & 15) == 0);
    g->kcap = std::max<uint32_t>(32, pow2ceil(k));
    g->flush_at = std::max<uint32_t>(1, g->kcap / 2);
    g->sort_n = pow2ceil(g->kcap + g->flush_at + SCAN_WARPS * 32);
    g->smem = scan_smem_bytes(order, g->vec4, dim, g->sort_n);
    g->variant = (order == 0 && g->vec4) ? (gather ? ctx->gather_variant : ctx->scan_variant) : 0;
    g->gather = gather;
    g->seq_variant = (order == 1 && g->vec4 && gather) ? ctx->seq_gather_variant : 0;
    if (g->seq_variant > 0) {
        const ScanVariant &sv = kSeqGatherVariants[g->seq_variant];
        int occ = 0;
        pqv::ScanParams dummy{};
        PQV_TRY(seq_gather_variant_dispatch(g->seq_variant, dummy, 0, g->smem, nullptr, &occ));
        if (occ < 1) return fail(PQV_ELIMIT, "scan variant does not fit");
        occ = std::min(occ, ctx->occ_override > 0 ? ctx->occ_override : sv.minb);
        const u64 NGv = (n + 31) / 32;
        g->grid = (uint32_t)std::min<u64>((u64)D.sm_count * occ, std::max<u64>(NGv, 1));
        return PQV_OK;
    }
    if (g->variant > 0) {
        const ScanVariant &sv = gather ? kGatherVariants[g->variant] : kVariants[g->variant];
        g->smem = (size_t)((dim + 3u) & ~3u) * 4 + (size_t)SCAN_WARPS * 32 * (32 * sv.cbv + 4) * 4 + (size_t)g->sort_n * 8;
        int occ = 0;
        pqv::ScanParams dummy{};
        if (gather) PQV_TRY(gather_variant_dispatch(g->variant, dummy, 0, g->smem, nullptr, &occ));
        else PQV_TRY(scan_variant_dispatch(g->variant, dummy, 0, g->smem, nullptr, &occ));
        if (occ < 1) return fail(PQV_ELIMIT, "scan variant does not fit");
        occ = std::min(occ, ctx->occ_override > 0 ? ctx->occ_override : sv.minb);
        const u64 NGv = (n + 31) / 32;
        g->grid = (uint32_t)std::min<u64>((u64)D.sm_count * occ, std::max<u64>(NGv, 1));
        return PQV_OK;
    }
    if (g->smem > 227 * 1024) return fail(PQV_ELIMIT, "scan needs %zu bytes of shared memory (dim=%u k=%u)", g->smem, dim, k);
    int occ = 0;
    PQV_TRY(SCAN_DISPATCH(scan_occupancy_t, order, g->vec4, gather, g->smem, &occ));
    if (occ < 1) return fail(PQV_ELIMIT, "scan kernel does not fit on an SM (smem %zu)", g->smem);
    int want = ctx->occ_override > 0 ? ctx->occ_override : 2;
    occ = std::min(occ, want);
    const u64 NG = (n + 31) / 32;
    g->grid = (uint32_t)std::min<u64>((u64)D.sm_count * occ, std::max<u64>(NG, 1));
    return PQV_OK;
}

// Enqueue scan + prefix merge + entrant filter on D.stream.  Device outputs: D.final_topk (kcap keys) or
// `final_out` if given, and entrant keys appended to `ent_out` ([0] = running count).
int enqueue_scan(pqv_ctx *ctx, DeviceState &D, const float *d_data, const uint32_t *d_row_ids, u64 n, uint32_t dim,
                 const float *d_query, uint32_t k, int order, uint32_t pos_base, const u64 *d_carry, u64 *final_out,
                 u64 *ent_out, uint32_t ent_out_cap, bool time_it, ScanGeom *geom_out, const u64 *n_dev = nullptr,
                 float *dist_out = nullptr) {
    ScanGeom g;
    PQV_TRY(scan_geometry(ctx, D, d_data, n, dim, k, order, d_row_ids != nullptr, &g));
    PQV_TRY(D.cta_topk.ensure((size_t)g.grid * g.kcap));
    PQV_TRY(D.ent.ensure((size_t)n + 64));
    PQV_TRY(D.ent_count.ensure(g.grid));
    PQV_TRY(D.gthr.ensure(g.grid));
    pqv::ScanParams p{};
    p.data = d_data;
    p.query = d_query;
    p.row_ids = d_row_ids;
    p.n = n;
    p.n_dev = n_dev;
    p.dim = dim;
    p.k = k;
    p.kcap = g.kcap;
    p.sort_n = g.sort_n;
    p.flush_at = g.flush_at;
    p.cap_bits = 0xFFFFFFFFu;
    p.carry = d_carry;
    p.pos_base = pos_base;
    p.cta_topk = D.cta_topk.p;
    p.ent = D.ent.p;
    p.ent_count = D.ent_count.p;
    p.dist_out = dist_out;
    if (time_it) CU_TRY(cudaEventRecord(D.ev[0], D.stream));
    if (g.seq_variant > 0) PQV_TRY(seq_gather_variant_dispatch(g.seq_variant, p, g.grid, g.smem, D.stream, nullptr));
    else if (g.variant > 0 && g.gather) PQV_TRY(gather_variant_dispatch(g.variant, p, g.grid, g.smem, D.stream, nullptr));
    else if (g.variant > 0) PQV_TRY(scan_variant_dispatch(g.variant, p, g.grid, g.smem, D.stream, nullptr));
    else PQV_TRY(SCAN_DISPATCH(scan_launch_t, order, g.vec4, d_row_ids != nullptr, p, g.grid, g.smem, D.stream));
    if (time_it) CU_TRY(cudaEventRecord(D.ev[1], D.stream));
    // two-level exclusive "top-k scan" over the per-CTA lists, then per-CTA threshold + entrant filter
    const uint32_t n_groups = (g.grid + pqv::MERGE_GROUP - 1) / pqv::MERGE_GROUP;
    PQV_TRY(D.within_prefix.ensure((size_t)g.grid * g.kcap));
    PQV_TRY(D.group_total.ensure((size_t)n_groups * g.kcap));
    PQV_TRY(D.group_prefix.ensure((size_t)n_groups * g.kcap));
    const size_t msmem = (size_t)2 * g.kcap * 8;
    pqv::topk_seq_merge_kernel<<<n_groups, g.kcap, msmem, D.stream>>>(D.cta_topk.p, g.grid, pqv::MERGE_GROUP, k, g.kcap,
                                                                      nullptr, D.within_prefix.p, D.group_total.p);
    CU_TRY(cudaGetLastError());
    pqv::topk_seq_merge_kernel<<<1, g.kcap, msmem, D.stream>>>(D.group_total.p, n_groups, n_groups, k, g.kcap, d_carry,
                                                               D.group_prefix.p, final_out);
    CU_TRY(cudaGetLastError());
    pqv::entrant_filter_kernel<<<g.grid, 256, msmem, D.stream>>>(D.ent.p, D.ent_count.p, D.group_prefix.p,
                                                                 D.within_prefix.p, k, g.kcap, D.gthr.p, n, n_dev,
                                                                 ent_out, ent_out_cap);
    CU_TRY(cudaGetLastError());
    if (time_it) CU_TRY(cudaEventRecord(D.ev[2], D.stream));
    if (geom_out) *geom_out = g;
    return PQV_OK;
}

// Fetch the entrant keys accumulated in ent_out (device) into `out` (appends).  Requires the stream
// to be idle on return.  Returns PQV_ELIMIT via *overflow when the device buffer was too small.
int fetch_entrants(DeviceState &D, u64 *d_ent_out, uint32_t ent_out_cap, std::vector<u64> &out, bool *overflow) {
    PQV_TRY(D.h_ent_out.ensure((size_t)ENT_FIRST_CHUNK + 1));
    const uint32_t first = std::min<uint32_t>(ENT_FIRST_CHUNK, ent_out_cap);
    CU_TRY(cudaMemcpyAsync(D.h_ent_out.p, d_ent_out, ((size_t)first + 1) * 8, cudaMemcpyDeviceToHost, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    const u64 count = D.h_ent_out.p[0];
    *overflow = count > ent_out_cap;
    if (*overflow) return PQV_OK;
    const size_t base = out.size();
    out.resize(base + count);
    const u64 got = std::min<u64>(count, first);
    memcpy(out.data() + base, D.h_ent_out.p + 1, got * 8);
    if (count > got) {
        CU_TRY(cudaMemcpyAsync(out.data() + base + got, d_ent_out + 1 + got, (count - got) * 8, cudaMemcpyDeviceToHost,
                               D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
    }
    return PQV_OK;
}

Dataset *find_dataset(pqv_ctx *ctx, u64 h) {
    auto it = ctx->datasets.find(h);
    return it == ctx->datasets.end() ? nullptr : &it->second;
}

// any_k: the entry point answers k > PQV_MAX_K as well (topk_one's full-replay path: every candidate's distance comes back
// and the reference loop runs on the host -- slow, but the reference accepts any k)
int check_topk_args(uint32_t k, uint32_t dim, uint32_t flags, bool any_k = false) {
    if (k == 0) return fail(PQV_EINVAL, "k must be > 0");  // src/ivf/search.rs:67
    if (k > PQV_MAX_K && !any_k) return fail(PQV_ELIMIT, "k = %u exceeds PQV_MAX_K = %u", k, PQV_MAX_K);
    if (dim == 0) return fail(PQV_EINVAL, "Embedding dimension must be > 0");  // src/ivf/mod.rs:61
    if (dim > PQV_MAX_DIM) return fail(PQV_ELIMIT, "dim = %u exceeds PQV_MAX_DIM = %u", dim, PQV_MAX_DIM);
    if (flags & ~(PQV_SUM_SEQ | PQV_SQRT | PQV_TIES_BY_POSITION)) return fail(PQV_EINVAL, "unknown flags 0x%x", flags);
    return PQV_OK;
}

// final (distance, position)-ordered keys -> outputs (PQV_TIES_BY_POSITION mode)
size_t emit_by_position(std::vector<u64> &keys, const RowMap &row_of, uint32_t k, uint32_t flags,
                        uint32_t *out_rows, float *out_dist) {
    std::sort(keys.begin(), keys.end());
    size_t n = 0;
    for (u64 key : keys) {
        if (key == pqv::KEY_MAX || n >= k) break;
        float d = key_dist(key);
        out_rows[n] = row_of(key_pos(key));
        out_dist[n] = (flags & PQV_SQRT) ? sqrtf(d) : d;
        ++n;
    }
    return n;
}

// One query over a resident dataset (all shards), brute force or gathered.  Host inputs/outputs.
// row_ids: host candidate list (gather API) or null.  d_cand: candidate list already on the device (IVF path;
// then row_fn maps positions to row ids on the host).  Neither => brute force over every resident row.
int topk_one(pqv_ctx *ctx, Dataset &ds, const float *query, const uint32_t *row_ids, u64 n_ids, uint32_t k,
             uint32_t flags, uint32_t *out_rows, float *out_dist, uint32_t *out_count,
             std::vector<u64> *entrants_out = nullptr, uint32_t pos_offset = 0, const uint32_t *d_cand = nullptr,
             const std::function<uint32_t(uint32_t)> *row_fn = nullptr, u64 limit_rows = 0) {
    const int order = (flags & PQV_SUM_SEQ) ? 1 : 0;
    const bool gather = row_ids != nullptr || d_cand != nullptr;
    // k above what the in-kernel selection holds: the scan runs with k = PQV_MAX_K only to produce the per-candidate distance
    // log, and the reference loop (any k) is replayed over every candidate on the host -- the path NaN distances take
    const bool big_k = k > PQV_MAX_K;
    const uint32_t k_user = k;
    if (big_k) k = PQV_MAX_K;
    static const bool trace = getenv("PQV_TRACE") != nullptr;
    double tt[4] = {0, 0, 0, 0};
    if (trace) tt[0] = trace_now_ms();
    RowMap row_of;
    row_of.ids = row_ids;
    row_of.fn = row_fn;
    std::vector<u64> entrants;
    std::vector<u64> finals;
    pqv_timing tm{};
    struct Launched {
        DeviceState *D;
        ScanGeom g;
        uint32_t cap;
        u64 n;
        const float *d_data;
        const uint32_t *d_ids;
        uint32_t pos_base;
        const std::vector<uint32_t> *gpos;  // multi-shard gather: position in the caller's candidate sequence of local candidate i
    };
    std::vector<Launched> launched;

    // A gathered search over a table spread over several devices: every shard scans the candidates it owns (local row
    // indices, in sequence order) and their keys are moved back to the positions of the caller's sequence, so the replay
    // below sees exactly what a single device would have produced.
    const bool split = gather && ds.shards.size() > 1;
    std::vector<std::vector<uint32_t>> loc, gpos;
    if (split) {
        if (d_cand) return fail(PQV_EINVAL, "device-resident candidate lists need a single-device dataset");
        loc.resize(ds.shards.size());
        gpos.resize(ds.shards.size());
        for (u64 i = 0; i < n_ids; ++i) {
            const uint32_t r = row_ids[i];
            size_t si = 0;
            while (si + 1 < ds.shards.size() && r >= ds.shards[si].first_row + ds.shards[si].n_rows) ++si;
            loc[si].push_back((uint32_t)(r - ds.shards[si].first_row));
            gpos[si].push_back((uint32_t)i);
        }
    }

    for (size_t si = 0; si < ds.shards.size(); ++si) {
        Shard &sh = ds.shards[si];
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        u64 n = 0;
        const uint32_t *d_ids = nullptr;
        if (split) {
            n = loc[si].size();
            if (n) {
                PQV_TRY(D.d_row_ids.ensure(n));
                CU_TRY(cudaMemcpyAsync(D.d_row_ids.p, loc[si].data(), n * 4, cudaMemcpyHostToDevice, D.stream));
                d_ids = D.d_row_ids.p;
            }
        } else if (gather) {
            n = n_ids;
            if (n && d_cand) {
                d_ids = d_cand;
            } else if (n) {
                PQV_TRY(D.d_row_ids.ensure(n));
                CU_TRY(cudaMemcpyAsync(D.d_row_ids.p, row_ids, n * 4, cudaMemcpyHostToDevice, D.stream));
                d_ids = D.d_row_ids.p;
            }
        } else {
            n = sh.n_rows;
            if (limit_rows) {  // brute force over the first limit_rows rows of the table only
                if (sh.first_row >= limit_rows) continue;
                n = std::min<u64>(n, limit_rows - sh.first_row);
            }
        }
        if (n == 0) continue;
        PQV_TRY(D.d_query.ensure(ds.dim));
        PQV_TRY(D.h_query.ensure(ds.dim));
        memcpy(D.h_query.p, query, (size_t)ds.dim * 4);
        CU_TRY(cudaMemcpyAsync(D.d_query.p, D.h_query.p, (size_t)ds.dim * 4, cudaMemcpyHostToDevice, D.stream));
        PQV_TRY(D.final_topk.ensure(PQV_MAX_K));
        PQV_TRY(D.ent_out.ensure((size_t)(1u << 16) + 1));
        const uint32_t cap = (uint32_t)std::min<size_t>(D.ent_out.cap - 1, 0xFFFFFFF0u);
        CU_TRY(cudaMemsetAsync(D.ent_out.p, 0, 8, D.stream));
        ScanGeom g;
        const uint32_t pos_base = gather ? 0u : (uint32_t)sh.first_row + pos_offset;
        if (big_k) PQV_TRY(D.d_dist.ensure(n));
        PQV_TRY(enqueue_scan(ctx, D, sh.d_data, d_ids, n, ds.dim, D.d_query.p, k, order, pos_base, nullptr, D.final_topk.p,
                             D.ent_out.p, cap, launched.empty(), &g, nullptr,  // the first launch carries the timing events
                             big_k ? D.d_dist.p : nullptr));
        launched.push_back({&D, g, cap, n, sh.d_data, d_ids, pos_base, split ? &gpos[si] : nullptr});
    }
    // keys of one launch: local candidate index -> position in the caller's sequence
    auto to_sequence = [](u64 *keys, size_t n_keys, const std::vector<uint32_t> *map) {
        if (!map) return;
        for (size_t i = 0; i < n_keys; ++i)
            if (keys[i] != pqv::KEY_MAX) keys[i] = (keys[i] & 0xFFFFFFFF00000000ull) | (*map)[key_pos(keys[i])];
    };
    if (trace) tt[1] = trace_now_ms();
    // collect
    for (auto &L : launched) {
        DeviceState &D = *L.D;
        DevGuard guard(D.dev);
        if (big_k) {
            CU_TRY(cudaStreamSynchronize(D.stream));  // the distance log is read below
        } else if (flags & PQV_TIES_BY_POSITION) {
            PQV_TRY(D.h_final.ensure(PQV_MAX_K));
            CU_TRY(cudaMemcpyAsync(D.h_final.p, D.final_topk.p, (size_t)L.g.kcap * 8, cudaMemcpyDeviceToHost, D.stream));
            CU_TRY(cudaStreamSynchronize(D.stream));
            to_sequence(D.h_final.p, L.g.kcap, L.gpos);
            finals.insert(finals.end(), D.h_final.p, D.h_final.p + L.g.kcap);
        } else {
            const size_t e0 = entrants.size();
            bool overflow = false;
            PQV_TRY(fetch_entrants(D, D.ent_out.p, L.cap, entrants, &overflow));
            if (overflow) {
                // adversarial ordering (most rows enter the heap): rerun the filter with a buffer that can
                // hold every row; the scan outputs are still valid in D.ent / D.gthr.
                const u64 need = D.h_ent_out.p[0];
                PQV_TRY(D.ent_out.ensure((size_t)need + 1));
                const uint32_t cap2 = (uint32_t)std::min<size_t>(D.ent_out.cap - 1, 0xFFFFFFF0u);
                CU_TRY(cudaMemsetAsync(D.ent_out.p, 0, 8, D.stream));
                pqv::entrant_filter_kernel<<<L.g.grid, 256, (size_t)2 * L.g.kcap * 8, D.stream>>>(
                    D.ent.p, D.ent_count.p, D.group_prefix.p, D.within_prefix.p, k, L.g.kcap, D.gthr.p, L.n, nullptr, D.ent_out.p,
                    cap2);
                CU_TRY(cudaGetLastError());
                PQV_TRY(fetch_entrants(D, D.ent_out.p, cap2, entrants, &overflow));
                if (overflow) return fail(PQV_ECUDA, "entrant buffer overflow after regrow");
            }
            to_sequence(entrants.data() + e0, entrants.size() - e0, L.gpos);
        }
    }
    if (trace) tt[2] = trace_now_ms();
    if (!launched.empty()) {
        DeviceState &D = *launched[0].D;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, D.ev[0], D.ev[1]);
        cudaEventElapsedTime(&b, D.ev[1], D.ev[2]);
        tm.scan_ms = a;
        tm.post_ms = b;
        tm.total_ms = a + b;
        tm.scan_bytes = launched[0].n * (u64)ds.dim * 4;
        tm.launches = 4 * (uint32_t)launched.size();
        tm.grid = launched[0].g.grid;
    }
    size_t cnt;
    // A NaN distance among the entrants: the reference heap no longer behaves like a threshold (src/ivf/search.rs:119-126
    // with `partial_cmp -> Equal`), so nothing short of its own loop over EVERY candidate, in order, reproduces it.  The scan is
    // repeated with a per-candidate distance log and the loop is replayed literally on the host (or, for a rank's half of a
    // sharded search, every row of the slice becomes a candidate).
    if (big_k || (!(flags & PQV_TIES_BY_POSITION) && any_nan_key(entrants.data(), entrants.size()))) {
        std::vector<float> dist;
        std::vector<uint32_t> poses;
        for (auto &L : launched) {
            DeviceState &D = *L.D;
            DevGuard guard(D.dev);
            if (!big_k) {  // (big k: the first scan already wrote the log)
                PQV_TRY(D.d_dist.ensure(L.n));
                CU_TRY(cudaMemsetAsync(D.ent_out.p, 0, 8, D.stream));
                PQV_TRY(enqueue_scan(ctx, D, L.d_data, L.d_ids, L.n, ds.dim, D.d_query.p, k, order, L.pos_base, nullptr, D.final_topk.p,
                                     D.ent_out.p, L.cap, false, nullptr, nullptr, D.d_dist.p));
            }
            const size_t base = dist.size();
            dist.resize(base + L.n);
            CU_TRY(cudaMemcpyAsync(dist.data() + base, D.d_dist.p, L.n * 4, cudaMemcpyDeviceToHost, D.stream));
            CU_TRY(cudaStreamSynchronize(D.stream));
            for (u64 i = 0; i < L.n; ++i) poses.push_back(L.gpos ? (*L.gpos)[i] : L.pos_base + (uint32_t)i);
        }
        if (split) {  // the reference loop walks the caller's sequence: back into that order
            std::vector<u64> ord(poses.size());
            for (size_t i = 0; i < poses.size(); ++i) ord[i] = ((u64)poses[i] << 32) | i;
            std::sort(ord.begin(), ord.end());
            std::vector<float> d2(dist.size());
            for (size_t i = 0; i < ord.size(); ++i) {
                d2[i] = dist[(uint32_t)ord[i]];
                poses[i] = (uint32_t)(ord[i] >> 32);
            }
            dist.swap(d2);
        }
        tm.entrants = (uint32_t)dist.size();
        ctx->last = tm;
        if (entrants_out) {
            entrants_out->resize(dist.size());
            for (size_t i = 0; i < dist.size(); ++i) {
                uint32_t b;
                memcpy(&b, &dist[i], 4);
                (*entrants_out)[i] = ((u64)b << 32) | poses[i];
            }
            return PQV_OK;
        }
        if (flags & PQV_TIES_BY_POSITION) {  // (big k only) the k smallest by (distance, position)
            std::vector<u64> keys(dist.size());
            for (size_t i = 0; i < dist.size(); ++i) {
                uint32_t b;
                memcpy(&b, &dist[i], 4);
                keys[i] = ((u64)b << 32) | poses[i];
            }
            *out_count = (uint32_t)emit_by_position(keys, row_of, k_user, flags, out_rows, out_dist);
            return PQV_OK;
        }
        *out_count = (uint32_t)replay_ordered(
            dist.size(), [&](size_t i) { return ReplayItem{dist[i], row_of(poses[i])}; }, k_user, flags, out_rows, out_dist);
        return PQV_OK;
    }
    if (entrants_out) {
        tm.entrants = (uint32_t)entrants.size();
        entrants_out->swap(entrants);
        ctx->last = tm;
        return PQV_OK;
    }
    if (flags & PQV_TIES_BY_POSITION) {
        cnt = emit_by_position(finals, row_of, k, flags, out_rows, out_dist);
    } else {
        tm.entrants = (uint32_t)entrants.size();
        cnt = replay_reference_heap(entrants, row_of, k, flags, out_rows, out_dist);
    }
    *out_count = (uint32_t)cnt;
    ctx->last = tm;
    if (trace) {
        tt[3] = trace_now_ms();
        fprintf(stderr, "[pqv trace] topk_one: enqueue %.1f us, wait+fetch %.1f, replay %.1f (kernels %.1f)\n",
                (tt[1] - tt[0]) * 1e3, (tt[2] - tt[1]) * 1e3, (tt[3] - tt[2]) * 1e3, tm.total_ms * 1e3);
    }
    return PQV_OK;
}

}  // namespace

#include "pqv_kmeanspp.cuh"
#include "pqv_tc_host.cuh"

static void pqv_free_all_indexes(pqv_ctx *ctx);

// Rows `ids` (global, any order) of a table spread over several devices, as one dense n x dim block on device D0 (d_out):
// every shard gathers the rows it owns on its own device, the pieces travel to D0 (peer copy) and are put in the caller's
// order there.  The k-means sample of pqv_ivf_build (src/ivf/index.rs:222-242) over a multi-device table.
static int gather_rows_multi(pqv_ctx *ctx, Dataset &ds, const uint32_t *ids, u64 n_ids, DeviceState &D0, float *d_out) {
    const uint32_t dim = ds.dim;
    const size_t ns = ds.shards.size();
    std::vector<std::vector<uint32_t>> loc(ns);
    std::vector<uint32_t> inv(n_ids);  // d_out[j] = staging[inv[j]]
    std::vector<u64> cnt(ns, 0), off(ns + 1, 0);
    std::vector<uint8_t> owner(n_ids);
    for (u64 i = 0; i < n_ids; ++i) {
        size_t si = 0;
        while (si + 1 < ns && ids[i] >= ds.shards[si].first_row + ds.shards[si].n_rows) ++si;
        owner[i] = (uint8_t)si;
        ++cnt[si];
    }
    for (size_t si = 0; si < ns; ++si) off[si + 1] = off[si] + cnt[si];
    for (size_t si = 0; si < ns; ++si) loc[si].reserve(cnt[si]);
    for (u64 i = 0; i < n_ids; ++i) {
        const size_t si = owner[i];
        inv[i] = (uint32_t)(off[si] + loc[si].size());
        loc[si].push_back((uint32_t)(ids[i] - ds.shards[si].first_row));
    }
    DevBuf<float> staging;
    DevBuf<uint32_t> d_inv;
    {
        DevGuard g0(D0.dev);
        PQV_TRY(staging.ensure((size_t)n_ids * dim));
        PQV_TRY(d_inv.ensure(n_ids));
    }
    int rc = PQV_OK;
    for (size_t si = 0; si < ns && rc == PQV_OK; ++si) {
        if (!cnt[si]) continue;
        Shard &sh = ds.shards[si];
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        rc = D.d_row_ids.ensure(cnt[si]);
        if (!rc) rc = D.d_tmp_rows.ensure((size_t)cnt[si] * dim);
        if (rc) break;
        cudaError_t e = cudaMemcpyAsync(D.d_row_ids.p, loc[si].data(), cnt[si] * 4, cudaMemcpyHostToDevice, D.stream);
        if (e == cudaSuccess) {
            pqv::gather_rows_kernel<<<D.sm_count * 8, 256, 0, D.stream>>>(sh.d_data, D.d_row_ids.p, cnt[si], dim, D.d_tmp_rows.p);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess)
            e = cudaMemcpyPeerAsync(staging.p + off[si] * dim, D0.dev, D.d_tmp_rows.p, D.dev, (size_t)cnt[si] * dim * 4, D.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
        if (e != cudaSuccess) rc = fail(PQV_ECUDA, "gathering rows of shard %zu failed: %s", si, cudaGetErrorString(e));
    }
    if (rc == PQV_OK) {
        DevGuard g0(D0.dev);
        cudaError_t e = cudaMemcpyAsync(d_inv.p, inv.data(), n_ids * 4, cudaMemcpyHostToDevice, D0.stream);
        if (e == cudaSuccess) {
            pqv::gather_rows_kernel<<<D0.sm_count * 8, 256, 0, D0.stream>>>(staging.p, d_inv.p, n_ids, dim, d_out);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(D0.stream);
        if (e != cudaSuccess) rc = fail(PQV_ECUDA, "placing the gathered rows failed: %s", cudaGetErrorString(e));
    }
    {
        DevGuard g0(D0.dev);
        staging.release();
        d_inv.release();
    }
    return rc;
}

// A batch over a table spread over several devices of ONE context (one process driving all GPUs): every shard answers the
// whole batch over its own rows in one tensor-core pass (batch_topk in raw mode: the k + 1 smallest exact keys per query,
// positions = global rows), each from its own host thread so the passes overlap, and the per-shard lists are merged as the
// per-rank lists of the process-per-GPU form are (pqv_merge_batch_keys).  Queries the merge cannot decide (exact ties, a
// shard that declined) stay unhandled: the caller runs them through the multi-shard single-query scan.
static int batch_topk_sharded(pqv_ctx *ctx, Dataset &ds, const float *queries, uint32_t n_queries, uint32_t k, uint32_t flags,
                              uint32_t *out_row_idx, float *out_dist, uint32_t *out_count, std::vector<uint8_t> &handled) {
    const size_t ns = ds.shards.size();
    const size_t kp = (size_t)k + 1;
    pqv_batch_timing total{};
    for (uint32_t q0 = 0; q0 < n_queries; q0 += BATCH_MAX_QUERIES) {
        const uint32_t nq = std::min(BATCH_MAX_QUERIES, n_queries - q0);
        if (nq < BATCH_MIN_QUERIES) break;
        std::vector<u64> keys(ns * nq * kp, pqv::KEY_MAX);
        std::vector<uint32_t> counts(ns * nq, 0xFFFFFFFFu);
        std::vector<pqv_batch_timing> bts(ns);
        std::vector<int> rcs(ns, PQV_OK);
        std::vector<std::string> errs(ns);
        auto run = [&](size_t si) {
            Shard &sh = ds.shards[si];
            if (sh.n_rows == 0) {
                for (uint32_t q = 0; q < nq; ++q) counts[si * nq + q] = 0;
                return;
            }
            DeviceState &D = ctx->devs[sh.di];
            DevGuard guard(D.dev);
            std::vector<uint8_t> part;
            if ((reinterpret_cast<uintptr_t>(sh.d_data) & 15) != 0) return;  // counts stay "undecided"
            rcs[si] = batch_topk(ctx, D, ds, sh.n_rows, ds.dim, queries + (size_t)q0 * ds.dim, nq, k, flags, nullptr, nullptr, nullptr,
                                 part, keys.data() + si * nq * kp, counts.data() + si * nq, (uint32_t)sh.first_row, nullptr, &sh,
                                 &bts[si]);
            if (rcs[si] != PQV_OK) errs[si] = g_err;
        };
        std::vector<std::thread> th;
        for (size_t si = 1; si < ns; ++si) th.emplace_back(run, si);
        run(0);
        for (auto &t : th) t.join();
        for (size_t si = 0; si < ns; ++si)
            if (rcs[si] != PQV_OK) {
                g_err = errs[si];
                return rcs[si];
            }
        std::vector<uint8_t> need(nq, 0);
        PQV_TRY(pqv_merge_batch_keys(reinterpret_cast<const uint64_t *>(keys.data()), counts.data(), (uint32_t)ns, nq, k, flags, out_row_idx + (size_t)q0 * k,
                                     out_dist + (size_t)q0 * k, out_count + q0, need.data()));
        for (uint32_t q = 0; q < nq; ++q) handled[q0 + q] = need[q] ? 0 : 1;
        total.queries += nq;
        total.rows = ds.n_rows;
        for (size_t si = 0; si < ns; ++si) {  // device times: the slowest shard of each phase (the shards run side by side)
            const pqv_batch_timing &b = bts[si];
            total.declined |= b.declined;
            total.candidates += b.candidates;
            total.sample_rows = std::max(total.sample_rows, b.sample_rows);
        }
        double pm = 0, sm = 0, fm = 0, rm = 0, tmx = 0;
        for (const auto &b : bts) {
            pm = std::max(pm, b.prep_ms), sm = std::max(sm, b.sample_ms), fm = std::max(fm, b.filter_ms);
            rm = std::max(rm, b.rerank_ms), tmx = std::max(tmx, b.total_ms);
        }
        total.prep_ms += pm, total.sample_ms += sm, total.filter_ms += fm, total.rerank_ms += rm, total.total_ms += tmx;
        for (uint32_t q = 0; q < nq; ++q) total.tie_queries += need[q] ? 1u : 0u;
    }
    ctx->last_batch = total;
    return PQV_OK;
}

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *pqv_last_error(void) { return g_err.c_str(); }
const char *pqv_version(void) { return "pq-vector-b200 0.3.0 (sm_100a)"; }

int pqv_init(pqv_ctx **out, const int *device_ids, int n_devices) {
    if (!out) return fail(PQV_EINVAL, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PQV_ENODEV, "no CUDA device available (%s); libpqv has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> ids;
    if (device_ids == nullptr || n_devices == 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        ids.push_back(cur);
    } else {
        for (int i = 0; i < n_devices; ++i) {
            if (device_ids[i] < 0 || device_ids[i] >= count) return fail(PQV_ENODEV, "device id %d out of range [0,%d)", device_ids[i], count);
            ids.push_back(device_ids[i]);
        }
    }
    pqv_ctx *ctx = new pqv_ctx();
    if (const char *s = getenv("PQV_SCAN_CTAS_PER_SM")) ctx->occ_override = atoi(s);
    if (const char *s = getenv("PQV_SCAN_VARIANT")) ctx->scan_variant = std::max(0, std::min(12, atoi(s)));
    if (const char *s = getenv("PQV_GATHER_VARIANT")) ctx->gather_variant = std::max(0, std::min(6, atoi(s)));
    if (const char *s = getenv("PQV_SEQ_GATHER_VARIANT")) ctx->seq_gather_variant = std::max(0, std::min(5, atoi(s)));
    for (int id : ids) {
        DeviceState D;
        D.dev = id;
        DevGuard guard(id);
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10) {
            delete ctx;
            return fail(PQV_ENODEV, "device %d is sm_%d%d; libpqv is built for sm_100a only", id, prop.major, prop.minor);
        }
        D.sm_count = prop.multiProcessorCount;
        CU_TRY(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&D.copy_stream, cudaStreamNonBlocking));
        for (auto &ev : D.ev) CU_TRY(cudaEventCreate(&ev));
        for (auto &ev : D.ev_shadow) CU_TRY(cudaEventCreate(&ev));
        ctx->devs.push_back(D);
    }
    *out = ctx;
    return PQV_OK;
}

static void peer_release(pqv_ctx *ctx) {
    PeerExchange &px = ctx->peer;
    if (!px.local && px.peers.empty()) return;
    DevGuard guard(ctx->devs[0].dev);
    cudaStreamSynchronize(ctx->devs[0].stream);
    for (uint32_t d = 0; d < px.peers.size(); ++d)
        if (px.peers[d] && px.peers[d] != px.local) cudaIpcCloseMemHandle(px.peers[d]);
    if (px.local) cudaFree(px.local);
    if (px.d_peers) cudaFree(px.d_peers);
    if (px.d_timeout) cudaFree(px.d_timeout);
    px.h_block.release();
    px = PeerExchange{};
}

void pqv_destroy(pqv_ctx *ctx) {
    if (!ctx) return;
    peer_release(ctx);
    for (auto &kv : ctx->streams) {
        DevGuard guard(ctx->devs[0].dev);
        stream_state_free(kv.second);
    }
    if (ctx->stream_cache) {
        DevGuard guard(ctx->devs[0].dev);
        stream_state_free(ctx->stream_cache);
        ctx->stream_cache = nullptr;
    }
    pqv_free_all_indexes(ctx);
    for (auto &kv : ctx->datasets)
        for (auto &sh : kv.second.shards) {
            DevGuard guard(ctx->devs[sh.di].dev);
            if (sh.d_data) cudaFree(sh.d_data);
            sh.drop_derived();
        }
    for (auto &D : ctx->devs) {
        DevGuard guard(D.dev);
        cudaStreamSynchronize(D.stream);
        D.ent_rows.release();
        D.ivf_info.release();
        D.csr_counts.release();
        D.csr_totals.release();
        D.h_ent_rows.release();
        D.h_ivf_info.release();
        if (D.pool) {
            D.pool->shutdown();
            delete D.pool;
            D.pool = nullptr;
        }
        for (auto &L : D.lanes) {
            if (L.st) cudaStreamDestroy(L.st);
            for (auto &e : L.ev)
                if (e) cudaEventDestroy(e);
            L.buf[0].release();
            L.buf[1].release();
        }
        D.lanes.clear();
        D.vt_bitmap.release();
        D.vt_mask.release();
        D.vt_sums.release();
        D.ad_query.release();
        D.ad_col.release();
        D.ad_out_dist.release();
        D.ad_out_row.release();
        D.ad_state.release();
        D.h_ad_dist.release();
        D.h_ad_row.release();
        D.h_ad_state.release();
        D.d_query.release();
        D.cta_topk.release();
        D.ent.release();
        D.final_topk.release();
        D.ent_out.release();
        D.within_prefix.release();
        D.group_total.release();
        D.group_prefix.release();
        D.ent_count.release();
        D.gthr.release();
        D.d_row_ids.release();
        D.d_assign.release();
        D.d_vec.release();
        D.d_tmp_rows.release();
        D.d_centroids.release();
        D.d_dist.release();
        D.tc_wc.release();
        D.ts_half.release();
        D.ts_stats.release();
        D.ts_mu.release();
        D.ts_g.release();
        for (auto &e : D.ev_shadow)
            if (e) cudaEventDestroy(e);
        D.tc_bp.release();
        D.tc_mu.release();
        D.tc_cn.release();
        D.tc_stats.release();
        D.tc_pairs.release();
        D.tc_best.release();
        D.tc_u32.release();
        D.tc_amb_rows.release();
        D.tc_ovf_rows.release();
        D.tc_perm.release();
        D.ts_mean_part.release();
        D.tc_okeys.release();
        D.tb_Q.release();
        D.tb_Qp.release();
        D.tb_qf.release();
        D.tb_U.release();
        D.tb_u32.release();
        D.tb_info.release();
        D.tb_cand.release();
        D.tb_seg.release();
        D.tb_keys.release();
        D.h_batch_keys.release();
        D.tie_dmat.release();
        D.tie_qsel.release();
        D.tie_ent.release();
        D.h_tie_ent.release();
        D.h_tie_seg.release();
        D.h_tie_qsel.release();
        D.h_ent_out.release();
        D.h_final.release();
        D.h_query.release();
        for (auto &ev : D.ev)
            if (ev) cudaEventDestroy(ev);
        cudaStreamDestroy(D.stream);
        cudaStreamDestroy(D.copy_stream);
    }
    delete ctx;
}

int pqv_device_count(pqv_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

// ---- datasets ----------------------------------------------------------------------------------
int pqv_dataset_create(pqv_ctx *ctx, uint32_t dim, uint64_t n_rows_hint, uint64_t *out_handle) {
    if (!ctx || !out_handle) return fail(PQV_EINVAL, "null argument");
    if (dim == 0) return fail(PQV_EINVAL, "Embedding dimension must be > 0");
    if (n_rows_hint > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row ids are u32: %llu rows do not fit", (unsigned long long)n_rows_hint);
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset ds;
    ds.dim = dim;
    const size_t nd = ctx->devs.size();
    const u64 hint = std::max<u64>(n_rows_hint, 1);
    const u64 per = (hint + nd - 1) / nd;
    for (size_t i = 0; i < nd; ++i) {
        Shard sh;
        sh.di = (int)i;
        sh.cap_rows = per;
        sh.first_row = per * i;
        DevGuard guard(ctx->devs[i].dev);
        cudaError_t e = cudaMalloc((void **)&sh.d_data, (size_t)per * dim * sizeof(float));
        if (e != cudaSuccess) {
            for (auto &s2 : ds.shards) cudaFree(s2.d_data);
            return fail(PQV_ENOMEM, "cudaMalloc of %.2f GB for a dataset shard failed: %s", per * (double)dim * 4 / 1e9,
                        cudaGetErrorString(e));
        }
        ds.shards.push_back(sh);
    }
    const u64 h = ctx->next_handle++;
    ctx->datasets[h] = ds;
    *out_handle = h;
    return PQV_OK;
}

// grow the last shard (single-device datasets only) so appends past the hint keep working
static int grow_shard(pqv_ctx *ctx, Dataset &ds, Shard &sh, u64 need_rows) {
    DeviceState &D = ctx->devs[sh.di];
    DevGuard guard(D.dev);
    u64 new_cap = std::max<u64>(need_rows, sh.cap_rows + sh.cap_rows / 2 + 1024);
    float *nd = nullptr;
    cudaError_t e = cudaMalloc((void **)&nd, (size_t)new_cap * ds.dim * sizeof(float));
    if (e != cudaSuccess) return fail(PQV_ENOMEM, "growing a dataset shard to %llu rows failed: %s", (unsigned long long)new_cap, cudaGetErrorString(e));
    CU_TRY(cudaMemcpyAsync(nd, sh.d_data, (size_t)sh.n_rows * ds.dim * 4, cudaMemcpyDeviceToDevice, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    cudaFree(sh.d_data);
    sh.d_data = nd;
    sh.cap_rows = new_cap;
    sh.shadow.drop();  // sized for the old allocation; rebuilt on the next tensor-core pass
    return PQV_OK;
}

// Host -> HBM copy of a large PAGEABLE buffer.  A plain cudaMemcpy of pageable memory is staged by the driver through one
// bounce buffer (11-16 GB/s measured on this box); here T host threads each run their own double-buffered lane -- copy an
// 8 MB chunk into page-locked staging, cudaMemcpyAsync it, fill the other buffer while that DMA runs -- so the CPU copies
// proceed in parallel and the copy engine always has a transfer queued (benchmarks/probe_append.py).  Page-locked
// sources (cudaHostRegister'ed Arrow buffers) skip the staging: the DMA reads them directly.

static int append_staged(DeviceState &D, unsigned char *dst, const unsigned char *src, size_t bytes) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t lanes = std::min<size_t>(APPEND_MAX_LANES, std::max<size_t>(1, hw / 2));
    // a record batch (a few MB) is cut into one piece per lane so that the host copies still run in parallel
    const size_t chunk = std::min(APPEND_CHUNK, std::max(APPEND_CHUNK_MIN, (bytes / lanes + 4095) & ~(size_t)4095));
    const size_t n_chunks = (bytes + chunk - 1) / chunk;
    if (D.lanes.size() < lanes) D.lanes.resize(lanes);
    for (size_t t = 0; t < lanes; ++t) {
        DeviceState::AppendLane &L = D.lanes[t];
        if (!L.st) CU_TRY(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
        for (auto &e : L.ev)
            if (!e) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PQV_TRY(L.buf[0].ensure(APPEND_CHUNK));
        PQV_TRY(L.buf[1].ensure(APPEND_CHUNK));
    }
    if (!D.pool) {
        D.pool = new AppendPool();
        for (size_t t = 0; t < lanes; ++t) D.pool->th.emplace_back(&AppendPool::worker, D.pool, t, D.dev);
    }
    AppendPool &P = *D.pool;
    const size_t T = std::min(lanes, n_chunks);
    {
        std::unique_lock<std::mutex> lk(P.m);
        P.D = &D;
        P.dst = dst;
        P.src = src;
        P.bytes = bytes;
        P.n_chunks = n_chunks;
        P.chunk = chunk;
        P.T = T;
        P.pending = T;
        for (auto &e : P.errs) e = cudaSuccess;
        P.job_id++;
        P.cv_job.notify_all();
        P.cv_done.wait(lk, [&] { return P.pending == 0; });
    }
    for (size_t t = 0; t < T; ++t)
        if (P.errs[t] != cudaSuccess) return fail(PQV_ECUDA, "staged host->device copy failed: %s", cudaGetErrorString(P.errs[t]));
    return PQV_OK;
}

int pqv_dataset_append(pqv_ctx *ctx, uint64_t handle, const float *values, uint64_t n_rows) {
    if (ctx) ctx->batch_state.valid = false;
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (n_rows == 0) return PQV_OK;
    if (!values) return fail(PQV_EINVAL, "values is null");
    if (ds->n_rows + n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row ids are u32");
    u64 done = 0;
    while (done < n_rows) {
        // the shard that owns global row ds->n_rows
        Shard *sh = nullptr;
        for (auto &s : ds->shards)
            if (s.n_rows < s.cap_rows && s.first_row + s.n_rows == ds->n_rows) {
                sh = &s;
                break;
            }
        if (!sh) {
            Shard &last = ds->shards.back();
            if (last.first_row + last.n_rows != ds->n_rows) return fail(PQV_ECUDA, "dataset shard bookkeeping is inconsistent");
            PQV_TRY(grow_shard(ctx, *ds, last, last.n_rows + (n_rows - done)));
            sh = &last;
        }
        const u64 take = std::min<u64>(n_rows - done, sh->cap_rows - sh->n_rows);
        DeviceState &D = ctx->devs[sh->di];
        DevGuard guard(D.dev);
        const size_t bytes = (size_t)take * ds->dim * 4;
        bool staged = false;
        if (bytes >= APPEND_STAGED_MIN && getenv("PQV_APPEND_DIRECT") == nullptr) {
            cudaPointerAttributes pa{};
            const bool pinned = cudaPointerGetAttributes(&pa, values) == cudaSuccess && pa.type != cudaMemoryTypeUnregistered;
            cudaGetLastError();  // an unregistered pointer may leave a sticky-less error code behind on old drivers
            if (!pinned) {
                CU_TRY(cudaStreamSynchronize(D.stream));
                PQV_TRY(append_staged(D, reinterpret_cast<unsigned char *>(sh->d_data + sh->n_rows * ds->dim),
                                      reinterpret_cast<const unsigned char *>(values + done * ds->dim), bytes));
                staged = true;
            }
        }
        if (!staged) {
            CU_TRY(cudaMemcpyAsync(sh->d_data + sh->n_rows * ds->dim, values + done * ds->dim, bytes, cudaMemcpyHostToDevice,
                                   D.stream));
            CU_TRY(cudaStreamSynchronize(D.stream));  // values is only borrowed for the call
        }
        sh->n_rows += take;
        sh->norms_rows = 0;  // cached norms no longer cover the shard
        ds->n_rows += take;
        done += take;
    }
    return PQV_OK;
}

int pqv_dataset_rows(pqv_ctx *ctx, uint64_t handle, uint64_t *out_rows, uint32_t *out_dim) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (out_rows) *out_rows = ds->n_rows;
    if (out_dim) *out_dim = ds->dim;
    return PQV_OK;
}

int pqv_dataset_drop(pqv_ctx *ctx, uint64_t handle) {
    if (ctx) ctx->batch_state.valid = false;
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    for (auto &sh : ds->shards) {
        DevGuard guard(ctx->devs[sh.di].dev);
        cudaStreamSynchronize(ctx->devs[sh.di].stream);
        if (sh.d_data) cudaFree(sh.d_data);
        sh.drop_derived();
    }
    ctx->datasets.erase(handle);
    {
        std::lock_guard<std::mutex> lk2(ctx->co.m);
        ctx->co.dims.erase(handle);
    }
    return PQV_OK;
}

int pqv_dataset_fill_synthetic(pqv_ctx *ctx, uint64_t handle, uint64_t n_rows, uint64_t seed, uint64_t stream_first_row) {
    if (ctx) ctx->batch_state.valid = false;
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    u64 total_cap = 0;
    for (auto &sh : ds->shards) total_cap += sh.cap_rows;
    if (n_rows > total_cap) return fail(PQV_EINVAL, "n_rows %llu exceeds the dataset capacity %llu", (unsigned long long)n_rows, (unsigned long long)total_cap);
    u64 left = n_rows;
    ds->n_rows = 0;
    for (auto &sh : ds->shards) {
        const u64 take = std::min<u64>(left, sh.cap_rows);
        sh.n_rows = take;
        sh.norms_rows = 0;
        sh.shadow.rows = 0;  // rows rewritten in place: the 16-bit shadow is rebuilt on the next tensor-core pass
        left -= take;
        ds->n_rows += take;
        if (!take) continue;
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        pqv::synth_fill_kernel<<<D.sm_count * 8, 256, 0, D.stream>>>(sh.d_data, (stream_first_row + sh.first_row) * ds->dim, take * ds->dim, seed);
        CU_TRY(cudaGetLastError());
    }
    for (auto &sh : ds->shards) {
        DevGuard guard(ctx->devs[sh.di].dev);
        CU_TRY(cudaStreamSynchronize(ctx->devs[sh.di].stream));
    }
    return PQV_OK;
}

int pqv_dataset_read(pqv_ctx *ctx, uint64_t handle, uint64_t first_row, uint64_t n_rows, float *out) {
    if (!ctx || (!out && n_rows)) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (first_row + n_rows > ds->n_rows) return fail(PQV_EINVAL, "rows [%llu,%llu) out of range (%llu rows)", (unsigned long long)first_row, (unsigned long long)(first_row + n_rows), (unsigned long long)ds->n_rows);
    for (auto &sh : ds->shards) {
        const u64 lo = std::max<u64>(first_row, sh.first_row), hi = std::min<u64>(first_row + n_rows, sh.first_row + sh.n_rows);
        if (lo >= hi) continue;
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        CU_TRY(cudaMemcpyAsync(out + (lo - first_row) * ds->dim, sh.d_data + (lo - sh.first_row) * ds->dim,
                               (size_t)(hi - lo) * ds->dim * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
    }
    return PQV_OK;
}

int pqv_dataset_read_rows(pqv_ctx *ctx, uint64_t handle, const uint32_t *row_ids, uint64_t n_ids, float *out) {
    if (!ctx || (n_ids && (!row_ids || !out))) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (n_ids == 0) return PQV_OK;
    for (u64 i = 0; i < n_ids; ++i)
        if (row_ids[i] >= ds->n_rows) return fail(PQV_EINVAL, "row %u out of range (%llu rows)", row_ids[i], (unsigned long long)ds->n_rows);
    Shard &sh = ds->shards[0];
    DeviceState &D = ctx->devs[sh.di];
    DevGuard guard(D.dev);
    PQV_TRY(D.d_row_ids.ensure(n_ids));
    PQV_TRY(D.d_tmp_rows.ensure((size_t)n_ids * ds->dim));
    if (ds->shards.size() > 1) {  // rows spread over several devices: collected on the first one
        PQV_TRY(gather_rows_multi(ctx, *ds, row_ids, n_ids, D, D.d_tmp_rows.p));
        CU_TRY(cudaMemcpyAsync(out, D.d_tmp_rows.p, (size_t)n_ids * ds->dim * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
        return PQV_OK;
    }
    CU_TRY(cudaMemcpyAsync(D.d_row_ids.p, row_ids, n_ids * 4, cudaMemcpyHostToDevice, D.stream));
    pqv::gather_rows_kernel<<<D.sm_count * 8, 256, 0, D.stream>>>(sh.d_data, D.d_row_ids.p, n_ids, ds->dim, D.d_tmp_rows.p);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, D.d_tmp_rows.p, (size_t)n_ids * ds->dim * 4, cudaMemcpyDeviceToHost, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    return PQV_OK;
}

// ---- top-k -------------------------------------------------------------------------------------
int pqv_l2_topk(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k, uint32_t flags,
                uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    if (n_queries && (!queries || !out_row_idx || !out_dist || !out_count)) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    PQV_TRY(check_topk_args(k, ds->dim, flags, true));
    const bool big_k = k > PQV_MAX_K;  // answered query by query through topk_one's full replay
    // several queries: one tensor-core pass over the table answers every query whose result does not hinge on the
    // reference heap's layout (pqv_tc.cuh); the rest -- and small batches -- take the single-query scan
    std::vector<uint8_t> handled(n_queries, 0);
    ctx->last_batch = pqv_batch_timing{};
    ctx->batch_state.valid = false;
    if (!big_k && n_queries && ds->n_rows && batch_path_applies(*ds, ds->shards[0].d_data, n_queries, k)) {
        DeviceState &D = ctx->devs[ds->shards[0].di];
        DevGuard guard(D.dev);
        // at most BATCH_MAX_QUERIES queries per pass over the table (the pass' scratch grows with the batch)
        pqv_batch_timing total{};
        std::vector<uint8_t> part;
        for (uint32_t q0 = 0; q0 < n_queries; q0 += BATCH_MAX_QUERIES) {
            const uint32_t nq = std::min(BATCH_MAX_QUERIES, n_queries - q0);
            if (nq < BATCH_MIN_QUERIES) break;  // a short tail takes the single-query scans
            PQV_TRY(batch_topk(ctx, D, *ds, ds->n_rows, ds->dim, queries + (size_t)q0 * ds->dim, nq, k, flags,
                               out_row_idx + (size_t)q0 * k, out_dist + (size_t)q0 * k, out_count + q0, part));
            for (uint32_t i = 0; i < nq; ++i) handled[q0 + i] = part[i];
            const pqv_batch_timing &b = ctx->last_batch;
            total.queries += b.queries;
            total.declined |= b.declined;
            total.tie_queries += b.tie_queries;
            total.tie_batched += b.tie_batched;
            total.rows = b.rows;
            total.sample_rows = b.sample_rows;
            total.candidates += b.candidates;
            total.prep_ms += b.prep_ms;
            total.sample_ms += b.sample_ms;
            total.filter_ms += b.filter_ms;
            total.rerank_ms += b.rerank_ms;
            total.total_ms += b.total_ms;
        }
        ctx->last_batch = total;
    }
    if (!big_k && n_queries && ds->n_rows && ds->shards.size() > 1 && !(flags & PQV_TIES_BY_POSITION) &&
        batch_path_applies(*ds, ds->shards[0].d_data, n_queries, k, true))
        PQV_TRY(batch_topk_sharded(ctx, *ds, queries, n_queries, k, flags, out_row_idx, out_dist, out_count, handled));
    for (uint32_t q = 0; q < n_queries; ++q)
        if (!handled[q])
            PQV_TRY(topk_one(ctx, *ds, queries + (size_t)q * ds->dim, nullptr, 0, k, flags, out_row_idx + (size_t)q * k,
                             out_dist + (size_t)q * k, out_count + q));
    return PQV_OK;
}

int pqv_last_batch_timing(pqv_ctx *ctx, pqv_batch_timing *out) {
    if (!ctx || !out) return fail(PQV_EINVAL, "null argument");
    *out = ctx->last_batch;
    return PQV_OK;
}

int pqv_l2_topk_gather(pqv_ctx *ctx, uint64_t handle, const float *query, const uint32_t *row_ids, uint64_t n_ids,
                       uint32_t k, uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!ctx || !query || !out_row_idx || !out_dist || !out_count) return fail(PQV_EINVAL, "null argument");
    if (n_ids && !row_ids) return fail(PQV_EINVAL, "row_ids is null");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    PQV_TRY(check_topk_args(k, ds->dim, flags, true));
    if (n_ids > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "candidate positions are u32");
    for (u64 i = 0; i < n_ids; ++i)
        if (row_ids[i] >= ds->n_rows) return fail(PQV_EINVAL, "row id %u at position %llu is out of range (%llu rows)", row_ids[i], (unsigned long long)i, (unsigned long long)ds->n_rows);
    if (n_ids == 0) {
        *out_count = 0;
        return PQV_OK;
    }
    static const uint32_t dummy = 0;
    return topk_one(ctx, *ds, query, row_ids ? row_ids : &dummy, n_ids, k, flags, out_row_idx, out_dist, out_count);
}


int pqv_l2_topk_candidates(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags,
                           uint32_t pos_base, uint64_t *out_keys, uint64_t cap, uint64_t *out_count) {
    if (!ctx || !query || !out_count || (cap && !out_keys)) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (flags & PQV_TIES_BY_POSITION) return fail(PQV_EINVAL, "candidates are only defined for the reference tie order");
    if ((u64)pos_base + ds->n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "global row ids are u32");
    std::vector<u64> ent;
    PQV_TRY(topk_one(ctx, *ds, query, nullptr, 0, k, flags, nullptr, nullptr, nullptr, &ent, pos_base));
    *out_count = ent.size();
    if (ent.size() > cap) return fail(PQV_ELIMIT, "%zu candidate keys do not fit the caller's buffer of %llu", ent.size(), (unsigned long long)cap);
    if (!ent.empty()) memcpy(out_keys, ent.data(), ent.size() * 8);
    return PQV_OK;
}

int pqv_replay_candidates(const uint64_t *keys, uint64_t n_keys, const uint32_t *row_ids, uint32_t k, uint32_t flags,
                          uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if ((n_keys && !keys) || !out_row_idx || !out_dist || !out_count) return fail(PQV_EINVAL, "null argument");
    if (k == 0) return fail(PQV_EINVAL, "k must be > 0");
    std::vector<u64> ent(keys, keys + n_keys);
    RowMap row_of;
    row_of.ids = row_ids;
    *out_count = (uint32_t)replay_reference_heap(ent, row_of, k, flags, out_row_idx, out_dist);
    return PQV_OK;
}

static int batch_keys_locked(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k, uint32_t flags,
                             uint32_t pos_base, uint64_t *out_keys, uint32_t *out_count);
int pqv_l2_topk_batch_keys(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k,
                           uint32_t flags, uint32_t pos_base, uint64_t *out_keys, uint32_t *out_count) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    if (n_queries && (!queries || !out_keys || !out_count)) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return batch_keys_locked(ctx, handle, queries, n_queries, k, flags, pos_base, out_keys, out_count);
}
static int batch_keys_locked(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k, uint32_t flags,
                             uint32_t pos_base, uint64_t *out_keys, uint32_t *out_count) {
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (flags & PQV_TIES_BY_POSITION) return fail(PQV_EINVAL, "batch keys are only defined for the reference tie order");
    if ((u64)pos_base + ds->n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "global row ids are u32");
    for (uint32_t q = 0; q < n_queries; ++q) out_count[q] = 0xFFFFFFFFu;
    ctx->last_batch = pqv_batch_timing{};
    ctx->batch_state.valid = false;
    if (n_queries && ds->n_rows == 0) {  // an empty slice contributes nothing
        for (uint32_t q = 0; q < n_queries; ++q) out_count[q] = 0;
        return PQV_OK;
    }
    if (n_queries > BATCH_MAX_QUERIES)
        return fail(PQV_ELIMIT, "at most %u queries per pqv_l2_topk_batch_keys call (got %u)", BATCH_MAX_QUERIES, n_queries);
    if (n_queries && batch_path_applies(*ds, ds->shards[0].d_data, n_queries, k)) {
        DeviceState &D = ctx->devs[ds->shards[0].di];
        DevGuard guard(D.dev);
        std::vector<uint8_t> handled(n_queries, 0);
        PQV_TRY(batch_topk(ctx, D, *ds, ds->n_rows, ds->dim, queries, n_queries, k, flags, nullptr, nullptr, nullptr, handled,
                           reinterpret_cast<u64 *>(out_keys), out_count, pos_base));
        if (ctx->batch_state.valid) ctx->batch_state.handle = handle;
    }
    return PQV_OK;
}

static int tie_candidates_locked(pqv_ctx *ctx, uint64_t handle, uint32_t q_index, const float *query, std::vector<u64> &ent);
int pqv_l2_topk_batch_tie_candidates(pqv_ctx *ctx, uint64_t handle, uint32_t q_index, const float *query, uint64_t *out_keys,
                                     uint64_t cap, uint64_t *out_count) {
    if (!ctx || !query || !out_count || (cap && !out_keys)) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::vector<u64> ent;
    PQV_TRY(tie_candidates_locked(ctx, handle, q_index, query, ent));
    *out_count = ent.size();
    if (ent.size() > cap) return fail(PQV_ELIMIT, "%zu candidate keys do not fit the caller's buffer of %llu", ent.size(), (unsigned long long)cap);
    if (!ent.empty()) memcpy(out_keys, ent.data(), ent.size() * 8);
    return PQV_OK;
}
static int tie_candidates_locked(pqv_ctx *ctx, uint64_t handle, uint32_t q_index, const float *query, std::vector<u64> &ent) {
    pqv_ctx::BatchState &bs = ctx->batch_state;
    if (!bs.valid || bs.handle != handle) return fail(PQV_EINVAL, "no batched pass of this dataset is pending (call pqv_l2_topk_batch_keys first)");
    if (q_index >= bs.nq) return fail(PQV_EINVAL, "query index %u out of range (batch of %u)", q_index, bs.nq);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    // admissions of the reference heap inside this slice: positions < S from the exact scan of that prefix alone,
    // positions >= S are all among the query's candidates of the batched pass
    ent.clear();
    uint32_t dummy = 0;
    PQV_TRY(topk_one(ctx, *ds, query, nullptr, 0, bs.k, bs.flags, nullptr, nullptr, &dummy, &ent, bs.pos_base, nullptr, nullptr, bs.S));
    DeviceState &D = ctx->devs[bs.dev_index];
    DevGuard guard(D.dev);
    const uint32_t cq = std::min<uint32_t>(bs.seg_count[q_index], bs.cap_q);
    std::vector<u64> seg(cq);
    if (cq) CU_TRY(cudaMemcpy(seg.data(), D.tb_seg.p + (size_t)q_index * bs.cap_q, (size_t)cq * 8, cudaMemcpyDeviceToHost));
    for (u64 key : seg)
        if (key_pos(key) >= bs.S) ent.push_back(key + bs.pos_base);
    return PQV_OK;
}

int pqv_merge_batch_keys(const uint64_t *keys, const uint32_t *counts, uint32_t n_ranks, uint32_t n_queries, uint32_t k,
                         uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count,
                         uint8_t *out_needs_replay) {
    if (n_queries && n_ranks && (!keys || !counts)) return fail(PQV_EINVAL, "null argument");
    if (n_queries && (!out_row_idx || !out_dist || !out_count || !out_needs_replay)) return fail(PQV_EINVAL, "null argument");
    if (k == 0) return fail(PQV_EINVAL, "k must be > 0");
    const size_t kp = (size_t)k + 1;
    // queries are independent: large batches are merged by up to four host threads
    const uint32_t n_thr = n_queries >= 512 ? (uint32_t)std::min<size_t>(std::min<size_t>(tie_threads(), 4), n_queries / 256) : 1u;  // spawning costs ~25 us each
    auto merge_range = [&](uint32_t q_begin, uint32_t q_end) {
    std::vector<u64> all;
    for (uint32_t q = q_begin; q < q_end; ++q) {
        out_count[q] = 0;
        out_needs_replay[q] = 0;
        all.clear();
        for (uint32_t r = 0; r < n_ranks; ++r) {
            const uint32_t c = counts[(size_t)r * n_queries + q];
            if (c == 0xFFFFFFFFu || c > kp) {  // some slice could not decide this query
                out_needs_replay[q] = 1;
                break;
            }
            const u64 *src = reinterpret_cast<const u64 *>(keys) + ((size_t)r * n_queries + q) * kp;
            all.insert(all.end(), src, src + c);
        }
        if (out_needs_replay[q]) continue;
        // the k + 1 smallest keys of the whole table are among the slices' k + 1 smallest
        const size_t take = std::min<size_t>(kp, all.size());
        std::partial_sort(all.begin(), all.begin() + take, all.end());
        const size_t cnt = std::min<size_t>(k, all.size());
        bool tie = all.size() > k && (uint32_t)(all[k - 1] >> 32) == (uint32_t)(all[k] >> 32);  // kept set = heap layout's
        if (cnt && (uint32_t)(all[cnt - 1] >> 32) > 0x7F800000u) tie = true;                       // NaN: reference-specific
        for (size_t i = 0; i < cnt && !tie; ++i) {
            const float d = key_dist(all[i]);
            out_row_idx[(size_t)q * k + i] = key_pos(all[i]);
            out_dist[(size_t)q * k + i] = (flags & PQV_SQRT) ? sqrtf(d) : d;
            if (i && out_dist[(size_t)q * k + i] == out_dist[(size_t)q * k + i - 1]) tie = true;  // order = heap layout's
        }
        if (tie) out_needs_replay[q] = 1;
        else out_count[q] = (uint32_t)cnt;
    }
    };
    if (n_thr <= 1) {
        merge_range(0, n_queries);
    } else {
        std::vector<std::thread> th;
        const uint32_t per = (n_queries + n_thr - 1) / n_thr;
        for (uint32_t t = 1; t < n_thr; ++t) th.emplace_back(merge_range, std::min(t * per, n_queries), std::min((t + 1) * per, n_queries));
        merge_range(0, std::min(per, n_queries));
        for (auto &x : th) x.join();
    }
    return PQV_OK;
}

int pqv_peer_exchange_create(pqv_ctx *ctx, uint32_t world, uint32_t rank, uint32_t cap_keys, uint8_t *out_handle64) {
    if (!ctx || !out_handle64) return fail(PQV_EINVAL, "null argument");
    if (world == 0 || rank >= world || cap_keys == 0) return fail(PQV_EINVAL, "bad world / rank / capacity");
    if (ctx->devs.size() != 1) return fail(PQV_EINVAL, "the peer exchange is for one process per GPU (single-device context)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    std::lock_guard<std::mutex> lk(ctx->mu);
    peer_release(ctx);
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    PeerExchange &px = ctx->peer;
    px.world = world;
    px.rank = rank;
    px.cap = cap_keys;
    px.seq = 0;
    cudaError_t e = cudaMalloc((void **)&px.local, px.words() * 8);
    if (e != cudaSuccess) return fail(PQV_ENOMEM, "peer exchange buffer: %s", cudaGetErrorString(e));
    CU_TRY(cudaMemset(px.local, 0, px.words() * 8));  // flags 0: sequence numbers start at 1
    CU_TRY(cudaMalloc((void **)&px.d_peers, (size_t)world * sizeof(u64 *)));
    CU_TRY(cudaMalloc((void **)&px.d_timeout, 4));
    PQV_TRY(px.h_block.ensure((size_t)world * (1 + (size_t)cap_keys) + 2));  // flags, total, counts[world], keys
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, px.local));
    memcpy(out_handle64, &h, 64);
    return PQV_OK;
}

int pqv_peer_exchange_open(pqv_ctx *ctx, const uint8_t *handles) {
    if (!ctx || !handles) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PeerExchange &px = ctx->peer;
    if (!px.local) return fail(PQV_EINVAL, "pqv_peer_exchange_create has not been called");
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    px.peers.assign(px.world, nullptr);
    for (uint32_t d = 0; d < px.world; ++d) {
        if (d == px.rank) {
            px.peers[d] = px.local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)d * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(PQV_ECUDA, "cudaIpcOpenMemHandle(rank %u): %s", d, cudaGetErrorString(e));
        px.peers[d] = static_cast<u64 *>(p);
    }
    CU_TRY(cudaMemcpy(px.d_peers, px.peers.data(), (size_t)px.world * sizeof(u64 *), cudaMemcpyHostToDevice));
    px.ready = true;
    return PQV_OK;
}

// scan -> filter -> publish to every peer -> wait + pack: leaves [flags, total, counts[world], keys...] in px.h_block
// (page-locked, written by the pack kernel itself); ctx->mu held
static int p2p_collect(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags, uint32_t pos_base,
                       u64 *out_total, uint32_t *out_overflow) {
    PeerExchange &px = ctx->peer;
    if (!px.ready) return fail(PQV_EINVAL, "the peer exchange is not set up (pqv_peer_exchange_create / _open)");
    if (px.world > 64) return fail(PQV_ELIMIT, "at most 64 ranks per peer exchange");
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (flags & PQV_TIES_BY_POSITION) return fail(PQV_EINVAL, "candidates are only defined for the reference tie order");
    if (ds->shards.size() != 1) return fail(PQV_EINVAL, "single-device dataset expected");
    if ((u64)pos_base + ds->n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "global row ids are u32");
    Shard &sh = ds->shards[0];
    DeviceState &D = ctx->devs[sh.di];
    DevGuard guard(D.dev);
    const int order = (flags & PQV_SUM_SEQ) ? 1 : 0;
    const u64 seq = ++px.seq;  // every rank calls this the same number of times: the sequence numbers agree
    PQV_TRY(D.d_query.ensure(ds->dim));
    PQV_TRY(D.h_query.ensure(ds->dim));
    PQV_TRY(D.final_topk.ensure(PQV_MAX_K));
    PQV_TRY(D.ent_out.ensure((size_t)(1u << 16) + 1));
    const uint32_t cap_local = (uint32_t)std::min<size_t>(D.ent_out.cap - 1, 0xFFFFFFF0u);
    memcpy(D.h_query.p, query, (size_t)ds->dim * 4);
    CU_TRY(cudaMemcpyAsync(D.d_query.p, D.h_query.p, (size_t)ds->dim * 4, cudaMemcpyHostToDevice, D.stream));
    CU_TRY(cudaMemsetAsync(D.ent_out.p, 0, 8, D.stream));
    ScanGeom g;
    if (ds->n_rows) {
        PQV_TRY(enqueue_scan(ctx, D, sh.d_data, nullptr, ds->n_rows, ds->dim, D.d_query.p, k, order, pos_base, nullptr,
                             D.final_topk.p, D.ent_out.p, cap_local, true, &g));
    } else {
        CU_TRY(cudaEventRecord(D.ev[0], D.stream));
        CU_TRY(cudaEventRecord(D.ev[1], D.stream));
        CU_TRY(cudaEventRecord(D.ev[2], D.stream));
        g.grid = 0;
    }
    pqv::peer_publish_kernel<<<px.world, 256, 0, D.stream>>>(D.ent_out.p, px.cap, px.d_peers, px.rank, px.world, seq,
                                                             ds->n_rows ? D.final_topk.p : nullptr, k);
    pqv::peer_wait_pack_kernel<<<1, 256, 0, D.stream>>>(px.local, px.cap, px.world, seq, px.h_block.p, 0u);
    CU_TRY(cudaGetLastError());
    px.t_enqueued = trace_now_ms();
    CU_TRY(cudaStreamSynchronize(D.stream));
    const u64 *hb = px.h_block.p;
    if (hb[0] & 1ull) return fail(PQV_ECUDA, "peer exchange timed out waiting for the other ranks (sequence %llu)", (unsigned long long)seq);
    *out_overflow = (hb[0] & 2ull) ? 1u : 0u;  // some rank had more candidates than a slot holds: every rank sees it
    *out_total = hb[1];
    pqv_timing tm{};
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, D.ev[0], D.ev[1]);
    cudaEventElapsedTime(&b, D.ev[1], D.ev[2]);
    tm.scan_ms = a;
    tm.post_ms = b;
    tm.total_ms = a + b;
    tm.scan_bytes = ds->n_rows * (u64)ds->dim * 4;
    tm.launches = 6;
    tm.grid = g.grid;
    tm.entrants = (uint32_t)hb[1];
    ctx->last = tm;
    return PQV_OK;
}

int pqv_l2_topk_candidates_p2p(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags, uint32_t pos_base,
                               uint64_t *out_keys, uint64_t cap_total, uint64_t *out_count, uint32_t *out_overflow) {
    if (!ctx || !query || !out_keys || !out_count || !out_overflow) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    u64 total = 0;
    PQV_TRY(p2p_collect(ctx, handle, query, k, flags, pos_base, &total, out_overflow));
    *out_count = total;
    if (*out_overflow) return PQV_OK;
    if (any_nan_key(ctx->peer.h_block.p + 2 + ctx->peer.world, total)) {  // NaN distance: the collective path replays every row
        *out_overflow = 1;
        return PQV_OK;
    }
    if (total > cap_total) return fail(PQV_ELIMIT, "%llu candidate keys do not fit the caller's buffer of %llu", (unsigned long long)total, (unsigned long long)cap_total);
    memcpy(out_keys, ctx->peer.h_block.p + 2 + ctx->peer.world, total * 8);
    return PQV_OK;
}

// The whole sharded search of one rank in one call: pqv_l2_topk_candidates_p2p + pqv_replay_candidates.  Every rank calls
// it with the same query and gets the same, bit-exact (row_idx, distance) list.  *out_overflow = 1 (nothing else written):
// a rank had more candidates than a slot holds -- all ranks see it and take the collective path for this query.
int pqv_l2_topk_p2p(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags, uint32_t pos_base,
                    uint32_t *out_row_idx, float *out_dist, uint32_t *out_count, uint32_t *out_overflow) {
    if (!ctx || !query || !out_row_idx || !out_dist || !out_count || !out_overflow) return fail(PQV_EINVAL, "null argument");
    static const bool trace = getenv("PQV_TRACE") != nullptr;
    const double t_in = trace ? trace_now_ms() : 0;
    std::lock_guard<std::mutex> lk(ctx->mu);
    u64 total = 0;
    *out_count = 0;
    PQV_TRY(p2p_collect(ctx, handle, query, k, flags, pos_base, &total, out_overflow));
    const double t_sync = trace ? trace_now_ms() : 0;
    struct TraceOut {
        bool on;
        double t_in, t_sync;
        pqv_ctx *ctx;
        ~TraceOut() {
            if (on)
                fprintf(stderr, "[pqv trace] l2_topk_p2p rank %u: enqueue %.1f us, wait for the device %.1f us (scan %.1f + post %.1f on the device), replay %.1f us\n",
                        ctx->peer.rank, (ctx->peer.t_enqueued - t_in) * 1e3, (t_sync - ctx->peer.t_enqueued) * 1e3, ctx->last.scan_ms * 1e3,
                        ctx->last.post_ms * 1e3, (trace_now_ms() - t_sync) * 1e3);
        }
    } trace_out{trace, t_in, t_sync, ctx};
    if (*out_overflow) return PQV_OK;
    const u64 *keys = ctx->peer.h_block.p + 2 + ctx->peer.world;
    if (any_nan_key(keys, total)) {  // a NaN distance: only the loop over every row answers that -- the collective path does it
        *out_overflow = 1;
        return PQV_OK;
    }
    std::vector<u64> ent(keys, keys + total);
    *out_count = (uint32_t)replay_reference_heap(ent, RowMap{}, k, flags, out_row_idx, out_dist);
    return PQV_OK;
}

// `n_words` host words of this rank travel to every peer and come back as [rank][n_words] (same order on every rank) in
// px.h_block behind the header; raw = 1 keeps every word in place (a structured payload), raw = 0 packs candidate keys.
// ctx->mu held.  *out_overflow = 1: the payload does not fit a slot (nothing sent: every rank passes the same size).
static int p2p_exchange_words(pqv_ctx *ctx, DeviceState &D, const u64 *words, u64 n_words, uint32_t raw, u64 *out_total,
                              uint32_t *out_overflow) {
    PeerExchange &px = ctx->peer;
    *out_overflow = 0;
    *out_total = 0;
    if (n_words > px.cap) {
        *out_overflow = 1;
        return PQV_OK;
    }
    DevGuard guard(D.dev);
    const u64 seq = ++px.seq;
    PQV_TRY(D.ent_out.ensure((size_t)std::max<u64>(n_words, 1u << 16) + 1));
    PQV_TRY(D.h_ent_out.ensure((size_t)std::max<u64>(n_words, ENT_FIRST_CHUNK) + 1));
    D.h_ent_out.p[0] = n_words;
    if (n_words) memcpy(D.h_ent_out.p + 1, words, n_words * 8);
    CU_TRY(cudaMemcpyAsync(D.ent_out.p, D.h_ent_out.p, (n_words + 1) * 8, cudaMemcpyHostToDevice, D.stream));
    pqv::peer_publish_kernel<<<px.world, 256, 0, D.stream>>>(D.ent_out.p, px.cap, px.d_peers, px.rank, px.world, seq, nullptr, 1u);
    pqv::peer_wait_pack_kernel<<<1, 256, 0, D.stream>>>(px.local, px.cap, px.world, seq, px.h_block.p, raw);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(D.stream));
    const u64 *hb = px.h_block.p;
    if (hb[0] & 1ull) return fail(PQV_ECUDA, "peer exchange timed out waiting for the other ranks (sequence %llu)", (unsigned long long)seq);
    if (hb[0] & 2ull) *out_overflow = 1;
    *out_total = hb[1];
    return PQV_OK;
}

// A rank's whole sharded BATCH in one call (config C5 from one process per GPU): the tensor-core pass over this rank's slice
// (pqv_l2_topk_batch_keys), the per-query key lists of all ranks exchanged over NVLink peer memory, merged on the host
// (pqv_merge_batch_keys), and the queries the merge cannot decide (exact ties) replayed from their candidates
// (pqv_l2_topk_batch_tie_candidates + the same exchange + pqv_replay_candidates).  Every rank passes the same queries and
// receives the same bit-exact results.  *out_overflow = 1: the payload does not fit the exchange slots (create them with
// cap_keys >= n_queries * (k + 2)) or a rank's slice declined the batch -- all ranks see it and take the collective path.
int pqv_l2_topk_batch_p2p(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k, uint32_t flags,
                          uint32_t pos_base, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count, uint32_t *out_replayed,
                          uint32_t *out_overflow) {
    if (!ctx || !out_overflow) return fail(PQV_EINVAL, "null argument");
    if (n_queries && (!queries || !out_row_idx || !out_dist || !out_count)) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PeerExchange &px = ctx->peer;
    if (!px.ready) return fail(PQV_EINVAL, "the peer exchange is not set up (pqv_peer_exchange_create / _open)");
    *out_overflow = 0;
    if (out_replayed) *out_replayed = 0;
    if (n_queries == 0) return PQV_OK;
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (ds->shards.size() != 1) return fail(PQV_EINVAL, "single-device dataset expected");
    DeviceState &D = ctx->devs[ds->shards[0].di];
    const size_t kp = (size_t)k + 1, nq = n_queries;
    const u64 n_words = nq * kp + nq;
    static const bool trace = getenv("PQV_TRACE") != nullptr;
    double tt[5] = {trace_now_ms(), 0, 0, 0, 0};
    // this rank's pass: k + 1 keys per query + the counts (widened to one word each) behind them
    std::vector<u64> mine(n_words, pqv::KEY_MAX);
    std::vector<uint32_t> cnt32(nq);
    PQV_TRY(batch_keys_locked(ctx, handle, queries, n_queries, k, flags, pos_base, reinterpret_cast<uint64_t *>(mine.data()), cnt32.data()));
    for (size_t q = 0; q < nq; ++q) mine[nq * kp + q] = cnt32[q];
    u64 total = 0;
    tt[1] = trace_now_ms();
    PQV_TRY(p2p_exchange_words(ctx, D, mine.data(), n_words, 1u, &total, out_overflow));
    tt[2] = trace_now_ms();
    if (*out_overflow) return PQV_OK;
    if (total != n_words * px.world) {  // the ranks did not send the same shape: nothing to merge
        *out_overflow = 1;
        return PQV_OK;
    }
    const u64 *un = px.h_block.p + 2 + px.world;
    std::vector<u64> all_keys((size_t)px.world * nq * kp);
    std::vector<uint32_t> all_cnt((size_t)px.world * nq);
    uint32_t undecided = 0;
    for (uint32_t r = 0; r < px.world; ++r) {
        memcpy(all_keys.data() + (size_t)r * nq * kp, un + (size_t)r * n_words, nq * kp * 8);
        for (size_t q = 0; q < nq; ++q) {
            all_cnt[(size_t)r * nq + q] = (uint32_t)un[(size_t)r * n_words + nq * kp + q];
            undecided += all_cnt[(size_t)r * nq + q] == 0xFFFFFFFFu ? 1u : 0u;
        }
    }
    if (undecided > 64) {  // some slice declined the batch: hundreds of single scans belong to the caller's collective path
        *out_overflow = 1;
        return PQV_OK;
    }
    std::vector<uint8_t> need(nq, 0);
    PQV_TRY(pqv_merge_batch_keys(reinterpret_cast<const uint64_t *>(all_keys.data()), all_cnt.data(), px.world, n_queries, k, flags,
                                 out_row_idx, out_dist, out_count, need.data()));
    tt[3] = trace_now_ms();
    // flagged queries: deterministic on identical data, so every rank walks the same list in the same order
    const bool pending = ctx->batch_state.valid && ctx->batch_state.handle == handle;
    std::vector<u64> cand;
    for (uint32_t q = 0; q < n_queries; ++q) {
        if (!need[q]) continue;
        const float *qv = queries + (size_t)q * ds->dim;
        if (pending) {
            PQV_TRY(tie_candidates_locked(ctx, handle, q, qv, cand));
        } else {  // no batched pass on this rank (tiny / unaligned slice): candidates of the full single-query scan
            cand.clear();
            uint32_t dummy = 0;
            PQV_TRY(topk_one(ctx, *ds, qv, nullptr, 0, k, flags, nullptr, nullptr, &dummy, &cand, pos_base));
        }
        PQV_TRY(p2p_exchange_words(ctx, D, cand.data(), cand.size(), 0u, &total, out_overflow));
        if (*out_overflow) return PQV_OK;
        const u64 *keys = px.h_block.p + 2 + px.world;
        if (any_nan_key(keys, total)) {
            *out_overflow = 1;
            return PQV_OK;
        }
        std::vector<u64> ent(keys, keys + total);
        out_count[q] = (uint32_t)replay_reference_heap(ent, RowMap{}, k, flags, out_row_idx + (size_t)q * k, out_dist + (size_t)q * k);
        if (out_replayed) ++*out_replayed;
    }
    if (trace)
        fprintf(stderr, "[pqv trace] batch_p2p rank %u: pass %.2f ms (device %.2f), exchange (incl. waiting for the peers) %.2f, merge %.2f, "
                        "tie replays %.2f\n", px.rank, tt[1] - tt[0], ctx->last_batch.total_ms, tt[2] - tt[1], tt[3] - tt[2],
                trace_now_ms() - tt[3]);
    return PQV_OK;
}

int pqv_last_timing(pqv_ctx *ctx, pqv_timing *out) {
    if (!ctx || !out) return fail(PQV_EINVAL, "null argument");
    *out = ctx->last;
    return PQV_OK;
}

int pqv_bench_scan(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags, uint32_t iters,
                   double *out_ms_per_scan) {
    if (!ctx || !query || !out_ms_per_scan || iters == 0) return fail(PQV_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    PQV_TRY(check_topk_args(k, ds->dim, flags));
    if (ds->shards.size() != 1 || ds->shards[0].n_rows == 0) return fail(PQV_EINVAL, "pqv_bench_scan needs a non-empty single-device dataset");
    Shard &sh = ds->shards[0];
    DeviceState &D = ctx->devs[sh.di];
    DevGuard guard(D.dev);
    const int order = (flags & PQV_SUM_SEQ) ? 1 : 0;
    PQV_TRY(D.d_query.ensure(ds->dim));
    CU_TRY(cudaMemcpyAsync(D.d_query.p, query, (size_t)ds->dim * 4, cudaMemcpyHostToDevice, D.stream));
    PQV_TRY(D.final_topk.ensure(PQV_MAX_K));
    PQV_TRY(D.ent_out.ensure((size_t)(1u << 20) + 1));
    const uint32_t cap = (uint32_t)(D.ent_out.cap - 1);
    std::vector<cudaEvent_t> evs((size_t)iters * 3);
    for (auto &e : evs) CU_TRY(cudaEventCreate(&e));
    ScanGeom g{};
    for (uint32_t it = 0; it < iters; ++it) {
        CU_TRY(cudaMemsetAsync(D.ent_out.p, 0, 8, D.stream));
        // same launch sequence as topk_one, with our own event triplet per iteration
        cudaEvent_t keep[3] = {D.ev[0], D.ev[1], D.ev[2]};
        D.ev[0] = evs[it * 3 + 0];
        D.ev[1] = evs[it * 3 + 1];
        D.ev[2] = evs[it * 3 + 2];
        int rc = enqueue_scan(ctx, D, sh.d_data, nullptr, sh.n_rows, ds->dim, D.d_query.p, k, order, 0, nullptr,
                              D.final_topk.p, D.ent_out.p, cap, true, &g);
        D.ev[0] = keep[0];
        D.ev[1] = keep[1];
        D.ev[2] = keep[2];
        if (rc) return rc;
    }
    CU_TRY(cudaStreamSynchronize(D.stream));
    double scan = 0, post = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, evs[it * 3], evs[it * 3 + 1]);
        cudaEventElapsedTime(&b, evs[it * 3 + 1], evs[it * 3 + 2]);
        scan += a;
        post += b;
    }
    float whole = 0;
    cudaEventElapsedTime(&whole, evs[0], evs[(size_t)iters * 3 - 1]);
    for (auto &e : evs) cudaEventDestroy(e);
    pqv_timing tm{};
    tm.scan_ms = scan / iters;
    tm.post_ms = post / iters;
    tm.total_ms = whole / iters;
    tm.scan_bytes = sh.n_rows * (u64)ds->dim * 4;
    tm.launches = 4;
    tm.grid = g.grid;
    ctx->last = tm;
    *out_ms_per_scan = tm.scan_ms;
    return PQV_OK;
}

// ---- k-means pieces ------------------------------------------------------------------------------
struct SweepTune {
    bool repeated = false;  // the same rows are swept again and again: the one-group-per-warp shape with L2 eviction hints
    u64 keep_groups = 0;    // groups [0, keep_groups) are loaded evict_last (kept in L2 for the next sweep), the rest evict_first
};
static int dist_launch(DeviceState &D, const float *d_data, const uint32_t *d_ids, u64 n, uint32_t dim, const float *d_vec,
                       float *d_out, int min_update, float *h_mirror = nullptr, const SweepTune *tune = nullptr) {
    const bool vec4 = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_data) & 15) == 0);
    const size_t smem = (size_t)((dim + 3u) & ~3u) * 4 + (size_t)SCAN_WARPS * pqv::TileCfg<0, true>::TILE_FLOATS * 4;
    const u64 NG = (n + 31) / 32;
    constexpr uint32_t WIDE_MAX_DIM = 4096;
    if (!h_mirror && vec4 && dim <= WIDE_MAX_DIM && NG <= (u64)D.sm_count * 2 && (reinterpret_cast<uintptr_t>(d_vec) & 15) == 0) {
        // short table (centroid ranking): one CTA per 32 rows instead of one warp (l2_dist_wide_kernel)
        uint32_t ts = ((dim >> 2) + 3u) & ~3u;
        if (((ts >> 2) & 1u) == 0) ts += 4;
        uint32_t tmax = ((WIDE_MAX_DIM >> 2) + 3u) & ~3u;
        if (((tmax >> 2) & 1u) == 0) tmax += 4;
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(pqv::l2_dist_wide_kernel), (size_t)32u * tmax * 4u));
        pqv::l2_dist_wide_kernel<<<(uint32_t)NG, 256, (size_t)32 * ts * 4, D.stream>>>(d_data, d_ids, n, dim, ts, d_vec, d_out,
                                                                                     min_update);
        CU_TRY(cudaGetLastError());
        return PQV_OK;
    }
    if (tune && tune->repeated && vec4 && d_ids && !h_mirror) {
        // repeated sweeps of one gathered set (k-means++: 1023 sweeps of a 154 MB init set): one group per warp, one warp per
        // CTA (the block scheduler evens out the SMs), 16 rows x 2 float4 per lane in flight, and the first keep_groups groups
        // loaded evict_last so that they are still in L2 for the next sweep (profiles/r02_kpp_probe_v*.jsonl: 27 -> 20 us per
        // sweep at 45 % of the set kept; the shape alone changed nothing, 60 % and more thrashes)
        auto kern = pqv::l2_dist_kernel<true, true, 1, 16, 2, true>;
        const size_t sm = (size_t)((dim + 3u) & ~3u) * 4 + (size_t)pqv::TileCfg<0, true, 2>::TILE_FLOATS * 4;
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), sm));
        kern<<<(uint32_t)NG, 32, sm, D.stream>>>(d_data, d_ids, n, dim, d_vec, d_out, min_update, nullptr, tune->keep_groups);
        CU_TRY(cudaGetLastError());
        return PQV_OK;
    }
    const uint32_t grid = (uint32_t)std::min<u64>((NG + SCAN_WARPS - 1) / SCAN_WARPS, (u64)D.sm_count * 4);
#define DIST_GO(V, G)                                                                                        \
    do {                                                                                                     \
        auto kern = pqv::l2_dist_kernel<V, G, SCAN_WARPS>;                                                   \
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem));                                \
        kern<<<grid, SCAN_WARPS * 32, smem, D.stream>>>(d_data, d_ids, n, dim, d_vec, d_out, min_update, h_mirror, 0); \
    } while (0)
    if (vec4) {
        if (d_ids) DIST_GO(true, true);
        else DIST_GO(true, false);
    } else {
        if (d_ids) DIST_GO(false, true);
        else DIST_GO(false, false);
    }
#undef DIST_GO
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

// resolve (handle, rows) to a device pointer for the first `n` rows; streams host rows into scratch
static int resolve_rows(pqv_ctx *ctx, uint64_t handle, const float *rows, u64 n, uint32_t dim, DeviceState **Dout,
                        const float **d_rows, Dataset **ds_out = nullptr) {
    if (ds_out) *ds_out = nullptr;
    if (rows) {
        DeviceState &D = ctx->devs[0];
        DevGuard guard(D.dev);
        PQV_TRY(D.d_tmp_rows.ensure((size_t)n * dim));
        CU_TRY(cudaMemcpyAsync(D.d_tmp_rows.p, rows, (size_t)n * dim * 4, cudaMemcpyHostToDevice, D.stream));
        *Dout = &D;
        *d_rows = D.d_tmp_rows.p;
        return PQV_OK;
    }
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu (and rows is null)", (unsigned long long)handle);
    if (ds->dim != dim) return fail(PQV_EINVAL, "dimension mismatch: dataset has %u, call has %u", ds->dim, dim);
    if (ds->shards.size() != 1) return fail(PQV_EINVAL, "k-means entry points need a single-device dataset");
    if (n > ds->n_rows) return fail(PQV_EINVAL, "n = %llu exceeds the dataset's %llu rows", (unsigned long long)n, (unsigned long long)ds->n_rows);
    *Dout = &ctx->devs[ds->shards[0].di];
    *d_rows = ds->shards[0].d_data;
    if (ds_out) *ds_out = ds;
    return PQV_OK;
}

// the 16-bit shadow of a resident table for an assignment sweep over its first n rows (null view: none -- layout, memory)
static int sweep_shadow(DeviceState &D, Dataset *ds, const float *d_rows, u64 n, ShadowView *sv, bool *have, bool *built,
                        Shard *sh = nullptr) {
    *have = false;
    *built = false;
    if (!ds || n < 2048 || !shadow_layout_ok(ds->dim, d_rows)) return PQV_OK;  // small sweeps take the SIMT kernel anyway
    const int rc = shard_shadow(D, *ds, sh ? *sh : ds->shards[0], sv, built);
    if (rc == PQV_OK) *have = true;
    else if (rc != PQV_ENOMEM) return rc;
    return PQV_OK;
}

int pqv_kmeans_assign(pqv_ctx *ctx, uint64_t handle, const float *rows, uint64_t n, uint32_t dim, const float *centroids,
                      uint32_t n_clusters, uint32_t *out_assign, uint64_t *out_sizes) {
    if (!ctx || !centroids || (!out_assign && n)) return fail(PQV_EINVAL, "null argument");
    if (dim == 0) return fail(PQV_EINVAL, "Embedding dimension must be > 0");
    if (n_clusters == 0) return fail(PQV_EINVAL, "Cluster count must be > 0");  // src/ivf/index.rs:24
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (out_sizes) memset(out_sizes, 0, sizeof(uint64_t) * n_clusters);
    if (n == 0) return PQV_OK;
    if (!rows) {
        // a resident table spread over several devices: every shard sweeps its own rows against its own copy of the
        // centroids; all sweeps are enqueued before the first one is waited for
        Dataset *mds = find_dataset(ctx, handle);
        if (mds && mds->shards.size() > 1) {
            if (mds->dim != dim) return fail(PQV_EINVAL, "dimension mismatch: dataset has %u, call has %u", mds->dim, dim);
            if (n > mds->n_rows) return fail(PQV_EINVAL, "n = %llu exceeds the dataset's %llu rows", (unsigned long long)n, (unsigned long long)mds->n_rows);
            struct Piece {
                DeviceState *D;
                u64 off, cnt;
                int path, kind;
                bool built;
                uint32_t h_counts[8];
            };
            std::vector<Piece> pieces;
            pieces.reserve(mds->shards.size());
            for (Shard &sh : mds->shards) {
                if (sh.first_row >= n || sh.n_rows == 0) continue;
                const u64 cnt = std::min<u64>(sh.n_rows, n - sh.first_row);
                DeviceState &D = ctx->devs[sh.di];
                DevGuard guard(D.dev);
                PQV_TRY(D.d_centroids.ensure((size_t)n_clusters * dim));
                CU_TRY(cudaMemcpyAsync(D.d_centroids.p, centroids, (size_t)n_clusters * dim * 4, cudaMemcpyHostToDevice, D.stream));
                PQV_TRY(D.d_assign.ensure(cnt));
                pieces.push_back(Piece{&D, sh.first_row, cnt, 0, 0, false, {0, 0, 0, 0, 0, 0, 0, 0}});
                Piece &P = pieces.back();
                ShadowView sv;
                bool have_sv = false, built = false, built2 = false;
                if (cnt == sh.n_rows) PQV_TRY(sweep_shadow(D, mds, sh.d_data, cnt, &sv, &have_sv, &built, &sh));
                PQV_TRY(assign_dispatch(D, sh.d_data, cnt, dim, D.d_centroids.p, n_clusters, D.d_assign.p, &P.path, true,
                                        have_sv ? &sv : nullptr, &P.kind, &built2));
                P.built = built || built2;
            }
            // read-backs only now: a copy into pageable memory holds the host until it is done, and the other shards' sweeps
            // should be running by then
            bool first = true;
            for (Piece &P : pieces) {
                DeviceState &D = *P.D;
                DevGuard guard(D.dev);
                if (P.path == ASSIGN_TC)
                    CU_TRY(cudaMemcpyAsync(P.h_counts, D.tc_u32.p + TC_COUNTS_OFFSET, sizeof P.h_counts, cudaMemcpyDeviceToHost, D.stream));
                CU_TRY(cudaMemcpyAsync(out_assign + P.off, D.d_assign.p, P.cnt * 4, cudaMemcpyDeviceToHost, D.stream));
                CU_TRY(cudaStreamSynchronize(P.D->stream));
                record_assign_timing(ctx, *P.D, P.path, P.cnt, P.h_counts, first, P.kind, P.built);
                first = false;
            }
            if (out_sizes)
                for (u64 i = 0; i < n; ++i) out_sizes[out_assign[i]]++;
            return PQV_OK;
        }
    }
    // host rows are streamed in bounded pieces so scratch stays small
    const u64 piece = rows ? std::max<u64>(1, (u64)(256u << 20) / ((u64)dim * 4)) : n;
    for (u64 off = 0; off < n; off += piece) {
        const u64 cnt = std::min<u64>(piece, n - off);
        DeviceState *D = nullptr;
        const float *d_rows = nullptr;
        Dataset *rds = nullptr;
        PQV_TRY(resolve_rows(ctx, handle, rows ? rows + off * dim : nullptr, rows ? cnt : n, dim, &D, &d_rows, &rds));
        DevGuard guard(D->dev);
        if (!rows) d_rows += off * dim;
        PQV_TRY(D->d_centroids.ensure((size_t)n_clusters * dim));
        if (off == 0)
            CU_TRY(cudaMemcpyAsync(D->d_centroids.p, centroids, (size_t)n_clusters * dim * 4, cudaMemcpyHostToDevice, D->stream));
        PQV_TRY(D->d_assign.ensure(cnt));
        int path = 0, kind = 0;
        ShadowView sv;
        bool have_sv = false, built = false, built2 = false;
        if (!rows) PQV_TRY(sweep_shadow(*D, rds, d_rows, cnt, &sv, &have_sv, &built));
        PQV_TRY(assign_dispatch(*D, d_rows, cnt, dim, D->d_centroids.p, n_clusters, D->d_assign.p, &path, true,
                                have_sv ? &sv : nullptr, &kind, &built2));
        uint32_t h_counts[2] = {0, 0};
        if (path == ASSIGN_TC)
            CU_TRY(cudaMemcpyAsync(h_counts, D->tc_u32.p + TC_COUNTS_OFFSET, sizeof h_counts, cudaMemcpyDeviceToHost, D->stream));
        CU_TRY(cudaMemcpyAsync(out_assign + off, D->d_assign.p, cnt * 4, cudaMemcpyDeviceToHost, D->stream));
        CU_TRY(cudaStreamSynchronize(D->stream));
        record_assign_timing(ctx, *D, path, cnt, h_counts, off == 0, kind, built || built2);
    }
    if (out_sizes)
        for (u64 i = 0; i < n; ++i) out_sizes[out_assign[i]]++;
    return PQV_OK;
}

int pqv_last_assign_timing(pqv_ctx *ctx, pqv_assign_timing *out) {
    if (!ctx || !out) return fail(PQV_EINVAL, "null argument");
    *out = ctx->last_assign;
    return PQV_OK;
}

int pqv_bench_assign(pqv_ctx *ctx, uint64_t handle, uint64_t n, const float *centroids, uint32_t n_clusters, uint32_t iters,
                     pqv_assign_timing *out, uint32_t *out_assign) {
    if (!ctx || !centroids || !out) return fail(PQV_EINVAL, "null argument");
    if (n_clusters == 0) return fail(PQV_EINVAL, "Cluster count must be > 0");
    if (iters == 0 || n == 0) return fail(PQV_EINVAL, "iters and n must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    DeviceState *D = nullptr;
    const float *d_rows = nullptr;
    Dataset *rds = nullptr;
    PQV_TRY(resolve_rows(ctx, handle, nullptr, n, ds->dim, &D, &d_rows, &rds));
    DevGuard guard(D->dev);
    const uint32_t dim = ds->dim;
    PQV_TRY(D->d_centroids.ensure((size_t)n_clusters * dim));
    PQV_TRY(D->d_assign.ensure(n));
    CU_TRY(cudaMemcpyAsync(D->d_centroids.p, centroids, (size_t)n_clusters * dim * 4, cudaMemcpyHostToDevice, D->stream));
    pqv_assign_timing acc{};
    for (uint32_t it = 0; it < iters; ++it) {
        int path = 0, kind = 0;
        ShadowView sv;
        bool have_sv = false, built = false, built2 = false;
        PQV_TRY(sweep_shadow(*D, rds, d_rows, n, &sv, &have_sv, &built));  // the table's shadow: built by the first sweep only
        PQV_TRY(assign_dispatch(*D, d_rows, n, dim, D->d_centroids.p, n_clusters, D->d_assign.p, &path, true,
                                have_sv ? &sv : nullptr, &kind, &built2));
        uint32_t h_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (path == ASSIGN_TC)
            CU_TRY(cudaMemcpyAsync(h_counts, D->tc_u32.p + TC_COUNTS_OFFSET, sizeof h_counts, cudaMemcpyDeviceToHost, D->stream));
        CU_TRY(cudaStreamSynchronize(D->stream));
        if (path == ASSIGN_TC && getenv("PQV_TRACE"))
            fprintf(stderr, "[pqv trace] assign filter: %u ambiguous rows, %u rows to the exact scan (group store full %u, non-finite %u, "
                            "empty window %u, pair buffer full %u), %u pairs\n",
                    h_counts[0], h_counts[1], h_counts[4], h_counts[5], h_counts[6], h_counts[7], h_counts[2]);
        record_assign_timing(ctx, *D, path, n, h_counts, true, kind, built || built2);
        const pqv_assign_timing &t = ctx->last_assign;
        acc.path = t.path;
        acc.kind = t.kind;
        acc.shadow_ms += t.shadow_ms;  // not averaged: the one-time build, if it fell into this call
        acc.rows = t.rows;
        acc.ambiguous_rows = t.ambiguous_rows;
        acc.overflow_rows = t.overflow_rows;
        acc.prep_ms += t.prep_ms / iters;
        acc.filter_ms += t.filter_ms / iters;
        acc.recheck_ms += t.recheck_ms / iters;
        acc.pair_ms += t.pair_ms / iters;
        acc.total_ms += t.total_ms / iters;
    }
    if (out_assign) {
        CU_TRY(cudaMemcpyAsync(out_assign, D->d_assign.p, n * 4, cudaMemcpyDeviceToHost, D->stream));
        CU_TRY(cudaStreamSynchronize(D->stream));
    }
    ctx->last_assign = acc;
    *out = acc;
    return PQV_OK;
}

int pqv_min_dist_update(pqv_ctx *ctx, uint64_t handle, const float *rows, const uint64_t *row_sel, uint64_t n_sel,
                        uint32_t dim, const float *centroid, int init, float *inout_min_dist) {
    if (!ctx || !centroid || (!inout_min_dist && n_sel)) return fail(PQV_EINVAL, "null argument");
    if (dim == 0) return fail(PQV_EINVAL, "Embedding dimension must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n_sel == 0) return PQV_OK;
    u64 n_table = n_sel;
    if (row_sel) {
        n_table = 0;
        for (u64 i = 0; i < n_sel; ++i) n_table = std::max<u64>(n_table, row_sel[i] + 1);
        if (n_table > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row ids are u32");
    }
    DeviceState *D = nullptr;
    const float *d_rows = nullptr;
    bool gathered = false;
    if (!rows) {
        // a resident table spread over several devices: the selected rows are collected on the first device, in selection
        // order, and swept there as a dense block
        Dataset *mds = find_dataset(ctx, handle);
        if (mds && mds->shards.size() > 1) {
            if (mds->dim != dim) return fail(PQV_EINVAL, "dimension mismatch: dataset has %u, call has %u", mds->dim, dim);
            if (n_table > mds->n_rows) return fail(PQV_EINVAL, "n = %llu exceeds the dataset's %llu rows", (unsigned long long)n_table, (unsigned long long)mds->n_rows);
            if (n_sel > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row ids are u32");
            std::vector<uint32_t> ids32(n_sel);
            for (u64 i = 0; i < n_sel; ++i) ids32[i] = row_sel ? (uint32_t)row_sel[i] : (uint32_t)i;
            D = &ctx->devs[mds->shards[0].di];
            DevGuard g0(D->dev);
            PQV_TRY(D->d_tmp_rows.ensure((size_t)n_sel * dim));
            PQV_TRY(gather_rows_multi(ctx, *mds, ids32.data(), n_sel, *D, D->d_tmp_rows.p));
            d_rows = D->d_tmp_rows.p;
            gathered = true;
        }
    }
    if (!gathered) PQV_TRY(resolve_rows(ctx, handle, rows, n_table, dim, &D, &d_rows));
    DevGuard guard(D->dev);
    const uint32_t *d_ids = nullptr;
    if (row_sel && !gathered) {
        std::vector<uint32_t> ids32(n_sel);
        for (u64 i = 0; i < n_sel; ++i) ids32[i] = (uint32_t)row_sel[i];
        PQV_TRY(D->d_row_ids.ensure(n_sel));
        CU_TRY(cudaMemcpyAsync(D->d_row_ids.p, ids32.data(), n_sel * 4, cudaMemcpyHostToDevice, D->stream));
        CU_TRY(cudaStreamSynchronize(D->stream));  // ids32 is a local
        d_ids = D->d_row_ids.p;
    }
    PQV_TRY(D->d_vec.ensure(dim));
    PQV_TRY(D->d_dist.ensure(n_sel));
    CU_TRY(cudaMemcpyAsync(D->d_vec.p, centroid, (size_t)dim * 4, cudaMemcpyHostToDevice, D->stream));
    if (!init) CU_TRY(cudaMemcpyAsync(D->d_dist.p, inout_min_dist, n_sel * 4, cudaMemcpyHostToDevice, D->stream));
    PQV_TRY(dist_launch(*D, d_rows, d_ids, n_sel, dim, D->d_vec.p, D->d_dist.p, init ? 0 : 1));
    CU_TRY(cudaMemcpyAsync(inout_min_dist, D->d_dist.p, n_sel * 4, cudaMemcpyDeviceToHost, D->stream));
    CU_TRY(cudaStreamSynchronize(D->stream));
    return PQV_OK;
}

int pqv_centroid_rank(pqv_ctx *ctx, const float *centroids, uint32_t n_clusters, uint32_t dim, const float *queries,
                      uint32_t n_queries, uint32_t nprobe, uint32_t *out_cluster_ids, uint32_t *out_nprobe_eff) {
    if (!ctx || !centroids || (n_queries && (!queries || !out_cluster_ids))) return fail(PQV_EINVAL, "null argument");
    if (dim == 0) return fail(PQV_EINVAL, "Embedding dimension must be > 0");
    if (n_clusters == 0) return fail(PQV_EINVAL, "Cluster count must be > 0");
    if (nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");  // src/ivf/search.rs:72
    std::lock_guard<std::mutex> lk(ctx->mu);
    const uint32_t np = std::min(nprobe, n_clusters);  // src/ivf/index.rs:131
    if (out_nprobe_eff) *out_nprobe_eff = np;
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    PQV_TRY(D.d_centroids.ensure((size_t)n_clusters * dim));
    PQV_TRY(D.d_vec.ensure(dim));
    PQV_TRY(D.d_dist.ensure(n_clusters));
    CU_TRY(cudaMemcpyAsync(D.d_centroids.p, centroids, (size_t)n_clusters * dim * 4, cudaMemcpyHostToDevice, D.stream));
    std::vector<float> dist(n_clusters);
    std::vector<uint32_t> idx(n_clusters);
    auto rank_one = [&](uint32_t q) -> int {
        CU_TRY(cudaMemcpyAsync(D.d_vec.p, queries + (size_t)q * dim, (size_t)dim * 4, cudaMemcpyHostToDevice, D.stream));
        // index.rs:138 squared_l2_distance(query, centroid); (q-c)^2 == (c-q)^2 bit for bit
        PQV_TRY(dist_launch(D, D.d_centroids.p, nullptr, n_clusters, dim, D.d_vec.p, D.d_dist.p, 0));
        CU_TRY(cudaMemcpyAsync(dist.data(), D.d_dist.p, (size_t)n_clusters * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
        for (uint32_t i = 0; i < n_clusters; ++i) idx[i] = i;
        // index.rs:143 stable sort, partial_cmp -> Equal for NaN
        merge_sort_stable(idx, [&](uint32_t a, uint32_t b) { return dist[a] < dist[b]; });
        memcpy(out_cluster_ids + (size_t)q * np, idx.data(), (size_t)np * 4);
        return PQV_OK;
    };
    // batches: all nq x C distances in one launch (exact order), one CTA per query ranks on the device
    uint32_t cp2 = 32;
    while (cp2 < n_clusters) cp2 <<= 1;
    const bool batched = n_queries >= 2 && (size_t)cp2 * 8 <= 128 * 1024 && getenv("PQV_RANK_BATCH_OFF") == nullptr;
    if (!batched) {
        for (uint32_t q = 0; q < n_queries; ++q) PQV_TRY(rank_one(q));
        return PQV_OK;
    }
    const bool vec4 = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(D.d_centroids.p) & 15) == 0);
    const size_t smem = (size_t)((dim + 3u) & ~3u) * 4 + (size_t)SCAN_WARPS * pqv::TileCfg<0, true>::TILE_FLOATS * 4;
    const u64 NG = ((u64)n_clusters + 31) / 32;
    auto *k_vec = pqv::l2_dist_batch_kernel<true, SCAN_WARPS>;
    auto *k_sca = pqv::l2_dist_batch_kernel<false, SCAN_WARPS>;
    PQV_TRY(ensure_dyn_smem(vec4 ? reinterpret_cast<const void *>(k_vec) : reinterpret_cast<const void *>(k_sca), smem));
    PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(pqv::rank_batch_kernel), (size_t)cp2 * 8));
    constexpr uint32_t CHUNK = 16384;  // queries per launch (gridDim.y limit, scratch size)
    const uint32_t chunk_max = std::min(n_queries, CHUNK);
    PQV_TRY(D.d_tmp_rows.ensure((size_t)chunk_max * dim));
    PQV_TRY(D.d_dist.ensure((size_t)chunk_max * n_clusters));
    PQV_TRY(D.d_assign.ensure((size_t)chunk_max * np));
    PQV_TRY(D.d_row_ids.ensure(chunk_max));
    std::vector<uint32_t> flags(chunk_max);
    std::vector<uint32_t> redo;
    for (uint32_t q0 = 0; q0 < n_queries; q0 += CHUNK) {
        const uint32_t nq = std::min(CHUNK, n_queries - q0);
        CU_TRY(cudaMemcpyAsync(D.d_tmp_rows.p, queries + (size_t)q0 * dim, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, D.stream));
        const dim3 grid((uint32_t)((NG + SCAN_WARPS - 1) / SCAN_WARPS), nq);
        if (vec4) k_vec<<<grid, SCAN_WARPS * 32, smem, D.stream>>>(D.d_centroids.p, n_clusters, dim, D.d_tmp_rows.p, D.d_dist.p);
        else k_sca<<<grid, SCAN_WARPS * 32, smem, D.stream>>>(D.d_centroids.p, n_clusters, dim, D.d_tmp_rows.p, D.d_dist.p);
        pqv::rank_batch_kernel<<<nq, 1024, (size_t)cp2 * 8, D.stream>>>(D.d_dist.p, n_clusters, cp2, np, D.d_assign.p, D.d_row_ids.p);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(out_cluster_ids + (size_t)q0 * np, D.d_assign.p, (size_t)nq * np * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaMemcpyAsync(flags.data(), D.d_row_ids.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
        for (uint32_t i = 0; i < nq; ++i)
            if (flags[i]) redo.push_back(q0 + i);
    }
    for (uint32_t q : redo) PQV_TRY(rank_one(q));  // NaN distances: the reference comparator on the host
    return PQV_OK;
}

// ---- streaming top-k (VectorTopKExec) --------------------------------------------------------------
int pqv_topk_stream_begin(pqv_ctx *ctx, uint32_t dim, const float *query, uint32_t k, uint32_t flags, uint64_t *out_stream) {
    if (!ctx || !query || !out_stream) return fail(PQV_EINVAL, "null argument");
    PQV_TRY(check_topk_args(k, dim, flags));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    StreamState *s = ctx->stream_cache ? ctx->stream_cache : new StreamState();
    const bool recycled = ctx->stream_cache != nullptr;
    ctx->stream_cache = nullptr;
    s->dim = dim;
    s->k = k;
    s->flags = flags;
    s->rows_pushed = s->ent_rows_bound = 0;
    s->cur = s->carry_cur = 0;
    s->host_entrants.clear();
    s->any = false;
    s->log_ok = true;
    int rc = s->d_query.ensure(dim);
    if (!rc) rc = s->carry[0].ensure(PQV_MAX_K);
    if (!rc) rc = s->carry[1].ensure(PQV_MAX_K);
    if (!rc) rc = s->ent_acc.ensure((size_t)(1u << 22) + 1);
    if (rc) {
        stream_state_free(s);
        return rc;
    }
    CU_TRY(cudaMemcpyAsync(s->d_query.p, query, (size_t)dim * 4, cudaMemcpyHostToDevice, D.stream));
    CU_TRY(cudaMemsetAsync(s->carry[0].p, 0xFF, (size_t)PQV_MAX_K * 8, D.stream));
    CU_TRY(cudaMemsetAsync(s->ent_acc.p, 0, 8, D.stream));
    CU_TRY(cudaStreamSynchronize(D.stream));
    if (!recycled) {
        for (auto &ev : s->scan_done) CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&s->copy_done, cudaEventDisableTiming));
    }
    const u64 h = ctx->next_handle++;
    ctx->streams[h] = s;
    *out_stream = h;
    return PQV_OK;
}

static int stream_drain(DeviceState &D, StreamState *s) {
    bool overflow = false;
    PQV_TRY(fetch_entrants(D, s->ent_acc.p, (uint32_t)(s->ent_acc.cap - 1), s->host_entrants, &overflow));
    if (overflow) return fail(PQV_ECUDA, "stream entrant accumulator overflow (internal bound violated)");
    CU_TRY(cudaMemsetAsync(s->ent_acc.p, 0, 8, D.stream));
    s->ent_rows_bound = 0;
    return PQV_OK;
}

static int stream_push_impl(pqv_ctx *ctx, uint64_t stream, const void *values, uint64_t n_rows, bool f64) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->streams.find(stream);
    if (it == ctx->streams.end()) return fail(PQV_EHANDLE, "unknown stream handle %llu", (unsigned long long)stream);
    StreamState *s = it->second;
    if (n_rows == 0) return PQV_OK;
    if (!values) return fail(PQV_EINVAL, "values is null");
    if (s->rows_pushed + n_rows > 0xFFFFFFFFull) return fail(PQV_ELIMIT, "row positions are u32");
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    const int cur = s->cur;
    const size_t elems = (size_t)n_rows * s->dim;
    // every row could be an entrant in the worst case: drain the accumulator before it can overflow
    if (s->ent_rows_bound + n_rows > s->ent_acc.cap - 1) {
        PQV_TRY(stream_drain(D, s));
        if (n_rows > s->ent_acc.cap - 1) {
            CU_TRY(cudaStreamSynchronize(D.stream));
            PQV_TRY(s->ent_acc.ensure((size_t)n_rows + 1));
            CU_TRY(cudaMemsetAsync(s->ent_acc.p, 0, 8, D.stream));
        }
    }
    // staging[cur] was last read by the scan two pushes ago
    if (s->any) CU_TRY(cudaStreamWaitEvent(D.copy_stream, s->scan_done[cur], 0));
    if (elems > s->staging[cur].cap) {
        CU_TRY(cudaEventSynchronize(s->scan_done[cur]));
        PQV_TRY(s->staging[cur].ensure(elems));
    }
    if (f64) {
        CU_TRY(cudaStreamSynchronize(D.copy_stream));
        CU_TRY(cudaStreamSynchronize(D.stream));
        PQV_TRY(s->staging64.ensure(elems));
        CU_TRY(cudaMemcpyAsync(s->staging64.p, values, elems * 8, cudaMemcpyHostToDevice, D.copy_stream));
        pqv::narrow_f64_kernel<<<D.sm_count * 4, 256, 0, D.copy_stream>>>(s->staging64.p, s->staging[cur].p, elems);
        CU_TRY(cudaGetLastError());
    } else {
        // a pageable batch (ordinary Arrow buffers) goes through the multi-lane pinned staging of append_staged: 30-40 GB/s
        // instead of the driver's 11-17 GB/s bounce copy; a page-locked one is DMA'd directly
        bool staged = false;
        const size_t bytes = elems * 4;
        if (bytes >= APPEND_STAGED_MIN && getenv("PQV_APPEND_DIRECT") == nullptr) {
            cudaPointerAttributes pa{};
            const bool pinned = cudaPointerGetAttributes(&pa, values) == cudaSuccess && pa.type != cudaMemoryTypeUnregistered;
            cudaGetLastError();
            if (!pinned) {
                if (s->any) CU_TRY(cudaEventSynchronize(s->scan_done[cur]));  // the scan that last read staging[cur]
                PQV_TRY(append_staged(D, reinterpret_cast<unsigned char *>(s->staging[cur].p),
                                      reinterpret_cast<const unsigned char *>(values), bytes));
                staged = true;
            }
        }
        if (!staged) CU_TRY(cudaMemcpyAsync(s->staging[cur].p, values, bytes, cudaMemcpyHostToDevice, D.copy_stream));
    }
    CU_TRY(cudaEventRecord(s->copy_done, D.copy_stream));
    CU_TRY(cudaStreamWaitEvent(D.stream, s->copy_done, 0));
    const int order = (s->flags & PQV_SUM_SEQ) ? 1 : 0;
    const int cc = s->carry_cur;
    float *log_at = nullptr;
    if (s->log_ok && s->rows_pushed + n_rows <= STREAM_LOG_MAX_ROWS) {
        if (s->rows_pushed + n_rows > s->dist_log.cap) {  // grow (geometric), keeping what is logged so far
            DevBuf<float> bigger;
            const size_t want = std::max<size_t>((size_t)(s->rows_pushed + n_rows) * 2, (size_t)1 << 22);
            if (bigger.ensure(want) == PQV_OK) {
                if (s->rows_pushed)
                    CU_TRY(cudaMemcpyAsync(bigger.p, s->dist_log.p, (size_t)s->rows_pushed * 4, cudaMemcpyDeviceToDevice, D.stream));
                CU_TRY(cudaStreamSynchronize(D.stream));
                s->dist_log.release();
                s->dist_log = bigger;
            } else {
                s->log_ok = false;  // no room for the log: the stream still answers, NaN distances from the entrants alone
            }
        }
        if (s->log_ok) log_at = s->dist_log.p + s->rows_pushed;
    } else {
        s->log_ok = false;
    }
    PQV_TRY(enqueue_scan(ctx, D, s->staging[cur].p, nullptr, n_rows, s->dim, s->d_query.p, s->k, order,
                         (uint32_t)s->rows_pushed, s->carry[cc].p, s->carry[cc ^ 1].p, s->ent_acc.p,
                         (uint32_t)(s->ent_acc.cap - 1), false, nullptr, nullptr, log_at));
    CU_TRY(cudaEventRecord(s->scan_done[cur], D.stream));
    s->carry_cur ^= 1;
    s->cur ^= 1;
    s->any = true;
    s->rows_pushed += n_rows;
    s->ent_rows_bound += n_rows;
    // the caller's buffer is only borrowed for this call: wait for the H2D copy (not for the scan)
    CU_TRY(cudaEventSynchronize(s->copy_done));
    return PQV_OK;
}

int pqv_topk_stream_push(pqv_ctx *ctx, uint64_t stream, const float *values, uint64_t n_rows) {
    return stream_push_impl(ctx, stream, values, n_rows, false);
}
int pqv_topk_stream_push_f64(pqv_ctx *ctx, uint64_t stream, const double *values, uint64_t n_rows) {
    return stream_push_impl(ctx, stream, values, n_rows, true);
}

int pqv_topk_stream_finish(pqv_ctx *ctx, uint64_t stream, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!ctx || !out_row_idx || !out_dist || !out_count) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->streams.find(stream);
    if (it == ctx->streams.end()) return fail(PQV_EHANDLE, "unknown stream handle %llu", (unsigned long long)stream);
    StreamState *s = it->second;
    DeviceState &D = ctx->devs[0];
    DevGuard guard(D.dev);
    int rc = PQV_OK;
    size_t cnt = 0;
    if (s->flags & PQV_TIES_BY_POSITION) {
        std::vector<u64> keys(PQV_MAX_K);
        cudaError_t e = cudaMemcpyAsync(keys.data(), s->carry[s->carry_cur].p, (size_t)PQV_MAX_K * 8, cudaMemcpyDeviceToHost, D.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
        if (e != cudaSuccess) rc = fail(PQV_ECUDA, "stream finish copy failed: %s", cudaGetErrorString(e));
        else cnt = emit_by_position(keys, RowMap{}, s->k, s->flags, out_row_idx, out_dist);
    } else {
        rc = stream_drain(D, s);
        if (!rc && s->log_ok && any_nan_key(s->host_entrants.data(), s->host_entrants.size())) {
            // a NaN distance: the reference loop over every pushed row, from the distance log (see topk_one)
            std::vector<float> dist(s->rows_pushed);
            cudaError_t e = cudaMemcpyAsync(dist.data(), s->dist_log.p, (size_t)s->rows_pushed * 4, cudaMemcpyDeviceToHost, D.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
            if (e != cudaSuccess) rc = fail(PQV_ECUDA, "stream distance log copy failed: %s", cudaGetErrorString(e));
            else
                cnt = replay_ordered(
                    dist.size(), [&](size_t i) { return ReplayItem{dist[i], (uint32_t)i}; }, s->k, s->flags, out_row_idx, out_dist);
        } else if (!rc) {
            cnt = replay_reference_heap(s->host_entrants, RowMap{}, s->k, s->flags, out_row_idx, out_dist);
        }
    }
    cudaStreamSynchronize(D.stream);
    cudaStreamSynchronize(D.copy_stream);
    ctx->streams.erase(it);
    if (ctx->stream_cache) stream_state_free(s);  // one retired set is kept
    else ctx->stream_cache = s;
    if (rc) return rc;
    *out_count = (uint32_t)cnt;
    return PQV_OK;
}

}  // extern "C"

#include "pqv_ivf_impl.cuh"
#include "pqv_adist_impl.cuh"
