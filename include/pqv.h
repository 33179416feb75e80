/*
 * pqv.h -- C ABI of libpqv.so: the B200 (sm_100a) replacement for pq-vector's brute-force
 * squared-L2 / top-k / IVF-assign hot path.
 *
 * The reference (XiangpengHao/pq-vector @ 808b90d) has no FFI boundary of its own: the path is
 * in-process Rust.  These are the entry points a Rust `extern "C"` block would bind at the call
 * sites listed beside each function (all paths relative to the reference root); INTEGRATION.md
 * shows the binding.  Plain pointers and sizes only; no CUDA or torch types cross this boundary.
 *
 * Conventions
 *  - every function returns 0 (PQV_OK) or a PQV_E* code; pqv_last_error() gives the message of
 *    the last failure on the calling thread (valid until the next call on that thread).
 *  - all input pointers are HOST pointers, borrowed for the duration of the call; all outputs are
 *    caller-allocated host buffers.  The library never retains or frees caller memory.
 *  - row ids are u32 everywhere (reference: src/ivf/index.rs:13, src/ivf/search.rs:43).
 *  - results are bit-identical to the reference loops: same f32 summation order, no FMA, and the
 *    same bounded-BinaryHeap / stable-sort tie behaviour.
 *  - there is no CPU fallback: without a CUDA device pqv_init fails with PQV_ENODEV.
 */
#ifndef PQV_H
#define PQV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PQV_API __attribute__((visibility("default")))
#else
#define PQV_API
#endif

#define PQV_OK        0
#define PQV_EINVAL    1  /* bad argument (k == 0, dim mismatch, null pointer, ...)            */
#define PQV_ENODEV    2  /* no usable CUDA device / device id out of range                    */
#define PQV_ECUDA     3  /* a CUDA runtime call failed; message carries cudaGetErrorString    */
#define PQV_ENOMEM    4  /* device or pinned-host allocation failed                            */
#define PQV_EHANDLE   5  /* unknown dataset / stream handle                                    */
#define PQV_ELIMIT    6  /* over an implementation limit (k > PQV_MAX_K, dim > PQV_MAX_DIM)    */

#define PQV_MAX_K    1024u   /* per-query k handled by the in-kernel selection; pqv_l2_topk,
                                 pqv_l2_topk_gather and pqv_ivf_search also take larger k (every
                                 candidate's distance is logged and the reference loop replayed on
                                 the host: exact, slower), the other entry points return PQV_ELIMIT */
#define PQV_MAX_DIM  16384u  /* query staged in shared memory                                 */

/* flags for the top-k calls */
#define PQV_SUM_UNROLL4  0u  /* src/ivf/index.rs:461-480   sum += ((d0^2+d1^2)+d2^2)+d3^2 + scalar tail   */
#define PQV_SUM_SEQ      1u  /* src/df_vector/exec.rs:529-533   dist += diff*diff, sequential               */
#define PQV_SQRT         2u  /* apply sqrt to the kept distances BEFORE the final stable sort and return
                                them (src/ivf/search.rs:129-140); without it squared distances are sorted
                                and returned (src/df_vector/exec.rs:269-274)                                */
#define PQV_ROW_ORDER    8u  /* pqv_ivf_search_batch only: candidates are visited in ascending row order (the
                                RowSelection order of VectorTopKExec, src/df_vector/access.rs:107-176) instead of
                                list-rank order (src/ivf/index.rs:57-63); matters among bit-equal distances only */
#define PQV_TIES_BY_POSITION 4u /* skip the reference heap replay: order strictly by (distance, candidate
                                   position).  Differs from the reference only among bit-equal distances.  */

typedef struct pqv_ctx pqv_ctx;

/* Lifetime.  device_ids == NULL && n_devices == 0 selects the current device.  With n_devices > 1
 * a dataset's rows are split into contiguous ranges, one per device, and ONE host
 * process drives all of them through the calls below: pqv_l2_topk (single queries: every shard scans
 * its rows; batches: one tensor-core pass per shard, each from its own host thread, keys merged),
 * pqv_l2_topk_gather / pqv_ivf_search(_batch) / pqv_vector_topk_indexed(_batch) (candidates split by
 * owning shard, keys moved back to the caller's sequence positions before the heap replay),
 * pqv_kmeans_assign over the resident table, pqv_ivf_build (sample gathered to the first device,
 * rows assigned where they live), pqv_array_distance(_topk,_topk_filtered) (every shard fills its
 * slice of the column / selects its own k smallest, merged by (distance, row)) and
 * pqv_dataset_read(_rows) -- all with the results of a single device, bit for bit
 * (tests/test_gpu_multi_device.py).  The per-rank entry points further down
 * (*_candidates, *_keys, *_p2p) are the other form: one process per GPU, each with its own context.
 * pqv_kmeans_train and pqv_min_dist_update over a resident table collect the rows they need on the
 * first device and run there. */
PQV_API int  pqv_init(pqv_ctx **out, const int *device_ids, int n_devices);
PQV_API void pqv_destroy(pqv_ctx *ctx);
PQV_API const char *pqv_last_error(void);
PQV_API const char *pqv_version(void);
PQV_API int  pqv_device_count(pqv_ctx *ctx);

/* ---- dataset residency ---------------------------------------------------------------------
 * Replaces the per-query Parquet re-read of src/ivf/search.rs:155-244 (read_embeddings_for_rows)
 * and the `Embeddings` copy built in src/ivf/parquet.rs:281-292: the dense row-major N x dim f32
 * block (the child values buffer of the Arrow List<Float32> column) is appended once, batch by
 * batch, and stays in HBM.  Row ids are assigned in append order starting at 0. */
PQV_API int pqv_dataset_create(pqv_ctx *ctx, uint32_t dim, uint64_t n_rows_hint, uint64_t *out_handle);
PQV_API int pqv_dataset_append(pqv_ctx *ctx, uint64_t handle, const float *values, uint64_t n_rows);
PQV_API int pqv_dataset_rows(pqv_ctx *ctx, uint64_t handle, uint64_t *out_rows, uint32_t *out_dim);
PQV_API int pqv_dataset_drop(pqv_ctx *ctx, uint64_t handle);
/* bench/test helper: fill rows [0, n_rows) on the device with the counter-based uniform[0,1)
 * generator (distribution of benches/bench_util.rs:29-41; stream documented in DESIGN.md); local row r
 * gets row (stream_first_row + r) of the seed's stream, so ranks can hold slices of one global table. */
PQV_API int pqv_dataset_fill_synthetic(pqv_ctx *ctx, uint64_t handle, uint64_t n_rows, uint64_t seed,
                                       uint64_t stream_first_row);
/* read rows back (tests, and fetching the k winners) */
PQV_API int pqv_dataset_read(pqv_ctx *ctx, uint64_t handle, uint64_t first_row, uint64_t n_rows, float *out);
/* rows row_ids[0 .. n_ids) in that order (gathered on the device, one copy back): the k winners of a search, or a rank's
 * share of the k-means training sample (src/ivf/index.rs:234-239) */
PQV_API int pqv_dataset_read_rows(pqv_ctx *ctx, uint64_t handle, const uint32_t *row_ids, uint64_t n_ids, float *out);

/* ---- brute-force and gathered top-k ---------------------------------------------------------
 * pqv_l2_topk        replaces the re-rank loop of src/ivf/search.rs:112-141 when every row is a
 *                    candidate (nprobe >= n_clusters), one call per query.
 * pqv_l2_topk_gather replaces the same loop for an IVF candidate list (`rows_to_check`,
 *                    src/ivf/search.rs:100): row_ids in candidate order.
 * Output i of query q is at out_*[q*k + i]; out_count[q] <= k results are valid, ascending. */
PQV_API int pqv_l2_topk(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k,
                uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count);
/* With n_queries >= 4 (dim % 4 == 0) pqv_l2_topk answers the whole batch in one tensor-core pass
 * over the table (per shard, when the table is spread over several devices) (tcgen05 tf32 filter + exact re-rank of the survivors); every query's output is
 * still identical to its own single-query call.  PQV_BATCH=off in the environment disables the batched pass. */
typedef struct {
    uint32_t queries;      /* batch size of the last pqv_l2_topk call (0: batched pass not used)                        */
    uint32_t declined;     /* 1: batch abandoned (non-finite inputs, candidate buffers full) -> single-query scans      */
    uint32_t tie_queries;  /* queries whose order hinges on the reference heap layout: the heap is replayed for them    */
    uint32_t tie_batched;  /* of those: resolved together from one pass over the sample prefix (PQV_TIE_BATCH=off: none) */
    uint64_t rows;
    uint64_t sample_rows;  /* rows of the threshold pass                                                               */
    uint64_t candidates;   /* (row, query) pairs re-evaluated exactly                                                  */
    double prep_ms;        /* query rounding + row norms                                                               */
    double sample_ms;      /* tcgen05 pass over the sample rows + per-query threshold selection                        */
    double filter_ms;      /* tcgen05 pass over all rows                                                               */
    double rerank_ms;      /* exact distances of the candidates + per-query top-k selection                            */
    double total_ms;
} pqv_batch_timing;
PQV_API int pqv_last_batch_timing(pqv_ctx *ctx, pqv_batch_timing *out);
PQV_API int pqv_l2_topk_gather(pqv_ctx *ctx, uint64_t handle, const float *query, const uint32_t *row_ids,
                       uint64_t n_ids, uint32_t k, uint32_t flags, uint32_t *out_row_idx, float *out_dist,
                       uint32_t *out_count);

/* ---- streaming top-k over RecordBatches ------------------------------------------------------
 * Replaces VectorTopKExec::topk_from_batches / update_topk_heap (src/df_vector/exec.rs:257-277,
 * 457-482): begin with the query, push each batch's dense values buffer (n_rows x dim f32, the
 * caller has already dropped null / wrong-length rows as exec.rs:496-498, 526-528 do), finish to
 * get the k winners as indices into the pushed row sequence (push order).  Batches are copied
 * host->device on a copy stream overlapped with the scan of the previous batch. */
PQV_API int pqv_topk_stream_begin(pqv_ctx *ctx, uint32_t dim, const float *query, uint32_t k, uint32_t flags,
                          uint64_t *out_stream);
PQV_API int pqv_topk_stream_push(pqv_ctx *ctx, uint64_t stream, const float *values, uint64_t n_rows);
PQV_API int pqv_topk_stream_push_f64(pqv_ctx *ctx, uint64_t stream, const double *values, uint64_t n_rows);
PQV_API int pqv_topk_stream_finish(pqv_ctx *ctx, uint64_t stream, uint32_t *out_row_idx, float *out_dist,
                           uint32_t *out_count);

/* ---- k-means / IVF build pieces ---------------------------------------------------------------
 * pqv_kmeans_assign   replaces the assignment sweeps src/ivf/index.rs:193-201 (final, all N rows)
 *                     and :398-424 (Lloyd): out_assign[i] = argmin_c dist(row i, centroid c), first
 *                     minimum wins (index.rs:251).  rows == NULL scans the resident dataset
 *                     `handle`; otherwise `rows` (n x dim, host) is streamed through.
 *                     out_sizes (may be NULL) = per-cluster counts (index.rs:421).
 * pqv_min_dist_update replaces one k-means++ sweep src/ivf/index.rs:344-370: for each selected row
 *                     s: d = dist(row, centroid); init ? slot = d : (d < slot ? slot = d).  The
 *                     f32 sums of index.rs:359-370 stay with the caller (thread-count dependent).
 * pqv_centroid_rank   replaces find_closest_centroids src/ivf/index.rs:130-149: per query the
 *                     min(nprobe, C) closest cluster ids, stable ascending. Returns nprobe_eff. */
PQV_API int pqv_kmeans_assign(pqv_ctx *ctx, uint64_t handle, const float *rows, uint64_t n, uint32_t dim,
                      const float *centroids, uint32_t n_clusters, uint32_t *out_assign, uint64_t *out_sizes);
PQV_API int pqv_min_dist_update(pqv_ctx *ctx, uint64_t handle, const float *rows, const uint64_t *row_sel,
                        uint64_t n_sel, uint32_t dim, const float *centroid, int init, float *inout_min_dist);
PQV_API int pqv_centroid_rank(pqv_ctx *ctx, const float *centroids, uint32_t n_clusters, uint32_t dim,
                      const float *queries, uint32_t n_queries, uint32_t nprobe, uint32_t *out_cluster_ids,
                      uint32_t *out_nprobe_eff);


/* ---- IVF index: build / blob / search ------------------------------------------------------------
 * Host-side mirror of the reference's IVF layer with every distance on the GPU.
 * pqv_ivf_build      build_ivf_index + k_means, src/ivf/index.rs:152-214, 323-457, over a resident dataset:
 *                    sizing rules (:161-174, :332), k-means++ (sweeps = pqv_min_dist_update's kernel; the f32
 *                    partial sums use `sum_workers` chunks, 0 = hardware threads, as :259-265, :356-370),
 *                    Lloyd (assign kernel + exact-order centroid update :436-453), final assignment of all rows
 *                    (:189-206).  The RNG stream is not rand's StdRng.
 * pqv_ivf_to_bytes / pqv_ivf_from_bytes   IvfIndex::to_bytes / from_bytes, src/ivf/index.rs:65-128 (same bytes).
 * pqv_ivf_candidate_rows  IvfIndex::candidate_rows, src/ivf/index.rs:57-63.
 * pqv_ivf_search     TopkBuilder::topk, src/ivf/search.rs:83-142, with the index and the table resident in HBM
 *                    (replaces read_index_from_parquet + read_embeddings_for_rows + the re-rank loop). */
PQV_API int pqv_ivf_build(pqv_ctx *ctx, uint64_t handle, uint32_t n_clusters_or_0, uint32_t max_iters, uint64_t seed,
                          uint32_t sum_workers, uint64_t *out_index);
/* The pieces of pqv_ivf_build for a table that is sharded over several processes: training stays on one
 * GPU, only the final assignment of all rows is sharded.
 * pqv_ivf_sample_rows  (pure host) the sizing rules of src/ivf/index.rs:161-174 and the training-sample draw of :222-242 for
 *                      a table of n_rows: *out_clusters = C, out_rows[0 .. *out_n) = the global row ids pqv_ivf_build would
 *                      train on, in its order.  out_rows == NULL just reports the sizes.
 * pqv_kmeans_train     k_means (src/ivf/index.rs:323-457) over ALL rows of the dataset `handle` (the gathered sample):
 *                      out_centroids[n_clusters * dim]; identical to what pqv_ivf_build computes from the same sample rows.
 * Every rank then calls pqv_kmeans_assign on its slice with these centroids; the lists are the per-rank lists concatenated in
 * rank order (ascending row ids, index.rs:202-206). */
PQV_API int pqv_ivf_sample_rows(uint64_t n_rows, uint32_t n_clusters_or_0, uint64_t seed, uint32_t *out_rows, uint64_t cap,
                                uint64_t *out_n, uint32_t *out_clusters);
PQV_API int pqv_kmeans_train(pqv_ctx *ctx, uint64_t handle, uint32_t n_clusters, uint32_t max_iters, uint64_t seed,
                             uint32_t sum_workers, float *out_centroids, uint32_t *out_iters);
PQV_API int pqv_ivf_build_stats(pqv_ctx *ctx, uint64_t index, uint32_t *out_lloyd_iters, double *out_ms4);
PQV_API int pqv_ivf_from_bytes(pqv_ctx *ctx, const uint8_t *bytes, uint64_t len, uint64_t *out_index);
PQV_API int pqv_ivf_to_bytes(pqv_ctx *ctx, uint64_t index, uint8_t *out, uint64_t cap, uint64_t *out_len);
PQV_API int pqv_ivf_info(pqv_ctx *ctx, uint64_t index, uint32_t *out_dim, uint32_t *out_clusters, uint64_t *out_ids);
PQV_API int pqv_ivf_drop(pqv_ctx *ctx, uint64_t index);
PQV_API int pqv_ivf_candidate_rows(pqv_ctx *ctx, uint64_t index, const float *query, uint32_t nprobe,
                                   uint32_t *out_rows, uint64_t cap, uint64_t *out_n);
PQV_API int pqv_ivf_search(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k,
                           uint32_t nprobe, uint32_t flags, uint32_t *out_row_idx, float *out_dist,
                           uint32_t *out_count);

/* ---- one process per GPU -----------------------------------------------------
 * Each rank owns a contiguous slice of the rows.  pqv_l2_topk_candidates scans the rank's resident
 * slice and returns the heap-entrant candidate keys (bits(squared distance) << 32 | global position,
 * global position = pos_base + local row): a superset of every row the reference's BinaryHeap
 * (src/ivf/search.rs:115-127) would admit while walking this slice.  The ranks exchange these few KB
 * with ONE all-gather and each calls pqv_replay_candidates on the union, which replays the reference
 * loop in global position order -> the same bit-exact result on every rank.
 * If more than `cap` keys exist the call returns PQV_ELIMIT with *out_count = the number needed. */
PQV_API int pqv_l2_topk_candidates(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags,
                                   uint32_t pos_base, uint64_t *out_keys, uint64_t cap, uint64_t *out_count);
PQV_API int pqv_replay_candidates(const uint64_t *keys, uint64_t n_keys, const uint32_t *row_ids, uint32_t k,
                                  uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count);

/* The same exchange over NVLink peer memory instead of a collective library: each rank allocates an exchange buffer
 * (pqv_peer_exchange_create -> a 64-byte CUDA IPC handle), the handles travel once through whatever channel the host
 * side has (torch.distributed here), pqv_peer_exchange_open maps the peers' buffers.  pqv_l2_topk_candidates_p2p then
 * runs scan -> filter -> a tail kernel that WRITES this rank's candidates into every peer's buffer and publishes a
 * sequence flag -> a kernel that waits for the peers' flags and packs the union of the live keys into page-locked host
 * memory: no collective launch, no copy call and no extra host round trip per query.  All ranks must call it the same number of times, in the same order.
 * out_keys[0 .. *out_count) = the union (no particular order; entrants of later ranks that the global heap cannot admit --
 * distance not below the smallest final k-th distance of the ranks before them -- are already dropped); *out_overflow = 1
 * (nothing written) when some rank had more than cap_keys candidates or a NaN distance showed up -- every rank sees the same
 * data, so all of them fall back to pqv_l2_topk_candidates together. */
PQV_API int pqv_peer_exchange_create(pqv_ctx *ctx, uint32_t world, uint32_t rank, uint32_t cap_keys, uint8_t *out_handle64);
PQV_API int pqv_peer_exchange_open(pqv_ctx *ctx, const uint8_t *handles /* world x 64 bytes, rank order */);
PQV_API int pqv_l2_topk_candidates_p2p(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags,
                                       uint32_t pos_base, uint64_t *out_keys, uint64_t cap_total, uint64_t *out_count,
                                       uint32_t *out_overflow);
/* One rank's whole sharded search in one call: the exchange above followed by the reference heap replay over the union
 * (= pqv_l2_topk_candidates_p2p + pqv_replay_candidates; TopkBuilder::search's re-rank, src/ivf/search.rs:112-141, with the
 * rows spread over the ranks).  Every rank passes the same query and receives the same bit-exact (row_idx, distance) list;
 * row_idx = pos_base + local row.  *out_overflow = 1 (nothing else written) as above. */
PQV_API int pqv_l2_topk_p2p(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags, uint32_t pos_base,
                            uint32_t *out_row_idx, float *out_dist, uint32_t *out_count, uint32_t *out_overflow);
/* The batched form (config C5 with one process per GPU): this rank's tensor-core pass over its slice (pqv_l2_topk_batch_keys
 * below), the per-query key lists of all ranks exchanged over the same peer buffers, merged on the host
 * (pqv_merge_batch_keys), and the queries the merge cannot decide replayed from their candidates
 * (pqv_l2_topk_batch_tie_candidates + the exchange + pqv_replay_candidates) -- one call, the same bit-exact results on every
 * rank (out_*[q*k + i], out_count[q]; *out_replayed = queries that needed the heap replay).  The exchange slots must hold
 * n_queries * (k + 2) words (pqv_peer_exchange_create's cap_keys); *out_overflow = 1 otherwise, or when a rank's slice
 * declined the batch / a NaN distance showed up -- all ranks see it and take the collective path together. */
PQV_API int pqv_l2_topk_batch_p2p(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k,
                                  uint32_t flags, uint32_t pos_base, uint32_t *out_row_idx, float *out_dist,
                                  uint32_t *out_count, uint32_t *out_replayed, uint32_t *out_overflow);

/* Batched variant (config C5: many queries, rows sharded over the ranks).  pqv_l2_topk_batch_keys answers the batch
 * over this rank's slice in one tensor-core pass and returns, per query, the k + 1 smallest
 * exact keys of the slice: out_keys[q*(k+1) + i] ascending, out_count[q] of them valid; out_count[q] = 0xFFFFFFFF when
 * the slice could not decide the query (non-finite inputs, candidate buffers full, batch too small for the pass).
 * The ranks exchange keys and counts with ONE all-gather ([rank][query][k+1] u64 and [rank][query] u32) and each calls
 * pqv_merge_batch_keys (pure host): the k + 1 smallest keys of the whole table are among the slices' k + 1 smallest,
 * so a query is final when its k-th and (k+1)-th smallest distances differ and its k returned values are pairwise
 * distinct -- then the reference loop (src/ivf/search.rs:112-141, src/df_vector/exec.rs:257-277) can only produce this
 * answer.  out_needs_replay[q] = 1 marks the queries whose answer hinges on the reference heap's layout: those go
 * through pqv_l2_topk_candidates + pqv_replay_candidates. */
PQV_API int pqv_l2_topk_batch_keys(pqv_ctx *ctx, uint64_t handle, const float *queries, uint32_t n_queries, uint32_t k,
                                   uint32_t flags, uint32_t pos_base, uint64_t *out_keys, uint32_t *out_count);
/* Cheaper candidates for a query pqv_merge_batch_keys flagged: valid right after pqv_l2_topk_batch_keys on the same
 * dataset (the batched pass leaves every query's exact-distance candidates on the device).  Returns a superset of the rows
 * the reference heap admits inside this slice -- exact scan of the first rows of the slice + the query's candidates
 * behind them (0.65 ms instead of a full scan) -- as keys for pqv_replay_candidates, like pqv_l2_topk_candidates. */
PQV_API int pqv_l2_topk_batch_tie_candidates(pqv_ctx *ctx, uint64_t handle, uint32_t q_index, const float *query,
                                             uint64_t *out_keys, uint64_t cap, uint64_t *out_count);
PQV_API int pqv_merge_batch_keys(const uint64_t *keys, const uint32_t *counts, uint32_t n_ranks, uint32_t n_queries,
                                 uint32_t k, uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count,
                                 uint8_t *out_needs_replay);

/* ---- measurement hooks (bench.py / ncu) ------------------------------------------------------- */
typedef struct {
    double scan_ms;        /* CUDA-event time of the last distance+select kernel (on the library's stream) */
    double post_ms;        /* prefix-merge + entrant-filter kernels                                       */
    double total_ms;       /* first launch -> last device op of the last call                              */
    uint64_t scan_bytes;   /* algorithmic bytes of that scan: rows * dim * 4                               */
    uint32_t launches;     /* kernels launched by the last call                                            */
    uint32_t entrants;     /* heap-entrant candidates replayed on the host by the last call                */
    uint32_t grid;         /* CTAs of the scan kernel                                                      */
    uint32_t reserved;
} pqv_timing;
PQV_API int pqv_last_timing(pqv_ctx *ctx, pqv_timing *out);
/* what the last pqv_kmeans_assign / pqv_bench_assign did.  path 1 = tcgen05 filter
 * (tensor cores decide every row whose candidate window holds one centroid; the rest are re-evaluated in the
 * reference's exact f32 order), path 0 = exact SIMT kernel for every (row, centroid) pair. */
typedef struct {
    uint32_t path;            /* 0 = exact SIMT, 1 = tcgen05 filter + exact re-check                        */
    uint32_t kind;            /* tcgen05 operand: 1 = the rows' fp16 shadow (kind::f16), 0 = f32 rows (kind::tf32) */
    uint64_t rows;
    uint64_t ambiguous_rows;  /* rows with 2..4 candidates after the filter (exact chain over those only)  */
    uint64_t overflow_rows;   /* rows sent to the full exact scan (non-finite norms, > 4 candidates)       */
    double prep_ms;           /* centroid centring/rounding + row norms                                    */
    double filter_ms;         /* the tcgen05 kernel (path 1) or the SIMT kernel (path 0), CUDA events      */
    double recheck_ms;        /* exact re-evaluation kernels (pairs + overflow scan + finalize)            */
    double pair_ms;           /* of which: the (row, candidate) pair kernel                                */
    double total_ms;          /* prep + filter + recheck                                                   */
    double shadow_ms;         /* building the rows' 16-bit shadow, when this call had to (resident tables: once
                                 per table, not per sweep; not part of total_ms)                           */
} pqv_assign_timing;
PQV_API int pqv_last_assign_timing(pqv_ctx *ctx, pqv_assign_timing *out);
/* device-resident loop for roofline timing of the assignment sweep (src/ivf/index.rs:189-206): `iters` sweeps of
 * the first n rows of the resident dataset against `centroids` with inputs and outputs in HBM; the mean times
 * land in *out (path as pqv_kmeans_assign would pick it, or forced by the PQV_ASSIGN=simt|tc environment variable).
 * out_assign (may be NULL) receives the assignments of the last sweep. */
PQV_API int pqv_bench_assign(pqv_ctx *ctx, uint64_t handle, uint64_t n, const float *centroids, uint32_t n_clusters,
                             uint32_t iters, pqv_assign_timing *out, uint32_t *out_assign);
/* device-resident loop for roofline timing: runs `iters` scans of query 0 back to back with inputs and
 * outputs in HBM and returns the mean kernel time (CUDA events on the launch stream). */
PQV_API int pqv_bench_scan(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags,
                   uint32_t iters, double *out_ms_per_scan);

/* Batched IVF search: n_queries independent TopkBuilder::search calls (src/ivf/search.rs:83-142) -- or, with
 * PQV_ROW_ORDER | PQV_SUM_SEQ, VectorTopKExec executions without cap and filter (src/df_vector/exec.rs:207-277) -- over
 * one resident table + index, answered by ONE tensor-core pass over the table restricted per query to the clusters it
 * probes.  Output layout as pqv_l2_topk (out_*[q*k + i], out_count[q]); every query's result is
 * identical to its own single-query call (ties, NaN rankings and declined batches take the single-query pipeline). */
PQV_API int pqv_ivf_search_batch(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *queries, uint32_t n_queries,
                         uint32_t k, uint32_t nprobe, uint32_t flags, uint32_t *out_row_idx, float *out_dist,
                         uint32_t *out_count);

/* Per-rank half of a SHARDED batched IVF search (config C5 with the index): `index` holds the lists cut to this rank's
 * row range (local ids), pos_base = first global row of the slice.  Returns per query the k + 1 smallest exact keys
 * (bits(squared distance) << 32 | pos_base + local row) among the probed rows of the slice, layout and meaning as
 * pqv_l2_topk_batch_keys (out_count[q] = 0xFFFFFFFF: undecided here).  The ranks exchange keys and counts with ONE
 * all-gather and call pqv_merge_batch_keys; queries it flags go through pqv_ivf_search_candidates + the replay. */
PQV_API int pqv_ivf_search_batch_keys(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *queries,
                              uint32_t n_queries, uint32_t k, uint32_t nprobe, uint32_t flags, uint32_t pos_base,
                              uint64_t *out_keys, uint32_t *out_count);

/* IVF search with the rows sharded over ranks ("the index is replicated; each rank filters candidate
 * ids to its row range").  Each rank loads the index restricted to its slice (same centroids, every list cut to the rank's
 * row range with local ids: pq_vector_b200/sharded.py) and calls pqv_ivf_search_candidates: the per-rank half of
 * pqv_ivf_search -- ranking, expansion, gathered scan, entrant filter -- returning the heap-entrant keys
 * (bits(squared distance) << 32 | position in THIS rank's candidate sequence), their local row ids and the probed
 * clusters in rank order (out_probe, min(nprobe, C) entries).  Lists are ascending and slices contiguous, so the global
 * candidate sequence of src/ivf/index.rs:57-63 is, list by list, rank 0's part, rank 1's part, ...: the ranks translate
 * positions with the per-list per-rank counts, exchange the keys with ONE all-gather and replay the reference heap
 * (pqv_replay_candidates).  PQV_ELIMIT with *out_count = the number needed when `cap` is too small. */
PQV_API int pqv_ivf_search_candidates(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k,
                              uint32_t nprobe, uint32_t flags, uint64_t *out_keys, uint32_t *out_rows, uint64_t cap,
                              uint64_t *out_count, uint32_t *out_probe, uint32_t *out_nprobe_eff);

/* ---- VectorTopKExec over a resident, indexed table in ONE call ------------------------------------
 * Replaces execute_with_candidates + topk_from_batches (src/df_vector/exec.rs:207-277) when the file's embedding column
 * and index are resident (pqv_dataset_*, pqv_ivf_from_bytes): the index's candidate rows for `nprobe`
 * (src/df_vector/index_exec.rs:159-163), of which the first `max_candidates` in rank order count (CandidateCursor over
 * one file, src/df_vector/access.rs:193-243; PQV_NO_CANDIDATE_CAP = `None`, no cap, options.rs:10-11; 0 = `Some(0)`: no
 * row is fetched and the result is empty, src/df_vector/exec.rs:222-223), visited in ascending row order (the
 * RowSelection of access.rs:107-176), rows whose bit in `row_mask` is clear dropped BEFORE scoring (the FilterExec of the
 * scan subtree, src/df_vector/tests.rs:151-241), then the bounded heap of k.  Use flags = PQV_SUM_SEQ for the operator's
 * own arithmetic (exec.rs:529-533, squared distances out).  row_mask: NULL = every row passes; else ceil(n_rows / 8)
 * bytes, bit (r & 7) of byte (r >> 3) = row r passes (an Arrow boolean buffer).  out_candidate_rows / out_rows_scored
 * (may be NULL) are the plan counters `candidate_rows` and `embeddings_fetched` of the reference's snapshots.
 * Ranking, candidate bitmap, ordered compaction, gathered scan and top-k run in one host<->device round trip. */
#define PQV_NO_CANDIDATE_CAP 0xFFFFFFFFFFFFFFFFull  /* VectorTopKOptions::max_candidates == None */
PQV_API int pqv_vector_topk_indexed(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k,
                            uint32_t nprobe, uint32_t flags, uint64_t max_candidates, const uint8_t *row_mask,
                            uint32_t *out_row_idx, float *out_dist, uint32_t *out_count, uint64_t *out_candidate_rows,
                            uint64_t *out_rows_scored);

/* n_queries executions of the operator that share one scan-subtree filter (row_mask as above, may be NULL) and have no
 * candidate cap -- e.g. the same SQL text with different literals -- answered by one tensor-core pass over the table with
 * the probe sets and the filter bitmap applied inside the pass; every query's result equals its own
 * pqv_vector_topk_indexed call.  Output layout as pqv_l2_topk. */
PQV_API int pqv_vector_topk_indexed_batch(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *queries,
                                  uint32_t n_queries, uint32_t k, uint32_t nprobe, uint32_t flags, const uint8_t *row_mask,
                                  uint32_t *out_row_idx, float *out_dist, uint32_t *out_count);

/* ---- coalescing front door for concurrent single-query callers ---------------------------------
 * The reference API is single-query (src/ivf/search.rs:49-54 `query: &[f32]`, src/df_vector/exec.rs:43) and its callers
 * are concurrent tasks (TopkBuilder::search is async; VectorTopKExec::execute is a stream::once per plan,
 * src/df_vector/exec.rs:387).  pqv_l2_topk_coalesced has the single-query contract of pqv_l2_topk (n_queries = 1) and may
 * be called from any number of threads: the first caller runs; calls arriving while a pass over the table is in flight
 * queue up and are answered together by ONE batched pass (pqv_l2_topk with n_queries = batch) as
 * soon as the running one finishes.  Every caller still receives exactly the output of its own single-query call.
 * Requests with a different (dataset, k, flags) are not mixed into one batch.  The calling thread blocks (wrap in
 * spawn_blocking on the Rust side).  pqv_coalesce_config: max_batch (default 1024) bounds a batch; window_us > 0 lets
 * the leading caller linger that long for a burst to assemble (default 0: batches form only behind a running pass). */
PQV_API int pqv_l2_topk_coalesced(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags,
                          uint32_t *out_row_idx, float *out_dist, uint32_t *out_count);
/* the same front door for TopkBuilder::search callers (src/ivf/search.rs:76-80 is an async fn): the single-query contract of
 * pqv_ivf_search; concurrent calls with the same (table, index, k, nprobe, flags) are answered by one pqv_ivf_search_batch. */
PQV_API int pqv_ivf_search_coalesced(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k,
                             uint32_t nprobe, uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count);
PQV_API int pqv_coalesce_config(pqv_ctx *ctx, uint32_t max_batch, uint32_t window_us);
PQV_API int pqv_coalesce_stats(pqv_ctx *ctx, uint64_t *out_queries, uint64_t *out_batches, uint64_t *out_max_batch);

/* ---- the un-indexed `array_distance` arm ---------------------------------------------------------
 * `ORDER BY array_distance(col, [..]) LIMIT k` over a file WITHOUT an embedded index is left alone by pq-vector's
 * optimizer rule (src/df_vector/physical.rs:198-214), so DataFusion's built-in UDF runs per row under SortExec(TopK)
 * (datafusion-functions-nested 52.1.0, Cargo.lock:1041-1042 -- not vendored in the reference; call sites
 * benches/query.rs:79-81, examples/datafusion_sql.rs:54-55, src/df_vector/tests.rs:77-80).  Upstream semantics as
 * published: both lists cast to Float64, sum of (a - b)^2 folded in element order in f64, sqrt; a length mismatch is an
 * error.  PARITY UNPINNED: the reference's tests never reach the UDF.
 *   pqv_array_distance       the UDF itself: out[i] = distance(row i, query) as Float64, one value per resident row.
 *   pqv_array_distance_topk  UDF + SortExec(TopK): the k smallest by f64 total order (NaN last); ties by ascending row
 *                            (the stock operator leaves that order unspecified).  out_count = min(k, rows).
 * metric: PQV_METRIC_L2 = Euclidean (above).  PQV_METRIC_COSINE = 1 - a.b / (sqrt(a.a) sqrt(b.b)), the three sums
 * folded in element order in f64 -- additive: the reference has no cosine distance. */
#define PQV_METRIC_L2      0u
#define PQV_METRIC_COSINE  1u
PQV_API int pqv_array_distance(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric,
                       double *out /* n_rows */);
PQV_API int pqv_array_distance_topk(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric,
                            uint32_t k, uint32_t *out_row_idx, double *out_dist, uint32_t *out_count);
/* the same under a WHERE clause: row_mask = the filter of the scan subtree evaluated over the table's rows (an Arrow boolean
 * buffer: bit (r & 7) of byte (r >> 3) = row r passes; NULL = no filter); filtered rows never reach the sort. */
PQV_API int pqv_array_distance_topk_filtered(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len,
                                     uint32_t metric, uint32_t k, const uint8_t *row_mask, uint32_t *out_row_idx,
                                     double *out_dist, uint32_t *out_count);

#ifdef __cplusplus
}
#endif
#endif /* PQV_H */
