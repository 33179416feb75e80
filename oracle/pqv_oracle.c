/*
 * pqv_oracle.c -- CPU restatement of pq-vector's squared-L2 / top-k / IVF-assign hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and only as the checker
 * (or as the timed CPU baseline).  Nothing under pq_vector_b200/ links, imports or calls it.
 *
 * The reference (XiangpengHao/pq-vector @ 808b90d) is safe scalar Rust and cannot be compiled in
 * this image (no cargo/rustc, 295 un-vendored crates), so this file restates its algorithm in C.
 * Every function cites the reference file:line it follows.  Build with
 *     gcc -O3 -ffp-contract=off -fno-fast-math
 * so that no FMA is formed and no float re-association happens: rustc never contracts or
 * re-associates f32 arithmetic, hence the bit patterns produced here equal the Rust ones.
 *
 * Parity pinning: checked in tests/test_oracle.py against every known-answer the reference's own
 * tests hold for this path (index.rs:487-493 -> 27.0; df_vector/tests.rs ids [5,2] and [3,4];
 * index.rs:495-511 blob round trip; vldb snapshot candidate count 496) and against the vldb
 * top-10 table of SURVEY.md section 8c.  The Rust std `BinaryHeap` (push = sift_up, pop =
 * swap-with-last + sift_down_to_bottom + sift_up, into_iter = backing-vector order) is a
 * third-party (std) algorithm restated from its published source; the reference's tests do not
 * pin heap-layout tie order, so exact-tie output order is "restated, not reference-pinned".
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PQO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* distance kernels                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* src/ivf/index.rs:459-480  squared_l2_distance: 4-wide chunks, sum += ((d0^2+d1^2)+d2^2)+d3^2,
 * then a scalar tail `sum += d*d`.  No FMA, fixed order. */
PQO_API float pqo_squared_l2_unroll4(const float *a, const float *b, size_t len) {
    float sum = 0.0f;
    size_t i = 0;
    while (i + 4 <= len) {
        float d0 = a[i] - b[i];
        float d1 = a[i + 1] - b[i + 1];
        float d2 = a[i + 2] - b[i + 2];
        float d3 = a[i + 3] - b[i + 3];
        sum += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        i += 4;
    }
    while (i < len) {
        float d = a[i] - b[i];
        sum += d * d;
        i += 1;
    }
    return sum;
}

/* src/df_vector/exec.rs:524-535  compute_distance_values (Float32 arm): diff = value - q;
 * dist += diff*diff, strictly sequential. */
PQO_API float pqo_squared_l2_seq(const float *values, const float *query, size_t len) {
    float dist = 0.0f;
    for (size_t i = 0; i < len; ++i) {
        float diff = values[i] - query[i];
        dist += diff * diff;
    }
    return dist;
}

/* src/df_vector/exec.rs:536-547  Float64 arm: value narrowed to f32 BEFORE the subtract. */
PQO_API float pqo_squared_l2_seq_f64(const double *values, const float *query, size_t len) {
    float dist = 0.0f;
    for (size_t i = 0; i < len; ++i) {
        float diff = (float)values[i] - query[i];
        dist += diff * diff;
    }
    return dist;
}

static inline float pqo_dist(int order, const float *a, const float *b, size_t len) {
    /* order 0: index.rs:461 (a = query/vec, b = vec/centroid as the call site passes them);
     * order 1: exec.rs:529 (a = stored values, b = query). */
    return order == 0 ? pqo_squared_l2_unroll4(a, b, len) : pqo_squared_l2_seq(a, b, len);
}

/* Fill out[i] = distance(query, row i) for a dense row-major block; operand order as in
 * src/ivf/search.rs:117 (`squared_l2_distance(query, vec)`) for order 0 and exec.rs:531
 * (`value - q`) for order 1. */
PQO_API void pqo_distances(const float *rows, uint64_t n_rows, uint32_t dim, const float *query,
                           int order, float *out) {
    for (uint64_t i = 0; i < n_rows; ++i) {
        const float *v = rows + i * (uint64_t)dim;
        out[i] = order == 0 ? pqo_squared_l2_unroll4(query, v, dim) : pqo_squared_l2_seq(v, query, dim);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Rust std::collections::BinaryHeap<HeapItem>, max-heap on distance                            */
/* HeapItem ordering: src/ivf/search.rs:18-38 (partial_cmp, NaN -> Equal); TopKRow is the same */
/* (src/df_vector/exec.rs:435-455).                                                            */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t row_idx;
    float distance;
} pqo_item;

/* `a <= b` under Ord::cmp = partial_cmp().unwrap_or(Equal): true unless a > b; NaN compares Equal. */
static inline int item_le(const pqo_item *a, const pqo_item *b) { return !(a->distance > b->distance); }

typedef struct {
    pqo_item *data;
    size_t len;
} pqo_heap;

/* alloc::collections::binary_heap::BinaryHeap::sift_up(start, pos) */
static size_t heap_sift_up(pqo_heap *h, size_t start, size_t pos) {
    pqo_item elt = h->data[pos]; /* the Hole */
    while (pos > start) {
        size_t parent = (pos - 1) / 2;
        if (item_le(&elt, &h->data[parent])) break;
        h->data[pos] = h->data[parent];
        pos = parent;
    }
    h->data[pos] = elt;
    return pos;
}

/* BinaryHeap::sift_down_to_bottom(0): walk the hole to a leaf always taking the greater child
 * (right child when left <= right), then sift_up from there. */
static void heap_sift_down_to_bottom(pqo_heap *h, size_t pos) {
    size_t end = h->len;
    size_t start = pos;
    pqo_item elt = h->data[pos];
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0; /* end.saturating_sub(2) */
    while (child <= lim && end >= 2) {
        child += (size_t)item_le(&h->data[child], &h->data[child + 1]);
        h->data[pos] = h->data[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1) {
        h->data[pos] = h->data[child];
        pos = child;
    }
    h->data[pos] = elt;
    heap_sift_up(h, start, pos);
}

static void heap_push(pqo_heap *h, pqo_item it) {
    size_t old_len = h->len;
    h->data[h->len++] = it;
    heap_sift_up(h, 0, old_len);
}

static void heap_pop(pqo_heap *h) {
    /* self.data.pop().map(|mut item| { if !self.is_empty() { swap(&mut item, &mut self.data[0]);
     *   self.sift_down_to_bottom(0) } item }) */
    pqo_item last = h->data[--h->len];
    if (h->len > 0) {
        h->data[0] = last;
        heap_sift_down_to_bottom(h, 0);
    }
}

/* stable merge sort ascending by distance, comparator partial_cmp().unwrap_or(Equal)
 * (src/ivf/search.rs:136-140, src/df_vector/exec.rs:270-274, src/ivf/index.rs:143).
 * `slice::sort_by` is a stable sort; for a consistent order its result does not depend on the
 * algorithm, so a textbook top-down merge sort reproduces it. */
static void stable_sort_items(pqo_item *a, pqo_item *tmp, size_t n) {
    if (n < 2) return;
    size_t mid = n / 2;
    stable_sort_items(a, tmp, mid);
    stable_sort_items(a + mid, tmp, n - mid);
    size_t i = 0, j = mid, o = 0;
    while (i < mid && j < n) {
        /* take right only if right < left strictly; ties keep left first (stability) */
        if (a[j].distance < a[i].distance) tmp[o++] = a[j++];
        else tmp[o++] = a[i++];
    }
    while (i < mid) tmp[o++] = a[i++];
    while (j < n) tmp[o++] = a[j++];
    memcpy(a, tmp, n * sizeof(pqo_item));
}

/* The bounded-heap loop shared by src/ivf/search.rs:112-141 (TopkBuilder, do_sqrt = 1) and
 * src/df_vector/exec.rs:257-277 + 457-482 (VectorTopKExec, do_sqrt = 0): push while len < k,
 * otherwise replace the root only when `distance < top.distance`; collect in backing-vector
 * order; (sqrt); stable sort ascending.  `dist[i]` are the precomputed squared distances of the
 * candidates in candidate order and `row_ids[i]` their row ids (NULL => row id = i).
 * Returns the number of results (<= k). */
PQO_API size_t pqo_heap_topk(const float *dist, const uint32_t *row_ids, uint64_t n_cand, size_t k,
                             int do_sqrt, uint32_t *out_rows, float *out_dist) {
    if (k == 0) return 0;
    pqo_heap h;
    h.data = (pqo_item *)malloc((k + 1) * sizeof(pqo_item));
    h.len = 0;
    for (uint64_t i = 0; i < n_cand; ++i) {
        pqo_item it;
        it.row_idx = row_ids ? row_ids[i] : (uint32_t)i;
        it.distance = dist[i];
        if (h.len < k) {
            heap_push(&h, it);
        } else if (it.distance < h.data[0].distance) {
            heap_pop(&h);
            heap_push(&h, it);
        }
    }
    size_t n = h.len;
    if (do_sqrt)
        for (size_t i = 0; i < n; ++i) h.data[i].distance = sqrtf(h.data[i].distance);
    pqo_item *tmp = (pqo_item *)malloc((n + 1) * sizeof(pqo_item));
    stable_sort_items(h.data, tmp, n);
    for (size_t i = 0; i < n; ++i) {
        out_rows[i] = h.data[i].row_idx;
        out_dist[i] = h.data[i].distance;
    }
    free(tmp);
    free(h.data);
    return n;
}

/* src/ivf/search.rs:112-141 in one call: `vectors` are the candidate vectors in candidate order
 * (what read_embeddings_for_rows returns), distances with squared_l2_distance(query, vec). */
PQO_API size_t pqo_topk_rerank(const float *query, const float *vectors, const uint32_t *row_ids,
                               uint64_t n_cand, uint32_t dim, size_t k, int order, int do_sqrt,
                               uint32_t *out_rows, float *out_dist) {
    float *d = (float *)malloc((n_cand + 1) * sizeof(float));
    pqo_distances(vectors, n_cand, dim, query, order, d);
    size_t n = pqo_heap_topk(d, row_ids, n_cand, k, do_sqrt, out_rows, out_dist);
    free(d);
    return n;
}

/* Same, but the candidate vectors are gathered from a dense table by row id (what the reference
 * obtains through Parquet row selection). */
PQO_API size_t pqo_topk_rerank_gather(const float *query, const float *table, const uint32_t *row_ids,
                                      uint64_t n_cand, uint32_t dim, size_t k, int order, int do_sqrt,
                                      uint32_t *out_rows, float *out_dist) {
    float *d = (float *)malloc((n_cand + 1) * sizeof(float));
    for (uint64_t i = 0; i < n_cand; ++i) {
        const float *v = table + (uint64_t)row_ids[i] * dim;
        d[i] = order == 0 ? pqo_squared_l2_unroll4(query, v, dim) : pqo_squared_l2_seq(v, query, dim);
    }
    size_t n = pqo_heap_topk(d, row_ids, n_cand, k, do_sqrt, out_rows, out_dist);
    free(d);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* IVF: nearest centroid, assignment sweeps, centroid ranking                                  */
/* ------------------------------------------------------------------------------------------ */

/* src/ivf/index.rs:244-257 nearest_centroid: strict '<' => lowest index wins ties; NaN/inf never
 * win => cluster 0. */
PQO_API uint32_t pqo_nearest_centroid(const float *vec, const float *centroids, uint32_t n_clusters,
                                      uint32_t dim) {
    uint32_t best_cluster = 0;
    float best_dist = INFINITY;
    for (uint32_t i = 0; i < n_clusters; ++i) {
        float dist = pqo_squared_l2_unroll4(vec, centroids + (size_t)i * dim, dim);
        if (dist < best_dist) {
            best_dist = dist;
            best_cluster = i;
        }
    }
    return best_cluster;
}

typedef struct {
    const float *data;
    const float *centroids;
    uint32_t dim, n_clusters;
    uint64_t start, end;
    uint32_t *out;
} assign_job;

static void *assign_worker(void *p) {
    assign_job *j = (assign_job *)p;
    for (uint64_t r = j->start; r < j->end; ++r)
        j->out[r] = pqo_nearest_centroid(j->data + r * (uint64_t)j->dim, j->centroids, j->n_clusters, j->dim);
    return NULL;
}

/* src/ivf/index.rs:189-206 (final assignment) and :395-424 (Lloyd assignment): every row vs every
 * centroid, split over `workers` contiguous chunks of ceil(len/workers) rows exactly as
 * parallel_ranges / parallel_chunks_mut do (index.rs:267-320).  The assignment of a row does not
 * depend on the split, only the timing does. */
PQO_API void pqo_assign(const float *data, uint64_t n, uint32_t dim, const float *centroids,
                        uint32_t n_clusters, uint32_t *out_assign, int workers) {
    if (n == 0) return;
    if (workers < 1) workers = 1;
    if ((uint64_t)workers > n) workers = (int)n;
    uint64_t chunk = (n + (uint64_t)workers - 1) / (uint64_t)workers;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)workers);
    assign_job *jobs = (assign_job *)malloc(sizeof(assign_job) * (size_t)workers);
    int started = 0;
    for (int w = 0; w < workers; ++w) {
        uint64_t s = (uint64_t)w * chunk;
        if (s >= n) break;
        uint64_t e = s + chunk < n ? s + chunk : n;
        jobs[w] = (assign_job){data, centroids, dim, n_clusters, s, e, out_assign};
        pthread_create(&th[w], NULL, assign_worker, &jobs[w]);
        started++;
    }
    for (int w = 0; w < started; ++w) pthread_join(th[w], NULL);
    free(th);
    free(jobs);
}

/* src/ivf/index.rs:202-206: inverted lists = per-cluster row ids in ascending order (chunks are
 * appended in chunk order and each chunk walks rows ascending).  CSR output: offsets[C+1], ids[n]. */
PQO_API void pqo_inverted_lists(const uint32_t *assign, uint64_t n, uint32_t n_clusters,
                                uint64_t *offsets, uint32_t *ids) {
    memset(offsets, 0, sizeof(uint64_t) * ((size_t)n_clusters + 1));
    for (uint64_t i = 0; i < n; ++i) offsets[assign[i] + 1]++;
    for (uint32_t c = 0; c < n_clusters; ++c) offsets[c + 1] += offsets[c];
    uint64_t *cur = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n_clusters);
    memcpy(cur, offsets, sizeof(uint64_t) * (size_t)n_clusters);
    for (uint64_t i = 0; i < n; ++i) ids[cur[assign[i]]++] = (uint32_t)i;
    free(cur);
}

/* src/ivf/index.rs:130-149 find_closest_centroids: C distances squared_l2_distance(query,
 * centroid), stable sort ascending (NaN -> Equal), take min(nprobe, C). Returns count. */
PQO_API uint32_t pqo_find_closest_centroids(const float *query, const float *centroids,
                                            uint32_t n_clusters, uint32_t dim, uint32_t nprobe,
                                            uint32_t *out_clusters) {
    if (nprobe > n_clusters) nprobe = n_clusters;
    pqo_item *cd = (pqo_item *)malloc(sizeof(pqo_item) * ((size_t)n_clusters + 1));
    pqo_item *tmp = (pqo_item *)malloc(sizeof(pqo_item) * ((size_t)n_clusters + 1));
    for (uint32_t i = 0; i < n_clusters; ++i) {
        cd[i].row_idx = i;
        cd[i].distance = pqo_squared_l2_unroll4(query, centroids + (size_t)i * dim, dim);
    }
    stable_sort_items(cd, tmp, n_clusters);
    for (uint32_t i = 0; i < nprobe; ++i) out_clusters[i] = cd[i].row_idx;
    free(cd);
    free(tmp);
    return nprobe;
}

/* src/ivf/index.rs:57-63 candidate_rows: the inverted lists of the nprobe closest clusters
 * concatenated in rank order.  Lists given as CSR.  `out_rows` must hold offsets[C] ids. */
PQO_API uint64_t pqo_candidate_rows(const float *query, const float *centroids, uint32_t n_clusters,
                                    uint32_t dim, const uint64_t *offsets, const uint32_t *ids,
                                    uint32_t nprobe, uint32_t *out_rows) {
    uint32_t *cl = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n_clusters + 1));
    uint32_t np = pqo_find_closest_centroids(query, centroids, n_clusters, dim, nprobe, cl);
    uint64_t o = 0;
    for (uint32_t r = 0; r < np; ++r) {
        uint32_t c = cl[r];
        uint64_t len = offsets[c + 1] - offsets[c];
        memcpy(out_rows + o, ids + offsets[c], len * sizeof(uint32_t));
        o += len;
    }
    free(cl);
    return o;
}

/* ------------------------------------------------------------------------------------------ */
/* k-means pieces (src/ivf/index.rs:323-457).  RNG-driven choices (rand 0.8.5 StdRng) are NOT   */
/* restated: they are passed in by the caller (SURVEY F8/H6 - the reference itself is not      */
/* reproducible across machines at that level); parity is pinned at the sweep level.           */
/* ------------------------------------------------------------------------------------------ */

/* src/ivf/index.rs:344-352: min_distances[slot] = squared_l2_distance(vec(row), centroid0). */
PQO_API void pqo_min_dist_init(const float *data, const uint64_t *row_sel, uint64_t n_sel, uint32_t dim,
                               const float *centroid, float *min_dist) {
    for (uint64_t s = 0; s < n_sel; ++s) {
        uint64_t r = row_sel ? row_sel[s] : s;
        min_dist[s] = pqo_squared_l2_unroll4(data + r * (uint64_t)dim, centroid, dim);
    }
}

/* src/ivf/index.rs:354-370: one k-means++ sweep against the newest centroid:
 * dist < slot => slot = dist; local_sum += slot per chunk of ceil(n/workers) slots; the
 * returned total is the sum of the per-chunk partial sums in chunk order (Iterator::sum from
 * 0.0), i.e. it depends on `workers` exactly like the reference depends on
 * available_parallelism(). */
PQO_API float pqo_min_dist_update(const float *data, const uint64_t *row_sel, uint64_t n_sel,
                                  uint32_t dim, const float *centroid, float *min_dist, int workers) {
    if (n_sel == 0) return 0.0f;
    if (workers < 1) workers = 1;
    if ((uint64_t)workers > n_sel) workers = (int)n_sel;
    uint64_t chunk = (n_sel + (uint64_t)workers - 1) / (uint64_t)workers;
    float total = 0.0f;
    for (uint64_t s0 = 0; s0 < n_sel; s0 += chunk) {
        uint64_t s1 = s0 + chunk < n_sel ? s0 + chunk : n_sel;
        float local_sum = 0.0f;
        for (uint64_t s = s0; s < s1; ++s) {
            uint64_t r = row_sel ? row_sel[s] : s;
            float dist = pqo_squared_l2_unroll4(data + r * (uint64_t)dim, centroid, dim);
            if (dist < min_dist[s]) min_dist[s] = dist;
            local_sum += min_dist[s];
        }
        total += local_sum;
    }
    return total;
}

/* src/ivf/index.rs:372-383: pick the first slot whose running f32 cumsum reaches
 * threshold = u * total (u drawn by the caller). Returns n_sel if none (the reference then leaves
 * the centroid at zero - the loop simply ends without a break). */
PQO_API uint64_t pqo_kmeanspp_pick(const float *min_dist, uint64_t n_sel, float threshold) {
    float cumsum = 0.0f;
    for (uint64_t s = 0; s < n_sel; ++s) {
        cumsum += min_dist[s];
        if (cumsum >= threshold) return s;
    }
    return n_sel;
}

/* src/ivf/index.rs:395-430 Lloyd assignment step: assign + `changed` + per-cluster sizes. */
PQO_API uint64_t pqo_lloyd_assign(const float *data, uint64_t n, uint32_t dim, const float *centroids,
                                  uint32_t n_clusters, uint32_t *assign_inout, uint64_t *sizes,
                                  int workers) {
    uint32_t *next = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    pqo_assign(data, n, dim, centroids, n_clusters, next, workers);
    uint64_t changed = 0;
    memset(sizes, 0, sizeof(uint64_t) * (size_t)n_clusters);
    for (uint64_t i = 0; i < n; ++i) {
        if (assign_inout[i] != next[i]) changed++;
        assign_inout[i] = next[i];
        sizes[next[i]]++;
    }
    free(next);
    return changed;
}

/* src/ivf/index.rs:436-453 centroid update: zero, serial row-order sums, divide by size when
 * size > 0 (empty cluster collapses to the origin). */
PQO_API void pqo_centroid_update(const float *data, uint64_t n, uint32_t dim, const uint32_t *assign,
                                 const uint64_t *sizes, uint32_t n_clusters, float *centroids) {
    memset(centroids, 0, sizeof(float) * (size_t)n_clusters * dim);
    for (uint64_t i = 0; i < n; ++i) {
        float *c = centroids + (size_t)assign[i] * dim;
        const float *v = data + i * (uint64_t)dim;
        for (uint32_t j = 0; j < dim; ++j) c[j] += v[j];
    }
    for (uint32_t j = 0; j < n_clusters; ++j) {
        if (sizes[j] > 0) {
            float size = (float)sizes[j];
            for (uint32_t d = 0; d < dim; ++d) centroids[(size_t)j * dim + d] /= size;
        }
    }
}

/* src/ivf/index.rs:161-174, 332: sizing rules.  out[0]=n_clusters, out[1]=sample_size,
 * out[2]=init_sample_size(sample_size).  Returns 0 ok, 1 zero vectors, 2 clusters > vectors. */
PQO_API int pqo_build_sizes(uint64_t n_vectors, uint64_t n_clusters_or_0, uint64_t *out) {
    if (n_vectors == 0) return 1;
    uint64_t c = n_clusters_or_0 ? n_clusters_or_0 : (uint64_t)ceil(sqrt((double)n_vectors));
    if (c > n_vectors) return 2;
    uint64_t sample = n_vectors / 20;
    if (sample < 1) sample = 1;
    if (sample > 100000) sample = 100000;
    if (sample < c) sample = c;
    if (sample > n_vectors) sample = n_vectors;
    uint64_t init = sample < 50000 ? sample : 50000;
    if (init < c) init = c;
    out[0] = c;
    out[1] = sample;
    out[2] = init;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* index blob: src/ivf/index.rs:65-128                                                         */
/* LE: u32 dim, u32 C, f32[C*dim], then per cluster u32 len + u32[len]                         */
/* ------------------------------------------------------------------------------------------ */

PQO_API uint64_t pqo_index_blob_size(uint32_t dim, uint32_t n_clusters, const uint64_t *offsets) {
    return 8 + (uint64_t)n_clusters * dim * 4 + (uint64_t)n_clusters * 4 + offsets[n_clusters] * 4;
}

PQO_API uint64_t pqo_index_to_bytes(uint32_t dim, uint32_t n_clusters, const float *centroids,
                                    const uint64_t *offsets, const uint32_t *ids, uint8_t *out) {
    uint8_t *p = out;
    memcpy(p, &dim, 4); p += 4;
    memcpy(p, &n_clusters, 4); p += 4;
    memcpy(p, centroids, (size_t)n_clusters * dim * 4); p += (size_t)n_clusters * dim * 4;
    for (uint32_t c = 0; c < n_clusters; ++c) {
        uint32_t len = (uint32_t)(offsets[c + 1] - offsets[c]);
        memcpy(p, &len, 4); p += 4;
        memcpy(p, ids + offsets[c], (size_t)len * 4); p += (size_t)len * 4;
    }
    return (uint64_t)(p - out);
}

/* Two-pass parse: call with centroids/offsets/ids NULL to get dim, C, total ids; then again with
 * buffers. Returns 0 ok, 1 "IVF index buffer too small" (index.rs:88-90), 2 zero dim/clusters,
 * 3 truncated body (the reference panics/errs on the slice conversion). */
PQO_API int pqo_index_from_bytes(const uint8_t *bytes, uint64_t len, uint32_t *dim, uint32_t *n_clusters,
                                 uint64_t *n_ids, float *centroids, uint64_t *offsets, uint32_t *ids) {
    if (len < 8) return 1;
    uint32_t d, c;
    memcpy(&d, bytes, 4);
    memcpy(&c, bytes + 4, 4);
    if (d == 0 || c == 0) return 2;
    uint64_t off = 8;
    uint64_t cbytes = (uint64_t)d * c * 4;
    if (off + cbytes > len) return 3;
    if (centroids) memcpy(centroids, bytes + off, cbytes);
    off += cbytes;
    uint64_t total = 0;
    for (uint32_t i = 0; i < c; ++i) {
        if (off + 4 > len) return 3;
        uint32_t l;
        memcpy(&l, bytes + off, 4);
        off += 4;
        if (off + (uint64_t)l * 4 > len) return 3;
        if (offsets) offsets[i] = total;
        if (ids) memcpy(ids + total, bytes + off, (size_t)l * 4);
        off += (uint64_t)l * 4;
        total += l;
    }
    if (offsets) offsets[c] = total;
    *dim = d;
    *n_clusters = c;
    *n_ids = total;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* synthetic data: value distribution of benches/bench_util.rs:29-41 (`rng.gen::<f32>()` =     */
/* (u32 >> 8) * 2^-24, uniform [0,1) with 24-bit mantissa) from a counter-based generator keyed */
/* (seed, element index) so CPU and GPU regenerate identical rows.  The stream itself is ours  */
/* (splitmix64 finaliser), not ChaCha12: the reference publishes no data, only the distribution.*/
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t pqo_synth_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

PQO_API void pqo_synth_fill(float *out, uint64_t first_elem, uint64_t n_elems, uint64_t seed) {
    for (uint64_t i = 0; i < n_elems; ++i)
        out[i] = (float)(pqo_synth_u32(seed, first_elem + i) >> 8) * (1.0f / 16777216.0f);
}

/* ------------------------------------------------------------------------------------------ */
/* timed CPU baselines for bench.py                                                            */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const float *rows;
    const float *query;
    uint32_t dim;
    int order;
    uint64_t start, end;
    size_t k;
    uint32_t *rows_out;
    float *dist_out;
    size_t n_out;
} scan_job;

static void *scan_worker(void *p) {
    scan_job *j = (scan_job *)p;
    uint64_t n = j->end - j->start;
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    for (uint64_t i = 0; i < n; ++i) ids[i] = (uint32_t)(j->start + i);
    j->n_out = pqo_topk_rerank(j->query, j->rows + j->start * (uint64_t)j->dim, ids, n, j->dim, j->k,
                               j->order, 0, j->rows_out, j->dist_out);
    free(ids);
    return NULL;
}

/* "Charitable" CPU baseline: the reference's re-rank loop is serial
 * (src/ivf/search.rs:115-127); this splits the rows over `workers` threads, each running the same
 * loop with its own heap, then re-ranks the <= workers*k survivors serially.  With workers = 1 it
 * IS the reference loop.  Squared distances out (no sqrt). */
PQO_API size_t pqo_scan_topk_mt(const float *rows, uint64_t n_rows, uint32_t dim, const float *query,
                                size_t k, int order, int workers, uint32_t *out_rows, float *out_dist) {
    if (n_rows == 0 || k == 0) return 0;
    if (workers < 1) workers = 1;
    if ((uint64_t)workers > n_rows) workers = (int)n_rows;
    uint64_t chunk = (n_rows + (uint64_t)workers - 1) / (uint64_t)workers;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)workers);
    scan_job *jobs = (scan_job *)calloc((size_t)workers, sizeof(scan_job));
    int started = 0;
    for (int w = 0; w < workers; ++w) {
        uint64_t s = (uint64_t)w * chunk;
        if (s >= n_rows) break;
        uint64_t e = s + chunk < n_rows ? s + chunk : n_rows;
        jobs[w] = (scan_job){rows, query, dim, order, s, e, k, NULL, NULL, 0};
        jobs[w].rows_out = (uint32_t *)malloc(sizeof(uint32_t) * k);
        jobs[w].dist_out = (float *)malloc(sizeof(float) * k);
        if (workers == 1) scan_worker(&jobs[w]);
        else pthread_create(&th[w], NULL, scan_worker, &jobs[w]);
        started++;
    }
    if (workers > 1)
        for (int w = 0; w < started; ++w) pthread_join(th[w], NULL);
    /* survivors in ascending row order per chunk is not guaranteed (they are distance-sorted);
     * re-rank them by a final heap pass over (chunk order, distance order). */
    size_t total = 0;
    for (int w = 0; w < started; ++w) total += jobs[w].n_out;
    float *d = (float *)malloc(sizeof(float) * (total + 1));
    uint32_t *r = (uint32_t *)malloc(sizeof(uint32_t) * (total + 1));
    size_t o = 0;
    for (int w = 0; w < started; ++w) {
        memcpy(d + o, jobs[w].dist_out, sizeof(float) * jobs[w].n_out);
        memcpy(r + o, jobs[w].rows_out, sizeof(uint32_t) * jobs[w].n_out);
        o += jobs[w].n_out;
        free(jobs[w].rows_out);
        free(jobs[w].dist_out);
    }
    size_t n = pqo_heap_topk(d, r, total, k, 0, out_rows, out_dist);
    free(d);
    free(r);
    free(th);
    free(jobs);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* the un-indexed `array_distance` arm -- PARITY UNPINNED           */
/* ------------------------------------------------------------------------------------------ */
/* DataFusion's built-in UDF (crate datafusion-functions-nested 52.1.0, Cargo.lock:1041-1042; the source is NOT under
 * /root/reference).  Call sites in the reference: benches/query.rs:79-81, examples/datafusion_sql.rs:54-55,
 * src/df_vector/tests.rs:77-80 (there the optimizer rule plans it away).  Restated from the published upstream
 * algorithm (distance.rs, compute_array_distance): both lists are cast to Float64, then
 *     sum_squares = zip(a, b).map(|(x, y)| { let d = x - y; d * d }).sum::<f64>();   result = sum_squares.sqrt()
 * i.e. a sequential f64 fold in element order.  No golden vector of the reference reaches this function (all three SQL
 * tests run with the optimizer rule), so parity of this arm is UNPINNED; tests check it against an independent numpy
 * float64 cumulative sum. */
PQO_API double pqo_array_distance(const float *row, const double *query, size_t len) {
    double sum = 0.0;
    for (size_t i = 0; i < len; ++i) {
        double d = (double)row[i] - query[i];
        sum += d * d;
    }
    return sqrt(sum);
}

/* Cosine distance: additive (the reference has none).  1 - a.b / (sqrt(a.a) * sqrt(b.b)), sequential f64. */
PQO_API double pqo_cosine_distance(const float *row, const double *query, size_t len) {
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (size_t i = 0; i < len; ++i) {
        double x = (double)row[i];
        dot += x * query[i];
        na += x * x;
    }
    for (size_t i = 0; i < len; ++i) nb += query[i] * query[i];
    return 1.0 - dot / (sqrt(na) * sqrt(nb));
}

PQO_API void pqo_array_distance_column(const float *rows, uint64_t n, uint32_t dim, const double *query, int metric,
                                       double *out) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = metric ? pqo_cosine_distance(rows + i * (uint64_t)dim, query, dim)
                        : pqo_array_distance(rows + i * (uint64_t)dim, query, dim);
}

/* f64 total order with every NaN last (how DataFusion's sort treats NaN for ASC), as an unsigned key */
static inline uint64_t f64_ordered_bits(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    if ((b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) b = 0x7FF8000000000000ull;
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
typedef struct {
    uint64_t key;
    uint32_t row;
} adist_item;
static int adist_cmp(const void *a, const void *b) {
    const adist_item *x = (const adist_item *)a, *y = (const adist_item *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row ? 1 : 0);
}
/* UDF + SortExec(TopK): k smallest by (f64 total order, row).  The stock TopK leaves the order among equal keys
 * unspecified; ascending row is this build's convention. */
PQO_API size_t pqo_array_distance_topk(const float *rows, uint64_t n, uint32_t dim, const double *query, int metric,
                                       size_t k, uint32_t *out_rows, double *out_dist) {
    double *col = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    adist_item *it = (adist_item *)malloc(sizeof(adist_item) * (size_t)(n + 1));
    pqo_array_distance_column(rows, n, dim, query, metric, col);
    for (uint64_t i = 0; i < n; ++i) {
        it[i].key = f64_ordered_bits(col[i]);
        it[i].row = (uint32_t)i;
    }
    qsort(it, (size_t)n, sizeof(adist_item), adist_cmp);
    size_t m = k < n ? k : (size_t)n;
    for (size_t i = 0; i < m; ++i) {
        out_rows[i] = it[i].row;
        out_dist[i] = col[it[i].row];
    }
    free(col);
    free(it);
    return m;
}

PQO_API const char *pqo_build_flags(void) { return "gcc -O3 -ffp-contract=off -fno-fast-math"; }
