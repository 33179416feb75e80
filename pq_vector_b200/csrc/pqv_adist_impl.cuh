// pqv_adist_impl.cuh -- host side of the entry points that sit next to the indexed path (included at the end of
// pqv_capi.cu; same translation unit so it shares pqv_ctx / DeviceState):
//
//   pqv_array_distance / pqv_array_distance_topk[_filtered]   the un-indexed `array_distance` arm:
//       DataFusion's built-in UDF + SortExec(TopK), reached from benches/query.rs:79-81 and
//       examples/datafusion_sql.rs:54-55 when no optimizer rule is registered (kernels: pqv_adist.cuh); the filtered form
//       takes the WHERE clause of the scan subtree as a row bitmap.
//   pqv_l2_topk_coalesced / pqv_ivf_search_coalesced           the coalescing front door SURVEY section 8b asks for: the
//       reference API is single-query (search.rs:49-54, exec.rs:43) and its callers are concurrent tokio tasks;
//       calls that arrive while a pass over the table is running are answered together by ONE batched tensor-core pass
//       (pqv_l2_topk with n_queries > 1, or pqv_ivf_search_batch), each caller still receiving exactly its own
//       single-query result.
#pragma once

namespace {

int adist_launch(DeviceState &D, const float *d_data, u64 n, uint32_t dim, const double *h_query, uint32_t metric,
                 double *d_out) {
    PQV_TRY(D.ad_query.ensure(dim));
    CU_TRY(cudaMemcpyAsync(D.ad_query.p, h_query, (size_t)dim * 8, cudaMemcpyHostToDevice, D.stream));
    double qn2 = 0.0;
    if (metric == PQV_METRIC_COSINE)
        for (uint32_t i = 0; i < dim; ++i) qn2 += h_query[i] * h_query[i];  // sequential f64 fold (host code is built without FMA)
    const bool qsmem = dim <= pqv::ADIST_QSMEM_MAX_DIM;
    const uint32_t dim_pad = (dim + pqv::ADIST_BLK - 1) / pqv::ADIST_BLK * pqv::ADIST_BLK;
    const size_t smem = (size_t)pqv::ADIST_WARPS * pqv::ADIST_TILE_BYTES + (qsmem ? (size_t)dim_pad * 8 : 0);
    const u64 groups = (n + 31) / 32;
    const bool vec2 = (dim % 2 == 0) && ((uintptr_t)d_data % 8 == 0);
    const int variant = (metric == PQV_METRIC_COSINE ? 4 : 0) | (vec2 ? 2 : 0) | (qsmem ? 1 : 0);
    cudaError_t attr_err = cudaSuccess;
#define PQV_ADIST_CASE(V, METRIC, VEC2, QS)                                                                                  \
    case V: {                                                                                                                \
        auto *fn = pqv::array_distance_kernel<METRIC, VEC2, QS>;                                                             \
        static int occ = 0;  /* CTAs per SM at this shared-memory size; the grid is persistent */                           \
        static size_t occ_smem = 0;                                                                                          \
        PQV_TRY(ensure_dyn_smem(reinterpret_cast<const void *>(fn), smem));                                                  \
        if (!occ || occ_smem != smem) {                                                                                      \
            attr_err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, pqv::ADIST_WARPS * 32, smem);                 \
            if (attr_err != cudaSuccess) break;                                                                              \
            if (occ < 1) occ = 1;                                                                                            \
            occ_smem = smem;                                                                                                 \
        }                                                                                                                    \
        const uint32_t grid = (uint32_t)std::max<u64>(                                                                       \
            1, std::min<u64>((u64)D.sm_count * occ, (groups + pqv::ADIST_WARPS - 1) / pqv::ADIST_WARPS));                    \
        fn<<<grid, pqv::ADIST_WARPS * 32, smem, D.stream>>>(d_data, n, dim, D.ad_query.p, qn2, d_out);                      \
        break;                                                                                                               \
    }
    switch (variant) {
        PQV_ADIST_CASE(0, pqv::ADIST_L2, false, false)
        PQV_ADIST_CASE(1, pqv::ADIST_L2, false, true)
        PQV_ADIST_CASE(2, pqv::ADIST_L2, true, false)
        PQV_ADIST_CASE(3, pqv::ADIST_L2, true, true)
        PQV_ADIST_CASE(4, pqv::ADIST_COSINE, false, false)
        PQV_ADIST_CASE(5, pqv::ADIST_COSINE, false, true)
        PQV_ADIST_CASE(6, pqv::ADIST_COSINE, true, false)
        PQV_ADIST_CASE(7, pqv::ADIST_COSINE, true, true)
    }
#undef PQV_ADIST_CASE
    CU_TRY(attr_err);
    CU_TRY(cudaGetLastError());
    return PQV_OK;
}

int adist_resolve(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric, Dataset **ds_out) {
    Dataset *ds = find_dataset(ctx, handle);
    if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
    if (!query) return fail(PQV_EINVAL, "query is null");
    // upstream: exec_err!("Both arrays must have the same length") per row
    if (query_len != ds->dim) return fail(PQV_EINVAL, "Both arrays must have the same length (row %u, literal %u)", ds->dim, query_len);
    if (metric != PQV_METRIC_L2 && metric != PQV_METRIC_COSINE) return fail(PQV_EINVAL, "unknown metric %u", metric);
    *ds_out = ds;
    return PQV_OK;
}

}  // namespace

extern "C" {

int pqv_array_distance(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric, double *out) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = nullptr;
    PQV_TRY(adist_resolve(ctx, handle, query, query_len, metric, &ds));
    if (ds->n_rows == 0) return PQV_OK;
    if (!out) return fail(PQV_EINVAL, "out is null");
    // a table spread over several devices: every shard (its own device state, stream and scratch) fills its slice of the
    // column; all passes are enqueued before the first one is waited for
    for (Shard &sh : ds->shards) {
        if (sh.n_rows == 0) continue;
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        PQV_TRY(D.ad_col.ensure(sh.n_rows));
        PQV_TRY(adist_launch(D, sh.d_data, sh.n_rows, ds->dim, query, metric, D.ad_col.p));
        CU_TRY(cudaMemcpyAsync(out + sh.first_row, D.ad_col.p, (size_t)sh.n_rows * 8, cudaMemcpyDeviceToHost, D.stream));
    }
    for (Shard &sh : ds->shards) {
        if (sh.n_rows == 0) continue;
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        CU_TRY(cudaStreamSynchronize(D.stream));
    }
    return PQV_OK;
}

static int adist_topk_impl(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric, uint32_t k,
                           const uint8_t *row_mask, uint32_t *out_row_idx, double *out_dist, uint32_t *out_count) {
    if (!ctx || !out_count) return fail(PQV_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dataset *ds = nullptr;
    PQV_TRY(adist_resolve(ctx, handle, query, query_len, metric, &ds));
    if (k == 0 || k > PQV_MAX_K) return fail(PQV_EINVAL, "k must be in [1, %d]", PQV_MAX_K);
    if (!out_row_idx || !out_dist) return fail(PQV_EINVAL, "null argument");
    *out_count = 0;
    if (ds->n_rows == 0) return PQV_OK;
    // Every shard (one per device of the context) selects its own k smallest; the host merges them by (distance, row).  All
    // passes are enqueued before the first one is waited for.
    struct Part {
        DeviceState *D;
        u64 first_row;
        uint32_t k_eff;
    };
    std::vector<Part> parts;
    const u64 total_mask_bytes = (ds->n_rows + 7) / 8;
    for (Shard &sh : ds->shards) {
        const u64 n = sh.n_rows;
        if (n == 0) continue;
        DeviceState &D = ctx->devs[sh.di];
        DevGuard guard(D.dev);
        // the scan subtree's filter: rows whose bit is clear never reach the sort (FilterExec below SortExec)
        const uint32_t *d_mask = nullptr;
        u64 live = n;
        if (row_mask) {
            // this shard's bits [first_row, first_row + n) of the caller's bitmap, moved down to bit 0
            const u64 nb = (n + 7) / 8, b0 = sh.first_row >> 3;
            const unsigned shft = (unsigned)(sh.first_row & 7);
            std::vector<uint8_t> m(nb);
            for (u64 j = 0; j < nb; ++j) {
                unsigned v = row_mask[b0 + j] >> shft;
                if (shft && b0 + j + 1 < total_mask_bytes) v |= (unsigned)row_mask[b0 + j + 1] << (8 - shft);
                m[j] = (uint8_t)v;
            }
            if (n & 7) m[nb - 1] &= (uint8_t)((1u << (n & 7)) - 1u);
            live = 0;
            for (u64 j = 0; j < nb; ++j) live += (u64)__builtin_popcount(m[j]);
            if (live == 0) continue;
            const u64 n_words = (n + 31) / 32;
            PQV_TRY(D.vt_bitmap.ensure(n_words));
            CU_TRY(cudaMemsetAsync(D.vt_bitmap.p + (n_words - 1), 0, 4, D.stream));
            CU_TRY(cudaMemcpyAsync(D.vt_bitmap.p, m.data(), (size_t)nb, cudaMemcpyHostToDevice, D.stream));
            CU_TRY(cudaStreamSynchronize(D.stream));  // `m` goes out of scope
            d_mask = D.vt_bitmap.p;
        }
        const uint32_t k_eff = (uint32_t)std::min<u64>(k, live);
        PQV_TRY(D.ad_col.ensure(n));
        PQV_TRY(D.ad_state.ensure(1));
        PQV_TRY(D.ad_out_dist.ensure(k_eff));
        PQV_TRY(D.ad_out_row.ensure(k_eff));
        PQV_TRY(D.h_ad_dist.ensure(PQV_MAX_K));
        PQV_TRY(D.h_ad_row.ensure(PQV_MAX_K));
        PQV_TRY(D.h_ad_state.ensure(1));
        CU_TRY(cudaEventRecord(D.ev[0], D.stream));
        PQV_TRY(adist_launch(D, sh.d_data, n, ds->dim, query, metric, D.ad_col.p));
        CU_TRY(cudaEventRecord(D.ev[1], D.stream));
        const uint32_t hgrid = (uint32_t)std::max<u64>(1, std::min<u64>((u64)D.sm_count * 8, (n + 255) / 256));
        pqv::sel_init_kernel<<<1, 256, 0, D.stream>>>(D.ad_state.p, k_eff);
        for (int pass = 0; pass < 12; ++pass) {
            pqv::sel_hist_kernel<<<hgrid, 256, 0, D.stream>>>(D.ad_col.p, n, pass, D.ad_state.p, d_mask);
            pqv::sel_pick_kernel<<<1, 256, 0, D.stream>>>(pass, D.ad_state.p);
        }
        pqv::sel_collect_kernel<<<hgrid, 256, 0, D.stream>>>(D.ad_col.p, n, D.ad_state.p, k_eff, D.ad_out_dist.p, D.ad_out_row.p, d_mask);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(D.ev[2], D.stream));
        CU_TRY(cudaMemcpyAsync(D.h_ad_dist.p, D.ad_out_dist.p, (size_t)k_eff * 8, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaMemcpyAsync(D.h_ad_row.p, D.ad_out_row.p, (size_t)k_eff * 4, cudaMemcpyDeviceToHost, D.stream));
        CU_TRY(cudaMemcpyAsync(D.h_ad_state.p, D.ad_state.p, sizeof(pqv::SelState), cudaMemcpyDeviceToHost, D.stream));
        parts.push_back(Part{&D, sh.first_row, k_eff});
    }
    std::vector<double> hd;
    std::vector<uint32_t> hr;
    ctx->last = pqv_timing{};
    for (Part &P : parts) {
        DeviceState &D = *P.D;
        DevGuard guard(D.dev);
        CU_TRY(cudaStreamSynchronize(D.stream));
        if (D.h_ad_state.p->out_count != P.k_eff)
            return fail(PQV_ECUDA, "array_distance top-k: selected %u keys, expected %u", D.h_ad_state.p->out_count, P.k_eff);
        float ms_scan = 0.f, ms_sel = 0.f;
        cudaEventElapsedTime(&ms_scan, D.ev[0], D.ev[1]);
        cudaEventElapsedTime(&ms_sel, D.ev[1], D.ev[2]);
        // the slowest shard is the call's device time
        if (ms_scan + ms_sel > ctx->last.total_ms) {
            ctx->last.scan_ms = ms_scan;
            ctx->last.post_ms = ms_sel;
            ctx->last.total_ms = ms_scan + ms_sel;
        }
        for (uint32_t i = 0; i < P.k_eff; ++i) {
            hd.push_back(D.h_ad_dist.p[i]);
            hr.push_back((uint32_t)(D.h_ad_row.p[i] + P.first_row));
        }
    }
    ctx->last.scan_bytes = ds->n_rows * (u64)ds->dim * 4;
    ctx->last.launches = 27 * (uint32_t)parts.size();
    // ascending by (f64 total order with NaN last, row)
    auto ordered = [](double d) {
        u64 b;
        memcpy(&b, &d, 8);
        if ((b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) b = 0x7FF8000000000000ull;
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    };
    std::vector<uint32_t> idx(hd.size());
    for (uint32_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
        const u64 ua = ordered(hd[a]), ub = ordered(hd[b]);
        return ua != ub ? ua < ub : hr[a] < hr[b];
    });
    const uint32_t n_out = (uint32_t)std::min<size_t>(k, idx.size());
    for (uint32_t i = 0; i < n_out; ++i) {
        out_row_idx[i] = hr[idx[i]];
        out_dist[i] = hd[idx[i]];
    }
    *out_count = n_out;
    return PQV_OK;
}

int pqv_array_distance_topk(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric, uint32_t k,
                            uint32_t *out_row_idx, double *out_dist, uint32_t *out_count) {
    return adist_topk_impl(ctx, handle, query, query_len, metric, k, nullptr, out_row_idx, out_dist, out_count);
}

int pqv_array_distance_topk_filtered(pqv_ctx *ctx, uint64_t handle, const double *query, uint32_t query_len, uint32_t metric,
                                     uint32_t k, const uint8_t *row_mask, uint32_t *out_row_idx, double *out_dist,
                                     uint32_t *out_count) {
    return adist_topk_impl(ctx, handle, query, query_len, metric, k, row_mask, out_row_idx, out_dist, out_count);
}

// ---- coalescing front door ---------------------------------------------------------------------------------------------
int pqv_coalesce_config(pqv_ctx *ctx, uint32_t max_batch, uint32_t window_us) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    if (max_batch == 0) return fail(PQV_EINVAL, "max_batch must be > 0");
    std::lock_guard<std::mutex> lk(ctx->co.m);
    ctx->co.max_batch = max_batch;
    ctx->co.window_us = window_us;
    return PQV_OK;
}

int pqv_coalesce_stats(pqv_ctx *ctx, uint64_t *out_queries, uint64_t *out_batches, uint64_t *out_max_batch) {
    if (!ctx) return fail(PQV_EINVAL, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->co.m);
    if (out_queries) *out_queries = ctx->co.n_queries;
    if (out_batches) *out_batches = ctx->co.n_batches;
    if (out_max_batch) *out_max_batch = ctx->co.max_seen;
    return PQV_OK;
}

// index == 0: brute force (pqv_l2_topk); otherwise an IVF search over (table, index) with `nprobe` (pqv_ivf_search)
static int coalesced_submit(pqv_ctx *ctx, uint64_t handle, uint64_t index, uint32_t nprobe, const float *query, uint32_t k,
                            uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!ctx || !query || !out_row_idx || !out_dist || !out_count) return fail(PQV_EINVAL, "null argument");
    if (index && nprobe == 0) return fail(PQV_EINVAL, "nprobe must be > 0");  // src/ivf/search.rs:72
    typedef pqv_ctx::CoalesceReq Req;
    pqv_ctx::Coalescer &co = ctx->co;
    // reject what the reference rejects per call (search.rs:66-74, 91-98) before the request can join a batch.  ctx->mu
    // is held by a running pass for its whole duration, so the dimension comes from the coalescer's own cache (handles
    // are never reused); only the first call per dataset looks it up under ctx->mu.
    uint32_t dim = 0;
    {
        std::lock_guard<std::mutex> lk(co.m);
        auto it = co.dims.find(handle);
        if (it != co.dims.end()) dim = it->second;
    }
    if (!dim) {
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            Dataset *ds = find_dataset(ctx, handle);
            if (!ds) return fail(PQV_EHANDLE, "unknown dataset handle %llu", (unsigned long long)handle);
            dim = ds->dim;
        }
        std::lock_guard<std::mutex> lk(co.m);
        co.dims[handle] = dim;
    }
    PQV_TRY(check_topk_args(k, dim, flags));
    Req me;
    me.handle = handle;
    me.index = index;
    me.nprobe = nprobe;
    me.k = k;
    me.flags = flags;
    me.query = query;
    me.rows = out_row_idx;
    me.dist = out_dist;
    me.count = out_count;

    std::unique_lock<std::mutex> lk(co.m);
    co.pending.push_back(&me);
    if (!co.leader_active) {
        co.leader_active = true;
        me.lead = true;
    } else {
        if (co.window_us) co.cv.notify_all();  // a lingering leader re-checks the batch size
        co.cv.wait(lk, [&] { return me.done || me.lead; });
    }
    if (me.done) {  // a leader answered this request
        if (me.status) g_err = me.err;
        return me.status;
    }
    // leader: `me` is the oldest pending request.  Optionally linger so a burst can assemble, then take every pending
    // request with the same (dataset, k, flags) in arrival order; the rest wait for the next leader.
    if (co.window_us && co.pending.size() < co.max_batch)
        co.cv.wait_for(lk, std::chrono::microseconds(co.window_us), [&] { return co.pending.size() >= co.max_batch; });
    std::vector<Req *> batch;
    std::deque<Req *> rest;
    for (Req *r : co.pending) {
        if (batch.size() < co.max_batch && r->handle == me.handle && r->index == me.index && r->nprobe == me.nprobe &&
            r->k == me.k && r->flags == me.flags)
            batch.push_back(r);
        else rest.push_back(r);
    }
    co.pending.swap(rest);
    lk.unlock();

    const uint32_t nq = (uint32_t)batch.size();
    int rc = PQV_OK;
    std::string err;
    try {  // nothing may escape before the followers are released below
        if (nq == 1) {
            rc = index ? pqv_ivf_search(ctx, handle, index, query, k, nprobe, flags, out_row_idx, out_dist, out_count)
                       : pqv_l2_topk(ctx, handle, query, 1, k, flags, out_row_idx, out_dist, out_count);
            if (rc) err = g_err;
        } else {
            std::vector<float> q((size_t)nq * dim);
            std::vector<uint32_t> rows((size_t)nq * k), counts(nq);
            std::vector<float> dist((size_t)nq * k);
            for (uint32_t i = 0; i < nq; ++i) memcpy(q.data() + (size_t)i * dim, batch[i]->query, (size_t)dim * 4);
            rc = index ? pqv_ivf_search_batch(ctx, handle, index, q.data(), nq, k, nprobe, flags, rows.data(), dist.data(), counts.data())
                       : pqv_l2_topk(ctx, handle, q.data(), nq, k, flags, rows.data(), dist.data(), counts.data());
            if (rc) err = g_err;
            else
                for (uint32_t i = 0; i < nq; ++i) {
                    memcpy(batch[i]->rows, rows.data() + (size_t)i * k, (size_t)counts[i] * 4);
                    memcpy(batch[i]->dist, dist.data() + (size_t)i * k, (size_t)counts[i] * 4);
                    *batch[i]->count = counts[i];
                }
        }
    } catch (const std::exception &e) {
        rc = PQV_ENOMEM;
        err = std::string("coalesced batch failed: ") + e.what();
    }

    lk.lock();
    co.n_queries += nq;
    co.n_batches += 1;
    co.max_seen = std::max<u64>(co.max_seen, nq);
    for (Req *r : batch) {
        if (r == &me) continue;
        r->status = rc;
        if (rc) r->err = err;
        r->done = true;  // last write to *r: its owner may return (and free it) as soon as the lock is released
    }
    if (co.pending.empty()) co.leader_active = false;
    else co.pending.front()->lead = true;
    lk.unlock();
    co.cv.notify_all();
    if (rc) g_err = err;
    return rc;
}


int pqv_l2_topk_coalesced(pqv_ctx *ctx, uint64_t handle, const float *query, uint32_t k, uint32_t flags, uint32_t *out_row_idx,
                          float *out_dist, uint32_t *out_count) {
    return coalesced_submit(ctx, handle, 0, 0, query, k, flags, out_row_idx, out_dist, out_count);
}

int pqv_ivf_search_coalesced(pqv_ctx *ctx, uint64_t handle, uint64_t index, const float *query, uint32_t k, uint32_t nprobe,
                             uint32_t flags, uint32_t *out_row_idx, float *out_dist, uint32_t *out_count) {
    if (!index) return fail(PQV_EHANDLE, "unknown index handle 0");
    return coalesced_submit(ctx, handle, index, nprobe, query, k, flags, out_row_idx, out_dist, out_count);
}

}  // extern "C"
