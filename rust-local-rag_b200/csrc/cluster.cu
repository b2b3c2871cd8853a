// cluster.cu -- rlr_cluster_*: ONE host process (the reference is one tokio process holding one
// `Arc<RwLock<RagEngine>>`, /root/reference/src/main.rs:140-167) driving a row-sharded chunk store
// on several GPUs of an NVSwitch box through the same calls as a single-GPU store.
//
// Per query the calling host thread enqueues, on one stream per GPU and with no host round trip
// in between:
//   every shard g : H2D of the normalised query (+ the shard's slice of the lexical pairs), then
//                   scan_topm_kernel with a ScanPost: the kernel's last CTA stores the shard's
//                   top-m list into the lane's mailbox in shard 0's HBM (peer stores over NVLink;
//                   peer access is enabled with cudaDeviceEnablePeerAccess, no IPC handles are
//                   needed inside one process) and publishes the sequence number;
//   shard 0 (root): mailbox_merge_kernel (waits in-kernel for the flags, merges, frees the slot),
//                   then MMR whose pairwise kernel dereferences peer pointers into the owning
//                   GPUs' HBM, then the D2H of the <= top_k result records.
// The host waits once, on the root's stream.  Shards are launched non-root first: the root owns
// fewer rows (tail-balanced plan) and can afford to start last.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "api_internal.hpp"

using namespace rlr_api;

namespace {

// one lane of a cluster: everything a query in flight needs on every GPU
struct ClusterCtx {
    std::vector<rlr_ctx *> c;          // one workspace + stream per shard, on that shard's device
    std::vector<rlr_ctx *> root_extra; // throughput mode: pool / MMR buffers of queries 1.. on the root (created on first use)
    struct LexBuf { uint32_t *h_rows = nullptr, *d_rows = nullptr; float *h_norm = nullptr, *d_norm = nullptr; };
    std::vector<std::vector<LexBuf>> lex_extra;   // throughput mode: [shard][query - 1] lexical pairs of queries 1.. (query 0 uses the ctx's)
    rlr_mailbox *mb = nullptr;         // in the root's HBM, private to this lane
    uint64_t seq = 0;                  // sequence numbers of this lane's mailbox start at 1
    uint32_t *h_status = nullptr;      // pinned: sticky mailbox status read back with every result
    // batched path: per-shard key lists (on the shard's GPU), the gathered lists + merged result on the root, pinned mirror
    std::vector<void *> batch_keys;
    void *batch_all = nullptr, *batch_out = nullptr, *batch_cnt = nullptr;
    unsigned long long *h_batch = nullptr;
    size_t batch_cap = 0;              // n_queries * m the buffers hold
};

thread_local std::vector<float> g_shard_scan_ms;

// merge + MMR tail of the root expressed in bytes of scanning (0.115 ms at ~7.3 TB/s, measured on B200,
// profiles/): the default plan gives the root that many fewer bytes so that all GPUs finish together
constexpr double kTailBytes = 0.115e-3 * 7.3e12;

} // namespace

struct rlr_cluster {
    uint32_t n = 0;                    // shards
    uint32_t dim = 0, pitch = 0, pitch16 = 0, flags = 0;
    uint64_t n_rows = 0;
    std::vector<int> device;
    std::vector<rlr_store *> shard;
    std::vector<uint64_t> row_base, shard_rows;
    rlr::PeerTable table32, table16;   // peer pointers of every shard (UVA; valid on the root after enabling peer access)
    std::mutex mu;
    std::vector<ClusterCtx *> free_ctx;
    std::atomic<uint64_t> launches{0};
};

namespace {

void cctx_free(rlr_cluster *cl, ClusterCtx *cc)
{
    if (!cc) return;
    for (size_t g = 0; g < cc->c.size(); ++g)
        if (cc->c[g]) { cudaSetDevice(cl->device[g]); ctx_free(cc->c[g]); }
    for (rlr_ctx *x : cc->root_extra) { cudaSetDevice(cl->device[0]); ctx_free(x); }
    for (size_t g = 0; g < cc->lex_extra.size(); ++g)
        for (auto &b : cc->lex_extra[g]) {
            cudaSetDevice(cl->device[g]);
            cudaFree(b.d_rows); cudaFree(b.d_norm); cudaFreeHost(b.h_rows); cudaFreeHost(b.h_norm);
        }
    for (size_t g = 0; g < cc->batch_keys.size(); ++g) if (cc->batch_keys[g]) { cudaSetDevice(cl->device[g]); cudaFree(cc->batch_keys[g]); }
    cudaSetDevice(cl->device[0]);
    cudaFree(cc->batch_all); cudaFree(cc->batch_out); cudaFree(cc->batch_cnt);
    if (cc->h_batch) cudaFreeHost(cc->h_batch);
    if (cc->mb) rlr_mailbox_close(cc->mb);
    if (cc->h_status) cudaFreeHost(cc->h_status);
    cudaGetLastError();
    delete cc;
}

int cctx_new(rlr_cluster *cl, ClusterCtx **out)
{
    ClusterCtx *cc = new ClusterCtx();
    cc->c.assign(cl->n, nullptr);
    for (uint32_t g = 0; g < cl->n; ++g) {
        int rc = ensure_device(cl->device[g]);
        if (rc == RLR_OK) rc = ctx_new(cl->shard[g], &cc->c[g]);
        if (rc != RLR_OK) { std::string keep = g_err; cctx_free(cl, cc); g_err = keep; return rc; }
    }
    if (cl->n > 1) {
        int rc = rlr_mailbox_create(cl->device[0], cl->n, RLR_MAX_M, 4, &cc->mb);
        if (rc != RLR_OK) { std::string keep = g_err; cctx_free(cl, cc); g_err = keep; return rc; }
    }
    cudaError_t e = cudaMallocHost(&cc->h_status, 2 * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaGetLastError(); cctx_free(cl, cc); return fail(RLR_ERR_OOM, "cudaMallocHost failed: %s", cudaGetErrorString(e)); }
    cc->h_status[0] = cc->h_status[1] = 0;
    *out = cc;
    return RLR_OK;
}

struct ClusterLease {
    rlr_cluster *cl;
    ClusterCtx *cc = nullptr;
    explicit ClusterLease(rlr_cluster *c) : cl(c) {}
    int acquire()
    {
        {
            std::lock_guard<std::mutex> lk(cl->mu);
            if (!cl->free_ctx.empty()) { cc = cl->free_ctx.back(); cl->free_ctx.pop_back(); }
        }
        if (cc) return RLR_OK;
        return cctx_new(cl, &cc);
    }
    ~ClusterLease()
    {
        if (cc) { std::lock_guard<std::mutex> lk(cl->mu); cl->free_ctx.push_back(cc); }
    }
};

int check_cluster(const rlr_cluster *c)
{
    if (!c) return fail(RLR_ERR_INVALID_ARG, "cluster is NULL");
    return RLR_OK;
}

uint32_t owner_of(const rlr_cluster *cl, uint64_t row)
{
    // shards are contiguous and few: upper_bound over row_base
    uint32_t g = static_cast<uint32_t>(std::upper_bound(cl->row_base.begin(), cl->row_base.end(), row) - cl->row_base.begin());
    return g == 0 ? 0 : g - 1;
}

// default plan: even blocks, the root short by the merge + MMR tail
void default_plan(uint64_t n_rows, uint32_t n, uint32_t row_bytes, std::vector<uint64_t> *rows)
{
    rows->assign(n, 0);
    if (n == 1) { (*rows)[0] = n_rows; return; }
    const double even = static_cast<double>(n_rows) / n;
    const double tail_rows = kTailBytes / row_bytes;
    double head = even - tail_rows * (n - 1) / n;
    if (head < even / 2) head = even / 2;          // small stores: the scan is not bandwidth-proportional anyway
    uint64_t h = static_cast<uint64_t>(head);
    if (h < 1) h = 1;
    (*rows)[0] = h;
    const uint64_t rest = n_rows - h;
    for (uint32_t g = 1; g < n; ++g) (*rows)[g] = rest * g / (n - 1) - rest * (g - 1) / (n - 1);
}

int enable_peers(const std::vector<int> &dev)
{
    const int root = dev[0];
    for (size_t g = 1; g < dev.size(); ++g) {
        const int d = dev[g];
        if (d == root) continue;
        for (int dir = 0; dir < 2; ++dir) {
            const int from = dir == 0 ? root : d, to = dir == 0 ? d : root;
            int can = 0;
            CU_TRY(cudaDeviceCanAccessPeer(&can, from, to));
            if (!can) return fail(RLR_ERR_UNSUPPORTED, "GPU %d cannot access GPU %d's memory (no peer access): a cluster needs NVLink/PCIe P2P", from, to);
            CU_TRY(cudaSetDevice(from));
            cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); return fail(RLR_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", from, to, cudaGetErrorString(e)); }
        }
    }
    return RLR_OK;
}

// the arguments of one search, after validation
struct Plan {
    uint32_t m = 0;          // records each shard delivers / the merged pool holds
    bool do_mmr = false;
    uint32_t top_k = 0;
    float lambda = 0.0f;
};

// :505-530 for every shard: (sorted local rows, score / GLOBAL max_lexical), later duplicates win
int stage_lex_shards(rlr_cluster *cl, ClusterCtx *cc, const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex,
                     std::vector<uint32_t> *n_out, uint32_t q = 0 /* throughput mode: which query's buffers */)
{
    n_out->assign(cl->n, 0);
    if (n_lex == 0) return RLR_OK;
    if (!lex_rows || !lex_scores) return fail(RLR_ERR_INVALID_ARG, "n_lex > 0 but lex_rows/lex_scores is NULL");
    if (n_lex > kLexCap) return fail(RLR_ERR_UNSUPPORTED, "n_lex %u exceeds %u", n_lex, kLexCap);
    float max_lexical = 0.0f; // fold(0.0_f32, f32::max).max(f32::EPSILON), :511-515
    for (uint32_t i = 0; i < n_lex; ++i) max_lexical = fmaxf(max_lexical, lex_scores[i]);
    max_lexical = fmaxf(max_lexical, 1.1920929e-07f);
    std::vector<std::pair<uint32_t, uint32_t>> order; // (global row, original index)
    order.reserve(n_lex);
    for (uint32_t i = 0; i < n_lex; ++i)
        if (lex_rows[i] < cl->n_rows) order.emplace_back(lex_rows[i], i);      // `if let Some(chunk)`, :525
    std::stable_sort(order.begin(), order.end(),
                     [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
    for (size_t i = 0; i < order.size(); ++i) {
        if (i + 1 < order.size() && order[i + 1].first == order[i].first) continue;
        const uint32_t g = owner_of(cl, order[i].first);
        rlr_ctx *c = cc->c[g];
        uint32_t *h_rows = q == 0 ? c->h_lex_rows : cc->lex_extra[g][q - 1].h_rows;
        float *h_norm = q == 0 ? c->h_lex_norm : cc->lex_extra[g][q - 1].h_norm;
        uint32_t &n = (*n_out)[g];
        h_rows[n] = static_cast<uint32_t>(order[i].first - cl->row_base[g]);
        h_norm[n] = lex_scores[order[i].second] / max_lexical; // :527-530
        ++n;
    }
    return RLR_OK;
}

int cluster_search(rlr_cluster *cl, const float *query, uint32_t dim, uint32_t flags, const Plan &pl,
                   const rlr_resolved_weights *w, const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex,
                   uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex, uint32_t *out_n)
{
    ClusterLease lease(cl);
    if (int rc = lease.acquire()) return rc;
    ClusterCtx *cc = lease.cc;
    rlr_ctx *r0 = cc->c[0];
    if (!query) return fail(RLR_ERR_INVALID_ARG, "query is NULL");
    if (dim != cl->dim)
        return fail(RLR_ERR_DIM_MISMATCH, "query has %u dims, store has %u (the reference would silently truncate, "
                    "src/rag_engine.rs:1778; this library refuses)", dim, cl->dim);
    // the query is staged ONCE in the root lane's pinned buffer; every GPU copies from there
    memcpy(r0->h_query, query, dim * sizeof(float));
    for (uint32_t i = 0; i < dim; ++i)
        if (!std::isfinite(r0->h_query[i])) return fail(RLR_ERR_NONFINITE, "query[%u] is not finite", i);
    if (!(flags & RLR_QUERY_PRENORMALIZED)) host_normalize(r0->h_query, dim);   // :494
    const size_t q_floats = ((cl->dim + 63u) & ~63u) + 128;
    std::vector<uint32_t> nl;
    if (int rc = stage_lex_shards(cl, cc, lex_rows, lex_scores, n_lex, &nl)) return rc;

    const bool timed = flags & RLR_WANT_TIMINGS;
    const uint64_t seq = ++cc->seq;
    const uint32_t slot = static_cast<uint32_t>(seq % cc->mb->ring);
    uint64_t launches = 0;
    for (uint32_t k = 0; k < cl->n; ++k) {
        const uint32_t g = (k + 1) % cl->n;            // 1, 2, ..., n-1, 0: the root (fewest rows) starts last
        rlr_store *s = cl->shard[g];
        rlr_ctx *c = cc->c[g];
        cudaStream_t st = c->stream;
        CU_TRY(cudaSetDevice(s->device));
        CU_TRY(cudaMemcpyAsync(c->d_query, r0->h_query, q_floats * sizeof(float), cudaMemcpyHostToDevice, st));
        if (nl[g]) {
            CU_TRY(cudaMemcpyAsync(c->d_lex_rows, c->h_lex_rows, nl[g] * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            CU_TRY(cudaMemcpyAsync(c->d_lex_norm, c->h_lex_norm, nl[g] * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        const bool half = s->use_half(flags);
        rlr::ScanArgs a;
        memset(&a, 0, sizeof a);
        rlr::scan_plan(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), half ? s->pitch16 : s->pitch, half, &a);
        a.tmap = half ? &s->tmap16 : &s->tmap;
        a.d_query = c->d_query;
        a.n_rows = static_cast<uint32_t>(s->n_rows);
        a.row_base = static_cast<uint32_t>(s->row_base);
        a.pitch = half ? s->pitch16 : s->pitch;
        a.w_embed = w->embedding; a.w_lex = w->lexical;
        a.d_lex_rows = c->d_lex_rows; a.d_lex_norm = c->d_lex_norm; a.n_lex = nl[g];
        a.m = pl.m;
        a.d_lists = c->d_lists; a.d_counts = c->d_counts; a.d_ticket = c->d_ticket; a.d_pub = c->d_pub;
        a.d_out = cc->mb->list(slot, g); a.d_out_n = cc->mb->count(slot, g);
        a.post.flag = cc->mb->flag(slot, g);
        a.post.consumed = cc->mb->consumed(slot);
        a.post.seq = seq; a.post.ring = cc->mb->ring; a.post.status = cc->mb->d_status;
        if (timed) CU_TRY(cudaEventRecord(c->ev[0], st));
        CU_TRY(rlr::scan_launch(a, st));
        if (timed) CU_TRY(cudaEventRecord(c->ev[1], st));
        ++launches;
    }
    // root tail (the root's device is current: it was launched last)
    rlr_store *s0 = cl->shard[0];
    cudaStream_t st0 = r0->stream;
    CU_TRY(rlr::mailbox_merge_launch(cc->mb->list(slot, 0), cc->mb->m_cap, cc->mb->flag(slot, 0), seq, cc->mb->consumed(slot),
                                     cl->n, pl.m, r0->d_pool, r0->d_pool_n, cc->mb->d_status, st0));
    ++launches;
    if (timed) CU_TRY(cudaEventRecord(r0->ev[2], st0));
    const uint8_t *d_res_blk = r0->d_pool_blk;          // [n | records]: one D2H for both
    uint32_t cap = pl.m;
    if (pl.do_mmr) {
        const bool half = s0->use_half(flags);
        rlr::MmrArgs a;
        memset(&a, 0, sizeof a);
        a.half = half;
        a.d_emb = nullptr; a.pitch = half ? cl->pitch16 : cl->pitch; a.dim = cl->dim;
        a.d_cands = r0->d_pool; a.d_n = r0->d_pool_n;
        a.use_rows = 1; a.p_cap = pl.m; a.top_k = pl.top_k; a.lambda = pl.lambda;
        a.d_tri = r0->d_tri; a.d_sel_pos = r0->d_sel_pos; a.d_sel_n = r0->d_sel_n; a.d_result = r0->d_result;
        a.max_smem_optin = s0->smem_optin;
        a.peers = half ? &cl->table16 : &cl->table32;
        a.d_gather = r0->d_gather;
        uint32_t l = 0;
        CU_TRY(rlr::mmr_launch(a, st0, &l));
        launches += l;
        d_res_blk = r0->d_result_blk;
        cap = std::min<uint32_t>(pl.m, std::max<uint32_t>(pl.top_k, 1));
    }
    if (timed) CU_TRY(cudaEventRecord(r0->ev[3], st0));
    CU_TRY(cudaMemcpyAsync(r0->h_result_blk, d_res_blk, 16 + cap * sizeof(rlr_cand), cudaMemcpyDeviceToHost, st0));
    CU_TRY(cudaStreamSynchronize(st0));
    cl->launches += launches;
    const uint32_t n = std::min(r0->h_result_n[0], cap);
    if (n == 0) {
        // every shard owns rows, so an empty result can only mean that a mailbox wait timed out (the kernels then
        // deliver an EMPTY list and set the sticky status word, scan_topm.cu / merge.cu): read it on this rare path only
        CU_TRY(cudaMemcpy(cc->h_status, cc->mb->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        return fail(RLR_ERR_CUDA, "the cluster delivered no result (mailbox status %u): a GPU did not deliver its list in time", cc->h_status[0]);
    }
    unpack(r0->h_result, n, out_rows, out_score, out_emb, out_lex);
    *out_n = n;
    if (timed) {
        rlr_timings t = {0, 0, 0, 0, 0};
        g_shard_scan_ms.assign(cl->n, 0.0f);
        for (uint32_t g = 0; g < cl->n; ++g) {
            cudaSetDevice(cl->device[g]);
            cudaEventElapsedTime(&g_shard_scan_ms[g], cc->c[g]->ev[0], cc->c[g]->ev[1]);
            t.scan_ms = std::max(t.scan_ms, g_shard_scan_ms[g]);
        }
        cudaSetDevice(cl->device[0]);
        cudaEventElapsedTime(&t.merge_ms, r0->ev[1], r0->ev[2]);     // root: end of its scan -> merged pool (includes waiting for the slowest GPU)
        if (pl.do_mmr) cudaEventElapsedTime(&t.mmr_ms, r0->ev[2], r0->ev[3]);
        cudaEventElapsedTime(&t.total_ms, r0->ev[0], r0->ev[3]);
        cudaGetLastError();
        t.launches = static_cast<uint32_t>(launches);
        g_timings = t;
    }
    return RLR_OK;
}

// throughput mode over the cluster: nq queries, ONE pass over every shard (query groups), nq posts per shard into
// consecutive mailbox slots, nq merges + MMRs on the root
int cluster_search_multi(rlr_cluster *cl, const float *queries, uint32_t nq, uint32_t dim, uint32_t flags, uint32_t pool,
                         bool do_mmr, uint32_t top_k, float lambda, const rlr_resolved_weights *w,
                         const uint32_t *const *lex_rows, const float *const *lex_scores, const uint32_t *n_lex, uint32_t *out_rows,
                         float *out_score, float *out_emb, float *out_lex, uint32_t *out_n)
{
    ClusterLease lease(cl);
    if (int rc = lease.acquire()) return rc;
    ClusterCtx *cc = lease.cc;
    rlr_ctx *r0 = cc->c[0];
    if (dim != cl->dim) return fail(RLR_ERR_DIM_MISMATCH, "queries have %u dims, store has %u", dim, cl->dim);
    while (cc->root_extra.size() + 1 < nq) {
        rlr_ctx *x = nullptr;
        if (int rc = ensure_device(cl->device[0])) return rc;
        if (int rc = ctx_new(cl->shard[0], &x)) return rc;
        cc->root_extra.push_back(x);
    }
    for (uint32_t q = 0; q < nq; ++q) {
        float *hq = r0->h_query + static_cast<size_t>(q) * rlr::kQueryCap;
        memcpy(hq, queries + static_cast<size_t>(q) * dim, dim * sizeof(float));
        for (uint32_t i = 0; i < dim; ++i)
            if (!std::isfinite(hq[i])) return fail(RLR_ERR_NONFINITE, "queries[%u][%u] is not finite", q, i);
        if (!(flags & RLR_QUERY_PRENORMALIZED)) host_normalize(hq, dim);
    }
    // lexical pairs, per query: routed to the owning shards, normalised by that query's global maximum (:505-530)
    std::vector<std::vector<uint32_t>> nl(nq, std::vector<uint32_t>(cl->n, 0));
    bool any_lex = false;
    for (uint32_t q = 0; q < nq; ++q) any_lex |= n_lex != nullptr && n_lex[q] != 0;
    if (any_lex) {
        if (!lex_rows || !lex_scores) return fail(RLR_ERR_INVALID_ARG, "n_lex > 0 but lex_rows/lex_scores is NULL");
        if (cc->lex_extra.size() < cl->n) cc->lex_extra.resize(cl->n);
        for (uint32_t g = 0; g < cl->n; ++g)
            while (cc->lex_extra[g].size() + 1 < nq) {
                if (int rc = ensure_device(cl->device[g])) return rc;
                ClusterCtx::LexBuf b;
                cudaError_t e = cudaMallocHost(&b.h_rows, kLexCap * sizeof(uint32_t));
                if (e == cudaSuccess) e = cudaMallocHost(&b.h_norm, kLexCap * sizeof(float));
                if (e == cudaSuccess) e = cudaMalloc(&b.d_rows, kLexCap * sizeof(uint32_t));
                if (e == cudaSuccess) e = cudaMalloc(&b.d_norm, kLexCap * sizeof(float));
                cc->lex_extra[g].push_back(b);          // owned (and freed) by the lane even when an allocation failed
                if (e != cudaSuccess) { cudaGetLastError(); return fail(RLR_ERR_OOM, "lexical staging allocation failed: %s", cudaGetErrorString(e)); }
            }
        for (uint32_t q = 0; q < nq; ++q)
            if (n_lex[q])
                if (int rc = stage_lex_shards(cl, cc, lex_rows[q], lex_scores[q], n_lex[q], &nl[q], q)) return rc;
    }
    const uint64_t seq0 = cc->seq + 1;
    cc->seq += nq;
    uint64_t launches = 0;
    for (uint32_t k = 0; k < cl->n; ++k) {
        const uint32_t g = (k + 1) % cl->n;
        rlr_store *s = cl->shard[g];
        rlr_ctx *c = cc->c[g];
        cudaStream_t st = c->stream;
        CU_TRY(cudaSetDevice(s->device));
        CU_TRY(cudaMemcpyAsync(c->d_query, r0->h_query, static_cast<size_t>(nq) * rlr::kQueryCap * sizeof(float), cudaMemcpyHostToDevice, st));
        for (uint32_t q = 0; q < nq; ++q) {
            if (!nl[q][g]) continue;
            const uint32_t *hr = q == 0 ? c->h_lex_rows : cc->lex_extra[g][q - 1].h_rows;
            const float *hn = q == 0 ? c->h_lex_norm : cc->lex_extra[g][q - 1].h_norm;
            uint32_t *dr = q == 0 ? c->d_lex_rows : cc->lex_extra[g][q - 1].d_rows;
            float *dn = q == 0 ? c->d_lex_norm : cc->lex_extra[g][q - 1].d_norm;
            CU_TRY(cudaMemcpyAsync(dr, hr, nl[q][g] * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            CU_TRY(cudaMemcpyAsync(dn, hn, nl[q][g] * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        const bool half = s->use_half(flags);
        rlr::ScanArgs a;
        memset(&a, 0, sizeof a);
        rlr::scan_plan(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), half ? s->pitch16 : s->pitch, half, &a, nq);
        if (a.grid <= 0) return fail(RLR_ERR_UNSUPPORTED, "%u query groups do not fit this store's row size in shared memory", nq);
        a.tmap = half ? &s->tmap16 : &s->tmap;
        a.n_rows = static_cast<uint32_t>(s->n_rows);
        a.row_base = static_cast<uint32_t>(s->row_base);
        a.pitch = half ? s->pitch16 : s->pitch;
        a.w_embed = w->embedding; a.w_lex = w->lexical;
        a.m = pool;
        a.d_lists = c->d_lists; a.d_counts = c->d_counts; a.d_ticket = c->d_ticket; a.d_pub = c->d_pub;
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t slot = static_cast<uint32_t>((seq0 + q) % cc->mb->ring);
            rlr::ScanGroupIO &io = a.groups.g[q];
            io.query = c->d_query + static_cast<size_t>(q) * rlr::kQueryCap;
            io.lex_rows = q == 0 ? c->d_lex_rows : (nl[q][g] ? cc->lex_extra[g][q - 1].d_rows : nullptr);
            io.lex_norm = q == 0 ? c->d_lex_norm : (nl[q][g] ? cc->lex_extra[g][q - 1].d_norm : nullptr);
            io.n_lex = nl[q][g];
            io.out = cc->mb->list(slot, g); io.out_n = cc->mb->count(slot, g);
            io.post.flag = cc->mb->flag(slot, g);
            io.post.consumed = cc->mb->consumed(slot);
            io.post.seq = seq0 + q; io.post.ring = cc->mb->ring; io.post.status = cc->mb->d_status;
        }
        CU_TRY(rlr::scan_launch(a, st));
        ++launches;
    }
    rlr_store *s0 = cl->shard[0];
    cudaStream_t st0 = r0->stream;
    const uint32_t cap = do_mmr ? std::min<uint32_t>(pool, std::max<uint32_t>(top_k, 1)) : pool;
    for (uint32_t q = 0; q < nq; ++q) {
        rlr_ctx *rq = q == 0 ? r0 : cc->root_extra[q - 1];
        const uint32_t slot = static_cast<uint32_t>((seq0 + q) % cc->mb->ring);
        CU_TRY(rlr::mailbox_merge_launch(cc->mb->list(slot, 0), cc->mb->m_cap, cc->mb->flag(slot, 0), seq0 + q, cc->mb->consumed(slot),
                                         cl->n, pool, rq->d_pool, rq->d_pool_n, cc->mb->d_status, st0));
        ++launches;
        const uint8_t *d_blk = rq->d_pool_blk;
        if (do_mmr) {
            const bool half = s0->use_half(flags);
            rlr::MmrArgs a;
            memset(&a, 0, sizeof a);
            a.half = half;
            a.pitch = half ? cl->pitch16 : cl->pitch; a.dim = cl->dim;
            a.d_cands = rq->d_pool; a.d_n = rq->d_pool_n;
            a.use_rows = 1; a.p_cap = pool; a.top_k = top_k; a.lambda = lambda;
            a.d_tri = rq->d_tri; a.d_sel_pos = rq->d_sel_pos; a.d_sel_n = rq->d_sel_n; a.d_result = rq->d_result;
            a.max_smem_optin = s0->smem_optin;
            a.peers = half ? &cl->table16 : &cl->table32;
            a.d_gather = rq->d_gather;
            uint32_t l = 0;
            CU_TRY(rlr::mmr_launch(a, st0, &l));
            launches += l;
            d_blk = rq->d_result_blk;
        }
        CU_TRY(cudaMemcpyAsync(rq->h_result_blk, d_blk, 16 + cap * sizeof(rlr_cand), cudaMemcpyDeviceToHost, st0));
    }
    CU_TRY(cudaStreamSynchronize(st0));
    cl->launches += launches;
    const uint32_t stride = std::max<uint32_t>(top_k, 1);
    for (uint32_t q = 0; q < nq; ++q) {
        rlr_ctx *rq = q == 0 ? r0 : cc->root_extra[q - 1];
        const uint32_t n = std::min(rq->h_result_n[0], cap);
        if (n == 0) {
            CU_TRY(cudaMemcpy(cc->h_status, cc->mb->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
            return fail(RLR_ERR_CUDA, "the cluster delivered no result for query %u (mailbox status %u)", q, cc->h_status[0]);
        }
        unpack(rq->h_result, n, out_rows + static_cast<size_t>(q) * stride, out_score ? out_score + static_cast<size_t>(q) * stride : nullptr,
               out_emb ? out_emb + static_cast<size_t>(q) * stride : nullptr, out_lex ? out_lex + static_cast<size_t>(q) * stride : nullptr);
        out_n[q] = n;
    }
    return RLR_OK;
}

} // namespace

// rlr_search_batch over the cluster (BASELINE configs[3] from ONE process): every GPU contracts the query batch
// against its shard (one host thread per GPU for the duration of the call: the per-shard call is host-synchronous),
// the root pulls the per-shard key lists over NVLink (cudaMemcpyPeerAsync) and merges them per query on the device.
RLR_EXPORT int rlr_cluster_search_batch(rlr_cluster *cl, const float *queries, uint32_t n_queries, uint32_t dim, uint32_t flags,
                                        uint32_t m, uint32_t *out_rows, float *out_scores, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (cl->n == 1) return rlr_search_batch(cl->shard[0], queries, n_queries, dim, flags, m, out_rows, out_scores, out_n);
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (n_queries == 0) return RLR_OK;
    if (!queries) return fail(RLR_ERR_INVALID_ARG, "queries is NULL");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    if (dim != cl->dim) return fail(RLR_ERR_DIM_MISMATCH, "queries have %u dims, store has %u", dim, cl->dim);
    if (static_cast<uint64_t>(cl->n) * m > 16384) return fail(RLR_ERR_UNSUPPORTED, "n_shards * m exceeds the merge capacity");
    ClusterLease lease(cl);
    if (int rc = lease.acquire()) return rc;
    ClusterCtx *cc = lease.cc;
    const size_t per = static_cast<size_t>(n_queries) * m;
    if (per > cc->batch_cap || cc->batch_keys.size() != cl->n) {
        for (size_t g = 0; g < cc->batch_keys.size(); ++g) if (cc->batch_keys[g]) { cudaSetDevice(cl->device[g]); cudaFree(cc->batch_keys[g]); }
        cc->batch_keys.assign(cl->n, nullptr);
        CU_TRY(cudaSetDevice(cl->device[0]));
        cudaFree(cc->batch_all); cudaFree(cc->batch_out); cudaFree(cc->batch_cnt);
        if (cc->h_batch) cudaFreeHost(cc->h_batch);
        cc->batch_all = cc->batch_out = cc->batch_cnt = nullptr; cc->h_batch = nullptr; cc->batch_cap = 0;
        for (uint32_t g = 0; g < cl->n; ++g) {
            CU_TRY(cudaSetDevice(cl->device[g]));
            CU_TRY(cudaMalloc(&cc->batch_keys[g], per * 8));
        }
        CU_TRY(cudaSetDevice(cl->device[0]));
        CU_TRY(cudaMalloc(&cc->batch_all, per * 8 * cl->n));
        CU_TRY(cudaMalloc(&cc->batch_out, per * 8));
        CU_TRY(cudaMalloc(&cc->batch_cnt, static_cast<size_t>(n_queries) * 4));
        CU_TRY(cudaMallocHost(&cc->h_batch, per * 8 + static_cast<size_t>(n_queries) * 4));
        cc->batch_cap = per;
    }
    // every shard in parallel: the contraction leaves rank-ordered keys (global rows) on the shard's GPU
    std::vector<int> rcs(cl->n, RLR_OK);
    std::vector<std::string> errs(cl->n);
    std::vector<std::thread> th;
    for (uint32_t g = 0; g < cl->n; ++g)
        th.emplace_back([&, g] {
            cudaSetDevice(cl->device[g]);
            rcs[g] = rlr_search_batch_device(cl->shard[g], queries, n_queries, dim, flags, m, cc->batch_keys[g], nullptr, cc->c[g]->stream);
            if (rcs[g] == RLR_OK && cudaStreamSynchronize(cc->c[g]->stream) != cudaSuccess) { cudaGetLastError(); rcs[g] = RLR_ERR_CUDA; }
            if (rcs[g] != RLR_OK) errs[g] = rlr_last_error();
        });
    for (auto &t : th) t.join();
    for (uint32_t g = 0; g < cl->n; ++g)
        if (rcs[g] != RLR_OK) return fail(rcs[g], "shard %u: %s", g, errs[g].c_str());
    CU_TRY(cudaSetDevice(cl->device[0]));
    cudaStream_t st0 = cc->c[0]->stream;
    for (uint32_t g = 0; g < cl->n; ++g)
        CU_TRY(cudaMemcpyPeerAsync(static_cast<uint8_t *>(cc->batch_all) + g * per * 8, cl->device[0], cc->batch_keys[g], cl->device[g], per * 8, st0));
    if (int rc = rlr_batch_merge_async(cl->shard[0], cc->batch_all, cl->n, n_queries, m, cc->batch_out, cc->batch_cnt, st0)) return rc;
    uint32_t *h_cnt = reinterpret_cast<uint32_t *>(cc->h_batch + per);
    CU_TRY(cudaMemcpyAsync(cc->h_batch, cc->batch_out, per * 8, cudaMemcpyDeviceToHost, st0));
    CU_TRY(cudaMemcpyAsync(h_cnt, cc->batch_cnt, static_cast<size_t>(n_queries) * 4, cudaMemcpyDeviceToHost, st0));
    CU_TRY(cudaStreamSynchronize(st0));
    for (uint32_t q = 0; q < n_queries; ++q) {
        const uint32_t n = std::min(h_cnt[q], m);
        out_n[q] = n;
        for (uint32_t i = 0; i < n; ++i) {
            const unsigned long long k = cc->h_batch[static_cast<size_t>(q) * m + i];
            out_rows[static_cast<size_t>(q) * m + i] = rlr::key_row(k);
            if (out_scores) {
                const uint32_t bits = rlr::bits_from_ord(static_cast<uint32_t>(k >> 32));
                memcpy(&out_scores[static_cast<size_t>(q) * m + i], &bits, 4);
            }
        }
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_search_mmr_multi(rlr_cluster *cl, const float *queries, uint32_t nq, uint32_t dim, uint32_t flags,
                                            uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                                            const uint32_t *const *lex_rows, const float *const *lex_scores, const uint32_t *n_lex,
                                            uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (!w) return fail(RLR_ERR_INVALID_ARG, "weights is NULL");
    if (nq == 0) return RLR_OK;
    if (nq > RLR_MAX_MULTI) return fail(RLR_ERR_UNSUPPORTED, "nq %u exceeds RLR_MAX_MULTI (%d)", nq, RLR_MAX_MULTI);
    if (!queries) return fail(RLR_ERR_INVALID_ARG, "queries is NULL");
    if (cl->n == 1)
        return rlr_search_mmr_multi(cl->shard[0], queries, nq, dim, flags, top_k, diversity_factor, w, lex_rows, lex_scores, n_lex,
                                    out_rows, out_score, out_emb, out_lex, out_n);
    if (nq == 1)
        return rlr_cluster_search_mmr(cl, queries, dim, flags, top_k, diversity_factor, w, lex_rows ? lex_rows[0] : nullptr,
                                      lex_scores ? lex_scores[0] : nullptr, n_lex ? n_lex[0] : 0, out_rows, out_score, out_emb, out_lex, out_n);
    for (uint32_t q = 0; q < nq; ++q) out_n[q] = 0;
    float lambda = diversity_factor;
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    const bool do_mmr = lambda != 0.0f;
    const uint64_t pool = do_mmr ? std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10) : std::max<uint32_t>(top_k, 1);
    if (pool > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "candidate pool %llu exceeds %d (top_k %u)", (unsigned long long)pool, RLR_MAX_M, top_k);
    if ((flags & RLR_SEARCH_F16) && !(cl->flags & (RLR_STORE_KEEP_F16 | RLR_STORE_F16_ONLY)))
        return fail(RLR_ERR_INVALID_ARG, "RLR_SEARCH_F16 but the store holds no f16 copy");
    return cluster_search_multi(cl, queries, nq, dim, flags, static_cast<uint32_t>(std::min<uint64_t>(pool, cl->n_rows)), do_mmr, top_k,
                                lambda, w, lex_rows, lex_scores, n_lex, out_rows, out_score, out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_cluster_create(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t n_rows, const float *rows,
                                  uint64_t host_pitch, uint32_t flags, const uint64_t *shard_rows, rlr_cluster **out)
{
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > RLR_MAX_SHARDS)
        return fail(RLR_ERR_INVALID_ARG, "n_devices %u not in 1..%d", n_devices, RLR_MAX_SHARDS);
    if (dim == 0 || dim > RLR_MAX_DIM) return fail(RLR_ERR_INVALID_ARG, "dim %u not in 1..%d", dim, RLR_MAX_DIM);
    if (n_rows >= (1ull << 32) - 1) return fail(RLR_ERR_UNSUPPORTED, "global rows must fit 32 bits");
    if (host_pitch == 0) host_pitch = dim;
    if (host_pitch < dim) return fail(RLR_ERR_INVALID_ARG, "host_pitch %llu < dim %u", (unsigned long long)host_pitch, dim);
    // every shard must own rows (a posting kernel needs at least one tile): tiny stores use fewer shards
    uint32_t n = n_devices;
    if (n_rows < n) n = static_cast<uint32_t>(std::max<uint64_t>(n_rows, 1));
    std::vector<uint64_t> plan;
    const uint32_t pitch = (dim + 31u) & ~31u, pitch16 = (dim + 63u) & ~63u;
    if (shard_rows && n == n_devices) {
        uint64_t sum = 0;
        for (uint32_t g = 0; g < n; ++g) {
            if (shard_rows[g] == 0 && n_rows != 0) return fail(RLR_ERR_INVALID_ARG, "shard_rows[%u] is 0: every shard must own rows", g);
            sum += shard_rows[g];
        }
        if (sum != n_rows) return fail(RLR_ERR_INVALID_ARG, "shard_rows sum to %llu, n_rows is %llu", (unsigned long long)sum, (unsigned long long)n_rows);
        plan.assign(shard_rows, shard_rows + n);
    } else {
        default_plan(n_rows, n, (flags & RLR_STORE_F16_ONLY) ? pitch16 * 2 : pitch * 4, &plan);
    }
    std::vector<int> dev(devices, devices + n);
    for (uint32_t g = 0; g < n; ++g)
        if (int rc = ensure_device(dev[g])) return rc;
    if (int rc = enable_peers(dev)) return rc;

    rlr_cluster *cl = new rlr_cluster();
    cl->n = n; cl->dim = dim; cl->pitch = pitch; cl->pitch16 = pitch16; cl->flags = flags; cl->n_rows = n_rows;
    cl->device = dev;
    cl->shard.assign(n, nullptr);
    cl->row_base.assign(n, 0);
    cl->shard_rows = plan;
    memset(&cl->table32, 0, sizeof cl->table32);
    memset(&cl->table16, 0, sizeof cl->table16);
    uint64_t base = 0;
    for (uint32_t g = 0; g < n; ++g) {
        cl->row_base[g] = base;
        int rc = rlr_store_create(dev[g], dim, plan[g], rows ? rows + base * host_pitch : nullptr, host_pitch, base, flags, &cl->shard[g]);
        if (rc != RLR_OK) { std::string keep = g_err; rlr_cluster_destroy(cl); g_err = keep; return rc; }
        base += plan[g];
    }
    for (uint32_t g = 0; g < n; ++g) {
        cl->table32.base[g] = cl->shard[g]->d_rows; cl->table16.base[g] = cl->shard[g]->d_rows16;
        cl->table32.row_base[g] = cl->table16.row_base[g] = static_cast<uint32_t>(cl->row_base[g]);
        cl->table32.n_rows[g] = cl->table16.n_rows[g] = static_cast<uint32_t>(plan[g]);
    }
    cl->table32.n = cl->table16.n = n;
    *out = cl;
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_destroy(rlr_cluster *cl)
{
    if (!cl) return RLR_OK;
    for (uint32_t g = 0; g < cl->n; ++g) {      // nothing of this cluster may still be running on any GPU
        cudaSetDevice(cl->device[g]);
        cudaDeviceSynchronize();
    }
    for (ClusterCtx *cc : cl->free_ctx) cctx_free(cl, cc);
    for (rlr_store *s : cl->shard) rlr_store_destroy(s);
    cudaGetLastError();
    delete cl;
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_info_get(const rlr_cluster *cl, rlr_cluster_info *out)
{
    if (int rc = check_cluster(cl)) return rc;
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    memset(out, 0, sizeof *out);
    out->n_rows = cl->n_rows; out->dim = cl->dim; out->pitch = cl->pitch; out->flags = cl->flags; out->n_shards = cl->n;
    for (uint32_t g = 0; g < cl->n; ++g) {
        out->device[g] = cl->device[g];
        out->row_base[g] = cl->row_base[g];
        out->shard_rows[g] = cl->shard_rows[g];
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_upload(rlr_cluster *cl, uint64_t row0, uint64_t n, const float *rows, uint64_t host_pitch)
{
    if (int rc = check_cluster(cl)) return rc;
    if (!rows && n) return fail(RLR_ERR_INVALID_ARG, "rows is NULL");
    if (row0 + n > cl->n_rows) return fail(RLR_ERR_INVALID_ARG, "rows [%llu,%llu) outside the store", (unsigned long long)row0, (unsigned long long)(row0 + n));
    if (host_pitch == 0) host_pitch = cl->dim;
    for (uint32_t g = 0; g < cl->n; ++g) {
        const uint64_t lo = std::max(row0, cl->row_base[g]), hi = std::min(row0 + n, cl->row_base[g] + cl->shard_rows[g]);
        if (lo >= hi) continue;
        if (int rc = rlr_store_upload(cl->shard[g], lo - cl->row_base[g], hi - lo, rows + (lo - row0) * host_pitch, host_pitch)) return rc;
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_read_rows(const rlr_cluster *cl, const uint32_t *rows, uint64_t n, float *out)
{
    if (int rc = check_cluster(cl)) return rc;
    if ((!rows || !out) && n) return fail(RLR_ERR_INVALID_ARG, "rows/out is NULL");
    for (uint64_t i = 0; i < n; ++i) {
        if (rows[i] >= cl->n_rows) return fail(RLR_ERR_INVALID_ARG, "row %u not in this store", rows[i]);
        const uint32_t g = owner_of(cl, rows[i]);
        if (int rc = rlr_store_read_rows(cl->shard[g], rows + i, 1, out + i * cl->dim)) return rc;
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_fill_synthetic(rlr_cluster *cl, int kind, uint64_t seed, uint64_t centroid_seed,
                                          uint32_t n_clusters, float sigma)
{
    if (int rc = check_cluster(cl)) return rc;
    for (uint32_t g = 0; g < cl->n; ++g)
        if (int rc = rlr_store_fill_synthetic(cl->shard[g], kind, seed, centroid_seed, n_clusters, sigma)) return rc;
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_search_topm(rlr_cluster *cl, const float *query, uint32_t dim, uint32_t flags,
                                       const rlr_resolved_weights *w, const uint32_t *lex_rows, const float *lex_scores,
                                       uint32_t n_lex, uint32_t m, uint32_t *out_rows, float *out_combined, float *out_emb,
                                       float *out_lex, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (cl->n == 1)
        return rlr_search_topm(cl->shard[0], query, dim, flags, w, lex_rows, lex_scores, n_lex, m, out_rows, out_combined,
                               out_emb, out_lex, out_n);
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (!w) return fail(RLR_ERR_INVALID_ARG, "weights is NULL");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    *out_n = 0;
    if ((flags & RLR_SEARCH_F16) && !(cl->flags & (RLR_STORE_KEEP_F16 | RLR_STORE_F16_ONLY)))
        return fail(RLR_ERR_INVALID_ARG, "RLR_SEARCH_F16 but the store holds no f16 copy");
    Plan pl;
    pl.m = static_cast<uint32_t>(std::min<uint64_t>(m, cl->n_rows));
    return cluster_search(cl, query, dim, flags, pl, w, lex_rows, lex_scores, n_lex, out_rows, out_combined, out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_cluster_embedding_candidates(rlr_cluster *cl, const float *query, uint32_t dim, uint32_t flags,
                                                uint32_t count, uint32_t *out_rows, float *out_score, uint32_t *out_n)
{
    // :438-447 raw dot, sort desc, take(count): w_embed = 1 makes combined == emb exactly
    if (!out_n) return fail(RLR_ERR_INVALID_ARG, "out_n is NULL");
    *out_n = 0;
    if (count == 0) return RLR_OK;
    const rlr_resolved_weights w = {1.0f, 0.0f, 0.0f, 0.0f};
    return rlr_cluster_search_topm(cl, query, dim, flags, &w, nullptr, nullptr, 0, count, out_rows, nullptr, out_score, nullptr, out_n);
}

RLR_EXPORT int rlr_cluster_search_mmr(rlr_cluster *cl, const float *query, uint32_t dim, uint32_t flags, uint32_t top_k,
                                      float diversity_factor, const rlr_resolved_weights *w, const uint32_t *lex_rows,
                                      const float *lex_scores, uint32_t n_lex, uint32_t *out_rows, float *out_score,
                                      float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (cl->n == 1)
        return rlr_search_mmr(cl->shard[0], query, dim, flags, top_k, diversity_factor, w, lex_rows, lex_scores, n_lex, out_rows,
                              out_score, out_emb, out_lex, out_n);
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (!w) return fail(RLR_ERR_INVALID_ARG, "weights is NULL");
    float lambda = diversity_factor;                                    // :725 f32::clamp (NaN stays NaN)
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    if (lambda == 0.0f)                                                 // :728-730 -> search(top_k), top_k.max(1) at :490
        return rlr_cluster_search_topm(cl, query, dim, flags, w, lex_rows, lex_scores, n_lex, std::max<uint32_t>(top_k, 1),
                                       out_rows, out_score, out_emb, out_lex, out_n);
    *out_n = 0;
    const uint64_t pool = std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10);   // :734
    if (pool > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "candidate pool %llu exceeds %d (top_k %u)", (unsigned long long)pool, RLR_MAX_M, top_k);
    if ((flags & RLR_SEARCH_F16) && !(cl->flags & (RLR_STORE_KEEP_F16 | RLR_STORE_F16_ONLY)))
        return fail(RLR_ERR_INVALID_ARG, "RLR_SEARCH_F16 but the store holds no f16 copy");
    if (cl->n_rows == 0) return RLR_OK;
    Plan pl;
    pl.m = static_cast<uint32_t>(std::min<uint64_t>(pool, cl->n_rows));
    pl.do_mmr = true; pl.top_k = top_k; pl.lambda = lambda;
    return cluster_search(cl, query, dim, flags, pl, w, lex_rows, lex_scores, n_lex, out_rows, out_score, out_emb, out_lex, out_n);
}

// ---- text queries over the cluster: BM25 scored on every shard's GPU (global statistics), ranked lists merged on the
// host, then the very same search as with caller-supplied pairs ----
uint32_t rlr_api_cluster_n(const rlr_cluster *cl) { return cl->n; }
rlr_store *rlr_api_cluster_shard(const rlr_cluster *cl, uint32_t i) { return cl->shard[i]; }

RLR_EXPORT int rlr_cluster_search_text_topm(rlr_cluster *cl, rlr_cluster_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                                            const rlr_resolved_weights *w, const uint32_t *query_terms, uint32_t n_terms, uint32_t m,
                                            uint32_t *out_rows, float *out_combined, float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (!ix || rlr_api_cluster_bm25_cluster(ix) != cl) return fail(RLR_ERR_INVALID_ARG, "the BM25 index is NULL or belongs to another cluster");
    if (cl->n == 1)        // one shard: the single-GPU path keeps everything on the device
        return rlr_search_text_topm(cl->shard[0], rlr_api_cluster_bm25_part(ix, 0), query, dim, flags, w, query_terms, n_terms, m,
                                    out_rows, out_combined, out_emb, out_lex, out_n);
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    std::vector<uint32_t> lr;
    std::vector<float> ls;
    if (int rc = rlr_api_cluster_bm25_score(ix, query_terms, n_terms, 5u * m, lr, ls)) return rc;      // lexical_index.score(query, top_k * 5), :505
    return rlr_cluster_search_topm(cl, query, dim, flags, w, lr.data(), ls.data(), static_cast<uint32_t>(lr.size()), m, out_rows,
                                   out_combined, out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_cluster_search_text_mmr(rlr_cluster *cl, rlr_cluster_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                                           uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                                           const uint32_t *query_terms, uint32_t n_terms, uint32_t *out_rows, float *out_score,
                                           float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (!ix || rlr_api_cluster_bm25_cluster(ix) != cl) return fail(RLR_ERR_INVALID_ARG, "the BM25 index is NULL or belongs to another cluster");
    if (cl->n == 1)
        return rlr_search_text_mmr(cl->shard[0], rlr_api_cluster_bm25_part(ix, 0), query, dim, flags, top_k, diversity_factor, w,
                                   query_terms, n_terms, out_rows, out_score, out_emb, out_lex, out_n);
    float lambda = diversity_factor;                                    // :725 f32::clamp (NaN stays NaN)
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    // search(m): m = top_k.max(1) without diversity (:728-730, :490), the candidate pool otherwise (:734)
    const uint64_t m = lambda == 0.0f ? std::max<uint64_t>(top_k, 1) : std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10);
    if (m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "candidate pool %llu exceeds %d (top_k %u)", (unsigned long long)m, RLR_MAX_M, top_k);
    std::vector<uint32_t> lr;
    std::vector<float> ls;
    if (int rc = rlr_api_cluster_bm25_score(ix, query_terms, n_terms, static_cast<uint32_t>(5u * m), lr, ls)) return rc;
    return rlr_cluster_search_mmr(cl, query, dim, flags, top_k, diversity_factor, w, lr.data(), ls.data(), static_cast<uint32_t>(lr.size()),
                                  out_rows, out_score, out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_cluster_mmr(rlr_cluster *cl, const uint32_t *cand_rows, const float *relevance, uint32_t p, uint32_t top_k,
                               float lambda, uint32_t flags, uint32_t *out_sel_pos, uint32_t *out_n)
{
    if (int rc = check_cluster(cl)) return rc;
    if (cl->n == 1) return rlr_mmr(cl->shard[0], cand_rows, relevance, p, top_k, lambda, flags, out_sel_pos, out_n);
    if (!out_sel_pos || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_sel_pos/out_n is NULL");
    *out_n = 0;
    if (p == 0) return RLR_OK; // :773-775
    if (!cand_rows || !relevance) return fail(RLR_ERR_INVALID_ARG, "cand_rows/relevance is NULL");
    if (p > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "p %u exceeds %d", p, RLR_MAX_M);
    for (uint32_t i = 0; i < p; ++i)
        if (cand_rows[i] >= cl->n_rows) return fail(RLR_ERR_INVALID_ARG, "cand_rows[%u]=%u not in this store", i, cand_rows[i]);
    ClusterLease lease(cl);
    if (int rc = lease.acquire()) return rc;
    rlr_ctx *c = lease.cc->c[0];
    rlr_store *s0 = cl->shard[0];
    CU_TRY(cudaSetDevice(s0->device));
    cudaStream_t st = c->stream;
    memcpy(c->h_u32, cand_rows, p * sizeof(uint32_t));
    c->h_u32[RLR_MAX_M] = p;
    memcpy(c->h_rel, relevance, p * sizeof(float));
    CU_TRY(cudaMemcpyAsync(c->d_rows_in, c->h_u32, p * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c->d_p_in, c->h_u32 + RLR_MAX_M, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c->d_rel_in, c->h_rel, p * sizeof(float), cudaMemcpyHostToDevice, st));
    const bool timed = flags & RLR_WANT_TIMINGS;
    if (timed) CU_TRY(cudaEventRecord(c->ev[2], st));
    const bool half = s0->use_half(flags);
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.half = half;
    a.d_emb = nullptr; a.pitch = half ? cl->pitch16 : cl->pitch; a.dim = cl->dim;
    a.d_cands = nullptr; a.d_n = c->d_p_in; a.d_rows = c->d_rows_in; a.d_rel = c->d_rel_in;
    a.use_rows = 1; a.p_cap = p; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = c->d_sel_pos; a.d_sel_n = c->d_sel_n; a.d_result = nullptr;
    a.max_smem_optin = s0->smem_optin;
    a.peers = half ? &cl->table16 : &cl->table32;
    a.d_gather = c->d_gather;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, st, &l));
    cl->launches += l;
    if (timed) CU_TRY(cudaEventRecord(c->ev[3], st));
    const uint32_t cap = std::min<uint32_t>(p, std::max<uint32_t>(top_k, 1));
    CU_TRY(cudaMemcpyAsync(c->h_u32, c->d_sel_pos, cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(c->h_u32 + RLR_MAX_M, c->d_sel_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const uint32_t n = std::min(c->h_u32[RLR_MAX_M], cap);
    memcpy(out_sel_pos, c->h_u32, n * sizeof(uint32_t));
    *out_n = n;
    if (timed) {
        rlr_timings t = {0, 0, 0, 0, 0};
        cudaEventElapsedTime(&t.mmr_ms, c->ev[2], c->ev[3]);
        t.total_ms = t.mmr_ms;
        t.launches = l;
        g_timings = t;
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_last_scan_ms(float *out_ms, uint32_t cap, uint32_t *out_n)
{
    if (!out_n) return fail(RLR_ERR_INVALID_ARG, "out_n is NULL");
    const uint32_t n = static_cast<uint32_t>(std::min<size_t>(cap, g_shard_scan_ms.size()));
    if (n && !out_ms) return fail(RLR_ERR_INVALID_ARG, "out_ms is NULL");
    for (uint32_t i = 0; i < n; ++i) out_ms[i] = g_shard_scan_ms[i];
    *out_n = static_cast<uint32_t>(g_shard_scan_ms.size());
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_launch_count(const rlr_cluster *cl, uint64_t *out)
{
    if (int rc = check_cluster(cl)) return rc;
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    uint64_t total = cl->launches.load();
    for (rlr_store *s : cl->shard)           // single-shard clusters forward to the store's pooled ctxs
        for (rlr_ctx *c : s->free_ctx) total += c->launches;
    *out = total;
    return RLR_OK;
}
