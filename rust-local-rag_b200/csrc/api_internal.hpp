// api_internal.hpp -- definitions shared by the translation units that implement the C ABI
// (api.cu: single-GPU stores; cluster.cu: one process driving the GPUs of a box).  Internal.
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels.cuh"

#define RLR_EXPORT extern "C" __attribute__((visibility("default")))

struct rlr_store {
    int device = 0;
    uint32_t dim = 0, pitch = 0, flags = 0;
    uint64_t n_rows = 0, row_base = 0;
    uint64_t capacity = 0;          // rows the device allocations can hold (>= n_rows)
    float *d_rows = nullptr;        // f32 matrix (absent for RLR_STORE_F16_ONLY)
    void *d_rows16 = nullptr;       // binary16 copy (RLR_STORE_KEEP_F16 / RLR_STORE_F16_ONLY)
    uint32_t pitch16 = 0;           // elements per binary16 row (dim rounded up to 64)
    void *d_rows_bf16 = nullptr;    // bfloat16 copy (RLR_STORE_KEEP_BF16): an operand of the batched contraction only
    CUtensorMap tmap, tmap16, tmap_bf16;
    CUtensorMap tmap_small;         // f32 rows with `rpt`-row boxes: small stores spread their rows over every SM
    uint32_t rpt = rlr::kScanRows;  // rows per scan tile on the latency path (scan_rows_per_tile)
    int sm_count = 0, smem_optin = 0;
    bool use_half(uint32_t flags) const { return d_rows == nullptr || ((flags & RLR_SEARCH_F16) && d_rows16 != nullptr); }
    std::mutex mu;
    std::vector<rlr_ctx *> free_ctx;
};

struct rlr_ctx {
    rlr_store *s = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float *d_query = nullptr;
    uint32_t *d_lex_rows = nullptr;
    float *d_lex_norm = nullptr;
    rlr_cand *d_lists = nullptr;
    uint32_t *d_counts = nullptr;
    uint32_t *d_ticket = nullptr;
    uint32_t *d_pub = nullptr;
    rlr_cand *d_tmp = nullptr;
    uint8_t *d_pool_blk = nullptr;  // [u32 n | pad to 16 B | RLR_MAX_M records]: d_pool_n / d_pool point into it
    rlr_cand *d_pool = nullptr;
    uint32_t *d_pool_n = nullptr;
    float *d_tri = nullptr;
    void *d_gather = nullptr;       // RLR_MAX_M x pitch f32 of scratch: pool rows gathered from peer GPUs for MMR
    uint32_t *d_sel_pos = nullptr;
    uint8_t *d_result_blk = nullptr; // same layout: d_sel_n / d_result point into it
    uint32_t *d_sel_n = nullptr;
    rlr_cand *d_result = nullptr;
    uint32_t *d_rows_in = nullptr;
    float *d_rel_in = nullptr;
    uint32_t *d_p_in = nullptr;
    // pinned host staging
    float *h_query = nullptr;
    uint32_t *h_lex_rows = nullptr;
    float *h_lex_norm = nullptr;
    uint8_t *h_result_blk = nullptr; // pinned mirror of a (count, records) block
    uint32_t *h_result_n = nullptr;
    rlr_cand *h_result = nullptr;   // RLR_MAX_M records
    uint32_t *h_u32 = nullptr;      // RLR_MAX_M + 8 words
    float *h_rel = nullptr;
    rlr::LatParams lat;             // latency path: the kernel's parameter block (query + lexical pairs + delivery)
    unsigned long long lat_seq = 0; // completion sequence number of this ctx's mapped result block
    uint64_t launches = 0;
    uint32_t n_lists_cap = 0;
    uint32_t search_flags = 0;      // RLR_SEARCH_F16 for the device-level entry points
    // batched path workspace (allocated on first use, grown on demand)
    void *batch_mem = nullptr;
    size_t batch_bytes = 0;
    float *h_batch_q = nullptr;     // pinned staging for the query batch
    size_t h_batch_q_bytes = 0;
    unsigned long long *h_batch_state = nullptr;
    size_t h_batch_state_bytes = 0;
};

struct rlr_mailbox {
    int device = 0;
    bool owner = false;
    uint32_t n_ranks = 0, m_cap = 0, ring = 0;
    uint8_t *base = nullptr;          // root's allocation (local on the root, an IPC mapping elsewhere)
    uint32_t *d_status = nullptr;     // local
    size_t bytes = 0;
    // layout: [0] consumed[ring] u64 (one word per slot: the last sequence number merged out of it)
    //         | [1024] flags[ring][n_ranks] u64 | counts[ring][n_ranks] u32 | (4 KB aligned) lists
    size_t flags_off() const { return 1024; }
    size_t counts_off() const { return flags_off() + static_cast<size_t>(ring) * n_ranks * 8; }
    size_t lists_off() const { return (counts_off() + static_cast<size_t>(ring) * n_ranks * 4 + 4095) & ~static_cast<size_t>(4095); }
    size_t total() const { return lists_off() + static_cast<size_t>(ring) * n_ranks * m_cap * sizeof(rlr_cand); }
    unsigned long long *consumed(uint32_t slot) const { return reinterpret_cast<unsigned long long *>(base) + slot; }
    unsigned long long *flag(uint32_t slot, uint32_t r) const { return reinterpret_cast<unsigned long long *>(base + flags_off()) + static_cast<size_t>(slot) * n_ranks + r; }
    uint32_t *count(uint32_t slot, uint32_t r) const { return reinterpret_cast<uint32_t *>(base + counts_off()) + static_cast<size_t>(slot) * n_ranks + r; }
    rlr_cand *list(uint32_t slot, uint32_t r) const { return reinterpret_cast<rlr_cand *>(base + lists_off()) + (static_cast<size_t>(slot) * n_ranks + r) * m_cap; }
};


namespace rlr_api {

constexpr uint32_t kLexCap = 8192;

extern thread_local std::string g_err;
extern thread_local rlr_timings g_timings;

int fail(int code, const char *fmt, ...);

#define CU_TRY(expr)                                                                             \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            cudaGetLastError();                                                                  \
            return ::rlr_api::fail(e__ == cudaErrorMemoryAllocation ? RLR_ERR_OOM : RLR_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                        \
    } while (0)

int ensure_device(int device);          // exists, is sm_100, kernels configured; leaves it current
int device_sm_count(int device);
int ctx_new(rlr_store *s, rlr_ctx **out);
void ctx_free(rlr_ctx *c);
void host_normalize(float *v, size_t n);   // src/rag_engine.rs:1763-1771
void unpack(const rlr_cand *h, uint32_t n, uint32_t *rows, float *score, float *emb, float *lex);

// RAII lease of a pooled ctx so that concurrent searches on one store are re-entrant.
struct CtxLease {
    rlr_store *s;
    rlr_ctx *c = nullptr;
    explicit CtxLease(rlr_store *st) : s(st) {}
    int acquire()
    {
        {
            std::lock_guard<std::mutex> lk(s->mu);
            if (!s->free_ctx.empty()) { c = s->free_ctx.back(); s->free_ctx.pop_back(); }
        }
        if (c) return RLR_OK;
        return ctx_new(s, &c);
    }
    ~CtxLease()
    {
        if (c) { std::lock_guard<std::mutex> lk(s->mu); s->free_ctx.push_back(c); }
    }
};

} // namespace rlr_api

// bm25.cu: the device BM25 index, used by the text-query entry points of api.cu
struct rlr_bm25;
int rlr_api_bm25_ws_acquire(rlr_bm25 *ix, void **out);
void rlr_api_bm25_ws_release(rlr_bm25 *ix, void *ws);
// corpus statistics of a sharded index: every shard scores with the GLOBAL N, avgdl and df (df[j]: query_terms[j])
struct rlr_api_bm25_global { uint64_t total_docs, total_length; const uint32_t *df; };
int rlr_api_bm25_enqueue(rlr_bm25 *ix, void *ws, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                         uint32_t *d_lex_rows, float *d_lex_norm, uint32_t lex_pad, uint32_t *d_desc_rows, float *d_desc_scores,
                         uint32_t *d_n, cudaStream_t st, bool *active, const rlr_api_bm25_global *gs = nullptr,
                         uint32_t *launches_out = nullptr);
rlr_store *rlr_api_bm25_store(rlr_bm25 *ix);

// cluster.cu <-> bm25.cu: a BM25 index over a sharded store (rlr_cluster_bm25_*)
struct rlr_cluster;
struct rlr_cluster_bm25;
uint32_t rlr_api_cluster_n(const rlr_cluster *cl);
rlr_store *rlr_api_cluster_shard(const rlr_cluster *cl, uint32_t i);
rlr_cluster *rlr_api_cluster_bm25_cluster(rlr_cluster_bm25 *ix);
rlr_bm25 *rlr_api_cluster_bm25_part(rlr_cluster_bm25 *ix, uint32_t i);
// LexicalIndex::score(query, limit) over every shard (device scoring with global statistics, host merge of the
// per-shard ranked lists): global rows + raw scores, score desc, ties to the lower row
int rlr_api_cluster_bm25_score(rlr_cluster_bm25 *ix, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                               std::vector<uint32_t> &rows, std::vector<float> &scores);
