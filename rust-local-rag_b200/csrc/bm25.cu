// bm25.cu -- SURVEY.md 8(f) N4: the reference's LexicalIndex::score (/root/reference/src/rag_engine.rs:2169-2227)
// on the device, feeding the blend of RagEngine::search (:505-532).
//
// Why: tools/bm25_cost.py measures the host-side BM25 pass (hash-map postings, like the reference's) at 18 ms per
// query for 10k chunks and 206 ms for 100k -- against 0.05 ms for the whole GPU search at 10k chunks.  For TEXT
// queries it is the dominant cost of the hot path at the reference's operating point.
//
// What stays on the host: strings.  The caller tokenizes (`tokenize`, :2242-2247) and keeps the term -> id
// dictionary; this index sees term ids, term frequencies and rows only.
//
// Arithmetic: bit-identical to the reference's f32 formula given its summation order.  The reference sums a
// document's per-term contributions in HashSet iteration order (random per process, :2194); the deterministic choice
// here -- as in the host-mirror twin and oracle/lexical.py -- is the order in which the caller lists the query terms
// (bytewise order of the term strings).  To keep that order WITHOUT atomics the kernel is document-parallel: one
// thread per row walks the query terms in order and looks each up in the row's sorted (term, tf) list (a forward
// index in CSR form), so `scores[doc] += contribution` happens in a fixed order per document.  idf needs ln(): it is
// computed on the host with the C library's logf (what Rust's f32::ln calls), one value per query term.
// Stores beyond kBm25InvertedMinRows keep the postings BY TERM instead and run one launch per query term, in the
// caller's order (bm25_term_pass_kernel): only the query's postings are read, and the per-document order is the same.
//
// Then `results.sort_by(score desc); truncate(limit)` (:2222-2226) without sorting n keys: an exact radix select of
// the limit-th largest rank key (score bits | ~row: ties go to the lower row, one of the reference's valid orders)
// in six 11-bit passes over the dense key array (shared-memory histograms, the last CTA of a pass picks the digit),
// a compaction of the <= limit survivors, and one small CTA that sorts them, divides by max(max, EPSILON) (:511-530)
// and emits them sorted by row -- the form the scan kernel's lexical lookup wants.  No host round trip anywhere:
// the scan is enqueued right behind on the same stream with n_lex = the padded list length.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "api_internal.hpp"

using namespace rlr_api;

namespace rlr {

constexpr uint32_t kBm25MaxTerms = 64;       // unique query terms per request
constexpr uint32_t kBm25MaxLimit = 8192;     // 5 * m: 1500 for top_k = 100 with MMR, 4500 for the reranker flow (m = 900)
constexpr uint32_t kDigitBits = 11;
constexpr uint32_t kBins = 1u << kDigitBits;
constexpr uint32_t kPasses = 6;              // 6 x 11 bits >= 64
constexpr uint64_t kBm25InvertedMinRows = 262144;   // beyond the latency path's stores the device index is inverted (see below)

struct Bm25Query {
    uint32_t n_terms;
    uint32_t term[kBm25MaxTerms];
    float idf[kBm25MaxTerms];
    float avg_doc_len;
};

// selection state, one per workspace (device memory)
struct Bm25Sel {
    unsigned long long prefix;     // digits decided so far, in place (low bits zero)
    uint32_t remaining;            // rank still to find inside the current prefix
    uint32_t n_hits;               // keys != 0
    uint32_t ticket;
    uint32_t n_out;                // survivors written by the compaction
    unsigned long long kstar;      // the limit-th largest key (1 when there are fewer hits than limit)
};

namespace {

// one thread per row: scores[doc] = sum over the query terms, in the caller's order (:2193-2219)
__global__ void bm25_score_kernel(const unsigned long long *__restrict__ off, const uint32_t *__restrict__ terms,
                                  const uint32_t *__restrict__ tfs, const uint32_t *__restrict__ doc_len, uint32_t n_rows,
                                  uint32_t row_base, const __grid_constant__ Bm25Query q, unsigned long long *__restrict__ keys)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const unsigned long long lo0 = off[r], hi0 = off[r + 1];
    unsigned long long key = 0ull;
    if (hi0 > lo0) {
        const float dl = static_cast<float>(doc_len[r]);
        if (dl != 0.0f) {                                               // `if doc_length == 0.0 { continue; }`
            const float k1 = 1.5f, b = 0.75f;
            // k1 * (1.0 - b + b * (doc_length / avg_doc_len)): the same for every term of this document
            const float x = mul_rn(k1, add_rn(sub_rn(1.0f, b), mul_rn(b, __fdiv_rn(dl, q.avg_doc_len))));
            float acc = 0.0f;
            bool hit = false;
            for (uint32_t j = 0; j < q.n_terms; ++j) {
                const uint32_t t = q.term[j];
                unsigned long long lo = lo0, hi = hi0;
                while (lo < hi) {
                    const unsigned long long mid = (lo + hi) >> 1;
                    if (terms[mid] < t) lo = mid + 1; else hi = mid;
                }
                if (lo < hi0 && terms[lo] == t) {
                    const float tf = static_cast<float>(tfs[lo]);
                    const float denom = add_rn(tf, x);
                    if (denom == 0.0f) continue;                          // `if denom == 0.0 { continue; }`
                    const float sc = __fdiv_rn(mul_rn(q.idf[j], mul_rn(tf, add_rn(k1, 1.0f))), denom);
                    acc = add_rn(acc, sc);                                // *entry.or_insert(0.0) += score
                    hit = true;
                }
            }
            if (hit) key = make_key(acc, row_base + r);
        }
    }
    keys[r] = key;
}

// one radix pass: histogram of the current digit among the keys that match the decided prefix; the last CTA picks
// the digit that holds the wanted rank
__global__ void __launch_bounds__(256)
bm25_select_pass_kernel(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t pass, uint32_t limit,
                        uint32_t *__restrict__ hist /* [kBins], zero on entry, zero on exit */, Bm25Sel *__restrict__ sel)
{
    __shared__ uint32_t h[kBins];
    __shared__ uint32_t s_last;
    const uint32_t hi_bit = 64u - pass * kDigitBits;                  // bits [shift, hi_bit) are this pass's digit
    const uint32_t shift = hi_bit > kDigitBits ? hi_bit - kDigitBits : 0u;
    const uint32_t width = hi_bit - shift;
    if (pass != 0 && sel->kstar != 0ull) return;                      // decided by an earlier pass (written by an earlier launch)
    const unsigned long long prefix = sel->prefix;
    const unsigned long long hi_mask = pass == 0 ? 0ull : ~0ull << hi_bit;
    for (uint32_t i = threadIdx.x; i < kBins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t my_hits = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k == 0ull) continue;
        ++my_hits;
        if ((k & hi_mask) != prefix) continue;
        atomicAdd(&h[static_cast<uint32_t>(k >> shift) & ((1u << width) - 1u)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kBins; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
    if (pass == 0 && my_hits) atomicAdd(&sel->n_hits, my_hits);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&sel->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: find, from the top, the bin that holds the wanted rank.  Parallel: thread t owns the 8 bins
    // [(255 - t) * 8, (255 - t) * 8 + 8) (thread 0 = the top bins), a block-wide scan of the per-thread sums tells each
    // thread how many keys sit above its bins, and exactly one thread finds the rank inside its own eight.
    for (uint32_t i = threadIdx.x; i < kBins; i += blockDim.x) h[i] = __ldcg(hist + i);
    __syncthreads();
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_want, s_skip;
    if (threadIdx.x == 0) {
        const uint32_t n_hits = *reinterpret_cast<volatile uint32_t *>(&sel->n_hits);
        s_skip = 0;
        if (pass == 0 && n_hits <= limit) {
            sel->kstar = 1ull;                                        // fewer hits than the limit: everything survives
            sel->remaining = 0;
            s_skip = 1;
        } else if (sel->kstar == 1ull) {
            s_skip = 1;
        }
        s_want = pass == 0 ? limit : sel->remaining;
        sel->ticket = 0;
    }
    __syncthreads();
    if (!s_skip) {
        const uint32_t t = threadIdx.x, lane = t & 31u, wp = t >> 5;
        const uint32_t base = (255u - t) * 8u;
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = h[base + 7 - i]; sum += c[i]; }      // c[0] = the highest of my bins
        uint32_t incl = sum;                                                      // inclusive scan over threads 0..t
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= static_cast<uint32_t>(o)) incl += v; }
        if (lane == 31) s_warp[wp] = incl;
        __syncthreads();
        uint32_t above = incl - sum;
        for (uint32_t w2 = 0; w2 < wp; ++w2) above += s_warp[w2];
        const uint32_t want = s_want;
        if (above < want && want <= above + sum) {                                // the rank falls inside my eight bins
            uint32_t w3 = want - above, d = base + 7, cd = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (w3 <= c[i]) { d = base + 7 - i; cd = c[i]; break; }
                w3 -= c[i];
            }
            const unsigned long long np = prefix | (static_cast<unsigned long long>(d) << shift);
            sel->prefix = np;
            sel->remaining = w3;
            // every digit decided -- or EVERY key of the chosen bin is wanted, so the cut falls right below the bin and the
            // smallest key with this prefix is a valid threshold: the remaining passes have nothing to do (the usual case
            // once the 32 score bits are decided and the scores at the cut are distinct)
            if (shift == 0 || cd == w3) sel->kstar = np;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0;      // ready for the next pass
}

__global__ void bm25_compact_kernel(const unsigned long long *__restrict__ keys, uint32_t n, Bm25Sel *__restrict__ sel,
                                    unsigned long long *__restrict__ out, uint32_t cap)
{
    const unsigned long long kstar = sel->kstar;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k != 0ull && k >= kstar) {
            const uint32_t slot = atomicAdd(&sel->n_out, 1u);
            if (slot < cap) out[slot] = k;
        }
    }
}

// Large stores: the same scores from an INVERTED index (postings by term, rows ascending), touching only the
// postings of the query's terms instead of every (row, term) pair of the corpus.  `scores[doc] += contribution` must
// happen in the caller's term order for every document, so the terms are processed by consecutive launches on one
// stream (a term lists a row once: no two threads of a launch touch the same score, no atomics), each posting doing
// exactly the arithmetic of bm25_score_kernel.  An untouched score keeps its 0xffffffff fill: the first
// contribution is added to 0.0 (`entry.or_insert(0.0)`), and "never touched" stays distinguishable from a sum of 0.0.
__global__ void bm25_term_pass_kernel(const uint32_t *__restrict__ pdocs, const uint32_t *__restrict__ ptfs, unsigned long long p_lo,
                                      unsigned long long p_hi, const uint32_t *__restrict__ doc_len, float idf, float avg_doc_len,
                                      float *__restrict__ scores)
{
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long p = p_lo + static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < p_hi; p += stride) {
        const uint32_t r = pdocs[p];
        const float dl = static_cast<float>(doc_len[r]);
        if (dl == 0.0f) continue;                                       // `if doc_length == 0.0 { continue; }`
        const float k1 = 1.5f, b = 0.75f;
        const float x = mul_rn(k1, add_rn(sub_rn(1.0f, b), mul_rn(b, __fdiv_rn(dl, avg_doc_len))));
        const float tf = static_cast<float>(ptfs[p]);
        const float denom = add_rn(tf, x);
        if (denom == 0.0f) continue;                                    // `if denom == 0.0 { continue; }`
        const float sc = __fdiv_rn(mul_rn(idf, mul_rn(tf, add_rn(k1, 1.0f))), denom);
        const float prev = scores[r];
        scores[r] = add_rn(__float_as_uint(prev) == 0xffffffffu ? 0.0f : prev, sc);   // *entry.or_insert(0.0) += score
    }
}

__global__ void bm25_keys_kernel(const float *__restrict__ scores, uint32_t n_rows, uint32_t row_base, unsigned long long *__restrict__ keys)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const float sc = scores[r];
    keys[r] = __float_as_uint(sc) == 0xffffffffu ? 0ull : make_key(sc, row_base + r);
}

__device__ void bitonic_smem_desc(unsigned long long *k, uint32_t n2)
{
    for (uint32_t size = 2; size <= n2; size <<= 1)
        for (uint32_t j = size >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < (n2 >> 1); i += blockDim.x) {
                const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const unsigned long long a = k[lo], b = k[hi];
                const bool desc = (lo & size) == 0;
                if ((a < b) == desc) { k[lo] = b; k[hi] = a; }
            }
            __syncthreads();
        }
}

// one CTA: sort the survivors by rank, emit (row, raw score) in that order (LexicalIndex::score's output), then
// (row ascending, score / max(max_score, EPSILON)) padded with sentinel rows for the scan kernel's lookup (:511-530)
__global__ void __launch_bounds__(1024)
bm25_finalize_kernel(const unsigned long long *__restrict__ sel_keys, Bm25Sel *__restrict__ sel, uint32_t limit, uint32_t row_base,
                     uint32_t *__restrict__ desc_rows, float *__restrict__ desc_scores, uint32_t *__restrict__ out_n,
                     uint32_t *__restrict__ lex_rows, float *__restrict__ lex_norm, uint32_t lex_pad)
{
    extern __shared__ unsigned long long k[];           // next_pow2(limit) keys
    __shared__ float s_max;
    uint32_t n = sel->n_out;
    if (n > limit) n = limit;
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    if (n2 < 2) n2 = 2;
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) k[i] = i < n ? sel_keys[i] : 0ull;
    __syncthreads();
    bitonic_smem_desc(k, n2);
    if (threadIdx.x == 0) {
        // lexical_scores.values().fold(0.0, f32::max).max(f32::EPSILON)
        float mx = n ? fmaxf(0.0f, key_score(k[0])) : 0.0f;
        s_max = fmaxf(mx, 1.1920929e-07f);
        if (out_n) *out_n = n;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (desc_rows) desc_rows[i] = key_row(k[i]);
        if (desc_scores) desc_scores[i] = key_score(k[i]);
    }
    __syncthreads();
    if (lex_rows == nullptr) return;
    // re-key by LOCAL row (descending sort of ~row == ascending rows), keep the score bits in the low half
    const float mx = s_max;
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {          // every thread rewrites only its own slots
        const unsigned long long kk = k[i];
        const uint32_t row = key_row(kk) - row_base;
        k[i] = (i < n && kk != 0ull) ? ((static_cast<unsigned long long>(~row) << 32) | __float_as_uint(__fdiv_rn(key_score(kk), mx))) : 0ull;
    }
    __syncthreads();
    bitonic_smem_desc(k, n2);
    for (uint32_t i = threadIdx.x; i < lex_pad; i += blockDim.x) {
        if (i < n) { lex_rows[i] = ~static_cast<uint32_t>(k[i] >> 32); lex_norm[i] = __uint_as_float(static_cast<uint32_t>(k[i])); }
        else { lex_rows[i] = 0xffffffffu; lex_norm[i] = 0.0f; }       // sentinels sort last and match no row
    }
}

} // namespace
} // namespace rlr

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Bm25Ws {                      // per-request device workspace
    unsigned long long *d_keys = nullptr; size_t keys_cap = 0;
    float *d_scores = nullptr; size_t scores_cap = 0;      // inverted-index path: per-row running sums
    uint32_t *d_hist = nullptr;
    rlr::Bm25Sel *d_sel = nullptr;
    unsigned long long *d_out = nullptr;
    uint32_t *d_desc_rows = nullptr; float *d_desc_scores = nullptr; uint32_t *d_n = nullptr;
    uint32_t *d_lex_rows = nullptr; float *d_lex_norm = nullptr;
    uint32_t *h_rows = nullptr; float *h_scores = nullptr; uint32_t *h_n = nullptr;      // pinned
    cudaStream_t stream = nullptr;
};

void ws_free(Bm25Ws *w)
{
    if (!w) return;
    cudaFree(w->d_keys); cudaFree(w->d_scores); cudaFree(w->d_hist); cudaFree(w->d_sel); cudaFree(w->d_out); cudaFree(w->d_desc_rows);
    cudaFree(w->d_desc_scores); cudaFree(w->d_n); cudaFree(w->d_lex_rows); cudaFree(w->d_lex_norm);
    cudaFreeHost(w->h_rows); cudaFreeHost(w->h_scores); cudaFreeHost(w->h_n);
    if (w->stream) cudaStreamDestroy(w->stream);
    cudaGetLastError();
    delete w;
}

} // namespace

struct rlr_bm25 {
    rlr_store *s = nullptr;
    int device = 0;                  // cached: destroy must not touch the store (it may already be gone)
    // numeric forward index on the host (rebuilt into CSR on the device when dirty): per local row, (term, tf) by term
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> docs;
    std::vector<uint32_t> doc_len;
    std::vector<uint32_t> df;        // documents holding term id t (== term_postings[t].len())
    uint64_t total_docs = 0, total_length = 0, n_terms_live = 0;
    bool dirty = true;
    unsigned long long *d_off = nullptr; uint32_t *d_terms = nullptr, *d_tfs = nullptr, *d_doclen = nullptr;
    // stores beyond kBm25InvertedMinRows: the device holds postings BY TERM instead (local rows ascending per term);
    // poff (host) are the term boundaries, read per query
    bool inverted = false;
    std::vector<unsigned long long> poff;
    uint32_t *d_pdocs = nullptr, *d_ptfs = nullptr;
    uint64_t dev_rows = 0;
    std::mutex mu;
    std::mutex sync_mu;              // the first queries after a mutation may arrive together: one of them rebuilds the CSR
    std::vector<Bm25Ws *> free_ws;
};

namespace {

int ws_acquire(rlr_bm25 *ix, Bm25Ws **out)
{
    {
        std::lock_guard<std::mutex> lk(ix->mu);
        if (!ix->free_ws.empty()) { *out = ix->free_ws.back(); ix->free_ws.pop_back(); return RLR_OK; }
    }
    Bm25Ws *w = new Bm25Ws();
    cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_hist, rlr::kBins * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(w->d_hist, 0, rlr::kBins * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&w->d_sel, sizeof(rlr::Bm25Sel));
    if (e == cudaSuccess) e = cudaMalloc(&w->d_out, rlr::kBm25MaxLimit * 8);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_desc_rows, rlr::kBm25MaxLimit * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_desc_scores, rlr::kBm25MaxLimit * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_n, 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_lex_rows, rlr::kBm25MaxLimit * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w->d_lex_norm, rlr::kBm25MaxLimit * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_rows, rlr::kBm25MaxLimit * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_scores, rlr::kBm25MaxLimit * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_n, 4);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaGetLastError(); ws_free(w); return fail(RLR_ERR_OOM, "bm25 workspace allocation failed: %s", cudaGetErrorString(e)); }
    *out = w;
    return RLR_OK;
}

void ws_release(rlr_bm25 *ix, Bm25Ws *w)
{
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->free_ws.push_back(w);
}

// (re)build the device CSR from the host forward index; mutators require exclusivity, like the store's
int bm25_sync(rlr_bm25 *ix)
{
    std::lock_guard<std::mutex> lk(ix->sync_mu);
    if (!ix->dirty) return RLR_OK;
    const uint64_t n = ix->s->n_rows;
    if (ix->docs.size() < n) { ix->docs.resize(n); ix->doc_len.resize(n, 0); }
    cudaFree(ix->d_off); cudaFree(ix->d_terms); cudaFree(ix->d_tfs); cudaFree(ix->d_doclen); cudaFree(ix->d_pdocs); cudaFree(ix->d_ptfs);
    ix->d_off = nullptr; ix->d_terms = ix->d_tfs = ix->d_doclen = ix->d_pdocs = ix->d_ptfs = nullptr;
    const char *const force = getenv("RLR_BM25_INDEX");                 // dev / test knob: "forward" | "inverted" whatever the size
    ix->inverted = force ? strcmp(force, "inverted") == 0 : n > rlr::kBm25InvertedMinRows;
    CU_TRY(cudaMalloc(&ix->d_doclen, std::max<uint64_t>(n, 1) * 4));
    if (n) CU_TRY(cudaMemcpy(ix->d_doclen, ix->doc_len.data(), n * 4, cudaMemcpyHostToDevice));
    if (!ix->inverted) {
        // forward index (CSR by row): one thread per row looks the query terms up in the row's sorted term list
        std::vector<unsigned long long> off(n + 1, 0);
        for (uint64_t r = 0; r < n; ++r) off[r + 1] = off[r] + ix->docs[r].size();
        const unsigned long long total = off[n];
        std::vector<uint32_t> terms(total), tfs(total);
        for (uint64_t r = 0; r < n; ++r) {
            unsigned long long o = off[r];
            for (auto &p : ix->docs[r]) { terms[o] = p.first; tfs[o] = p.second; ++o; }
        }
        CU_TRY(cudaMalloc(&ix->d_off, (n + 1) * 8));
        CU_TRY(cudaMalloc(&ix->d_terms, std::max<unsigned long long>(total, 1) * 4));
        CU_TRY(cudaMalloc(&ix->d_tfs, std::max<unsigned long long>(total, 1) * 4));
        CU_TRY(cudaMemcpy(ix->d_off, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
        if (total) {
            CU_TRY(cudaMemcpy(ix->d_terms, terms.data(), total * 4, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(ix->d_tfs, tfs.data(), total * 4, cudaMemcpyHostToDevice));
        }
    } else {
        // inverted index (CSR by term): rows are visited in ascending order, so every term's postings ascend by row
        const size_t vocab = ix->df.size();
        ix->poff.assign(vocab + 1, 0);
        for (size_t t = 0; t < vocab; ++t) ix->poff[t + 1] = ix->poff[t] + ix->df[t];
        const unsigned long long total = ix->poff[vocab];
        std::vector<uint32_t> pdocs(total), ptfs(total);
        std::vector<unsigned long long> cur(ix->poff.begin(), ix->poff.end() - 1);
        for (uint64_t r = 0; r < n; ++r)
            for (auto &p : ix->docs[r]) { const unsigned long long o = cur[p.first]++; pdocs[o] = static_cast<uint32_t>(r); ptfs[o] = p.second; }
        CU_TRY(cudaMalloc(&ix->d_pdocs, std::max<unsigned long long>(total, 1) * 4));
        CU_TRY(cudaMalloc(&ix->d_ptfs, std::max<unsigned long long>(total, 1) * 4));
        if (total) {
            CU_TRY(cudaMemcpy(ix->d_pdocs, pdocs.data(), total * 4, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(ix->d_ptfs, ptfs.data(), total * 4, cudaMemcpyHostToDevice));
        }
    }
    ix->dev_rows = n;
    ix->dirty = false;
    return RLR_OK;
}

} // namespace

// Enqueue LexicalIndex::score(query, limit) on `st`.  d_lex_rows / d_lex_norm (lex_pad entries, nullable) receive the
// (sorted local rows, score / max_lexical) form that the scan kernel blends; d_desc_* / d_n (nullable) the ranked list.
// *active = false when the query cannot match anything (no terms / empty index): nothing is enqueued.
int rlr_api_bm25_enqueue(rlr_bm25 *ix, void *ws_opaque, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                         uint32_t *d_lex_rows, float *d_lex_norm, uint32_t lex_pad, uint32_t *d_desc_rows, float *d_desc_scores,
                         uint32_t *d_n, cudaStream_t st, bool *active, const rlr_api_bm25_global *gs, uint32_t *launches_out)
{
    Bm25Ws *w = static_cast<Bm25Ws *>(ws_opaque);
    *active = false;
    if (limit == 0 || limit > rlr::kBm25MaxLimit) return fail(RLR_ERR_UNSUPPORTED, "bm25 limit %u not in 1..%u", limit, rlr::kBm25MaxLimit);
    if (n_terms && !query_terms) return fail(RLR_ERR_INVALID_ARG, "query_terms is NULL");
    if (int rc = bm25_sync(ix)) return rc;
    rlr_store *s = ix->s;
    const uint32_t n = static_cast<uint32_t>(s->n_rows);
    if (ix->total_docs == 0 || n == 0) return RLR_OK;                                  // :2170-2172 (a shard without documents has no hits)
    // one shard of a cluster scores with the statistics of the WHOLE corpus
    const uint64_t total_docs = gs ? gs->total_docs : ix->total_docs, total_length = gs ? gs->total_length : ix->total_length;
    rlr::Bm25Query q;
    memset(&q, 0, sizeof q);
    // avg_doc_len = total_length as f32 / total_docs as f32 (:2188)
    q.avg_doc_len = static_cast<float>(total_length) / static_cast<float>(total_docs);
    for (uint32_t j = 0; j < n_terms; ++j) {
        const uint32_t t = query_terms[j];
        bool dup = false;
        for (uint32_t i = 0; i < q.n_terms; ++i) dup |= q.term[i] == t;                // HashSet: unique terms
        const uint32_t df_t = gs ? gs->df[j] : (t < ix->df.size() ? ix->df[t] : 0u);
        if (dup || df_t == 0) continue;                                                // `if let Some(postings)`
        if (q.n_terms == rlr::kBm25MaxTerms) return fail(RLR_ERR_UNSUPPORTED, "more than %u unique query terms", rlr::kBm25MaxTerms);
        // idf = ((N - df + 0.5) / (df + 0.5)).ln().max(0.0), f32 throughout, the C library's logf (:2197-2200)
        const float df = static_cast<float>(df_t);
        volatile float num = static_cast<float>(total_docs) - df;
        num = num + 0.5f;
        volatile float den = df + 0.5f;
        volatile float ratio = num / den;
        q.term[q.n_terms] = t;
        q.idf[q.n_terms] = fmaxf(logf(ratio), 0.0f);
        ++q.n_terms;
    }
    if (q.n_terms == 0) return RLR_OK;
    if (w->keys_cap < n) {
        cudaFree(w->d_keys); w->d_keys = nullptr; w->keys_cap = 0;
        CU_TRY(cudaMalloc(&w->d_keys, static_cast<size_t>(s->capacity > n ? s->capacity : n) * 8));
        w->keys_cap = s->capacity > n ? s->capacity : n;
    }
    CU_TRY(cudaMemsetAsync(w->d_sel, 0, sizeof(rlr::Bm25Sel), st));
    uint32_t launches = 0;
    if (!ix->inverted) {
        rlr::bm25_score_kernel<<<(n + 255) / 256, 256, 0, st>>>(ix->d_off, ix->d_terms, ix->d_tfs, ix->d_doclen, n,
                                                                static_cast<uint32_t>(s->row_base), q, w->d_keys);
        launches = 1;
    } else {
        if (w->scores_cap < n) {
            cudaFree(w->d_scores); w->d_scores = nullptr; w->scores_cap = 0;
            CU_TRY(cudaMalloc(&w->d_scores, static_cast<size_t>(s->capacity > n ? s->capacity : n) * 4));
            w->scores_cap = s->capacity > n ? s->capacity : n;
        }
        CU_TRY(cudaMemsetAsync(w->d_scores, 0xff, static_cast<size_t>(n) * 4, st));
        for (uint32_t j = 0; j < q.n_terms; ++j) {                      // consecutive launches: the caller's term order per document
            const uint32_t t = q.term[j];
            if (t + 1 >= ix->poff.size()) continue;                     // a term no local document holds
            const unsigned long long p_lo = ix->poff[t], p_hi = ix->poff[t + 1];
            if (p_hi == p_lo) continue;
            const unsigned long long blocks = (p_hi - p_lo + 255) / 256;
            const uint32_t g = static_cast<uint32_t>(std::min<unsigned long long>(blocks, static_cast<unsigned long long>(s->sm_count) * 16ull));
            rlr::bm25_term_pass_kernel<<<g, 256, 0, st>>>(ix->d_pdocs, ix->d_ptfs, p_lo, p_hi, ix->d_doclen, q.idf[j], q.avg_doc_len, w->d_scores);
            ++launches;
        }
        rlr::bm25_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(w->d_scores, n, static_cast<uint32_t>(s->row_base), w->d_keys);
        ++launches;
    }
    const uint32_t grid = std::min<uint32_t>((n + 255) / 256, static_cast<uint32_t>(s->sm_count) * 4u);
    for (uint32_t pass = 0; pass < rlr::kPasses; ++pass)
        rlr::bm25_select_pass_kernel<<<grid, 256, 0, st>>>(w->d_keys, n, pass, limit, w->d_hist, w->d_sel);
    rlr::bm25_compact_kernel<<<grid, 256, 0, st>>>(w->d_keys, n, w->d_sel, w->d_out, rlr::kBm25MaxLimit);
    uint32_t lim2 = 2;
    while (lim2 < limit) lim2 <<= 1;
    {
        static std::mutex cfg_mu;
        static bool configured[64] = {false};
        std::lock_guard<std::mutex> lk(cfg_mu);
        if (!configured[s->device]) {
            CU_TRY(cudaFuncSetAttribute(rlr::bm25_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rlr::kBm25MaxLimit * 8));
            configured[s->device] = true;
        }
    }
    rlr::bm25_finalize_kernel<<<1, 1024, lim2 * 8, st>>>(w->d_out, w->d_sel, limit, static_cast<uint32_t>(s->row_base), d_desc_rows,
                                                  d_desc_scores, d_n, d_lex_rows, d_lex_norm, lex_pad);
    CU_TRY(cudaGetLastError());
    *active = true;
    if (launches_out) *launches_out = launches + rlr::kPasses + 2;
    return RLR_OK;
}

int rlr_api_bm25_ws_acquire(rlr_bm25 *ix, void **out)
{
    Bm25Ws *w = nullptr;
    if (int rc = ws_acquire(ix, &w)) return rc;
    *out = w;
    return RLR_OK;
}
void rlr_api_bm25_ws_release(rlr_bm25 *ix, void *w) { ws_release(ix, static_cast<Bm25Ws *>(w)); }
rlr_store *rlr_api_bm25_store(rlr_bm25 *ix) { return ix->s; }

RLR_EXPORT int rlr_bm25_create(rlr_store *s, rlr_bm25 **out)
{
    if (!s || !out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (int rc = ensure_device(s->device)) return rc;
    rlr_bm25 *ix = new rlr_bm25();
    ix->s = s;
    ix->device = s->device;
    *out = ix;
    return RLR_OK;
}

RLR_EXPORT int rlr_bm25_destroy(rlr_bm25 *ix)
{
    if (!ix) return RLR_OK;
    cudaSetDevice(ix->device);
    for (Bm25Ws *w : ix->free_ws) ws_free(w);
    cudaFree(ix->d_off); cudaFree(ix->d_terms); cudaFree(ix->d_tfs); cudaFree(ix->d_doclen); cudaFree(ix->d_pdocs); cudaFree(ix->d_ptfs);
    cudaGetLastError();
    delete ix;
    return RLR_OK;
}

namespace {
void bm25_drop(rlr_bm25 *ix, uint32_t row)           // remove_chunk, :2140-2167
{
    if (row >= ix->docs.size() || ix->docs[row].empty()) return;
    for (auto &p : ix->docs[row]) {
        if (p.first < ix->df.size() && ix->df[p.first] > 0 && --ix->df[p.first] == 0) --ix->n_terms_live;
    }
    const uint64_t len = ix->doc_len[row];
    ix->total_length = ix->total_length >= len ? ix->total_length - len : 0;
    if (ix->total_docs > 0) --ix->total_docs;
    if (ix->total_docs == 0) ix->total_length = 0;
    ix->docs[row].clear();
    ix->doc_len[row] = 0;
}
} // namespace

RLR_EXPORT int rlr_bm25_set_doc(rlr_bm25 *ix, uint32_t row, const uint32_t *term_ids, const uint32_t *term_freqs, uint32_t n_terms)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    if (n_terms && (!term_ids || !term_freqs)) return fail(RLR_ERR_INVALID_ARG, "term_ids/term_freqs is NULL");
    const uint64_t g = row;
    if (g < ix->s->row_base || g - ix->s->row_base >= ix->s->n_rows) return fail(RLR_ERR_INVALID_ARG, "row %u not in the store", row);
    const uint32_t r = static_cast<uint32_t>(g - ix->s->row_base);
    if (ix->docs.size() <= r) { ix->docs.resize(ix->s->n_rows); ix->doc_len.resize(ix->s->n_rows, 0); }
    bm25_drop(ix, r);                                   // add_chunk replaces an existing document (:2107-2109)
    ix->dirty = true;
    std::vector<std::pair<uint32_t, uint32_t>> d;
    uint64_t len = 0;
    for (uint32_t i = 0; i < n_terms; ++i)
        if (term_freqs[i]) { d.emplace_back(term_ids[i], term_freqs[i]); len += term_freqs[i]; }
    if (d.empty() || len == 0) return RLR_OK;           // no tokens: the document is not indexed (:2112-2114, :2122-2124)
    std::sort(d.begin(), d.end());
    for (size_t i = 1; i < d.size(); ++i)
        if (d[i].first == d[i - 1].first) return fail(RLR_ERR_INVALID_ARG, "term id %u listed twice for row %u", d[i].first, row);
    if (len > 0xffffffffull) return fail(RLR_ERR_UNSUPPORTED, "document too long");
    for (auto &p : d) {
        if (p.first >= ix->df.size()) ix->df.resize(static_cast<size_t>(p.first) + 1, 0);
        if (ix->df[p.first]++ == 0) ++ix->n_terms_live;
    }
    ix->docs[r] = std::move(d);
    ix->doc_len[r] = static_cast<uint32_t>(len);
    ++ix->total_docs;
    ix->total_length += len;
    return RLR_OK;
}

RLR_EXPORT int rlr_bm25_set_docs(rlr_bm25 *ix, uint32_t row0, uint32_t n_docs, const uint64_t *offsets, const uint32_t *term_ids,
                                 const uint32_t *term_freqs)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    if (n_docs && !offsets) return fail(RLR_ERR_INVALID_ARG, "offsets is NULL");
    for (uint32_t i = 0; i < n_docs; ++i) {
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0xffffffffull) return fail(RLR_ERR_INVALID_ARG, "offsets must ascend");
        const uint32_t nt = static_cast<uint32_t>(offsets[i + 1] - offsets[i]);
        if (int rc = rlr_bm25_set_doc(ix, row0 + i, nt ? term_ids + offsets[i] : nullptr, nt ? term_freqs + offsets[i] : nullptr, nt)) return rc;
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_bm25_remove_doc(rlr_bm25 *ix, uint32_t row)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    const uint64_t g = row;
    if (g < ix->s->row_base) return fail(RLR_ERR_INVALID_ARG, "row %u not in the store", row);
    bm25_drop(ix, static_cast<uint32_t>(g - ix->s->row_base));
    ix->dirty = true;
    return RLR_OK;
}

// follow rlr_store_remove_rows: `to` takes over the document of `from` (whose slot becomes empty)
RLR_EXPORT int rlr_bm25_move_doc(rlr_bm25 *ix, uint32_t from_row, uint32_t to_row)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    const uint64_t base = ix->s->row_base;
    if (from_row < base || to_row < base) return fail(RLR_ERR_INVALID_ARG, "row not in the store");
    const uint32_t f = static_cast<uint32_t>(from_row - base), t = static_cast<uint32_t>(to_row - base);
    const size_t need = static_cast<size_t>(std::max(f, t)) + 1;
    if (ix->docs.size() < need) { ix->docs.resize(need); ix->doc_len.resize(need, 0); }
    bm25_drop(ix, t);
    ix->docs[t] = std::move(ix->docs[f]);
    ix->docs[f].clear();
    ix->doc_len[t] = ix->doc_len[f];
    ix->doc_len[f] = 0;
    ix->dirty = true;
    return RLR_OK;
}

RLR_EXPORT int rlr_bm25_stats(const rlr_bm25 *ix, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    if (total_docs) *total_docs = ix->total_docs;
    if (total_length) *total_length = ix->total_length;
    if (n_terms) *n_terms = ix->n_terms_live;
    return RLR_OK;
}

RLR_EXPORT int rlr_bm25_score(rlr_bm25 *ix, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit, uint32_t *out_rows,
                              float *out_scores, uint32_t cap, uint32_t *out_n)
{
    if (!ix || !out_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    *out_n = 0;
    if (int rc = ensure_device(ix->s->device)) return rc;
    if (limit == 0 || limit > rlr::kBm25MaxLimit) return fail(RLR_ERR_UNSUPPORTED, "limit %u not in 1..%u", limit, rlr::kBm25MaxLimit);
    if (cap < limit && cap < ix->total_docs) return fail(RLR_ERR_INVALID_ARG, "output capacity %u < limit %u", cap, limit);
    void *wv = nullptr;
    if (int rc = rlr_api_bm25_ws_acquire(ix, &wv)) return rc;
    Bm25Ws *w = static_cast<Bm25Ws *>(wv);
    bool active = false;
    int rc = rlr_api_bm25_enqueue(ix, w, query_terms, n_terms, limit, nullptr, nullptr, 0, w->d_desc_rows, w->d_desc_scores, w->d_n,
                                  w->stream, &active);
    if (rc == RLR_OK && active) {
        cudaError_t e = cudaMemcpyAsync(w->h_n, w->d_n, 4, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(w->h_rows, w->d_desc_rows, limit * 4, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(w->h_scores, w->d_desc_scores, limit * 4, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
        if (e != cudaSuccess) { cudaGetLastError(); rc = fail(RLR_ERR_CUDA, "bm25 score failed: %s", cudaGetErrorString(e)); }
        else {
            const uint32_t n = std::min(std::min(w->h_n[0], limit), cap);
            if (n && (!out_rows || !out_scores)) rc = fail(RLR_ERR_INVALID_ARG, "output buffers are NULL");
            else {
                memcpy(out_rows, w->h_rows, n * 4);
                memcpy(out_scores, w->h_scores, n * 4);
                *out_n = n;
            }
        }
    }
    rlr_api_bm25_ws_release(ix, w);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// the same index over a sharded store (rlr_cluster): one rlr_bm25 per shard, global statistics at query time
// ---------------------------------------------------------------------------------------------------------------
struct rlr_cluster_bm25 {
    rlr_cluster *cl = nullptr;
    std::vector<rlr_bm25 *> parts;
};

rlr_cluster *rlr_api_cluster_bm25_cluster(rlr_cluster_bm25 *ix) { return ix->cl; }
rlr_bm25 *rlr_api_cluster_bm25_part(rlr_cluster_bm25 *ix, uint32_t i) { return ix->parts[i]; }

namespace {
// the shard that owns global row `row` (shards are contiguous row ranges)
rlr_bm25 *part_of(rlr_cluster_bm25 *ix, uint32_t row)
{
    for (rlr_bm25 *p : ix->parts)
        if (row >= p->s->row_base && row - p->s->row_base < p->s->n_rows) return p;
    return nullptr;
}
} // namespace

RLR_EXPORT int rlr_cluster_bm25_create(rlr_cluster *cl, rlr_cluster_bm25 **out)
{
    if (!cl || !out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    rlr_cluster_bm25 *ix = new rlr_cluster_bm25();
    ix->cl = cl;
    for (uint32_t i = 0; i < rlr_api_cluster_n(cl); ++i) {
        rlr_bm25 *p = nullptr;
        if (int rc = rlr_bm25_create(rlr_api_cluster_shard(cl, i), &p)) { rlr_cluster_bm25_destroy(ix); return rc; }
        ix->parts.push_back(p);
    }
    *out = ix;
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_bm25_destroy(rlr_cluster_bm25 *ix)
{
    if (!ix) return RLR_OK;
    for (rlr_bm25 *p : ix->parts) rlr_bm25_destroy(p);
    delete ix;
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_bm25_set_doc(rlr_cluster_bm25 *ix, uint32_t row, const uint32_t *term_ids, const uint32_t *term_freqs, uint32_t n_terms)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    rlr_bm25 *p = part_of(ix, row);
    if (!p) return fail(RLR_ERR_INVALID_ARG, "row %u not in the cluster", row);
    return rlr_bm25_set_doc(p, row, term_ids, term_freqs, n_terms);
}

RLR_EXPORT int rlr_cluster_bm25_set_docs(rlr_cluster_bm25 *ix, uint32_t row0, uint32_t n_docs, const uint64_t *offsets,
                                         const uint32_t *term_ids, const uint32_t *term_freqs)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    if (n_docs && !offsets) return fail(RLR_ERR_INVALID_ARG, "offsets is NULL");
    for (uint32_t i = 0; i < n_docs; ++i) {
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0xffffffffull) return fail(RLR_ERR_INVALID_ARG, "offsets must ascend");
        const uint32_t nt = static_cast<uint32_t>(offsets[i + 1] - offsets[i]);
        if (int rc = rlr_cluster_bm25_set_doc(ix, row0 + i, nt ? term_ids + offsets[i] : nullptr, nt ? term_freqs + offsets[i] : nullptr, nt)) return rc;
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_bm25_remove_doc(rlr_cluster_bm25 *ix, uint32_t row)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    rlr_bm25 *p = part_of(ix, row);
    if (!p) return fail(RLR_ERR_INVALID_ARG, "row %u not in the cluster", row);
    return rlr_bm25_remove_doc(p, row);
}

RLR_EXPORT int rlr_cluster_bm25_stats(const rlr_cluster_bm25 *ix, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "index is NULL");
    uint64_t docs = 0, len = 0, live = 0;
    size_t vocab = 0;
    for (const rlr_bm25 *p : ix->parts) { docs += p->total_docs; len += p->total_length; vocab = std::max(vocab, p->df.size()); }
    if (n_terms) {
        for (size_t t = 0; t < vocab; ++t) {
            bool any = false;
            for (const rlr_bm25 *p : ix->parts) any |= t < p->df.size() && p->df[t] != 0;
            live += any ? 1 : 0;
        }
        *n_terms = live;
    }
    if (total_docs) *total_docs = docs;
    if (total_length) *total_length = len;
    return RLR_OK;
}

int rlr_api_cluster_bm25_score(rlr_cluster_bm25 *ix, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                               std::vector<uint32_t> &rows, std::vector<float> &scores)
{
    rows.clear(); scores.clear();
    if (limit == 0 || limit > rlr::kBm25MaxLimit) return fail(RLR_ERR_UNSUPPORTED, "limit %u not in 1..%u", limit, rlr::kBm25MaxLimit);
    if (n_terms && !query_terms) return fail(RLR_ERR_INVALID_ARG, "query_terms is NULL");
    // the statistics of the whole corpus: N, total length, df of every query term (:2188, :2197)
    rlr_api_bm25_global gs = {0, 0, nullptr};
    std::vector<uint32_t> df(n_terms, 0);
    for (const rlr_bm25 *p : ix->parts) {
        gs.total_docs += p->total_docs; gs.total_length += p->total_length;
        for (uint32_t j = 0; j < n_terms; ++j) df[j] += query_terms[j] < p->df.size() ? p->df[query_terms[j]] : 0u;
    }
    gs.df = df.data();
    if (gs.total_docs == 0 || n_terms == 0) return RLR_OK;
    const size_t np = ix->parts.size();
    // every shard scores its documents and ranks its own `limit` best, driven by one host thread per shard (a shard's
    // request is ~13 launches and copies: in sequence they would cost more host time than the GPUs need) ...
    struct Hit { float score; uint32_t row; };
    std::vector<std::vector<Hit>> part_hits(np);
    std::vector<int> rcs(np, RLR_OK);
    std::vector<std::string> errs(np);
    const auto run = [&](size_t i) {
        rlr_bm25 *p = ix->parts[i];
        int rc = ensure_device(p->s->device);
        void *wv = nullptr;
        if (rc == RLR_OK) rc = rlr_api_bm25_ws_acquire(p, &wv);
        if (rc == RLR_OK) {
            Bm25Ws *w = static_cast<Bm25Ws *>(wv);
            bool act = false;
            rc = rlr_api_bm25_enqueue(p, w, query_terms, n_terms, limit, nullptr, nullptr, 0, w->d_desc_rows, w->d_desc_scores, w->d_n,
                                      w->stream, &act, &gs);
            if (rc == RLR_OK && act) {
                cudaError_t e = cudaMemcpyAsync(w->h_n, w->d_n, 4, cudaMemcpyDeviceToHost, w->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(w->h_rows, w->d_desc_rows, limit * 4, cudaMemcpyDeviceToHost, w->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(w->h_scores, w->d_desc_scores, limit * 4, cudaMemcpyDeviceToHost, w->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
                if (e != cudaSuccess) { cudaGetLastError(); rc = fail(RLR_ERR_CUDA, "bm25 score failed: %s", cudaGetErrorString(e)); }
                else {
                    const uint32_t n = std::min(w->h_n[0], limit);
                    part_hits[i].resize(n);
                    for (uint32_t k = 0; k < n; ++k) part_hits[i][k] = {w->h_scores[k], w->h_rows[k]};
                }
            }
            rlr_api_bm25_ws_release(p, w);
        }
        rcs[i] = rc;
        if (rc != RLR_OK) errs[i] = rlr_last_error();          // the message is thread-local: carry it to the caller
    };
    if (np == 1) run(0);
    else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < np; ++i) th.emplace_back(run, i);
        for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < np; ++i)
        if (rcs[i] != RLR_OK) return fail(rcs[i], "shard %zu: %s", i, errs[i].c_str());
    // ... then the global `limit` best are the best of the shards' lists: each list is already in rank order (score
    // descending, ties to the lower row), so a k-way merge of the <= 16 heads yields them in order
    const auto before = [](const Hit &a, const Hit &b) { return a.score > b.score || (a.score == b.score && a.row < b.row); };
    std::vector<size_t> head(np, 0);
    rows.reserve(limit); scores.reserve(limit);
    while (rows.size() < limit) {
        size_t best = np;
        for (size_t i = 0; i < np; ++i)
            if (head[i] < part_hits[i].size() && (best == np || before(part_hits[i][head[i]], part_hits[best][head[best]]))) best = i;
        if (best == np) break;
        const Hit &h = part_hits[best][head[best]++];
        rows.push_back(h.row); scores.push_back(h.score);
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_cluster_bm25_score(rlr_cluster_bm25 *ix, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                                      uint32_t *out_rows, float *out_scores, uint32_t cap, uint32_t *out_n)
{
    if (!ix || !out_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    *out_n = 0;
    std::vector<uint32_t> rows;
    std::vector<float> scores;
    if (int rc = rlr_api_cluster_bm25_score(ix, query_terms, n_terms, limit, rows, scores)) return rc;
    if (rows.size() > cap) return fail(RLR_ERR_INVALID_ARG, "output capacity %u < %zu results", cap, rows.size());
    if (!rows.empty() && (!out_rows || !out_scores)) return fail(RLR_ERR_INVALID_ARG, "output buffers are NULL");
    if (!rows.empty()) { memcpy(out_rows, rows.data(), rows.size() * 4); memcpy(out_scores, scores.data(), scores.size() * 4); }
    *out_n = static_cast<uint32_t>(rows.size());
    return RLR_OK;
}
