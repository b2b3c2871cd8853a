// merge.cu -- reduce several rank-ordered candidate lists to the best m.
//
// Completes the "stable sort desc + take(initial_k)" of RagEngine::search
// (/root/reference/src/rag_engine.rs:543-548) across the per-CTA lists of scan_topm.cu
// and, on the multi-GPU path, across the all-gathered per-GPU lists.
// Keys are unique (they embed the global row), so the merge is order-exact.
#include "common.cuh"
#include "kernels.cuh"

namespace rlr {

namespace {

constexpr int kMergeThreads = 1024;
constexpr uint32_t kMergeCap = 4096; // entries sorted per block

__global__ void __launch_bounds__(kMergeThreads, 1)
merge_kernel(const rlr_cand *__restrict__ in, uint32_t n_lists, uint32_t m, uint32_t lists_per_block,
             rlr_cand *__restrict__ out, uint32_t *__restrict__ out_n)
{
    extern __shared__ uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *src = reinterpret_cast<uint32_t *>(smem_raw + kMergeCap * 8);

    const uint32_t t = threadIdx.x;
    const uint32_t l0 = blockIdx.x * lists_per_block;
    uint32_t l1 = l0 + lists_per_block;
    if (l1 > n_lists) l1 = n_lists;
    const uint32_t n_in = (l1 - l0) * m;
    const rlr_cand *base = in + static_cast<size_t>(l0) * m;

    uint32_t n2 = 1;
    while (n2 < n_in) n2 <<= 1;
    for (uint32_t i = t; i < n2; i += kMergeThreads) {
        keys[i] = i < n_in ? base[i].key : 0ull;
        src[i] = i;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = t; i < (n2 >> 1); i += kMergeThreads) {
                const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const uint32_t hi = lo | j;
                const uint64_t a = keys[lo], b = keys[hi];
                const bool desc = (lo & k) == 0;
                if ((a < b) == desc) {
                    keys[lo] = b; keys[hi] = a;
                    const uint32_t sa = src[lo], sb = src[hi];
                    src[lo] = sb; src[hi] = sa;
                }
            }
            __syncthreads();
        }
    }
    rlr_cand *o = out + static_cast<size_t>(blockIdx.x) * m;
    for (uint32_t i = t; i < m; i += kMergeThreads) {
        rlr_cand c;
        c.key = i < n2 ? keys[i] : 0ull;
        if (c.key != 0ull) {
            const rlr_cand s = base[src[i]];
            c.emb = s.emb; c.lex = s.lex;
        } else {
            c.emb = 0.0f; c.lex = 0.0f;
        }
        o[i] = c;
    }
    if (out_n != nullptr && gridDim.x == 1) {
        // number of valid records among the first m (keys are sorted, zeros last)
        __shared__ uint32_t s_n;
        if (t == 0) s_n = 0;
        __syncthreads();
        uint32_t local = 0;
        for (uint32_t i = t; i < m && i < n2; i += kMergeThreads) local += keys[i] != 0ull;
        if (local) atomicAdd(&s_n, local);
        __syncthreads();
        if (t == 0) *out_n = s_n;
    }
}

// Few long sorted lists (the all-gathered per-GPU lists): rank by counting.  An element's
// global rank is its position in its own list plus, for every other list, the number of
// keys greater than it (binary search in shared memory).  Keys are unique, so ranks are a
// permutation and every record is written straight to its final slot: no sorting passes.
constexpr uint32_t kRankCap = 16384;   // keys held in shared memory (128 KB)

__global__ void __launch_bounds__(kMergeThreads, 1)
merge_rank_kernel(const rlr_cand *__restrict__ in, uint32_t n_lists, uint32_t m, rlr_cand *__restrict__ out,
                  uint32_t *__restrict__ out_n)
{
    extern __shared__ uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint32_t s_total;
    const uint32_t t = threadIdx.x;
    const uint32_t n = n_lists * m;
    if (t == 0) s_total = 0;
    for (uint32_t i = t; i < n; i += kMergeThreads) keys[i] = in[i].key;
    __syncthreads();
    // valid records per list (zero keys pad the tail)
    if (t < n_lists) {
        const uint64_t *l = keys + t * m;
        uint32_t lo = 0, hi = m;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (l[mid] != 0ull) lo = mid + 1; else hi = mid; }
        atomicAdd(&s_total, lo);
    }
    for (uint32_t e = t; e < n; e += kMergeThreads) {
        const uint64_t x = keys[e];
        if (x == 0ull) continue;
        const uint32_t j = e / m;
        uint32_t rank = e - j * m;
        for (uint32_t i = 0; i < n_lists && rank < m; ++i) {
            if (i == j) continue;
            const uint64_t *l = keys + i * m;
            uint32_t lo = 0, hi = m;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (l[mid] > x) lo = mid + 1; else hi = mid; }
            rank += lo;
        }
        if (rank < m) out[rank] = in[e];
    }
    __syncthreads();
    const uint32_t total = s_total < m ? s_total : m;
    for (uint32_t i = total + t; i < m; i += kMergeThreads) { rlr_cand z; z.key = 0; z.emb = 0.0f; z.lex = 0.0f; out[i] = z; }
    if (t == 0 && out_n != nullptr) *out_n = total;
}

// Root side of the fused exchange: the per-GPU lists are not all-gathered by
// a collective; every rank's scan kernel stores its list straight into this GPU's mailbox slot
// (NVLink peer stores) and then publishes the query's sequence number.  This kernel waits for
// the n_lists flags (acquire, system scope), merges by rank counting exactly as
// merge_rank_kernel does, and hands the slot back (`*consumed = seq`).
__global__ void __launch_bounds__(kMergeThreads, 1)
mailbox_merge_kernel(const rlr_cand *slot, uint32_t list_stride, const unsigned long long *flags,
                     unsigned long long seq, unsigned long long *consumed, uint32_t n_lists, uint32_t m,
                     rlr_cand *__restrict__ out, uint32_t *__restrict__ out_n, uint32_t *status)
{
    extern __shared__ uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint32_t s_total, s_bad;
    const uint32_t t = threadIdx.x;
    const uint32_t n = n_lists * m;
    if (t == 0) { s_total = 0; s_bad = 0; }
    __syncthreads();
    if (t < n_lists) {
        const unsigned long long t_start = globaltimer_ns_common();
        while (ld_acquire_sys_u64(flags + t) < seq) {
            if (globaltimer_ns_common() - t_start > kMailboxTimeoutNs) { *status = 2u; s_bad = 1; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (s_bad != 0) {
        // a rank never delivered: an EMPTY result plus the sticky status word, never a merge of stale lists
        for (uint32_t i = t; i < m; i += kMergeThreads) { rlr_cand z; z.key = 0; z.emb = 0.0f; z.lex = 0.0f; out[i] = z; }
        if (t == 0) {
            if (out_n != nullptr) *out_n = 0;
            st_release_sys_u64(consumed, seq);
        }
        return;
    }
    // the records were written by other GPUs: read them at L2 (never from a stale L1 line)
    for (uint32_t i = t; i < n; i += kMergeThreads) {
        const uint32_t j = i / m;
        keys[i] = __ldcg(reinterpret_cast<const unsigned long long *>(&slot[static_cast<size_t>(j) * list_stride + (i - j * m)].key));
    }
    __syncthreads();
    if (t < n_lists) {
        const uint64_t *l = keys + t * m;
        uint32_t lo = 0, hi = m;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (l[mid] != 0ull) lo = mid + 1; else hi = mid; }
        atomicAdd(&s_total, lo);
    }
    for (uint32_t e = t; e < n; e += kMergeThreads) {
        const uint64_t x = keys[e];
        if (x == 0ull) continue;
        const uint32_t j = e / m;
        uint32_t rank = e - j * m;
        for (uint32_t i = 0; i < n_lists && rank < m; ++i) {
            if (i == j) continue;
            const uint64_t *l = keys + i * m;
            uint32_t lo = 0, hi = m;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (l[mid] > x) lo = mid + 1; else hi = mid; }
            rank += lo;
        }
        if (rank < m) {
            const float2 el = __ldcg(reinterpret_cast<const float2 *>(&slot[static_cast<size_t>(j) * list_stride + (e - j * m)].emb));
            rlr_cand r;
            r.key = x; r.emb = el.x; r.lex = el.y;
            out[rank] = r;
        }
    }
    __syncthreads();
    const uint32_t total = s_total < m ? s_total : m;
    for (uint32_t i = total + t; i < m; i += kMergeThreads) { rlr_cand z; z.key = 0; z.emb = 0.0f; z.lex = 0.0f; out[i] = z; }
    if (t == 0) {
        if (out_n != nullptr) *out_n = total;
        st_release_sys_u64(consumed, seq);        // every read of the slot happened before the barrier above
    }
}

// Batched path across GPUs: query q's lists are in[(j * n_queries + q) * m ...], j < n_lists, each
// sorted descending with unique keys (they embed the global row) and zero padded.
constexpr int kBatchMergeThreads = 256;
__global__ void __launch_bounds__(kBatchMergeThreads)
batch_merge_kernel(const unsigned long long *__restrict__ in, uint32_t n_lists, uint32_t n_queries, uint32_t m,
                   unsigned long long *__restrict__ out, uint32_t *__restrict__ out_cnt)
{
    extern __shared__ uint8_t smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    __shared__ uint32_t s_total;
    const uint32_t t = threadIdx.x, q = blockIdx.x;
    const uint32_t n = n_lists * m;
    if (t == 0) s_total = 0;
    for (uint32_t i = t; i < n; i += kBatchMergeThreads) {
        const uint32_t j = i / m;
        keys[i] = in[(static_cast<size_t>(j) * n_queries + q) * m + (i - j * m)];
    }
    __syncthreads();
    if (t < n_lists) {
        const uint64_t *l = keys + t * m;
        uint32_t lo = 0, hi = m;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (l[mid] != 0ull) lo = mid + 1; else hi = mid; }
        atomicAdd(&s_total, lo);
    }
    unsigned long long *o = out + static_cast<size_t>(q) * m;
    for (uint32_t e = t; e < n; e += kBatchMergeThreads) {
        const uint64_t x = keys[e];
        if (x == 0ull) continue;
        const uint32_t j = e / m;
        uint32_t rank = e - j * m;
        for (uint32_t i = 0; i < n_lists && rank < m; ++i) {
            if (i == j) continue;
            const uint64_t *l = keys + i * m;
            uint32_t lo = 0, hi = m;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (l[mid] > x) lo = mid + 1; else hi = mid; }
            rank += lo;
        }
        if (rank < m) o[rank] = x;
    }
    __syncthreads();
    const uint32_t total = s_total < m ? s_total : m;
    for (uint32_t i = total + t; i < m; i += kBatchMergeThreads) o[i] = 0ull;
    if (t == 0 && out_cnt != nullptr) out_cnt[q] = total;
}

inline uint32_t lists_per_block(uint32_t m)
{
    uint32_t l = kMergeCap / m;
    return l < 2 ? 2 : l;
}

} // namespace

size_t merge_tmp_records(uint32_t n_lists, uint32_t m)
{
    const uint32_t lpb = lists_per_block(m);
    const size_t nb = (n_lists + lpb - 1) / lpb;
    return 2 * nb * m + m;
}

cudaError_t batch_merge_launch(const unsigned long long *d_lists, uint32_t n_lists, uint32_t n_queries, uint32_t m,
                               unsigned long long *d_out, uint32_t *d_out_cnt, cudaStream_t stream)
{
    if (static_cast<uint64_t>(n_lists) * m > kRankCap) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(batch_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRankCap * 8);
    if (e != cudaSuccess) return e;
    batch_merge_kernel<<<n_queries, kBatchMergeThreads, n_lists * m * 8, stream>>>(d_lists, n_lists, n_queries, m, d_out, d_out_cnt);
    return cudaGetLastError();
}

cudaError_t mailbox_merge_launch(const rlr_cand *d_slot, uint32_t list_stride, const unsigned long long *d_flags,
                                 unsigned long long seq, unsigned long long *d_consumed, uint32_t n_lists, uint32_t m,
                                 rlr_cand *d_out, uint32_t *d_out_n, uint32_t *d_status, cudaStream_t stream)
{
    if (static_cast<uint64_t>(n_lists) * m > kRankCap || n_lists > kMergeThreads) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(mailbox_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRankCap * 8);
    if (e != cudaSuccess) return e;
    mailbox_merge_kernel<<<1, kMergeThreads, n_lists * m * 8, stream>>>(d_slot, list_stride, d_flags, seq, d_consumed,
                                                                        n_lists, m, d_out, d_out_n, d_status);
    return cudaGetLastError();
}

cudaError_t merge_launch(const rlr_cand *d_lists, uint32_t n_lists, uint32_t m, rlr_cand *d_tmp, rlr_cand *d_out,
                         uint32_t *d_out_n, cudaStream_t stream, uint32_t *launches)
{
    const int smem = kMergeCap * 12;
    {
        // per-device function attributes (cheap; a process may drive several devices)
        cudaError_t e = cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(merge_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRankCap * 8);
        if (e != cudaSuccess) return e;
    }
    if (static_cast<uint64_t>(n_lists) * m <= kRankCap) {
        merge_rank_kernel<<<1, kMergeThreads, n_lists * m * 8, stream>>>(d_lists, n_lists, m, d_out, d_out_n);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    const uint32_t lpb = lists_per_block(m);
    const rlr_cand *cur = d_lists;
    uint32_t n = n_lists;
    const size_t half = ((n_lists + lpb - 1) / lpb) * static_cast<size_t>(m);
    int ping = 0;
    for (;;) {
        const uint32_t nb = (n + lpb - 1) / lpb;
        rlr_cand *dst = nb == 1 ? d_out : d_tmp + (ping ? half : 0);
        merge_kernel<<<nb, kMergeThreads, smem, stream>>>(cur, n, m, lpb, dst, nb == 1 ? d_out_n : nullptr);
        if (launches) ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (nb == 1) break;
        cur = dst;
        n = nb;
        ping ^= 1;
    }
    return cudaSuccess;
}

} // namespace rlr
