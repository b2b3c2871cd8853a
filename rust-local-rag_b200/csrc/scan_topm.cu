// scan_topm.cu -- kernels (1)+(2): exhaustive exact-order cosine scan fused with a
// per-CTA top-M, for sm_100a.
//
// Replaces the hot loop + sort + cut of RagEngine::search,
// /root/reference/src/rag_engine.rs:524-548 (dot_product per chunk :526, blend :531-532,
// stable sort desc :543, take(initial_k) :546), without ever writing an N-length score
// array to HBM.
//
// Design (HBM-bound):
//   * the store is a row-major f32 matrix with 128-byte-aligned rows.  A persistent CTA
//     per SM streams tiles of 128 rows through a ring of shared-memory stages with TMA
//     (cp.async.bulk.tensor.2d, SWIZZLE_128B, L2 evict-first), one producer lane.
//   * each of the 128 consumer threads owns ONE row of the tile and accumulates
//     q[i]*row[i] strictly left to right with separate mul/add roundings, i.e. the
//     reference's exact f32 result.  The 128B swizzle makes the per-thread row walk
//     (LDS.128 at row*128 + (chunk ^ (row&7))*16) bank-conflict free; q is a broadcast
//     read.  The FP32 pipe needs ~0.25 of its issue slots for this, so the sequential
//     order costs nothing against the HBM roofline.
//   * scores are blended (w_e*e + w_l*lex) and encoded as rank keys; a warp ballot +
//     one shared atomic appends the keys that beat the CTA's running M-th best to a
//     2048-entry buffer that is pruned with a bitonic sort when it fills.
//   * each CTA writes its best M records; the LAST CTA to finish (atomic ticket) merges
//     the <=148 lists in place (final_merge), so one launch yields the global top-M.
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "sort_regs.cuh"
#include "mmr_device.cuh"

namespace rlr {

namespace {

constexpr int R = kScanRows;
constexpr int CH = kScanChunks;
constexpr uint32_t kBoxBytes = R * 128;            // one TMA box: 128 rows x 128 B
constexpr uint32_t kStageBytes = CH * kBoxBytes;
constexpr uint32_t kGroupScratch = 32 * 1024;      // per query group: a slice of the idle TMA ring used as scratch by the tail
                                                   // (kTopBuf keys + payloads = 24 KB, then the per-CTA counts)

// Shared memory: the TMA ring, then one block per QUERY GROUP (query, candidate buffer, a line of scalars), then
// the barriers, the tile-id mailbox and (latency path) the lexical pairs.
struct SmemLayout {
    uint32_t stages_off, grp_off, grp_stride, q_rel, keys_rel, embs_rel, misc_rel, bars_off, tile_off, lex_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(int n_stages, uint32_t q_floats, bool lat = false, uint32_t nq = 1)
{
    SmemLayout L;
    uint32_t o = 0;
    L.stages_off = o; o += n_stages * kStageBytes;
    L.grp_off = o;
    L.q_rel = 0;
    L.keys_rel = q_floats * 4;
    L.embs_rel = L.keys_rel + kTopBuf * 8;
    L.misc_rel = L.embs_rel + kTopBuf * 4;
    L.grp_stride = L.misc_rel + 128;
    o += nq * L.grp_stride;
    L.bars_off = o;   o += n_stages * 16;
    L.tile_off = o;   o += 64;
    L.lex_off = o;    o += lat ? kLatLex * 8 : 0;      // latency path: the lexical pairs of the parameter block
    L.total = o;
    return L;
}

__device__ __forceinline__ float lex_lookup(const uint32_t *rows, const float *vals, uint32_t n, uint32_t row)
{
    // plain loads: the arrays live in global memory, or in shared memory on the latency path
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (rows[mid] < row) lo = mid + 1; else hi = mid;
    }
    return (lo < n && rows[lo] == row) ? vals[lo] : 0.0f;
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t x)
{
    return x <= 1 ? 1u : 1u << (32 - __clz(x - 1));
}

// Records written by OTHER CTAs of the same launch are read with ld.global.cg (L2, never
// the non-coherent path); the writers fence before taking their ticket.
__device__ __forceinline__ uint64_t ld_key(const rlr_cand *p)
{
    return __ldcg(reinterpret_cast<const unsigned long long *>(p));
}
__device__ __forceinline__ rlr_cand ld_cand(const rlr_cand *p)
{
    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
    rlr_cand r;
    r.key = (static_cast<uint64_t>(v.y) << 32) | v.x;
    r.emb = __uint_as_float(v.z);
    r.lex = __uint_as_float(v.w);
    return r;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// number of keys >= T in a descending list of `cnt` records
__device__ __forceinline__ uint32_t count_ge(const rlr_cand *list, uint32_t lo, uint32_t cnt, uint64_t T)
{
    uint32_t hi = cnt;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ld_key(list + mid) >= T) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// end (exclusive) of the run of records at positions >= c whose key is > TA.  The first four
// positions are probed with independent loads (one L2 round trip covers the common case).
__device__ __forceinline__ uint32_t extras_end(const rlr_cand *list, uint32_t c, uint32_t cnt, uint64_t TA)
{
    if (cnt <= c) return c;
    uint64_t k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (c + i < cnt) ? ld_key(list + c + i) : 0ull;
    uint32_t n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) n += (n == static_cast<uint32_t>(i) && k[i] > TA) ? 1u : 0u;
    if (n < 4 || c + 4 >= cnt) return c + n;
    return count_ge(list, c + 4, cnt, TA + 1);
}

// Reduce the L per-CTA lists (each sorted descending, `counts[j]` valid records) to the
// global best m, by the R consumer threads of the last CTA.  Keys are unique.
//   1. sample: the first c records of every list, sorted in shared memory; its m-th key
//      T_A is a lower bound of the global m-th key;
//   2. verify: any record outside the sample with key > T_A ("extra") is added and the
//      buffer re-sorted -- for evenly spread data there are none;
//   3. if the extras do not fit (adversarially skewed lists): exact bisection on the key
//      value for the global m-th key, then gather exactly m records.
__device__ void final_merge(uint64_t *keys, float *embs, uint32_t *s_counts, volatile uint32_t *s_cnt,
                            volatile uint32_t *s_aux, const rlr_cand *lists, const uint32_t *counts, uint32_t Ln,
                            uint32_t m, uint32_t row_base, const uint32_t *lex_rows,
                            const float *lex_norm, uint32_t n_lex, rlr_cand *__restrict__ out,
                            uint32_t *__restrict__ out_n, uint32_t t, unsigned long long *tr, uint32_t bar)
{
    // counts -> shared memory (one batched L2 round trip), total valid records
    if (t == 0) { *s_cnt = 0; *s_aux = 0; }
    named_bar_sync(bar, R);
    {
        uint32_t local = 0;
        for (uint32_t j = t; j < Ln; j += R) { const uint32_t cj = __ldcg(counts + j); s_counts[j] = cj; local += cj; }
        if (local) atomicAdd(const_cast<uint32_t *>(s_cnt), local);
    }
    named_bar_sync(bar, R);
    const uint32_t total = *s_cnt;
    const uint32_t m_out = total < m ? total : m;
    named_bar_sync(bar, R);

    // sample depth: the least c whose sample can hold m records, grown while it does not
    // change the padded (power-of-two) sort size
    uint32_t c = (m + Ln - 1) / Ln;
    while (c < m && Ln * (c + 1) <= next_pow2(Ln * c)) ++c;
    if (Ln * c > static_cast<uint32_t>(kTopBuf)) c = kTopBuf / Ln;
    if (c == 0) c = 1;
    const uint32_t nA = Ln * c;
    uint32_t n2 = next_pow2(nA);
#pragma unroll 4
    for (uint32_t i = t; i < n2; i += R) {
        const uint32_t j = i / c, p = i - j * c;
        const bool valid = (i < nA) && (p < s_counts[j < Ln ? j : 0]);
        rlr_cand r;
        r.key = 0; r.emb = 0.0f;
        if (valid) r = ld_cand(lists + static_cast<size_t>(j) * m + p);
        keys[i] = r.key; embs[i] = r.emb;
    }
    named_bar_sync(bar, R);
    if (tr != nullptr && t == 0) tr[1] = globaltimer_ns();
    bitonic_desc(keys, embs, n2, t, bar);
    if (tr != nullptr && t == 0) tr[2] = globaltimer_ns();

    uint32_t n_final = n2;                       // sorted entries currently in keys[]
    if (m_out > 0) {
        const uint64_t TA = (m_out <= nA) ? keys[m_out - 1] : 0ull;   // 0 => sample holds < m_out valid records
        const uint32_t base = (TA != 0ull) ? m_out : nA;
        named_bar_sync(bar, R);
        // count extras: records at positions >= c with key > TA
        if (t == 0) { *s_cnt = 0; *s_aux = 0; }
        named_bar_sync(bar, R);
        uint32_t my_extra = 0;
        for (uint32_t j = t; j < Ln; j += R) my_extra += extras_end(lists + static_cast<size_t>(j) * m, c, s_counts[j], TA) - c;
        if (my_extra) atomicAdd(const_cast<uint32_t *>(s_cnt), my_extra);
        named_bar_sync(bar, R);
        const uint32_t n_extra = *s_cnt;
        named_bar_sync(bar, R);
        if (tr != nullptr && t == 0) { tr[3] = globaltimer_ns(); tr[8] = (static_cast<unsigned long long>(total) << 32) | n_extra; tr[9] = (static_cast<unsigned long long>(c) << 32) | n2; }
        if (n_extra != 0 && base + n_extra <= kTopBuf) {
            // append extras after the kept prefix, re-sort
            for (uint32_t j = t; j < Ln; j += R) {
                const uint32_t cj = s_counts[j];
                if (cj <= c) continue;
                const rlr_cand *lj = lists + static_cast<size_t>(j) * m;
                const uint32_t e_end = extras_end(lj, c, cj, TA);
                if (e_end > c) {
                    const uint32_t slot = atomicAdd(const_cast<uint32_t *>(s_aux), e_end - c);
                    for (uint32_t p = c; p < e_end; ++p) { const rlr_cand r = ld_cand(lj + p); keys[base + slot + p - c] = r.key; embs[base + slot + p - c] = r.emb; }
                }
            }
            n2 = next_pow2(base + n_extra);
            named_bar_sync(bar, R);
            for (uint32_t i = base + n_extra + t; i < n2; i += R) keys[i] = 0;
            named_bar_sync(bar, R);
            bitonic_desc(keys, embs, n2, t, bar);
            n_final = n2;
        } else if (n_extra != 0) {
            // exact bisection for T* = the m_out-th largest key overall
            if (t == 0) *s_aux = 0;
            named_bar_sync(bar, R);
            {
                uint32_t hmax = 0;
                for (uint32_t j = t; j < Ln; j += R)
                    if (__ldcg(counts + j)) { const uint32_t h = static_cast<uint32_t>(ld_key(lists + static_cast<size_t>(j) * m) >> 32); hmax = h > hmax ? h : hmax; }
                atomicMax(const_cast<uint32_t *>(s_aux), hmax);
            }
            named_bar_sync(bar, R);
            uint64_t lo = 1, hi = (static_cast<uint64_t>(*s_aux) << 32) | 0xffffffffull;
            while (lo < hi) {
                const uint64_t mid = lo + ((hi - lo + 1) >> 1);
                named_bar_sync(bar, R);
                if (t == 0) *s_cnt = 0;
                named_bar_sync(bar, R);
                uint32_t cnt = 0;
                for (uint32_t j = t; j < Ln; j += R) cnt += count_ge(lists + static_cast<size_t>(j) * m, 0, __ldcg(counts + j), mid);
                if (cnt) atomicAdd(const_cast<uint32_t *>(s_cnt), cnt);
                named_bar_sync(bar, R);
                if (*s_cnt >= m_out) lo = mid; else hi = mid - 1;
            }
            named_bar_sync(bar, R);
            if (t == 0) *s_cnt = 0;
            named_bar_sync(bar, R);
            for (uint32_t j = t; j < Ln; j += R) {
                const rlr_cand *lj = lists + static_cast<size_t>(j) * m;
                const uint32_t e_end = count_ge(lj, 0, __ldcg(counts + j), lo);
                if (e_end) {
                    const uint32_t slot = atomicAdd(const_cast<uint32_t *>(s_cnt), e_end);
                    for (uint32_t p = 0; p < e_end; ++p) { const rlr_cand r = ld_cand(lj + p); keys[slot + p] = r.key; embs[slot + p] = r.emb; }
                }
            }
            n2 = next_pow2(m_out);
            named_bar_sync(bar, R);
            for (uint32_t i = m_out + t; i < n2; i += R) keys[i] = 0;
            named_bar_sync(bar, R);
            bitonic_desc(keys, embs, n2, t, bar);
            n_final = n2;
        }
    }
    if (tr != nullptr && t == 0) tr[4] = globaltimer_ns();
    for (uint32_t i = t; i < m; i += R) {
        rlr_cand r;
        if (i < m_out && i < n_final) {
            r.key = keys[i];
            r.emb = embs[i];
            r.lex = n_lex ? lex_lookup(lex_rows, lex_norm, n_lex, key_row(r.key) - row_base) : 0.0f;
        } else {
            r.key = 0; r.emb = 0.0f; r.lex = 0.0f;
        }
        out[i] = r;
    }
    if (t == 0) *out_n = m_out;
}

// number of keys in keys[0, n) strictly greater than `mine`: the position of `mine` in descending order (keys are
// unique).  Every thread walks the same addresses (shared-memory broadcast); for n <= R this beats a bitonic sort.
__device__ __forceinline__ uint32_t rank_desc(const uint64_t *keys, uint32_t n, uint64_t mine)
{
    uint32_t r = 0;
#pragma unroll 4
    for (uint32_t i = 0; i < n; ++i) r += keys[i] > mine ? 1u : 0u;
    return r;
}

// Latency path (m <= kLatFusePool): the exact top-m of the L sorted lists through a bound taken from the list HEADS.
// T = the m-th largest head is a lower bound of the global m-th key (m distinct records are >= T), and only the
// <= m lists whose head is >= T can hold records >= T: sort the heads (payload = list id), then look at those m
// lists only -- m * m <= 1024 records, all loaded at once -- and rank the survivors.  Two L2 round trips and one
// small register sort instead of final_merge's sample / verify / append sequence.  The pool is left in `out` (global) AND in `pool_s`
// (shared); *s_pool_n = its size.
__device__ __forceinline__ void merge_small(uint64_t *keys, float *embs, uint32_t *s_counts, volatile uint32_t *s_cnt,
                                         const rlr_cand *lists, const uint32_t *counts, uint32_t Ln, uint32_t m, uint32_t row_base,
                                         const uint32_t *lex_rows, const float *lex_norm, uint32_t n_lex,
                                         rlr_cand *__restrict__ out, uint32_t *__restrict__ out_n, rlr_cand *pool_s,
                                         volatile uint32_t *s_pool_n, uint32_t t, unsigned long long *tr, uint32_t bar)
{
    if (t == 0) *s_cnt = 0;
    named_bar_sync(bar, R);
    {
        // 256 head slots, two per thread, both loads (and both counts) in flight together: one L2 round trip
        static_assert(R == 128, "two head slots per thread");
        const uint32_t ja = t, jb = t + R;
        const uint32_t ca = ja < Ln ? __ldcg(counts + ja) : 0u, cb = jb < Ln ? __ldcg(counts + jb) : 0u;
        const uint64_t ka = ja < Ln ? ld_key(lists + static_cast<size_t>(ja) * m) : 0ull;      // independent of the counts
        const uint64_t kb = jb < Ln ? ld_key(lists + static_cast<size_t>(jb) * m) : 0ull;
        s_counts[ja] = ca; s_counts[jb] = cb;
        keys[ja] = ca ? ka : 0ull; keys[jb] = cb ? kb : 0ull;
        embs[ja] = __uint_as_float(ja); embs[jb] = __uint_as_float(jb);         // payload: the list a head belongs to
        if (ca + cb) atomicAdd(const_cast<uint32_t *>(s_cnt), ca + cb);
    }
    named_bar_sync(bar, R);
    const uint32_t total = *s_cnt;
    const uint32_t m_out = total < m ? total : m;
    named_bar_sync(bar, R);
    if (t == 0) { *s_cnt = 0; *s_pool_n = m_out; }
    if (m_out == 0) {
        for (uint32_t i = t; i < m; i += R) { rlr_cand r; r.key = 0; r.emb = 0.0f; r.lex = 0.0f; out[i] = r; }
        if (t == 0) *out_n = 0;
        named_bar_sync(bar, R);
        return;
    }
    if (tr != nullptr && t == 0) tr[1] = globaltimer_ns();
    bitonic_desc(keys, embs, 256u, t, bar);                       // register sort: faster than ranking 2 x 148 heads
    if (tr != nullptr && t == 0) tr[2] = globaltimer_ns();
    const uint64_t T = keys[m_out - 1];                           // 0: fewer than m_out non-empty lists => keep everything
    constexpr int kPer = (kLatFusePool * kLatFusePool + R - 1) / R;            // records per thread, all in flight at once
    rlr_cand rec[kPer];
    const uint32_t n_scan = m_out * m;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        const uint32_t idx = t + static_cast<uint32_t>(u) * R;
        const uint32_t li = idx < n_scan ? idx / m : 0u;
        const uint32_t p = idx - li * m;
        const uint32_t j = __float_as_uint(embs[li]);             // < 256
        const bool valid = idx < n_scan && keys[li] != 0ull && p < s_counts[j];
        const rlr_cand *src = lists + (valid ? static_cast<size_t>(j) * m + p : 0u);
        rec[u] = ld_cand(src);                                    // always a valid address; masked below
        if (!valid) rec[u].key = 0ull;
    }
    named_bar_sync(bar, R);                                       // the heads are read before keys[] is overwritten
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        if (rec[u].key != 0ull && rec[u].key >= T) {
            const uint32_t slot = atomicAdd(const_cast<uint32_t *>(s_cnt), 1u);
            keys[slot] = rec[u].key; embs[slot] = rec[u].emb;     // <= m * m <= kTopBuf survivors
        }
    }
    named_bar_sync(bar, R);
    const uint32_t n = *s_cnt;                                    // >= m_out
    if (tr != nullptr && t == 0) tr[3] = globaltimer_ns();
    if (n <= static_cast<uint32_t>(R)) {
        if (t < n) {
            const uint64_t k = keys[t];
            const uint32_t r = rank_desc(keys, n, k);
            if (r < m_out) {
                rlr_cand c;
                c.key = k; c.emb = embs[t];
                c.lex = n_lex ? lex_lookup(lex_rows, lex_norm, n_lex, key_row(k) - row_base) : 0.0f;
                pool_s[r] = c;
                out[r] = c;
            }
        }
    } else {
        const uint32_t n2 = next_pow2(n);
        for (uint32_t i = n + t; i < n2; i += R) keys[i] = 0;
        named_bar_sync(bar, R);
        bitonic_desc(keys, embs, n2, t, bar);
        for (uint32_t i = t; i < m_out; i += R) {
            rlr_cand c;
            c.key = keys[i]; c.emb = embs[i];
            c.lex = n_lex ? lex_lookup(lex_rows, lex_norm, n_lex, key_row(c.key) - row_base) : 0.0f;
            pool_s[i] = c;
            out[i] = c;
        }
    }
    for (uint32_t i = m_out + t; i < m; i += R) { rlr_cand c; c.key = 0; c.emb = 0.0f; c.lex = 0.0f; out[i] = c; }
    if (t == 0) *out_n = m_out;
    if (tr != nullptr && t == 0) tr[4] = globaltimer_ns();
    named_bar_sync(bar, R);                                       // pool_s / s_pool_n are visible to the CTA
}

// Fused MMR tail: where pool row r starts in the staging area.  Rows are skewed by 16-byte bank groups so that the
// LDS.128 of eight consecutive EVEN rows (and of eight consecutive ODD rows) -- what a wavefront of pairwise_small
// touches -- fall on eight different bank groups: even row 2b -> group b, odd row 2b+1 -> group b+4 (mod 8).
__device__ __forceinline__ uint32_t lat_row_off(uint32_t r, uint32_t row_stride)
{
    return r * row_stride + ((((r >> 1) + 4u * (r & 1u)) & 7u) << 4);
}

// All pairwise dot products of the P staged pool rows (dot_product, :1776-1779: strict index order, separate multiply
// and add roundings) into the triangle tri[j(j-1)/2 + i], i < j.  One thread owns the pairs (i0, j), (i0+1, j): two
// independent accumulation chains keep the FADD pipe busy (the dependent chain, 4 cycles per add, is the bound) for
// three instead of four row loads per 16-byte step.  Measured: an LDS.128 costs four shared-memory cycles per WARP
// whatever its active lanes, and this loop is bound by them -- so tiles are packed into as few warps as possible
// (P = 15: 56 tiles in two warps, 6 LDS.128 per step; one pair per thread needed four warps and 8).
__device__ __forceinline__ void pairwise_small(const uint8_t *rows_s, uint32_t row_stride, uint32_t vpr, uint32_t P,
                                            float *tri, uint32_t t)
{
    if (P < 2u) return;
    uint32_t n_tiles2 = 0;
    for (uint32_t j = 1; j < P; ++j) n_tiles2 += (j + 1u) / 2u;   // ceil(j / 2) tiles in column j
    for (uint32_t tile = t; tile < n_tiles2; tile += R) {         // packed: as few warps as possible (see above)
        uint32_t j = 1, c = 0;
        while (c + (j + 1u) / 2u <= tile) { c += (j + 1u) / 2u; ++j; }
        const uint32_t i0 = 2u * (tile - c), i1 = i0 + 1u;        // i0 < j always; i1 < j unless j is odd and this is its last tile
        const bool has_i1 = i1 < j;
        const uint8_t *a0 = rows_s + lat_row_off(i0, row_stride);
        const uint8_t *a1 = rows_s + lat_row_off(has_i1 ? i1 : i0, row_stride);
        const uint8_t *b0 = rows_s + lat_row_off(j, row_stride);
        float c0 = 0.0f, c1 = 0.0f;
#pragma unroll 8
        for (uint32_t v = 0; v < vpr; ++v) {
            const float4 x0 = *reinterpret_cast<const float4 *>(a0 + v * 16u);
            const float4 x1 = *reinterpret_cast<const float4 *>(a1 + v * 16u);
            const float4 y = *reinterpret_cast<const float4 *>(b0 + v * 16u);
            c0 = add_rn(c0, mul_rn(x0.x, y.x)); c1 = add_rn(c1, mul_rn(x1.x, y.x));
            c0 = add_rn(c0, mul_rn(x0.y, y.y)); c1 = add_rn(c1, mul_rn(x1.y, y.y));
            c0 = add_rn(c0, mul_rn(x0.z, y.z)); c1 = add_rn(c1, mul_rn(x1.z, y.z));
            c0 = add_rn(c0, mul_rn(x0.w, y.w)); c1 = add_rn(c1, mul_rn(x1.w, y.w));
        }
        const uint32_t row = j * (j - 1u) / 2u;
        tri[row + i0] = c0;
        if (has_i1) tri[row + i1] = c1;
    }
}

constexpr int kTopR = 8;   // a CTA publishes its r-th best score, r = ceil(m / grid) <= kTopR

// warp 0: fold the keys appended since the last call (keys[from, to)) into the CTA's sorted
// top-r and publish the r-th best score.  r rounds of "largest key below the previous one".
__device__ __forceinline__ void top_r_update_warp(const uint64_t *keys, uint32_t from, uint32_t to,
                                                  volatile uint64_t *top, volatile uint32_t *n_top, uint32_t r,
                                                  uint32_t lane, uint32_t *pub_slot)
{
    const uint32_t n_old = *n_top;
    const uint64_t mine_old = lane < n_old ? top[lane] : 0ull;
    __syncwarp();
    uint64_t bound = ~0ull, best = 0;
    uint32_t n = 0;
    for (uint32_t round = 0; round < r; ++round) {
        best = mine_old < bound ? mine_old : 0ull;
        for (uint32_t i = from + lane; i < to; i += 32) {
            const uint64_t k = keys[i];
            if (k < bound && k > best) best = k;
        }
        best = warp_max_u64(best);
        if (best == 0ull) break;
        if (lane == 0) top[round] = best;
        bound = best;
        ++n;
    }
    if (lane == 0) {
        *n_top = n;
        if (n >= r) *reinterpret_cast<volatile uint32_t *>(pub_slot) = static_cast<uint32_t>(best >> 32);
    }
}

// kHalf: the store holds IEEE binary16 rows (64 elements per 128-byte box row).  Elements are
// widened to f32 (exact) and accumulated in f32 in index order, i.e. the reference arithmetic
// applied to the f16-rounded store.
struct NoLat { uint32_t unused; };
template <bool kLat> struct LatSel { typedef NoLat type; };
template <> struct LatSel<true> { typedef LatParams type; };

// kLat: latency path for small f32 stores -- the query and the lexical pairs come from the parameter block `lp`
// (no H2D copy), and the last CTA delivers the result into mapped pinned host memory (lp.mode, see LatParams).
// NQ: QUERY GROUPS.  NQ > 1 answers NQ queries in ONE pass over the rows: the CTA has NQ groups of four consumer
// warps; every group owns one query (its own copy of the query, candidate buffer, thresholds, named barrier, per-CTA
// list and cross-CTA ticket) and all groups consume the SAME shared-memory stages -- each tile crosses HBM once and
// feeds NQ exact sequential chains per row.  (row, query) arithmetic is untouched, so results are bit-identical to
// NQ separate scans; HBM bytes per query drop by NQ.  Throughput mode only (rlr_search_mmr_multi).
template <bool kHalf, bool kLat, int NQ>
__global__ void __launch_bounds__(R * NQ + 64, 1)
scan_topm_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanGroups io,
                 uint32_t n_rows, uint32_t row_base, uint32_t n_chunks, float w_embed, float w_lex,
                 uint32_t m, uint32_t buf_cap, uint32_t r_pub, int n_stages, rlr_cand *g_lists_all, uint32_t *g_counts_all,
                 uint32_t *g_pub_all, uint32_t *g_tickets, uint32_t *g_tile_ctr,
                 unsigned long long *g_trace /* dev-only phase timestamps, may be null */,
                 uint32_t rpt /* rows per tile: a multiple of 8, <= R; the tensor map's box has this many rows */,
                 const __grid_constant__ typename LatSel<kLat>::type lp)
{
    static_assert(!kLat || NQ == 1, "the latency path answers one query per launch");
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B atoms are 1024 B: align the carve-up by hand.
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;

    constexpr uint32_t EPB = kHalf ? 64u : 32u;        // elements per 128-byte box row
    constexpr uint32_t kThreads = R * NQ + 64;
    constexpr uint32_t kConsWarps = (R / 32) * NQ;
    const uint32_t KB = (n_chunks + CH - 1) / CH;      // pipeline stages consumed per tile
    const uint32_t q_floats = KB * CH * EPB;
    const SmemLayout L = smem_layout(n_stages, q_floats, kLat, NQ);

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31;
    const uint32_t grp = warp < kConsWarps ? warp / (R / 32) : 0;        // query group of a consumer warp
    const uint32_t bar = 1 + grp;                                        // the group's named barrier (R threads)
    uint8_t *gs = smem + L.grp_off + grp * L.grp_stride;

    float *q_s = reinterpret_cast<float *>(gs + L.q_rel);
    uint64_t *keys = reinterpret_cast<uint64_t *>(gs + L.keys_rel);
    float *embs = reinterpret_cast<float *>(gs + L.embs_rel);
    volatile uint32_t *s_count = reinterpret_cast<volatile uint32_t *>(gs + L.misc_rel);
    volatile uint64_t *s_tau = reinterpret_cast<volatile uint64_t *>(gs + L.misc_rel + 8);
    volatile uint32_t *s_flag = reinterpret_cast<volatile uint32_t *>(gs + L.misc_rel + 16);
    volatile uint32_t *s_tau_g = reinterpret_cast<volatile uint32_t *>(gs + L.misc_rel + 20);
    volatile uint32_t *s_done = reinterpret_cast<volatile uint32_t *>(gs + L.misc_rel + 24);
    volatile uint32_t *s_ntop = reinterpret_cast<volatile uint32_t *>(gs + L.misc_rel + 28);
    volatile uint64_t *s_top = reinterpret_cast<volatile uint64_t *>(gs + L.misc_rel + 32);   // [kTopR]
    // tile-id mailbox, one slot per pipeline stage (<= 8): the slot belongs to whoever owns the
    // stage, so the producer can run any number of tiles ahead without overwriting an unread id
    volatile uint32_t *s_tile = reinterpret_cast<volatile uint32_t *>(smem + L.tile_off);
    const uint32_t stages_addr = smem_u32(smem + L.stages_off);
    const uint32_t full_bar = smem_u32(smem + L.bars_off);
    const uint32_t empty_bar = full_bar + n_stages * 8;

    // this group's slice of the per-launch global workspace and its outputs
    rlr_cand *g_lists = g_lists_all + static_cast<size_t>(grp) * gridDim.x * m;
    uint32_t *g_counts = g_counts_all + grp * gridDim.x;
    uint32_t *g_pub = g_pub_all + grp * gridDim.x;
    uint32_t *g_ticket = g_tickets + grp;
    rlr_cand *g_out = io.g[grp].out;
    uint32_t *g_out_n = io.g[grp].out_n;
    const ScanPost post = io.g[grp].post;

    const uint32_t n_tiles = (n_rows + rpt - 1) / rpt;
    const uint32_t *lex_rows = io.g[grp].lex_rows;
    const float *lex_norm = io.g[grp].lex_norm;
    const uint32_t n_lex = io.g[grp].n_lex;
    if constexpr (kLat) {
        if (lp.d_lex_rows != nullptr) {
            lex_rows = lp.d_lex_rows; lex_norm = lp.d_lex_norm;       // written on the device by the BM25 stage
        } else {
            uint32_t *lr = reinterpret_cast<uint32_t *>(smem + L.lex_off);
            float *ln = reinterpret_cast<float *>(smem + L.lex_off + kLatLex * 4);
            for (uint32_t i = tid; i < n_lex; i += kThreads) { lr[i] = lp.lex_rows[i]; ln[i] = lp.lex_norm[i]; }
            lex_rows = lr; lex_norm = ln;
        }
    }

    if (g_trace != nullptr && tid == 0) g_trace[4 * gridDim.x + 16 + blockIdx.x] = globaltimer_ns();   // kernel entry
    // The producer warp initialises the pipeline barriers itself, ARRIVES at the prologue barrier and starts
    // issuing TMA loads at once: the first tile is in flight while the consumers still stage the query.
    if (warp == kConsWarps) {
        if (lane == 0) {
            for (int s = 0; s < n_stages; ++s) {
                mbar_init(full_bar + s * 8, 1);
                mbar_init(empty_bar + s * 8, kConsWarps);      // every consumer warp of every group arrives
            }
            fence_mbar_init();
        }
        __syncwarp();
        named_bar_arrive(14, kThreads);
    }
    if (warp < kConsWarps) {
        const uint32_t tg = tid - grp * R;
        if (tg == 0) {
            *s_count = 0;
            *s_tau = 0;
            *s_flag = 0;
            *s_tau_g = 0;
            *s_done = 0;
            *s_ntop = 0;
        }
        if constexpr (kLat) {       // 16 bytes per constant-bank access: the index differs per lane, so accesses serialise
            const float4 *pq = reinterpret_cast<const float4 *>(lp.q);
            for (uint32_t i = tg; i < q_floats / 4u; i += R) reinterpret_cast<float4 *>(q_s)[i] = pq[i];
        }
        else { const float *gq = io.g[grp].query; for (uint32_t i = tg; i < q_floats; i += R) q_s[i] = gq[i]; }
    }
    if (warp != kConsWarps) named_bar_sync(14, kThreads);         // consumers + threshold warp wait; the producer only arrived
    if (g_trace != nullptr && tid == 0) g_trace[blockIdx.x] = globaltimer_ns();

    if (warp == kConsWarps) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&tmap);
            uint64_t pol = policy_evict_first();                  // a streaming pass: do not displace what others keep in L2
            if constexpr (kLat) { if (lp.keep_l2) pol = policy_evict_last(); }   // a store that fits L2 stays there between queries
            uint32_t stage = 0, phase = 0;
            // Tiles are handed out dynamically (first one static, then a global counter): SMs
            // do not get equal shares of HBM bandwidth, and with a static split the slow ones
            // finish ~40 % later than the fast ones.  Tile ids still ascend within a CTA.
            uint32_t tile = blockIdx.x;
            for (;;) {
                const bool has = tile < n_tiles;
                uint32_t next_tile = 0xffffffffu;
                if (has) next_tile = atomicAdd(g_tile_ctr, 1u) + gridDim.x;   // consumed a whole tile later
                mbar_wait(empty_bar + stage * 8, phase ^ 1);
                s_tile[stage] = has ? tile : 0xffffffffu;                    // published by the arrive below
                if (!has) { mbar_arrive(full_bar + stage * 8); break; }
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    if (kb != 0) mbar_wait(empty_bar + stage * 8, phase ^ 1);
                    mbar_arrive_expect_tx(full_bar + stage * 8, CH * rpt * 128u);
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        tma_load_2d(stages_addr + stage * kStageBytes + c * kBoxBytes, &tmap,
                                    static_cast<int32_t>((kb * CH + c) * EPB),
                                    static_cast<int32_t>(tile * rpt), full_bar + stage * 8, pol);
                    if (++stage == static_cast<uint32_t>(n_stages)) { stage = 0; phase ^= 1; }
                }
                tile = next_tile;
            }
        }
        return;
    }
    if (warp == kConsWarps + 1) {
        // ---------------- threshold warp: global lower bound of the m-th best score ----------------
        // Every CTA publishes the score of its r-th best row so far, r = ceil(m / grid).  All
        // CTAs then hold >= r rows at or above min_j pub[j], i.e. >= m rows in total, so the
        // global m-th best score is >= that minimum at any moment: rows strictly below it can
        // be dropped without ever entering a buffer.  Stale reads only make the bound looser.
        // One warp serves every query group.
        if (r_pub == 0) return;
        for (;;) {
            bool all_done = true;
#pragma unroll
            for (int g = 0; g < NQ; ++g) {
                uint8_t *gg = smem + L.grp_off + g * L.grp_stride + L.misc_rel;
                volatile uint32_t *tau_g = reinterpret_cast<volatile uint32_t *>(gg + 20);
                if (*reinterpret_cast<volatile uint32_t *>(gg + 24) != 0) continue;       // this group's consumers left the loop
                all_done = false;
                const volatile uint32_t *pub = g_pub_all + g * gridDim.x;
                uint32_t v = 0xffffffffu;
                for (uint32_t j = lane; j < gridDim.x; j += 32) { const uint32_t u = pub[j]; v = u < v ? u : v; }
                v = __reduce_min_sync(0xffffffffu, v);
                if (lane == 0 && v > *tau_g) *tau_g = v;
            }
            if (all_done) break;
            __nanosleep(400);
        }
        return;
    }

    // --------------------------------- consumers ---------------------------------
    const uint32_t t = tid - grp * R;             // row of the tile owned by this thread (within its query group)
    const uint32_t xr = (t & 7u) << 4;            // SWIZZLE_128B: 16B-chunk index ^= row & 7
    const uint8_t *stage0 = smem + L.stages_off + t * 128;
    const float4 *q4 = reinterpret_cast<const float4 *>(q_s);
    uint32_t stage = 0, phase = 0;
    uint64_t tau = 0;                             // key of the CTA's current M-th best
    uint32_t prev_cnt = 0, n_my_tiles = 0;

    for (;;) {
        mbar_wait(full_bar + stage * 8, phase);   // first stage of the next tile (or the end marker)
        const uint32_t tile = s_tile[stage];
        if (tile == 0xffffffffu) break;
        ++n_my_tiles;
        float acc = 0.0f;
        for (uint32_t kb = 0; kb < KB; ++kb) {
            if (kb != 0) mbar_wait(full_bar + stage * 8, phase);
            const uint8_t *sp = stage0 + stage * kStageBytes;
            const float4 *qp = q4 + kb * (CH * (EPB / 4));
#pragma unroll
            for (int c = 0; c < CH; ++c) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint8_t *src = sp + c * kBoxBytes + ((j << 4) ^ xr);
                    if constexpr (kHalf) {
                        const uint4 raw = *reinterpret_cast<const uint4 *>(src);      // 8 halves
                        const float4 w0 = qp[c * 16 + j * 2], w1 = qp[c * 16 + j * 2 + 1];
                        const float2 v0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
                        const float2 v1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                        const float2 v2 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.z));
                        const float2 v3 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.w));
                        acc = add_rn(acc, mul_rn(w0.x, v0.x));
                        acc = add_rn(acc, mul_rn(w0.y, v0.y));
                        acc = add_rn(acc, mul_rn(w0.z, v1.x));
                        acc = add_rn(acc, mul_rn(w0.w, v1.y));
                        acc = add_rn(acc, mul_rn(w1.x, v2.x));
                        acc = add_rn(acc, mul_rn(w1.y, v2.y));
                        acc = add_rn(acc, mul_rn(w1.z, v3.x));
                        acc = add_rn(acc, mul_rn(w1.w, v3.y));
                    } else {
                        const float4 v = *reinterpret_cast<const float4 *>(src);
                        const float4 w = qp[c * 8 + j];
                        acc = add_rn(acc, mul_rn(w.x, v.x));
                        acc = add_rn(acc, mul_rn(w.y, v.y));
                        acc = add_rn(acc, mul_rn(w.z, v.z));
                        acc = add_rn(acc, mul_rn(w.w, v.w));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + stage * 8);
            if (++stage == static_cast<uint32_t>(n_stages)) { stage = 0; phase ^= 1; }
        }

        // ---- blend (:531-532) and offer to the CTA's top-M ----
        const uint32_t row_local = tile * rpt + t;
        float lexv = 0.0f;
        if (n_lex) lexv = lex_lookup(lex_rows, lex_norm, n_lex, row_local);
        const float combined = add_rn(mul_rn(w_embed, acc), mul_rn(w_lex, lexv));
        const uint64_t key = make_key(combined, row_base + row_local);
        const uint32_t tau_g = *s_tau_g;          // ordered score bits; 0 while unknown
        const bool pass = (t < rpt) && (row_local < n_rows) && (key > tau) && (static_cast<uint32_t>(key >> 32) >= tau_g);
        const uint32_t mask = __ballot_sync(0xffffffffu, pass);
        if (mask) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(const_cast<uint32_t *>(s_count), __popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) {
                const uint32_t idx = base + __popc(mask & ((1u << lane) - 1u));
                keys[idx] = key;
                embs[idx] = acc;
            }
        }
        named_bar_sync(bar, R);
        uint32_t cnt = *s_count;
        if (warp == 0 && r_pub != 0 && cnt > prev_cnt)
            top_r_update_warp(keys, prev_cnt, cnt, s_top, s_ntop, r_pub, lane, g_pub + blockIdx.x);
        named_bar_sync(bar, R);
        if (cnt > buf_cap - R) {
            // prune: keep the best m, raise the threshold.  buf_cap (a power of two <= kTopBuf)
            // is sized so that one prune costs about as much HBM time as the TMA ring holds.
            for (uint32_t i = cnt + t; i < buf_cap; i += R) keys[i] = 0;
            named_bar_sync(bar, R);
            bitonic_desc(keys, embs, buf_cap, t, bar);
            if (t == 0) {
                const uint32_t kept = cnt < m ? cnt : m;
                *s_count = kept;
                *s_tau = cnt >= m ? keys[m - 1] : 0;
                if (r_pub != 0) {                 // the sorted prefix IS the top-r now
                    const uint32_t nt = kept < r_pub ? kept : r_pub;
                    for (uint32_t i = 0; i < nt; ++i) s_top[i] = keys[i];
                    *s_ntop = nt;
                    if (nt >= r_pub)
                        *reinterpret_cast<volatile uint32_t *>(g_pub + blockIdx.x) = static_cast<uint32_t>(keys[r_pub - 1] >> 32);
                }
            }
            named_bar_sync(bar, R);
            tau = *s_tau;
            cnt = *s_count;
            named_bar_sync(bar, R);             // reads done before the next tile's appends bump s_count
        }
        prev_cnt = cnt;
    }

    // ---- final: drop what the freshest global bound excludes, sort the rest, write the list ----
    if constexpr (NQ > 1) named_bar_sync(15, R * NQ);     // every group has consumed its last stage: the ring is idle for all
    if (g_trace != nullptr && tid == 0) g_trace[gridDim.x + blockIdx.x] = globaltimer_ns();
    if (t == 0) { *s_done = 1; *s_flag = 0xffffffffu; }
    named_bar_sync(bar, R);
    const uint32_t cnt = *s_count;
    if (r_pub != 0) {
        const volatile uint32_t *pub = g_pub;
        uint32_t v = 0xffffffffu;
        for (uint32_t j = t; j < gridDim.x; j += R) { const uint32_t u = pub[j]; v = u < v ? u : v; }
        v = __reduce_min_sync(0xffffffffu, v);
        if (lane == 0) atomicMin(const_cast<uint32_t *>(s_flag), v);
    }
    named_bar_sync(bar, R);                     // everyone has read cnt; the atomicMin results are in
    uint32_t tau_fin = *s_tau_g;
    if (r_pub != 0 && *s_flag > tau_fin) tau_fin = *s_flag;
    if (t == 0) *s_count = 0;
    named_bar_sync(bar, R);
    // the TMA ring is idle now (every issued load was consumed): reuse it as the compaction target
    uint8_t *scratch = smem + L.stages_off + grp * kGroupScratch;          // this group's slice of the idle ring
    uint64_t *keys2 = reinterpret_cast<uint64_t *>(scratch);
    float *embs2 = reinterpret_cast<float *>(scratch + kTopBuf * 8);
    for (uint32_t i0 = 0; i0 < cnt; i0 += R) {
        const uint32_t i = i0 + t;
        const uint64_t k = i < cnt ? keys[i] : 0ull;
        const bool keep_it = (i < cnt) && (static_cast<uint32_t>(k >> 32) >= tau_fin);
        const uint32_t mk = __ballot_sync(0xffffffffu, keep_it);
        if (mk) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(const_cast<uint32_t *>(s_count), __popc(mk));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep_it) {
                const uint32_t idx = base + __popc(mk & ((1u << lane) - 1u));
                keys2[idx] = k;
                embs2[idx] = embs[i];
            }
        }
    }
    named_bar_sync(bar, R);
    const uint32_t cnt2 = *s_count;
    const uint32_t keep = cnt2 < m ? cnt2 : m;
    rlr_cand *out = g_lists + static_cast<size_t>(blockIdx.x) * m;
    if (cnt2 <= static_cast<uint32_t>(R)) {
        // one entry per thread (small stores: a single short tile): its rank IS its place in the list, no sort
        if (t < cnt2) {
            const uint64_t k = keys2[t];
            const uint32_t r = rank_desc(keys2, cnt2, k);
            if (r < keep) {
                rlr_cand c;
                c.key = k;
                c.emb = embs2[t];
                c.lex = n_lex ? lex_lookup(lex_rows, lex_norm, n_lex, key_row(k) - row_base) : 0.0f;
                out[r] = c;
            }
        }
    } else {
        const uint32_t n2 = next_pow2(cnt2);
        for (uint32_t i = cnt2 + t; i < n2; i += R) keys2[i] = 0;
        named_bar_sync(bar, R);
        bitonic_desc(keys2, embs2, n2, t, bar);
        for (uint32_t i = t; i < keep; i += R) {       // records beyond `keep` are never read
            rlr_cand c;
            c.key = keys2[i];
            c.emb = embs2[i];
            c.lex = n_lex ? lex_lookup(lex_rows, lex_norm, n_lex, key_row(c.key) - row_base) : 0.0f;
            out[i] = c;
        }
    }
    if (t == 0) g_counts[blockIdx.x] = keep;
    if (g_trace != nullptr && tid == 0) {
        g_trace[2 * gridDim.x + blockIdx.x] = globaltimer_ns();
        g_trace[3 * gridDim.x + blockIdx.x] = (static_cast<unsigned long long>(n_my_tiles) << 48) | (static_cast<unsigned long long>(cnt) << 24) | cnt2;
    }

    // ---- cross-CTA merge, done by whichever CTA finishes last (no second launch) ----
    __threadfence();
    named_bar_sync(bar, R);
    if (t == 0) {
        const uint32_t ticket = atomicAdd(g_ticket, 1u);
        *s_flag = (ticket == gridDim.x - 1) ? 1u : 0u;
    }
    named_bar_sync(bar, R);
    const uint32_t is_last = *s_flag;
    named_bar_sync(bar, R);                     // s_flag is reused below: everyone reads it first
    if (is_last == 0) return;
    __threadfence();
    // stream-ordered launches reuse the ticket, the tile counter (every producer is done once ANY group's last CTA
    // gets here: all CTAs' consumers of that group have seen the end marker) ...
    if (t == 0) { *g_ticket = 0; if (grp == 0) *g_tile_ctr = 0; }
    for (uint32_t j = t; j < gridDim.x; j += R) g_pub[j] = 0;   // ... and the published bounds
    if (g_out == nullptr) return;
    unsigned long long *tr = (g_trace != nullptr && grp == 0) ? g_trace + 4 * gridDim.x : nullptr;
    if (tr != nullptr && t == 0) tr[0] = globaltimer_ns();
    if (post.flag != nullptr) {
        // fused exchange: g_out is a mailbox slot (possibly in a peer GPU's HBM); it is free once
        // the root has merged the query that used it `ring` sequence numbers ago.  On a timeout the
        // slot is NOT touched and the flag is not published: the root's merge then times out for this
        // query as well and delivers an empty result with a sticky status, never a stale list.
        if (t == 0) {
            uint32_t timed_out = 0;
            if (post.consumed != nullptr && post.seq > post.ring) {
                const unsigned long long t_start = globaltimer_ns();
                while (ld_acquire_sys_u64(post.consumed) + post.ring < post.seq) {
                    if (globaltimer_ns() - t_start > kMailboxTimeoutNs) { *post.status = 1u; timed_out = 1; break; }
                    __nanosleep(200);
                }
            }
            s_tile[0] = timed_out;            // the tile mailbox is idle: the producer has exited
        }
        named_bar_sync(bar, R);
        if (s_tile[0] != 0) return;
    }
    bool merged = false;
    if constexpr (kLat) {
        if (lp.mode == 2u && m <= kLatFusePool && gridDim.x <= 256u) {
            // the pool also lands in shared memory (the idle ring's first 512 bytes) for the fused MMR tail below
            merge_small(keys, embs, reinterpret_cast<uint32_t *>(scratch + kTopBuf * 12), s_count, g_lists, g_counts,
                        gridDim.x, m, row_base, lex_rows, lex_norm, n_lex, g_out, g_out_n,
                        reinterpret_cast<rlr_cand *>(smem + L.stages_off), s_flag, t, tr, bar);
            merged = true;
        }
    }
    if (!merged)
        final_merge(keys, embs, reinterpret_cast<uint32_t *>(scratch + kTopBuf * 12), s_count, s_flag, g_lists,
                    g_counts, gridDim.x, m, row_base, lex_rows, lex_norm, n_lex, g_out, g_out_n, t, tr, bar);
    if (post.flag != nullptr) {
        __threadfence_system();               // every writer: the records are visible system-wide ...
        named_bar_sync(bar, R);
        if (t == 0) st_release_sys_u64(post.flag, post.seq);   // ... before the flag says so
    }
    if constexpr (kLat) {
        if (lp.mode == 2u) {
            // ---- fused MMR tail (pool <= kLatFusePool): mmr_diversify (:767-839) by this CTA, no further launch ----
            // merge_small left the pool in shared memory (and in g_out).  Stage the pool's rows in the idle TMA
            // ring, compute every pairwise dot with the reference's sequential arithmetic, run the greedy loop, and
            // let thread 0 write the selection straight into the host's mapped result block.
            uint8_t *ring = smem + L.stages_off;
            rlr_cand *pool_s = reinterpret_cast<rlr_cand *>(ring);                                  // 32 x 16 B
            float *tri_s = reinterpret_cast<float *>(ring + 512);                                   // 496 floats -> 2 KB
            uint64_t (*s_best)[4] = reinterpret_cast<uint64_t (*)[4]>(ring + 512 + 2048);           // 64 B
            uint32_t (*s_besti)[4] = reinterpret_cast<uint32_t (*)[4]>(ring + 512 + 2048 + 64);     // 32 B
            uint8_t *rows_s = ring + 4096;
            const uint32_t row_stride = lp.pitch * 4u + 128u;             // + room for the per-row bank-group skew (lat_row_off)
            uint32_t P;
            if (merged) P = *s_flag;                                      // merge_small left the pool in pool_s already
            else {
                named_bar_sync(bar, R);                                   // g_out / g_out_n written by this CTA
                P = *reinterpret_cast<volatile uint32_t *>(g_out_n);
                if (t < P) pool_s[t] = g_out[t];
                named_bar_sync(bar, R);
            }
            const uint32_t vpr = lp.pitch / 4u;                           // 16-byte vectors per row
            // cp.async (LDGSTS, 16 B each, L2 only): every load of the pool's rows is in flight at once -- one L2 round
            // trip instead of one per loop iteration (6 us -> ~1 us for 15 x 3 KB)
            for (uint32_t idx = t; idx < P * vpr; idx += R) {
                const uint32_t r = idx / vpr, v = idx - r * vpr;
                const float *src = lp.g_rows + static_cast<size_t>(key_row(pool_s[r].key) - row_base) * lp.pitch + v * 4u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(rows_s + lat_row_off(r, row_stride) + v * 16u)), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            named_bar_sync(bar, R);
            if (tr != nullptr && t == 0) tr[5] = globaltimer_ns();
            pairwise_small(rows_s, row_stride, vpr, P, tri_s, t);
            named_bar_sync(bar, R);
            if (tr != nullptr && t == 0) tr[6] = globaltimer_ns();
            if (P == 0) { if (t == 0) *lp.result_n = 0; }
            else greedy_loop<1>(tri_s, pool_s, nullptr, P, lp.top_k, lp.lambda, lp.d_sel_pos, lp.result_n, lp.result, t, s_best, s_besti);
            if (tr != nullptr && t == 0) tr[10] = globaltimer_ns();
            // thread 0 alone wrote the records and the count: its release store orders them before the flag
            if (t == 0) st_release_sys_u64(lp.flag, lp.seq);
        } else if (lp.mode == 1u) {
            // the merged top-m went straight into the host's mapped block (g_out / g_out_n point there)
            __threadfence_system();
            named_bar_sync(bar, R);
            if (t == 0) st_release_sys_u64(lp.flag, lp.seq);
        }
    }
    if (tr != nullptr && t == 0) tr[7] = globaltimer_ns();
}

} // namespace

void scan_plan(int sm_count, int max_smem_optin, uint32_t n_rows, uint32_t pitch, int half, ScanArgs *a, uint32_t n_groups)
{
    const uint32_t epb = half ? 64u : 32u;
    const uint32_t n_chunks = pitch / epb;
    const uint32_t KB = (n_chunks + CH - 1) / CH;
    const uint32_t q_floats = KB * CH * epb;
    if (n_groups < 1) n_groups = 1;
    a->half = half;
    a->n_groups = n_groups;
    const uint32_t n_tiles = (n_rows + R - 1) / R;
    // tuning knob: leave a few SMs to other streams' small kernels (merge / MMR of the previous query)
    // (default 2: measured on a B200, the scan reads HBM just as fast from 140 SMs as from 148 -- 7.52 TB/s either
    // way -- and with two queries in flight the freed SMs let the previous query's tail overlap this scan)
    static const int reserved = getenv("RLR_SCAN_SMS_RESERVED") ? atoi(getenv("RLR_SCAN_SMS_RESERVED")) : 2;
    int grid = sm_count - (reserved > 0 && reserved < sm_count ? reserved : 0);
    if (static_cast<uint32_t>(grid) > n_tiles) grid = static_cast<int>(n_tiles);
    if (grid < 1) grid = 1;
    // as many stages as fit (1 KB slack for the manual 1024 B alignment); the tail of every query group uses
    // kGroupScratch bytes of the idle ring, so the ring must hold n_groups of those
    int stages = 8;
    while (stages > 2 && smem_layout(stages, q_floats, false, n_groups).total + 1024 > static_cast<uint32_t>(max_smem_optin)) --stages;
    a->grid = grid;
    a->n_stages = stages;
    a->buf_cap = 0;
    a->smem_bytes = static_cast<int>(smem_layout(stages, q_floats, false, n_groups).total + 1024);
    if (static_cast<uint32_t>(stages) * kStageBytes < n_groups * kGroupScratch) a->grid = 0;     // does not fit: the caller reports it
}

uint32_t scan_rows_per_tile(int sm_count, uint64_t n_rows)
{
    if (n_rows >= static_cast<uint64_t>(R) * sm_count) return R;
    uint64_t per = (n_rows + sm_count - 1) / sm_count;
    per = (per + 7) & ~7ull;
    if (per < 8) per = 8;
    return static_cast<uint32_t>(per > static_cast<uint64_t>(R) ? R : per);
}

void scan_plan_small(int sm_count, int max_smem_optin, uint32_t n_rows, uint32_t pitch, uint32_t rows_per_tile, ScanArgs *a)
{
    const uint32_t n_chunks = pitch / 32u;
    const uint32_t KB = (n_chunks + CH - 1) / CH;
    const uint32_t q_floats = KB * CH * 32u;
    a->half = 0;
    a->rows_per_tile = rows_per_tile;
    const uint32_t n_tiles = (n_rows + rows_per_tile - 1) / rows_per_tile;
    int grid = sm_count;                                   // nothing else runs beside a latency-path query: every SM
    if (static_cast<uint32_t>(grid) > n_tiles) grid = static_cast<int>(n_tiles);
    if (grid < 1) grid = 1;
    int stages = 8;
    while (stages > 2 && smem_layout(stages, q_floats, true).total + 1024 > static_cast<uint32_t>(max_smem_optin)) --stages;
    a->grid = grid;
    a->n_stages = stages;
    a->buf_cap = 0;
    a->smem_bytes = static_cast<int>(smem_layout(stages, q_floats, true).total + 1024);
}

cudaError_t scan_configure()
{
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
#define RLR_CFG(...) if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_topm_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin)
    RLR_CFG(false, false, 1); RLR_CFG(false, false, 2); RLR_CFG(false, false, 3);
    RLR_CFG(true, false, 1);  RLR_CFG(true, false, 2);  RLR_CFG(true, false, 3);
    RLR_CFG(false, true, 1);
#undef RLR_CFG
    return e;
}

// candidate-buffer capacity for a given m: room for m kept + a few tiles of new entries
static uint32_t pick_buf_cap(uint32_t m)
{
    uint32_t want = 2 * m > m + 2 * R ? 2 * m : m + 2 * R;
    uint32_t cap = 512;
    while (cap < want) cap <<= 1;
    if (cap > static_cast<uint32_t>(kTopBuf)) cap = kTopBuf;
    static const char *const bufcap_env = getenv("RLR_DEBUG_BUFCAP");
    if (const char *e = bufcap_env) {                       // tuning knob (power of two, >= m + R)
        const uint32_t v = static_cast<uint32_t>(atoi(e));
        if (v >= m + R && v <= static_cast<uint32_t>(kTopBuf) && (v & (v - 1)) == 0) cap = v;
    }
    return cap;
}

cudaError_t scan_launch(const ScanArgs &a, cudaStream_t stream)
{
    const uint32_t buf_cap = a.buf_cap ? a.buf_cap : pick_buf_cap(a.m);
    static const bool no_merge = getenv("RLR_DEBUG_NOMERGE") != nullptr;   // tuning knob: time the scan without final_merge
    static const bool no_global_tau = getenv("RLR_DEBUG_NOGLOBALTAU") != nullptr;
    uint32_t r_pub = (a.m + a.grid - 1) / a.grid;                    // r-th best published per CTA (0 = off)
    if (r_pub > static_cast<uint32_t>(kTopR) || a.d_pub == nullptr || no_global_tau) r_pub = 0;
    {   // one tile per CTA (latency path): a global bound cannot prune anything, its L2 round trip is pure cost
        const uint32_t rows_tile = a.rows_per_tile ? a.rows_per_tile : static_cast<uint32_t>(R);
        if (a.lat != nullptr && (a.n_rows + rows_tile - 1) / rows_tile <= static_cast<uint32_t>(a.grid)) r_pub = 0;
    }
    const uint32_t n_chunks = a.pitch / (a.half ? 64u : 32u);
    const uint32_t rpt = a.rows_per_tile ? a.rows_per_tile : static_cast<uint32_t>(R);
    const NoLat nolat = {0};
    const uint32_t nq = a.n_groups > 1 ? a.n_groups : 1;
    if (nq > static_cast<uint32_t>(kMaxQueryGroups) || a.grid <= 0) return cudaErrorInvalidValue;
    ScanGroups io;
    memset(&io, 0, sizeof io);
    if (nq == 1) {
        io.g[0].query = a.d_query; io.g[0].out = a.d_out; io.g[0].out_n = a.d_out_n; io.g[0].post = a.post;
        io.g[0].lex_rows = a.d_lex_rows; io.g[0].lex_norm = a.d_lex_norm; io.g[0].n_lex = a.n_lex;
    }
    else io = a.groups;
    if (no_merge) for (uint32_t g = 0; g < nq; ++g) io.g[g].out = nullptr;
#define RLR_SCAN_ARGS *a.tmap, io, a.n_rows, a.row_base, n_chunks, a.w_embed, a.w_lex, \
                      a.m, buf_cap, r_pub, a.n_stages, a.d_lists, a.d_counts, a.d_pub, a.d_ticket, a.d_ticket + 4, a.d_trace, rpt
    const int threads = R * static_cast<int>(nq) + 64;
    if (a.lat != nullptr && !a.half && nq == 1) {
        io.g[0].query = nullptr; io.g[0].lex_rows = nullptr; io.g[0].lex_norm = nullptr;     // they ride in the parameter block
        scan_topm_kernel<false, true, 1><<<a.grid, threads, a.smem_bytes, stream>>>(
            *a.tmap, io, a.n_rows, a.row_base, n_chunks, a.w_embed, a.w_lex,
            a.m, buf_cap, r_pub, a.n_stages, a.d_lists, a.d_counts, a.d_pub, a.d_ticket, a.d_ticket + 4, a.d_trace, rpt, *a.lat);
    } else if (a.half) {
        if (nq == 1) scan_topm_kernel<true, false, 1><<<a.grid, threads, a.smem_bytes, stream>>>(RLR_SCAN_ARGS, nolat);
        else if (nq == 2) scan_topm_kernel<true, false, 2><<<a.grid, threads, a.smem_bytes, stream>>>(RLR_SCAN_ARGS, nolat);
        else scan_topm_kernel<true, false, 3><<<a.grid, threads, a.smem_bytes, stream>>>(RLR_SCAN_ARGS, nolat);
    } else {
        if (nq == 1) scan_topm_kernel<false, false, 1><<<a.grid, threads, a.smem_bytes, stream>>>(RLR_SCAN_ARGS, nolat);
        else if (nq == 2) scan_topm_kernel<false, false, 2><<<a.grid, threads, a.smem_bytes, stream>>>(RLR_SCAN_ARGS, nolat);
        else scan_topm_kernel<false, false, 3><<<a.grid, threads, a.smem_bytes, stream>>>(RLR_SCAN_ARGS, nolat);
    }
#undef RLR_SCAN_ARGS
    return cudaGetLastError();
}

} // namespace rlr
