// mmr_device.cuh -- the greedy MMR selection loop (RagEngine::mmr_diversify, /root/reference/src/rag_engine.rs:767-839)
// as a device function, shared by mmr_greedy_kernel (mmr.cu) and the fused small-store tail of scan_topm_kernel.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace rlr {

// Greedy selection loop, run by FOUR warps (one per SM sub-partition): candidate i lives in
// thread (i % 128), slot (i / 128), with its relevance term, running max_sim and current
// position in `remaining` in registers.  Per selection: CPT shared-memory reads of the
// similarity triangle, CPT fmax/mul/sub, a local argmax, two redux.sync + ballot + shfl for the
// warp argmax, then ONE named barrier to exchange the four warp winners through a
// double-buffered shared slot.  The remaining warps of the CTA only help to stage the triangle.
constexpr int kLoopThreads = 128;       // threads that run greedy_loop (named barrier 2)

template <int CPT>
__device__ __forceinline__ void greedy_loop(const float *tri, const rlr_cand *__restrict__ cands,
                                            const float *__restrict__ rel_opt, uint32_t p, uint32_t top_k, float lambda,
                                            uint32_t *__restrict__ sel_pos, uint32_t *__restrict__ sel_n,
                                            rlr_cand *__restrict__ result, uint32_t tid, uint64_t (*s_best)[4],
                                            uint32_t (*s_besti)[4])
{
    const uint32_t lane = tid & 31, warp = tid >> 5;
    float rel_term[CPT], max_sim[CPT];
    uint32_t pos[CPT], tri_row[CPT];
    uint32_t alive = 0, rel_ok = 0;                        // bit s: slot s
    const float one_minus = sub_rn(1.0f, lambda);          // (1.0 - diversity_factor), :808
    const uint32_t i_max = p - 1;
#pragma unroll
    for (int s = 0; s < CPT; ++s) {
        const uint32_t i = tid + kLoopThreads * s;
        const uint32_t ic = i > i_max ? i_max : i;         // out-of-range slots are never alive
        float rel = 0.0f;
        max_sim[s] = 0.0f;                                 // fold(0.0_f32, max), :804
        pos[s] = i;
        if (i < p) {
            rel = rel_opt != nullptr ? rel_opt[i] : key_score(cands[i].key);
            if (i != 0) alive |= 1u << s;
            if (is_finite_f32(rel)) rel_ok |= 1u << s;     // :794-797
            if (i == p - 1 && i != 0) pos[s] = 0;          // swap_remove(0), :783
        }
        rel_term[s] = mul_rn(one_minus, rel);              // loop invariant, same rounding as :808
        tri_row[s] = ic * (ic - 1) / 2;
    }
    uint32_t n_rem = p - 1, n_sel = 1, last = 0, buf = 0;
    if (tid == 0) {
        sel_pos[0] = 0;
        if (result != nullptr) result[0] = cands[0];
    }

    // Written with selects, not short-circuit logic: a branch per slot would serialise the
    // independent per-candidate chains of an in-order warp.
    while (n_sel < top_k && n_rem > 0) {                   // :788
        float sim[CPT];
        const uint32_t tl = last * (last - 1) / 2;         // row of `last` in the triangle (for i < last)
#pragma unroll
        for (int s = 0; s < CPT; ++s) {
            uint32_t i = tid + kLoopThreads * s;
            i = i > i_max ? i_max : i;
            const uint32_t idx = i < last ? tl + i : (i == last ? 0u : tri_row[s] + last);
            sim[s] = tri[idx];
        }
        uint64_t best = 0;
        uint32_t best_i = 0;
        const uint32_t live = alive & rel_ok;              // :794-797
#pragma unroll
        for (int s = 0; s < CPT; ++s) {
            const bool upd = is_finite_f32(sim[s]) & (((alive >> s) & 1u) != 0);      // :803
            const float ms = fmaxf(max_sim[s], sim[s]);                               // :804
            max_sim[s] = upd ? ms : max_sim[s];
            const float mmr = sub_rn(rel_term[s], mul_rn(lambda, max_sim[s]));        // :808-809
            const bool ok = (((live >> s) & 1u) != 0) & is_finite_f32(mmr);           // :812
            const uint64_t key = (static_cast<uint64_t>(ord_f32(mmr)) << 32) | (0xffffffffu - pos[s]);
            const uint64_t k = ok ? key : 0ull;
            const bool better = k > best;                  // strict '>' on (score, lowest current position)
            best = better ? k : best;
            best_i = better ? tid + kLoopThreads * s : best_i;
        }
        const uint64_t wk = warp_max_u64(best);
        const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, best == wk)) - 1;
        const uint32_t wi = __shfl_sync(0xffffffffu, best_i, owner);
        if (lane == 0) { s_best[buf][warp] = wk; s_besti[buf][warp] = wi; }
        named_bar_sync(2, kLoopThreads);
        uint64_t gk = 0;
        uint32_t gi = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint64_t k = s_best[buf][w];
            const uint32_t ii = s_besti[buf][w];
            const bool better = k > gk;
            gk = better ? k : gk;
            gi = better ? ii : gi;
        }
        buf ^= 1u;
        if (gk == 0ull) break;                             // :819-822 (no finite candidate left)
        const uint32_t b_pos = 0xffffffffu - static_cast<uint32_t>(gk);
        // swap_remove(best_idx), :825: winner leaves, the last element moves into its slot
#pragma unroll
        for (int s = 0; s < CPT; ++s) {
            const uint32_t i = tid + kLoopThreads * s;
            const bool mv = (((alive >> s) & 1u) != 0) & (i != gi) & (pos[s] == n_rem - 1);
            pos[s] = mv ? b_pos : pos[s];
            alive &= ~((i == gi ? 1u : 0u) << s);
        }
        if (tid == 0) {
            sel_pos[n_sel] = gi;
            if (result != nullptr) result[n_sel] = cands[gi];
        }
        ++n_sel; --n_rem; last = gi;
    }
    if (tid == 0) *sel_n = n_sel;
}


} // namespace rlr
