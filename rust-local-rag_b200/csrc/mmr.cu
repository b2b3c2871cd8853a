// mmr.cu -- kernel (4): greedy Maximal-Marginal-Relevance selection for sm_100a.
//
// Replaces RagEngine::mmr_diversify, /root/reference/src/rag_engine.rs:767-839, and the
// embedding lookup in front of it (:742-753).
//
// The reference recomputes every candidate x selected dot product in every round
// (O(k^2 * P * D), :800-804).  Because `max` over a set of identically computed dots is
// order-free, the same result is obtained from
//   (a) ONE pass that computes every pairwise dot exactly (strict left-to-right f32 sum,
//       separate mul/add roundings == dot_product :1776-1779), spread over all SMs; then
//   (b) a greedy loop run by ONE warp: the similarity triangle sits in shared memory, each
//       lane keeps its candidates' relevance / running max_sim / position in registers,
//       and every selection is a warp-level argmax (redux.sync) -- no block barrier.
// swap_remove bookkeeping (:783,:825) is reproduced with a per-candidate "position in
// `remaining`" so that exact MMR-score ties resolve to the lowest CURRENT position,
// exactly what the strict '>' scan at :812 does.
#include "common.cuh"
#include "kernels.cuh"

namespace rlr {

namespace {

constexpr int T = 16;             // pair tile edge
constexpr int KC = 768;           // floats of every row staged per pass (whole row for dim <= 768)
constexpr int PADW = KC + 4;      // +16 B: conflict-free LDS.128 across 8 rows

__device__ __forceinline__ uint32_t cand_row(const rlr_cand *cands, const uint32_t *rows, uint32_t row_base,
                                             int use_rows, uint32_t i)
{
    if (rows != nullptr) return rows[i] - row_base;
    if (use_rows) return key_row(cands[i].key) - row_base;
    return i;
}

// All P(P-1)/2 pairwise dots, one 16x16 pair tile per CTA.  The 32 candidate rows of a tile
// are staged into shared memory in ONE shot with cp.async (LDGSTS, 16 B each, L2 only): a
// single DRAM round trip instead of one per column chunk; then every thread runs its own
// strictly sequential mul/add chain over the row pair.  Row stride +16 B keeps the LDS.128
// of eight different rows on eight different bank groups.
__global__ void __launch_bounds__(T * T)
mmr_pairwise_kernel(const float *__restrict__ emb, uint32_t pitch, const rlr_cand *__restrict__ cands,
                    const uint32_t *__restrict__ rows, const uint32_t *__restrict__ d_n, uint32_t row_base,
                    int use_rows, float *__restrict__ tri)
{
    const uint32_t bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    const uint32_t p = *d_n;
    if (bi * T >= p || bj * T >= p) return;

    extern __shared__ __align__(16) float pw_smem[];     // [2*T][KC + 4]
    __shared__ const float *rowptr[2 * T];

    const uint32_t tid = threadIdx.x;
    if (tid < 2 * T) {
        const uint32_t ci = (tid < T) ? bi * T + tid : bj * T + (tid - T);
        rowptr[tid] = ci < p ? emb + static_cast<size_t>(cand_row(cands, rows, row_base, use_rows, ci)) * pitch
                             : nullptr;
    }
    __syncthreads();

    const uint32_t ti = tid >> 4, tj = tid & 15;
    const float *a_row = pw_smem + ti * PADW;
    const float *b_row = pw_smem + (T + tj) * PADW;
    float acc = 0.0f;
    for (uint32_t base = 0; base < pitch; base += KC) {
        const uint32_t ncols = (pitch - base) < KC ? (pitch - base) : KC;   // multiple of 32
        const uint32_t n4 = ncols >> 2;
        for (uint32_t idx = tid; idx < 2 * T * n4; idx += T * T) {
            const uint32_t r = idx / n4, c4 = idx - r * n4;
            float *dst = pw_smem + r * PADW + c4 * 4;
            const float *src = rowptr[r];
            if (src != nullptr) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + base + c4 * 4)
                             : "memory");
            } else {
                *reinterpret_cast<float4 *>(dst) = make_float4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll 8
        for (uint32_t d = 0; d < ncols; d += 4) {
            const float4 a = *reinterpret_cast<const float4 *>(a_row + d);
            const float4 b = *reinterpret_cast<const float4 *>(b_row + d);
            acc = add_rn(acc, mul_rn(a.x, b.x));
            acc = add_rn(acc, mul_rn(a.y, b.y));
            acc = add_rn(acc, mul_rn(a.z, b.z));
            acc = add_rn(acc, mul_rn(a.w, b.w));
        }
        __syncthreads();
    }
    const uint32_t i = bi * T + ti, j = bj * T + tj;
    if (i < j && j < p) tri[static_cast<size_t>(j) * (j - 1) / 2 + i] = acc;
}

// Greedy selection loop, run by ONE warp: candidate i lives in lane (i & 31), slot (i >> 5),
// with its relevance, running max_sim and current position in `remaining` in registers.
// Per selection: CPL shared-memory reads of the similarity triangle, CPL fmax/mul/sub, a
// local argmax, two redux.sync + ballot + shfl for the warp argmax -- no block barrier.
// The other warps of the CTA only help to stage the triangle into shared memory.
constexpr int kGreedyThreads = 256;

// The selection loop proper.  `tri` is either the shared-memory copy (LDS) or the global
// triangle; the body is branch-free per slot (dead / out-of-range slots compute on a
// clamped index and are masked out of the argmax) so that the CPL loads issue back to back.
template <int CPL>
__device__ __forceinline__ void greedy_loop(const float *tri, const rlr_cand *__restrict__ cands,
                                            const float *__restrict__ rel_opt, uint32_t p, uint32_t top_k, float lambda,
                                            uint32_t *__restrict__ sel_pos, uint32_t *__restrict__ sel_n,
                                            rlr_cand *__restrict__ result, uint32_t lane)
{
    float rel[CPL], max_sim[CPL];
    uint32_t pos[CPL];
    uint32_t alive = 0, rel_ok = 0;                        // bit s: slot s
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
        const uint32_t i = lane + 32u * s;
        rel[s] = 0.0f;
        max_sim[s] = 0.0f;                                 // fold(0.0_f32, max), :804
        pos[s] = i;
        if (i < p) {
            rel[s] = rel_opt != nullptr ? rel_opt[i] : key_score(cands[i].key);
            if (i != 0) alive |= 1u << s;
            if (is_finite_f32(rel[s])) rel_ok |= 1u << s;  // :794-797
            if (i == p - 1 && i != 0) pos[s] = 0;          // swap_remove(0), :783
        }
    }
    uint32_t n_rem = p - 1, n_sel = 1, last = 0;
    if (lane == 0) {
        sel_pos[0] = 0;
        if (result != nullptr) result[0] = cands[0];
    }
    const float one_minus = sub_rn(1.0f, lambda);          // (1.0 - diversity_factor), :808
    const uint32_t i_max = p - 1;

    while (n_sel < top_k && n_rem > 0) {                   // :788
        float sim[CPL];
        const uint32_t tl = last * (last - 1) / 2;         // row offset of `last` in the triangle (last > i)
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            uint32_t i = lane + 32u * s;
            i = i > i_max ? i_max : i;                     // clamp: masked out below
            const uint32_t idx = i < last ? tl + i : (i == last ? 0u : i * (i - 1) / 2 + last);
            sim[s] = tri[idx];
        }
        uint32_t best_hi = 0, best_lo = 0, best_i = 0;
        const uint32_t live = alive & rel_ok;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            // dead slots keep updating a max_sim nobody reads; alive slots follow :803-804
            if (is_finite_f32(sim[s]) && (alive & (1u << s))) max_sim[s] = fmaxf(max_sim[s], sim[s]);
            const float mmr = sub_rn(mul_rn(one_minus, rel[s]), mul_rn(lambda, max_sim[s])); // :808-809
            const bool ok = ((live >> s) & 1u) && is_finite_f32(mmr);                        // :794, :812
            const uint32_t hi = ok ? ord_f32(mmr) : 0u;
            const uint32_t lo = 0xffffffffu - pos[s];
            const bool better = (hi > best_hi) || (hi == best_hi && hi != 0u && lo > best_lo);
            if (better) { best_hi = hi; best_lo = lo; best_i = lane + 32u * s; }
        }
        const uint32_t mh = __reduce_max_sync(0xffffffffu, best_hi);
        if (mh == 0) break;                                // :819-822 (no finite candidate left)
        const uint32_t ml = __reduce_max_sync(0xffffffffu, best_hi == mh ? best_lo : 0u);
        const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, best_hi == mh && best_lo == ml)) - 1;
        best_i = __shfl_sync(0xffffffffu, best_i, owner);
        const uint32_t b_pos = 0xffffffffu - ml;
        // swap_remove(best_idx), :825: winner leaves, the last element moves into its slot
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            const uint32_t i = lane + 32u * s;
            const bool is_alive = (alive >> s) & 1u;
            if (is_alive && i != best_i && pos[s] == n_rem - 1) pos[s] = b_pos;
            if (i == best_i) alive &= ~(1u << s);
        }
        if (lane == 0) {
            sel_pos[n_sel] = best_i;
            if (result != nullptr) result[n_sel] = cands[best_i];
        }
        ++n_sel; --n_rem; last = best_i;
    }
    if (lane == 0) *sel_n = n_sel;
}

template <int CPL>
__global__ void __launch_bounds__(kGreedyThreads, 1)
mmr_greedy_kernel(const float *__restrict__ tri_g, const rlr_cand *__restrict__ cands,
                  const float *__restrict__ rel_opt, const uint32_t *__restrict__ d_n, uint32_t top_k,
                  float lambda, int tri_in_smem, uint32_t *__restrict__ sel_pos, uint32_t *__restrict__ sel_n,
                  rlr_cand *__restrict__ result)
{
    extern __shared__ __align__(16) float tri_s[];
    const uint32_t p = *d_n;
    const uint32_t tid = threadIdx.x;
    if (p == 0) {
        if (tid == 0) *sel_n = 0;
        return;
    }
    if (tri_in_smem) {
        const uint32_t n_tri = p * (p - 1) / 2;
        const uint32_t n4 = n_tri >> 2;
        const float4 *g4 = reinterpret_cast<const float4 *>(tri_g);
        float4 *s4 = reinterpret_cast<float4 *>(tri_s);
        for (uint32_t x = tid; x < n4; x += kGreedyThreads) s4[x] = g4[x];
        for (uint32_t x = (n4 << 2) + tid; x < n_tri; x += kGreedyThreads) tri_s[x] = tri_g[x];
        __syncthreads();
        if (tid >= 32) return;
        greedy_loop<CPL>(tri_s, cands, rel_opt, p, top_k, lambda, sel_pos, sel_n, result, tid);
    } else {
        if (tid >= 32) return;
        greedy_loop<CPL>(tri_g, cands, rel_opt, p, top_k, lambda, sel_pos, sel_n, result, tid);
    }
}

__global__ void gather_kernel(const float *__restrict__ store, uint32_t pitch, uint32_t n_rows, uint32_t row_base,
                              const rlr_cand *__restrict__ cands, const uint32_t *__restrict__ d_n,
                              float *__restrict__ out)
{
    const uint32_t i = blockIdx.x;
    const uint32_t p = *d_n;
    float4 *o = reinterpret_cast<float4 *>(out + static_cast<size_t>(i) * pitch);
    const float4 *src = nullptr;
    if (i < p && cands[i].key != 0ull) {
        const uint32_t g = key_row(cands[i].key);
        if (g >= row_base && g - row_base < n_rows)
            src = reinterpret_cast<const float4 *>(store + static_cast<size_t>(g - row_base) * pitch);
    }
    for (uint32_t c = threadIdx.x; c < pitch / 4; c += blockDim.x)
        o[c] = src != nullptr ? __ldg(src + c) : make_float4(0, 0, 0, 0);
}

__global__ void gather_rows_kernel(const float *__restrict__ store, uint32_t pitch, const uint32_t *__restrict__ rows,
                                   float *__restrict__ out, uint32_t out_pitch)
{
    const uint32_t i = blockIdx.x;
    const float *src = store + static_cast<size_t>(rows[i]) * pitch;
    float *o = out + static_cast<size_t>(i) * out_pitch;
    for (uint32_t c = threadIdx.x; c < out_pitch; c += blockDim.x) o[c] = c < pitch ? src[c] : 0.0f;
}

} // namespace

cudaError_t mmr_configure()
{
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_pairwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             2 * T * PADW * static_cast<int>(sizeof(float)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_greedy_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 4 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_greedy_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 4 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mmr_greedy_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 4 * 1024);
}

cudaError_t mmr_launch(const MmrArgs &a, cudaStream_t stream, uint32_t *launches)
{
    if (a.p_cap == 0) return cudaErrorInvalidValue;
    const uint32_t nb = (a.p_cap + T - 1) / T;
    if (a.p_cap > 1) {
        mmr_pairwise_kernel<<<dim3(nb, nb), T * T, 2 * T * PADW * sizeof(float), stream>>>(
            a.d_emb, a.pitch, a.d_cands, a.d_rows, a.d_n, a.row_base, a.use_rows, a.d_tri);
        if (launches) ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const size_t tri_bytes = static_cast<size_t>(a.p_cap) * (a.p_cap - 1) / 2 * sizeof(float);
    const size_t smem_cap = static_cast<size_t>(a.max_smem_optin) - 4 * 1024;
    const int in_smem = tri_bytes <= smem_cap;
    const size_t smem = in_smem ? tri_bytes + 16 : 0;
#define RLR_GREEDY(CPL)                                                                                               \
    mmr_greedy_kernel<CPL><<<1, kGreedyThreads, smem, stream>>>(a.d_tri, a.d_cands, a.d_rel, a.d_n, a.top_k, a.lambda, \
                                                                in_smem, a.d_sel_pos, a.d_sel_n, a.d_result)
    if (a.p_cap <= 320) RLR_GREEDY(10);
    else if (a.p_cap <= 512) RLR_GREEDY(16);
    else RLR_GREEDY(32);
#undef RLR_GREEDY
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t gather_launch(const float *d_store, uint32_t pitch, uint32_t n_rows, uint32_t row_base,
                          const rlr_cand *d_cands, const uint32_t *d_n, uint32_t p_cap, float *d_out,
                          cudaStream_t stream)
{
    if (p_cap == 0) return cudaSuccess;
    gather_kernel<<<p_cap, 128, 0, stream>>>(d_store, pitch, n_rows, row_base, d_cands, d_n, d_out);
    return cudaGetLastError();
}

cudaError_t gather_rows_launch(const float *d_store, uint32_t pitch, const uint32_t *d_rows, uint32_t n,
                               float *d_out, uint32_t out_pitch, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    gather_rows_kernel<<<n, 128, 0, stream>>>(d_store, pitch, d_rows, d_out, out_pitch);
    return cudaGetLastError();
}

} // namespace rlr
