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
//   (b) a single-CTA greedy loop that keeps the similarity triangle in shared memory,
//       a running max_sim per candidate in a register, and does a warp-level argmax
//       (redux.sync) + one block barrier per selection.
// swap_remove bookkeeping (:783,:825) is reproduced with a per-candidate "position in
// `remaining`" so that exact MMR-score ties resolve to the lowest CURRENT position,
// exactly what the strict '>' scan at :812 does.
#include "common.cuh"
#include "kernels.cuh"

namespace rlr {

namespace {

constexpr int T = 16;             // pair tile edge
constexpr int KC = 64;            // floats per staged chunk
constexpr int PADW = KC + 4;      // +16 B: conflict-free LDS.128 across 8 rows

__device__ __forceinline__ uint32_t cand_row(const rlr_cand *cands, const uint32_t *rows, uint32_t row_base,
                                             int use_rows, uint32_t i)
{
    if (rows != nullptr) return rows[i] - row_base;
    if (use_rows) return key_row(cands[i].key) - row_base;
    return i;
}

__global__ void __launch_bounds__(T * T)
mmr_pairwise_kernel(const float *__restrict__ emb, uint32_t pitch, const rlr_cand *__restrict__ cands,
                    const uint32_t *__restrict__ rows, const uint32_t *__restrict__ d_n, uint32_t row_base,
                    int use_rows, float *__restrict__ tri)
{
    const uint32_t bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    const uint32_t p = *d_n;
    if (bi * T >= p || bj * T >= p) return;

    __shared__ __align__(16) float A[T][PADW];
    __shared__ __align__(16) float B[T][PADW];
    __shared__ const float *rowptr[2 * T];

    const uint32_t tid = threadIdx.x;
    if (tid < 2 * T) {
        const uint32_t ci = (tid < T) ? bi * T + tid : bj * T + (tid - T);
        rowptr[tid] = ci < p ? emb + static_cast<size_t>(cand_row(cands, rows, row_base, use_rows, ci)) * pitch
                             : nullptr;
    }
    __syncthreads();

    const uint32_t ti = tid >> 4, tj = tid & 15;
    // two float4 per thread per chunk: (row, col4) = (idx / 16, idx % 16), idx = tid, tid + 256
    const uint32_t r0 = tid >> 4, c0 = tid & 15;
    const float *p0 = rowptr[r0];
    const float *p1 = rowptr[r0 + T];

    float acc = 0.0f;
    float4 v0, v1;
    auto fetch = [&](uint32_t base) {
        const uint32_t col = base + c0 * 4;
        v0 = (p0 != nullptr && col < pitch) ? __ldg(reinterpret_cast<const float4 *>(p0 + col)) : make_float4(0, 0, 0, 0);
        v1 = (p1 != nullptr && col < pitch) ? __ldg(reinterpret_cast<const float4 *>(p1 + col)) : make_float4(0, 0, 0, 0);
    };
    fetch(0);
    for (uint32_t base = 0; base < pitch; base += KC) {
        *reinterpret_cast<float4 *>(&A[r0][c0 * 4]) = v0;
        *reinterpret_cast<float4 *>(&B[r0][c0 * 4]) = v1;
        __syncthreads();
        if (base + KC < pitch) fetch(base + KC);
#pragma unroll
        for (int d = 0; d < KC; d += 4) {
            const float4 a = *reinterpret_cast<const float4 *>(&A[ti][d]);
            const float4 b = *reinterpret_cast<const float4 *>(&B[tj][d]);
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

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t key)
{
    const uint32_t hi = static_cast<uint32_t>(key >> 32);
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t lo = (hi == mh) ? static_cast<uint32_t>(key) : 0u;
    const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
    return (static_cast<uint64_t>(mh) << 32) | ml;
}

__global__ void __launch_bounds__(1024, 1)
mmr_greedy_kernel(const float *__restrict__ tri_g, const rlr_cand *__restrict__ cands,
                  const float *__restrict__ rel_opt, const uint32_t *__restrict__ d_n, uint32_t top_k,
                  float lambda, int tri_in_smem, uint32_t *__restrict__ sel_pos, uint32_t *__restrict__ sel_n,
                  rlr_cand *__restrict__ result)
{
    extern __shared__ float tri_s[];
    __shared__ uint64_t part_key[2][32];
    __shared__ uint32_t part_idx[2][32];
    __shared__ uint32_t s_sel[RLR_MAX_M];

    const uint32_t p = *d_n;
    const uint32_t i = threadIdx.x;
    const uint32_t warp = i >> 5;
    const uint32_t n_warps = blockDim.x >> 5;
    if (p == 0) {
        if (i == 0) *sel_n = 0;
        return;
    }
    const float *tri = tri_g;
    if (tri_in_smem) {
        const uint32_t n_tri = p * (p - 1) / 2;
        for (uint32_t x = i; x < n_tri; x += blockDim.x) tri_s[x] = tri_g[x];
        tri = tri_s;
    }
    float rel = 0.0f;
    if (i < p) rel = rel_opt != nullptr ? rel_opt[i] : key_score(cands[i].key);
    const bool rel_ok = is_finite_f32(rel);                 // :794-797
    bool alive = (i < p) && (i != 0);
    uint32_t pos = i;
    if (i == p - 1 && i != 0) pos = 0;                      // swap_remove(0), :783
    uint32_t n_rem = p - 1, n_sel = 1, last = 0;
    if (i == 0) s_sel[0] = 0;
    const float one_minus = sub_rn(1.0f, lambda);           // (1.0 - diversity_factor), :808
    float max_sim = 0.0f;                                   // fold(0.0_f32, max), :804
    uint32_t buf = 0;
    __syncthreads();

    while (n_sel < top_k && n_rem > 0) {                    // :788
        uint64_t key = 0;
        if (alive) {
            const uint32_t a = i < last ? i : last, b = i < last ? last : i;
            const float sim = tri[b * (b - 1) / 2 + a];
            if (is_finite_f32(sim)) max_sim = fmaxf(max_sim, sim);   // :803-804
            if (rel_ok) {
                const float mmr = sub_rn(mul_rn(one_minus, rel), mul_rn(lambda, max_sim)); // :808-809
                if (is_finite_f32(mmr))                                                    // :812
                    key = (static_cast<uint64_t>(ord_f32(mmr)) << 32) | (0xffffffffu - pos);
            }
        }
        const uint64_t wk = warp_max_u64(key);
        if (wk != 0) {
            if (key == wk) { part_key[buf][warp] = wk; part_idx[buf][warp] = i; }
        } else if ((i & 31) == 0) {
            part_key[buf][warp] = 0;
        }
        __syncthreads();
        uint64_t best = 0;
        uint32_t best_i = 0;
        for (uint32_t w = 0; w < n_warps; ++w) {
            const uint64_t k = part_key[buf][w];
            if (k > best) { best = k; best_i = part_idx[buf][w]; }
        }
        if (best == 0) break;                                // :819-822
        const uint32_t b_pos = 0xffffffffu - static_cast<uint32_t>(best);
        if (i == best_i) alive = false;                      // swap_remove(best_idx), :825
        else if (alive && pos == n_rem - 1) pos = b_pos;
        if (i == 0) s_sel[n_sel] = best_i;
        ++n_sel; --n_rem; last = best_i; buf ^= 1;
    }
    __syncthreads();
    if (i == 0) *sel_n = n_sel;
    for (uint32_t x = i; x < n_sel; x += blockDim.x) {
        sel_pos[x] = s_sel[x];
        if (result != nullptr) result[x] = cands[s_sel[x]];
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
    return cudaFuncSetAttribute(mmr_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 16 * 1024);
}

cudaError_t mmr_launch(const MmrArgs &a, cudaStream_t stream, uint32_t *launches)
{
    if (a.p_cap == 0) return cudaErrorInvalidValue;
    const uint32_t nb = (a.p_cap + T - 1) / T;
    if (a.p_cap > 1) {
        mmr_pairwise_kernel<<<dim3(nb, nb), T * T, 0, stream>>>(a.d_emb, a.pitch, a.d_cands, a.d_rows, a.d_n,
                                                               a.row_base, a.use_rows, a.d_tri);
        if (launches) ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const size_t tri_bytes = static_cast<size_t>(a.p_cap) * (a.p_cap - 1) / 2 * sizeof(float);
    const size_t smem_cap = static_cast<size_t>(a.max_smem_optin) - 16 * 1024; // static arrays live there too
    const int in_smem = tri_bytes <= smem_cap;
    const uint32_t threads = ((a.p_cap + 31) / 32) * 32;
    mmr_greedy_kernel<<<1, threads, in_smem ? tri_bytes : 0, stream>>>(a.d_tri, a.d_cands, a.d_rel, a.d_n, a.top_k,
                                                                      a.lambda, in_smem, a.d_sel_pos, a.d_sel_n,
                                                                      a.d_result);
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
