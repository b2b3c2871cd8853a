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
#include <cstring>

#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "mmr_device.cuh"

namespace rlr {

namespace {

constexpr int T = 16;             // pair tile edge
constexpr int KC = 768;           // floats of every row staged per pass (whole row for dim <= 768)

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
template <bool kHalf>
__global__ void __launch_bounds__(T * T)
mmr_pairwise_kernel(const void *__restrict__ emb, uint32_t pitch, const rlr_cand *__restrict__ cands,
                    const uint32_t *__restrict__ rows, const uint32_t *__restrict__ d_n, uint32_t row_base,
                    int use_rows, float *__restrict__ tri, const __grid_constant__ PeerTable peers)
{
    constexpr uint32_t ESZ = kHalf ? 2u : 4u;             // bytes per stored element
    constexpr uint32_t EPV = 16u / ESZ;                   // elements per 16-byte vector
    constexpr uint32_t ROWB = KC * ESZ + 16u;             // staged row stride in bytes (+16: bank spread)
    const uint32_t bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    const uint32_t p = *d_n;
    if (bi * T >= p || bj * T >= p) return;

    extern __shared__ __align__(16) uint8_t pw_smem[];    // [2*T][ROWB]
    __shared__ const uint8_t *rowptr[2 * T];

    const uint32_t tid = threadIdx.x;
    if (tid < 2 * T) {
        const uint32_t ci = (tid < T) ? bi * T + tid : bj * T + (tid - T);
        const uint8_t *ptr = nullptr;
        if (ci < p) {
            if (peers.n != 0) {
                // global row -> owning shard -> peer-mapped pointer (a load over NVLink if not local)
                const uint32_t g = rows != nullptr ? rows[ci] : key_row(cands[ci].key);   // explicit rows are GLOBAL here
                for (uint32_t s = 0; s < peers.n; ++s)
                    if (g >= peers.row_base[s] && g - peers.row_base[s] < peers.n_rows[s])
                        ptr = static_cast<const uint8_t *>(peers.base[s]) + static_cast<size_t>(g - peers.row_base[s]) * pitch * ESZ;
            } else {
                ptr = static_cast<const uint8_t *>(emb) + static_cast<size_t>(cand_row(cands, rows, row_base, use_rows, ci)) * pitch * ESZ;
            }
        }
        rowptr[tid] = ptr;
    }
    __syncthreads();

    const uint32_t ti = tid >> 4, tj = tid & 15;
    const uint8_t *a_row = pw_smem + ti * ROWB;
    const uint8_t *b_row = pw_smem + (T + tj) * ROWB;
    float acc = 0.0f;
    for (uint32_t base = 0; base < pitch; base += KC) {
        const uint32_t ncols = (pitch - base) < KC ? (pitch - base) : KC;   // multiple of 32 (f32) / 64 (f16)
        const uint32_t nv = ncols / EPV;
        for (uint32_t idx = tid; idx < 2 * T * nv; idx += T * T) {
            const uint32_t r = idx / nv, cv = idx - r * nv;
            uint8_t *dst = pw_smem + r * ROWB + cv * 16;
            const uint8_t *src = rowptr[r];
            if (src != nullptr) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)),
                             "l"(src + static_cast<size_t>(base) * ESZ + cv * 16)
                             : "memory");
            } else {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll 8
        for (uint32_t v = 0; v < nv; ++v) {
            if constexpr (kHalf) {
                const uint4 ra = *reinterpret_cast<const uint4 *>(a_row + v * 16);
                const uint4 rb = *reinterpret_cast<const uint4 *>(b_row + v * 16);
                const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&wa[h]));
                    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&wb[h]));
                    acc = add_rn(acc, mul_rn(a.x, b.x));
                    acc = add_rn(acc, mul_rn(a.y, b.y));
                }
            } else {
                const float4 a = *reinterpret_cast<const float4 *>(a_row + v * 16);
                const float4 b = *reinterpret_cast<const float4 *>(b_row + v * 16);
                acc = add_rn(acc, mul_rn(a.x, b.x));
                acc = add_rn(acc, mul_rn(a.y, b.y));
                acc = add_rn(acc, mul_rn(a.z, b.z));
                acc = add_rn(acc, mul_rn(a.w, b.w));
            }
        }
        __syncthreads();
    }
    const uint32_t i = bi * T + ti, j = bj * T + tj;
    if (i < j && j < p) tri[static_cast<size_t>(j) * (j - 1) / 2 + i] = acc;
}

constexpr int kGreedyThreads = 256;

template <int CPT>
__global__ void __launch_bounds__(kGreedyThreads, 1)
mmr_greedy_kernel(const float *__restrict__ tri_g, const rlr_cand *__restrict__ cands,
                  const float *__restrict__ rel_opt, const uint32_t *__restrict__ d_n, uint32_t top_k,
                  float lambda, int tri_in_smem, uint32_t *__restrict__ sel_pos, uint32_t *__restrict__ sel_n,
                  rlr_cand *__restrict__ result, unsigned long long *done_flag, unsigned long long done_seq)
{
    extern __shared__ __align__(128) float tri_s[];
    __shared__ uint64_t s_best[2][4];
    __shared__ uint32_t s_besti[2][4];
    __shared__ __align__(8) uint64_t s_mbar;
    const uint32_t p = *d_n;
    const uint32_t tid = threadIdx.x;
    if (p == 0) {
        if (tid == 0) {
            *sel_n = 0;
            if (done_flag != nullptr) { __threadfence_system(); st_release_sys_u64(done_flag, done_seq); }
        }
        return;
    }
    if (tri_in_smem) {
        // stage the triangle with bulk async copies (UBLKCP): one thread, a handful of instructions
        const uint32_t n_tri = p * (p - 1) / 2;
        const uint32_t bytes16 = (n_tri * 4u) & ~15u;
        const uint32_t bar = smem_u32(&s_mbar);
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, bytes16);
            for (uint32_t off = 0; off < bytes16; off += 32768u) {
                const uint32_t n = bytes16 - off < 32768u ? bytes16 - off : 32768u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tri_s) + off), "l"(reinterpret_cast<const uint8_t *>(tri_g) + off), "r"(n), "r"(bar)
                             : "memory");
            }
        }
        for (uint32_t x = (bytes16 >> 2) + tid; x < n_tri; x += kGreedyThreads) tri_s[x] = tri_g[x];   // < 4 floats
        if (bytes16) mbar_wait(bar, 0);
        __syncthreads();
        if (tid >= kLoopThreads) return;
        greedy_loop<CPT>(tri_s, cands, rel_opt, p, top_k, lambda, sel_pos, sel_n, result, tid, s_best, s_besti);
    } else {
        if (tid >= kLoopThreads) return;
        greedy_loop<CPT>(tri_g, cands, rel_opt, p, top_k, lambda, sel_pos, sel_n, result, tid, s_best, s_besti);
    }
    // latency path: result / sel_n live in mapped pinned host memory and the host polls this flag (thread 0 wrote
    // both, so its system-scope fence orders them before the flag)
    if (tid == 0 && done_flag != nullptr) { __threadfence_system(); st_release_sys_u64(done_flag, done_seq); }
}

// Peer-memory MMR, step 0: bring every pool row ONCE from its owning GPU's HBM (NVLink loads through the peer
// table) into a dense local matrix.  The pairwise kernel tiles the P x P triangle 16 x 16, so it reads every row
// P/16 (= 19 for P = 300) times; over NVLink that was 18.6 MB of re-reads per query (0.06 ms of the root's tail)
// against 0.9 MB here.  Raw 16-byte copies: the element type does not matter.
__global__ void __launch_bounds__(256)
gather_peers_kernel(const __grid_constant__ PeerTable peers, const rlr_cand *__restrict__ cands, const uint32_t *__restrict__ rows,
                    const uint32_t *__restrict__ d_n, uint32_t row_bytes, uint8_t *__restrict__ out)
{
    const uint32_t i = blockIdx.x;
    if (i >= *d_n) return;
    const uint32_t g = rows != nullptr ? rows[i] : key_row(cands[i].key);      // GLOBAL row
    const uint8_t *src = nullptr;
    for (uint32_t s = 0; s < peers.n; ++s)
        if (g >= peers.row_base[s] && g - peers.row_base[s] < peers.n_rows[s])
            src = static_cast<const uint8_t *>(peers.base[s]) + static_cast<size_t>(g - peers.row_base[s]) * row_bytes;
    uint4 *dst = reinterpret_cast<uint4 *>(out + static_cast<size_t>(i) * row_bytes);
    for (uint32_t v = threadIdx.x; v < row_bytes / 16; v += blockDim.x)
        dst[v] = src != nullptr ? __ldcg(reinterpret_cast<const uint4 *>(src) + v) : make_uint4(0, 0, 0, 0);
}

template <bool kHalf>
__global__ void gather_kernel(const void *__restrict__ store, uint32_t pitch, uint32_t n_rows, uint32_t row_base,
                              const rlr_cand *__restrict__ cands, const uint32_t *__restrict__ d_n,
                              float *__restrict__ out, uint32_t out_pitch)
{
    const uint32_t i = blockIdx.x;
    const uint32_t p = *d_n;
    float *o = out + static_cast<size_t>(i) * out_pitch;
    const void *src = nullptr;
    if (i < p && cands[i].key != 0ull) {
        const uint32_t g = key_row(cands[i].key);
        if (g >= row_base && g - row_base < n_rows)
            src = static_cast<const uint8_t *>(store) + static_cast<size_t>(g - row_base) * pitch * (kHalf ? 2 : 4);
    }
    for (uint32_t c = threadIdx.x; c < out_pitch; c += blockDim.x) {
        float v = 0.0f;
        if (src != nullptr && c < pitch) {
            if constexpr (kHalf) v = __half2float(static_cast<const __half *>(src)[c]);
            else v = static_cast<const float *>(src)[c];
        }
        o[c] = v;
    }
}

template <bool kHalf>
__global__ void gather_rows_kernel(const void *__restrict__ store, uint32_t pitch, const uint32_t *__restrict__ rows,
                                   float *__restrict__ out, uint32_t out_pitch)
{
    const uint32_t i = blockIdx.x;
    const uint8_t *src = static_cast<const uint8_t *>(store) + static_cast<size_t>(rows[i]) * pitch * (kHalf ? 2 : 4);
    float *o = out + static_cast<size_t>(i) * out_pitch;
    for (uint32_t c = threadIdx.x; c < out_pitch; c += blockDim.x) {
        float v = 0.0f;
        if (c < pitch) {
            if constexpr (kHalf) v = __half2float(reinterpret_cast<const __half *>(src)[c]);
            else v = reinterpret_cast<const float *>(src)[c];
        }
        o[c] = v;
    }
}

} // namespace

cudaError_t mmr_configure()
{
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_pairwise_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * T * (KC * 4 + 16));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_pairwise_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * T * (KC * 2 + 16));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_greedy_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 4 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mmr_greedy_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 4 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mmr_greedy_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 4 * 1024);
}

cudaError_t mmr_launch(const MmrArgs &a, cudaStream_t stream, uint32_t *launches)
{
    if (a.p_cap == 0) return cudaErrorInvalidValue;
    const uint32_t nb = (a.p_cap + T - 1) / T;
    if (a.p_cap > 1) {
        PeerTable pt;
        memset(&pt, 0, sizeof pt);
        const void *emb = a.d_emb;
        const uint32_t *rows = a.d_rows;
        int use_rows = a.use_rows;
        if (a.peers != nullptr && a.d_gather != nullptr) {
            // rows live on several GPUs: one gather pass over NVLink, then the pairwise kernel reads local memory
            gather_peers_kernel<<<a.p_cap, 256, 0, stream>>>(*a.peers, a.d_cands, a.d_rows, a.d_n, a.pitch * (a.half ? 2u : 4u),
                                                             static_cast<uint8_t *>(a.d_gather));
            if (launches) ++*launches;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            emb = a.d_gather; rows = nullptr; use_rows = 0;       // candidate i is row i of the gathered matrix
        } else if (a.peers != nullptr) {
            pt = *a.peers;
        }
        if (a.half)
            mmr_pairwise_kernel<true><<<dim3(nb, nb), T * T, 2 * T * (KC * 2 + 16), stream>>>(
                emb, a.pitch, a.d_cands, rows, a.d_n, a.row_base, use_rows, a.d_tri, pt);
        else
            mmr_pairwise_kernel<false><<<dim3(nb, nb), T * T, 2 * T * (KC * 4 + 16), stream>>>(
                emb, a.pitch, a.d_cands, rows, a.d_n, a.row_base, use_rows, a.d_tri, pt);
        if (launches) ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const size_t tri_bytes = static_cast<size_t>(a.p_cap) * (a.p_cap - 1) / 2 * sizeof(float);
    const size_t smem_cap = static_cast<size_t>(a.max_smem_optin) - 4 * 1024;
    const int in_smem = tri_bytes <= smem_cap;
    const size_t smem = in_smem ? tri_bytes + 128 : 0;
#define RLR_GREEDY(CPL)                                                                                               \
    mmr_greedy_kernel<CPL><<<1, kGreedyThreads, smem, stream>>>(a.d_tri, a.d_cands, a.d_rel, a.d_n, a.top_k, a.lambda, \
                                                                in_smem, a.d_sel_pos, a.d_sel_n, a.d_result, a.done_flag, a.done_seq)
    if (a.p_cap <= 384) RLR_GREEDY(3);
    else if (a.p_cap <= 512) RLR_GREEDY(4);
    else RLR_GREEDY(8);
#undef RLR_GREEDY
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t gather_launch(const void *d_store, int half, uint32_t pitch, uint32_t n_rows, uint32_t row_base,
                          const rlr_cand *d_cands, const uint32_t *d_n, uint32_t p_cap, float *d_out, uint32_t out_pitch,
                          cudaStream_t stream)
{
    if (p_cap == 0) return cudaSuccess;
    if (half) gather_kernel<true><<<p_cap, 128, 0, stream>>>(d_store, pitch, n_rows, row_base, d_cands, d_n, d_out, out_pitch);
    else gather_kernel<false><<<p_cap, 128, 0, stream>>>(d_store, pitch, n_rows, row_base, d_cands, d_n, d_out, out_pitch);
    return cudaGetLastError();
}

cudaError_t gather_rows_launch(const void *d_store, int half, uint32_t pitch, const uint32_t *d_rows, uint32_t n,
                               float *d_out, uint32_t out_pitch, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    if (half) gather_rows_kernel<true><<<n, 128, 0, stream>>>(d_store, pitch, d_rows, d_out, out_pitch);
    else gather_rows_kernel<false><<<n, 128, 0, stream>>>(d_store, pitch, d_rows, d_out, out_pitch);
    return cudaGetLastError();
}

} // namespace rlr
