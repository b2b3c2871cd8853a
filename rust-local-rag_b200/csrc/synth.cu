// synth.cu -- synthetic unit-norm embedding rows generated on the device, plus the
// non-finite check for uploaded rows.
//
// Inputs for the benchmark configurations of BASELINE.json: the values
// come from a counter-based hash (bit-reproducible by oracle/rlr_oracle.c:orc_synth_rows)
// and every row is normalised with the reference's own arithmetic,
// /root/reference/src/rag_engine.rs:1763-1771 (sequential sum of squares, sqrt, true
// division per element), so a device-generated store is bit-identical to one the host
// would have normalised and uploaded.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace rlr {

namespace {

constexpr int kRowsPerBlock = 128;

__device__ __forceinline__ float synth_value(int kind, uint64_t seed, uint64_t centroid_seed, uint32_t n_clusters,
                                             float sigma, uint64_t row, uint32_t dim, uint32_t c)
{
    float u = hash_uniform(seed, row * dim + c);
    if (kind == RLR_SYNTH_CLUSTERED) {
        const uint64_t cl = row % n_clusters;
        const float ce = hash_uniform(centroid_seed, cl * dim + c);
        u = add_rn(ce, mul_rn(sigma, u));
    }
    return u;
}

// One thread per row (sequential arithmetic), 32-column tiles transposed through shared
// memory so that the global stores are 128-byte coalesced.
__global__ void __launch_bounds__(kRowsPerBlock)
synth_kernel(float *__restrict__ out, uint32_t pitch, __half *__restrict__ out16, uint32_t pitch16, uint32_t dim,
             uint64_t row_base, uint32_t n_rows, int kind, uint64_t seed, uint64_t centroid_seed, uint32_t n_clusters,
             float sigma)
{
    __shared__ float tile[kRowsPerBlock][33];
    const uint32_t t = threadIdx.x;
    const uint32_t r_local = blockIdx.x * kRowsPerBlock + t;
    const uint64_t row = row_base + r_local;
    const bool valid = r_local < n_rows;

    float norm_sq = 0.0f;
    if (valid)
        for (uint32_t c = 0; c < dim; ++c) {
            const float x = synth_value(kind, seed, centroid_seed, n_clusters, sigma, row, dim, c);
            norm_sq = add_rn(norm_sq, mul_rn(x, x));
        }
    const bool scale = norm_sq > 1e-20f;
    const float norm = __fsqrt_rn(norm_sq); // IEEE round-to-nearest

    const uint32_t warp = t >> 5, lane = t & 31;
    const uint32_t pmax = pitch > pitch16 ? pitch : pitch16;
    for (uint32_t c0 = 0; c0 < pmax; c0 += 32) {
        if (valid)
            for (uint32_t j = 0; j < 32; ++j) {
                const uint32_t c = c0 + j;
                float x = 0.0f;
                if (c < dim) {
                    x = synth_value(kind, seed, centroid_seed, n_clusters, sigma, row, dim, c);
                    if (scale) x = __fdiv_rn(x, norm);
                }
                tile[t][j] = x;
            }
        __syncwarp();
        // warp w writes its own 32 rows: lane = column
        for (uint32_t r = 0; r < 32; ++r) {
            const uint32_t rl = blockIdx.x * kRowsPerBlock + warp * 32 + r;
            if (rl < n_rows) {
                const float x = tile[warp * 32 + r][lane];
                if (out != nullptr && c0 < pitch) out[static_cast<size_t>(rl) * pitch + c0 + lane] = x;
                if (out16 != nullptr && c0 < pitch16)   // binary16 copy: round to nearest even
                    out16[static_cast<size_t>(rl) * pitch16 + c0 + lane] = __float2half_rn(x);
            }
        }
        __syncwarp();
    }
}

// normalize (/root/reference/src/rag_engine.rs:1763-1771) of stored rows, in place: one thread owns
// one row and sums x*x strictly left to right (mul and add rounded separately), then divides every
// element by sqrt(sum) when sum > 1e-20 -- the bits the reference's load path produces (:1678-1680).
// 32-column tiles go through shared memory so that global traffic is 128-byte coalesced.
__global__ void __launch_bounds__(kRowsPerBlock)
normalize_rows_kernel(float *__restrict__ rows, uint32_t pitch, uint32_t dim, uint64_t n_rows)
{
    __shared__ float tile[kRowsPerBlock][33];
    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint64_t r0 = static_cast<uint64_t>(blockIdx.x) * kRowsPerBlock;
    float norm_sq = 0.0f;
    for (uint32_t c0 = 0; c0 < dim; c0 += 32) {
        for (uint32_t r = 0; r < 32; ++r) {
            const uint64_t row = r0 + warp * 32 + r;
            tile[warp * 32 + r][lane] = (row < n_rows && c0 + lane < dim) ? rows[row * pitch + c0 + lane] : 0.0f;
        }
        __syncwarp();
        const uint32_t w = dim - c0 < 32 ? dim - c0 : 32;
        for (uint32_t j = 0; j < w; ++j) { const float x = tile[t][j]; norm_sq = add_rn(norm_sq, mul_rn(x, x)); }
        __syncwarp();
    }
    const bool scale = norm_sq > 1e-20f;
    const float norm = __fsqrt_rn(norm_sq);
    for (uint32_t c0 = 0; c0 < dim; c0 += 32) {
        for (uint32_t r = 0; r < 32; ++r) {
            const uint64_t row = r0 + warp * 32 + r;
            tile[warp * 32 + r][lane] = (row < n_rows && c0 + lane < dim) ? rows[row * pitch + c0 + lane] : 0.0f;
        }
        __syncwarp();
        for (uint32_t j = 0; j < 32; ++j) if (scale) tile[t][j] = __fdiv_rn(tile[t][j], norm);
        __syncwarp();
        for (uint32_t r = 0; r < 32; ++r) {
            const uint64_t row = r0 + warp * 32 + r;
            if (row < n_rows && c0 + lane < dim) rows[row * pitch + c0 + lane] = tile[warp * 32 + r][lane];
        }
        __syncwarp();
    }
}

// Same arithmetic for a FEW rows (a batch of queries): one warp per row.  The row is staged in shared
// memory with coalesced loads, lane 0 runs the sequential sum of squares (the order is what makes the
// bits the reference's), all lanes divide.
__global__ void __launch_bounds__(32)
normalize_row_per_warp_kernel(float *__restrict__ rows, uint32_t pitch, uint32_t dim)
{
    extern __shared__ float srow[];
    float *row = rows + static_cast<size_t>(blockIdx.x) * pitch;
    const uint32_t lane = threadIdx.x;
    for (uint32_t c = lane; c < dim; c += 32) srow[c] = row[c];
    __syncwarp();
    float norm_sq = 0.0f;
    if (lane == 0) {
#pragma unroll 8
        for (uint32_t c = 0; c < dim; ++c) { const float x = srow[c]; norm_sq = add_rn(norm_sq, mul_rn(x, x)); }
    }
    norm_sq = __shfl_sync(0xffffffffu, norm_sq, 0);
    if (norm_sq > 1e-20f) {
        const float norm = __fsqrt_rn(norm_sq);
        for (uint32_t c = lane; c < dim; c += 32) row[c] = __fdiv_rn(srow[c], norm);
    }
}

__global__ void finite_check_kernel(const float4 *__restrict__ v, uint64_t n4, uint32_t *flag)
{
    bool bad = false;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const float4 x = __ldg(v + i);
        bad |= !is_finite_f32(x.x) | !is_finite_f32(x.y) | !is_finite_f32(x.z) | !is_finite_f32(x.w);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

} // namespace

cudaError_t synth_launch(float *d_rows, uint32_t pitch, void *d_rows16, uint32_t pitch16, uint32_t dim, uint64_t row_base,
                         uint32_t n_rows, int kind, uint64_t seed, uint64_t centroid_seed, uint32_t n_clusters,
                         float sigma, cudaStream_t stream)
{
    if (n_rows == 0) return cudaSuccess;
    const uint32_t blocks = (n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    synth_kernel<<<blocks, kRowsPerBlock, 0, stream>>>(d_rows, d_rows ? pitch : 0, static_cast<__half *>(d_rows16),
                                                       d_rows16 ? pitch16 : 0, dim, row_base, n_rows, kind, seed,
                                                       centroid_seed, n_clusters ? n_clusters : 1, sigma);
    return cudaGetLastError();
}

namespace {
// f32 rows -> zero-padded binary16 rows (round to nearest even), one warp per row segment
__global__ void to_half_kernel(const float *__restrict__ src, uint32_t src_pitch, __half *__restrict__ dst,
                               uint32_t dst_pitch, uint32_t dim, uint64_t n_rows)
{
    for (uint64_t r = blockIdx.x; r < n_rows; r += gridDim.x)
        for (uint32_t c = threadIdx.x; c < dst_pitch; c += blockDim.x)
            dst[r * dst_pitch + c] = c < dim ? __float2half_rn(src[r * src_pitch + c]) : __float2half_rn(0.0f);
}
} // namespace

namespace {
__global__ void move_rows_kernel(uint8_t *rows, uint32_t row_bytes, const uint32_t *__restrict__ from,
                                 const uint32_t *__restrict__ to, uint32_t n)
{
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(rows + static_cast<size_t>(from[i]) * row_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(rows + static_cast<size_t>(to[i]) * row_bytes);
        for (uint32_t c = threadIdx.x; c < row_bytes / 16; c += blockDim.x) dst[c] = src[c];
    }
}
} // namespace

cudaError_t move_rows_launch(void *d_rows, uint32_t row_bytes, const uint32_t *d_from, const uint32_t *d_to, uint32_t n,
                             cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    move_rows_kernel<<<n < 4096 ? n : 4096, 128, 0, stream>>>(static_cast<uint8_t *>(d_rows), row_bytes, d_from, d_to, n);
    return cudaGetLastError();
}

cudaError_t to_half_launch(const float *d_src, uint32_t src_pitch, void *d_dst, uint32_t dst_pitch, uint32_t dim,
                           uint64_t n_rows, cudaStream_t stream)
{
    if (n_rows == 0) return cudaSuccess;
    const uint32_t blocks = static_cast<uint32_t>(n_rows < 148 * 32 ? n_rows : 148 * 32);
    to_half_kernel<<<blocks, 256, 0, stream>>>(d_src, src_pitch, static_cast<__half *>(d_dst), dst_pitch, dim, n_rows);
    return cudaGetLastError();
}

cudaError_t normalize_rows_launch(float *d_rows, uint32_t pitch, uint32_t dim, uint64_t n_rows, cudaStream_t stream)
{
    if (n_rows == 0) return cudaSuccess;
    if (n_rows <= 16384 && dim <= 8192) {      // a query batch: one warp per row keeps every SM busy
        normalize_row_per_warp_kernel<<<static_cast<unsigned>(n_rows), 32, dim * sizeof(float), stream>>>(d_rows, pitch, dim);
        return cudaGetLastError();
    }
    const uint64_t blocks = (n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
    normalize_rows_kernel<<<static_cast<unsigned>(blocks), kRowsPerBlock, 0, stream>>>(d_rows, pitch, dim, n_rows);
    return cudaGetLastError();
}

cudaError_t finite_check_launch(const float *d_rows, uint64_t n_floats, uint32_t *d_flag, cudaStream_t stream)
{
    if (n_floats == 0) return cudaSuccess;
    finite_check_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const float4 *>(d_rows), n_floats / 4, d_flag);
    return cudaGetLastError();
}

} // namespace rlr
