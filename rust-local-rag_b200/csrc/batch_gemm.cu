// batch_gemm.cu -- kernel (3): batched-query scoring as a dense contraction on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA), fused with a
// per-query threshold filter so that the Q x N score matrix is never written.
//
// Not in the reference (rust-local-rag answers one query at a time, src/rag_engine.rs:470);
// this is BASELINE.json config 4 / north_star kernel (3): S[q][row] = sum_d Q[q][d]*R[row][d]
// for a batch of queries against the binary16 copy of the chunk store, f32 accumulation.
// Scores are the tensor-core result (f16 inputs, f32 accumulate in MMA order): they differ
// from the exact sequential f32 path by <= ~2e-4 absolute on unit vectors;
// rlr_search_batch can re-score the shortlist exactly.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer: per k-chunk (64 halves = one 128-byte swizzle span) loads a
//            128-row store tile (A, 16 KB) and a 256-query tile (B, 32 KB) into a 4-stage ring
//   warp 1   MMA issuer: one lane issues 4 x tcgen05.mma (M=128, N=256, K=16) per k-chunk;
//            tcgen05.commit releases the smem stage / publishes the accumulator
//   warp 2   allocates / frees the 512 TMEM columns (two 128 x 256 f32 accumulators)
//   warps 4-7 epilogue: tcgen05.ld the accumulator (thread = store row, 32 queries at a time),
//            compare with the per-query threshold tau[q] held in shared memory, and append the
//            rare survivors (rank key = ordered score | ~row) to per-query lists in global memory
// A second tiny kernel (batch_prune_kernel) merges the appended candidates into each query's
// running top-M and raises tau[q]; the host runs the corpus in geometrically growing phases
// so that the expected number of survivors per phase stays below the list capacity.
#include <cstdio>
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "sort_regs.cuh"

namespace rlr {

namespace {

constexpr int kBM = 128;                  // store rows per tile (UMMA M)
constexpr int kBN = 256;                  // queries per tile   (UMMA N)
constexpr int kBK = 64;                   // halves per k-chunk (128 bytes: the swizzle span)
constexpr int kStagesB = 4;
constexpr uint32_t kABytes = kBM * 128;   // 16 KB
constexpr uint32_t kBBytes = kBN * 128;   // 32 KB
constexpr uint32_t kStageB = kABytes + kBBytes;
constexpr int kBatchThreads = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kCntStride = 32;     // one 128-byte line per query counter: same-line L2 atomics serialise

// ---- tcgen05 wrappers ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile in shared memory, rows of exactly 128 bytes, SWIZZLE_128B (what a TMA box
// {64 halves, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces): 8-row atoms of 1024 bytes.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused: 0)
//   bits [32,46) stride byte offset >> 4 = 1024 >> 4      bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = f16, both K-major, M = 128, N = 256
//   bits [4,6) c_format = 1 (F32); [7,10) a_format = 0 (F16); [10,13) b_format = 0 (F16);
//   bit 15 / 16 a/b major = 0 (K); bits [17,23) N >> 3; bits [24,29) M >> 4
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(kBN >> 3) << 17) | (static_cast<uint32_t>(kBM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_load_2d_nohint(uint32_t dst, const void *tmap, int32_t x, int32_t y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}

// Epilogue filter for 32 accumulator columns of one store row: a branch-free pass mask first
// (the tensor pipe is only ahead of the epilogue if the common "nothing survives" case costs a
// few dozen instructions), then the rare survivors are appended to their queries' lists.
__device__ __forceinline__ void epilogue_filter(const uint32_t (&v)[32], const float *tau32, uint32_t q0, bool row_ok,
                                                uint32_t inv_row, unsigned long long *__restrict__ app_keys,
                                                uint32_t *__restrict__ app_cnt, uint32_t cap, uint32_t *__restrict__ overflow)
{
    uint32_t mask = 0;
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        const float4 t = *reinterpret_cast<const float4 *>(tau32 + i4 * 4);       // broadcast LDS.128
        mask |= (__uint_as_float(v[i4 * 4 + 0]) >= t.x ? 1u : 0u) << (i4 * 4 + 0);
        mask |= (__uint_as_float(v[i4 * 4 + 1]) >= t.y ? 1u : 0u) << (i4 * 4 + 1);
        mask |= (__uint_as_float(v[i4 * 4 + 2]) >= t.z ? 1u : 0u) << (i4 * 4 + 2);
        mask |= (__uint_as_float(v[i4 * 4 + 3]) >= t.w ? 1u : 0u) << (i4 * 4 + 3);
    }
    if (!row_ok) mask = 0;
    const uint32_t colmask = __reduce_or_sync(0xffffffffu, mask);   // columns with a survivor in ANY lane
    if (colmask == 0) return;
    // Survivors: the 32 lanes of the warp hold 32 different rows of the SAME query column, so the
    // appends of one column are aggregated into one atomicAdd per warp and land in consecutive slots.
    const uint32_t lane = threadIdx.x & 31u;
    // pass 1: one atomicAdd per (warp, column with survivors); the up-to-32 atomics are independent,
    // so they are all in flight together (in the early, dense phases a serial chain of atomic round
    // trips per thread was the whole cost)
    uint32_t ballots[32], bases[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        ballots[i] = 0;
        bases[i] = 0;
        if (((colmask >> i) & 1u) == 0) continue;              // warp-uniform: most columns have no survivor
        ballots[i] = __ballot_sync(0xffffffffu, (mask >> i) & 1u);
        if (lane == static_cast<uint32_t>(__ffs(ballots[i]) - 1))
            bases[i] = atomicAdd(app_cnt + static_cast<size_t>(q0 + i) * kCntStride, static_cast<uint32_t>(__popc(ballots[i])));
    }
    // pass 2: consecutive slots per column
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint32_t b = ballots[i];
        if (b == 0) continue;                                   // warp-uniform
        const uint32_t base = __shfl_sync(0xffffffffu, bases[i], __ffs(b) - 1);
        if ((mask >> i) & 1u) {
            const uint32_t slot = base + __popc(b & ((1u << lane) - 1u));
            if (slot < cap) app_keys[static_cast<size_t>(q0 + i) * cap + slot] = (static_cast<unsigned long long>(ord_f32(__uint_as_float(v[i]))) << 32) | inv_row;
            else *overflow = 1u;
        }
    }
}

__global__ void __launch_bounds__(kBatchThreads, 1)
batch_gemm_topm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapQ,
                       uint32_t n_rows, uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_tiles,
                       uint32_t n_k, const float *__restrict__ tau, unsigned long long *__restrict__ app_keys,
                       uint32_t *__restrict__ app_cnt, uint32_t cap, uint32_t *__restrict__ overflow,
                       unsigned long long *dbg /* dev-only cycle counters, may be null */)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t stages_addr = smem_u32(smem);
    uint8_t *ctrl = smem + kStagesB * kStageB;
    long long dbg_wait_a = 0, dbg_wait_b = 0, dbg_work = 0, dbg_t0 = clock64();
    const uint32_t full_bar = smem_u32(ctrl);                 // [kStagesB]
    const uint32_t empty_bar = full_bar + kStagesB * 8;       // [kStagesB]
    const uint32_t tfull_bar = empty_bar + kStagesB * 8;      // [2]
    const uint32_t tempty_bar = tfull_bar + 16;               // [2]
    volatile uint32_t *s_tmem = reinterpret_cast<volatile uint32_t *>(ctrl + 128);
    float *tau_s = reinterpret_cast<float *>(ctrl + 256);     // [nq_tiles * kBN]

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nq_pad = nq_tiles * kBN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmapA);
        tma_prefetch_desc(&tmapQ);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStagesB; ++s) { mbar_init(full_bar + s * 8, 1); mbar_init(empty_bar + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + a * 8, 1); mbar_init(tempty_bar + a * 8, 4); }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t *>(s_tmem))), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = tid; i < nq_pad; i += kBatchThreads) tau_s[i] = tau[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t rt = tile0 + blockIdx.x; rt < tile1; rt += gridDim.x)
                for (uint32_t qt = 0; qt < nq_tiles; ++qt)
                    for (uint32_t kc = 0; kc < n_k; ++kc) {
                        mbar_wait(empty_bar + stage * 8, phase ^ 1);
                        mbar_arrive_expect_tx(full_bar + stage * 8, kStageB);
                        const uint32_t a_dst = stages_addr + stage * kStageB;
                        tma_load_2d_nohint(a_dst, &tmapA, static_cast<int32_t>(kc * kBK), static_cast<int32_t>(rt * kBM), full_bar + stage * 8);
                        tma_load_2d_nohint(a_dst + kABytes, &tmapQ, static_cast<int32_t>(kc * kBK), static_cast<int32_t>(qt * kBN), full_bar + stage * 8);
                        if (++stage == kStagesB) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t rt = tile0 + blockIdx.x; rt < tile1; rt += gridDim.x)
                for (uint32_t qt = 0; qt < nq_tiles; ++qt) {
                    long long c0 = clock64();
                    mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1);      // epilogue has drained this accumulator
                    dbg_wait_a += clock64() - c0;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kBN;
                    for (uint32_t kc = 0; kc < n_k; ++kc) {
                        long long c1 = clock64();
                        mbar_wait(full_bar + stage * 8, phase);
                        dbg_wait_b += clock64() - c1;
                        tc_fence_after();
                        const uint32_t a_addr = stages_addr + stage * kStageB;
                        const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + kABytes);
#pragma unroll
                        for (uint32_t k = 0; k < kBK / 16; ++k)        // advance 32 bytes (>>4 == 2) inside the swizzle span
                            tc_mma_f16(d_tmem, da + 2 * k, db + 2 * k, kIdesc, (kc | k) != 0 ? 1u : 0u);
                        tc_commit(empty_bar + stage * 8);                // smem stage free once these MMAs retire
                        if (++stage == kStagesB) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull_bar + acc * 8);                      // accumulator complete
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            if (dbg != nullptr && blockIdx.x == 0) { dbg[0] = dbg_wait_a; dbg[1] = dbg_wait_b; dbg[2] = clock64() - dbg_t0; }
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue ------------------------------
        const uint32_t w = warp - 4;                                     // TMEM lanes 32w .. 32w+31
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t rt = tile0 + blockIdx.x; rt < tile1; rt += gridDim.x) {
            const uint32_t row_local = rt * kBM + w * 32 + lane;
            const bool row_ok = row_local < n_rows;
            const uint32_t inv_row = ~(row_base + row_local);
            for (uint32_t qt = 0; qt < nq_tiles; ++qt) {
                long long c0 = clock64();
                mbar_wait(tfull_bar + acc * 8, acc_phase);
                long long c1 = clock64();
                dbg_wait_a += c1 - c0;
                tc_fence_after();
#pragma unroll 1
                for (uint32_t c = 0; c < kBN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((w * 32u) << 16) + acc * kBN + c * 32, v);
                    const uint32_t q0 = qt * kBN + c * 32;
                    epilogue_filter(v, tau_s + q0, q0, row_ok, inv_row, app_keys, app_cnt, cap, overflow);
                }
                dbg_work += clock64() - c1;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + acc * 8);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    }
    if (dbg != nullptr && blockIdx.x == 0 && tid == 128) { dbg[3] = dbg_wait_a; dbg[4] = dbg_work; }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256-row x 256-query
// tile.  Each CTA stages only ITS 128 store rows (A half) and ITS 128 queries (B half) -- the
// pair's tensor cores read both halves -- so the L2->SM operand traffic per flop drops by a
// third against the 1-CTA kernel (32 KB instead of 48 KB per k-chunk per CTA for the same
// 128 x 256 x 64 of math per CTA).  The leader CTA (rank 0) issues the MMAs; every TMA load of
// either CTA completes on the leader's `full` barrier; tcgen05.commit multicasts to both CTAs'
// `empty` / `tmem_full` barriers; the peer's epilogue warps arrive remotely on the leader's
// `tmem_empty` barrier.
// ------------------------------------------------------------------------------------------
constexpr int kStages2 = 6;
constexpr uint32_t kHalfBBytes = 128 * 128;            // 128 queries x 128 B
constexpr uint32_t kStage2 = kABytes + kHalfBBytes;    // 32 KB per CTA per k-chunk
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;            // clears the CTA-rank bit of a shared::cluster address
constexpr uint32_t kIdesc2 = (1u << 4) | (static_cast<uint32_t>(kBN >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void *tmap, int32_t x, int32_t y, uint32_t leader_bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(leader_bar)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar)   // arrives on the barrier at this offset in BOTH CTAs
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBatchThreads, 1)
batch_gemm2_topm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapQ,
                        uint32_t n_rows, uint32_t row_base, uint32_t pair0, uint32_t pair1, uint32_t nq_tiles,
                        uint32_t n_k, const float *__restrict__ tau, unsigned long long *__restrict__ app_keys,
                        uint32_t *__restrict__ app_cnt, uint32_t cap, uint32_t *__restrict__ overflow,
                        unsigned long long *dbg /* dev-only cycle counters of cluster 0, may be null */)
{
    extern __shared__ uint8_t smem_raw[];
    const bool trace = dbg != nullptr && blockIdx.x < 2;
    long long dbg_a = 0, dbg_b = 0, dbg_w = 0;
    const long long dbg_t0 = clock64();
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t stages_addr = smem_u32(smem);
    uint8_t *ctrl = smem + kStages2 * kStage2;
    const uint32_t full_bar = smem_u32(ctrl);                 // [kStages2] (used in the leader)
    const uint32_t empty_bar = full_bar + kStages2 * 8;       // [kStages2] (one per CTA)
    const uint32_t tfull_bar = empty_bar + kStages2 * 8;      // [2]        (one per CTA)
    const uint32_t tempty_bar = tfull_bar + 16;               // [2]        (used in the leader)
    volatile uint32_t *s_tmem = reinterpret_cast<volatile uint32_t *>(ctrl + 160);
    float *tau_s = reinterpret_cast<float *>(ctrl + 256);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;
    const uint32_t nq_pad = nq_tiles * kBN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmapA);
        tma_prefetch_desc(&tmapQ);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages2; ++s) { mbar_init(full_bar + s * 8, 1); mbar_init(empty_bar + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + a * 8, 1); mbar_init(tempty_bar + a * 8, 8); }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t *>(s_tmem))), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = tid; i < nq_pad; i += kBatchThreads) tau_s[i] = tau[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // both CTAs' barriers + TMEM exist
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs; completion on the LEADER's full barrier) ----------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t rp = pair0 + pair; rp < pair1; rp += n_pairs_grid)
                for (uint32_t qt = 0; qt < nq_tiles; ++qt)
                    for (uint32_t kc = 0; kc < n_k; ++kc) {
                        const long long c0 = trace ? clock64() : 0;
                        mbar_wait(empty_bar + stage * 8, phase ^ 1);
                        if (trace) dbg_a += clock64() - c0;
                        const uint32_t lead_full = (full_bar + stage * 8) & kPeerMask;
                        if (rank == 0) mbar_arrive_expect_tx(full_bar + stage * 8, 2 * kStage2);   // both CTAs' bytes
                        const uint32_t a_dst = stages_addr + stage * kStage2;
                        tma_load_2d_2sm(a_dst, &tmapA, static_cast<int32_t>(kc * kBK), static_cast<int32_t>((rp * 2 + rank) * kBM), lead_full);
                        tma_load_2d_2sm(a_dst + kABytes, &tmapQ, static_cast<int32_t>(kc * kBK), static_cast<int32_t>(qt * kBN + rank * 128), lead_full);
                        if (++stage == kStages2) { stage = 0; phase ^= 1; }
                    }
            if (trace) { dbg[8 * rank + 5] = dbg_a; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t rp = pair0 + pair; rp < pair1; rp += n_pairs_grid)
                for (uint32_t qt = 0; qt < nq_tiles; ++qt) {
                    const long long c0 = trace ? clock64() : 0;
                    mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1);       // both CTAs' epilogues drained it
                    if (trace) dbg_a += clock64() - c0;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kBN;
                    for (uint32_t kc = 0; kc < n_k; ++kc) {
                        const long long c1 = trace ? clock64() : 0;
                        mbar_wait(full_bar + stage * 8, phase);
                        if (trace) dbg_b += clock64() - c1;
                        tc_fence_after();
                        const uint32_t a_addr = stages_addr + stage * kStage2;
                        const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + kABytes);
#pragma unroll
                        for (uint32_t k = 0; k < kBK / 16; ++k)
                            tc_mma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, kIdesc2, (kc | k) != 0 ? 1u : 0u);
                        tc_commit_2sm(empty_bar + stage * 8);             // frees this stage in BOTH CTAs
                        if (++stage == kStages2) { stage = 0; phase ^= 1; }
                    }
                    tc_commit_2sm(tfull_bar + acc * 8);                   // accumulator ready in BOTH CTAs
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            if (trace) { dbg[0] = dbg_a; dbg[1] = dbg_b; dbg[2] = clock64() - dbg_t0; }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (each CTA: its own 128 rows) ----------------
        const uint32_t w = warp - 4;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t rp = pair0 + pair; rp < pair1; rp += n_pairs_grid) {
            const uint32_t row_local = (rp * 2 + rank) * kBM + w * 32 + lane;
            const bool row_ok = row_local < n_rows;
            const uint32_t inv_row = ~(row_base + row_local);
            for (uint32_t qt = 0; qt < nq_tiles; ++qt) {
                const long long c0 = trace ? clock64() : 0;
                mbar_wait(tfull_bar + acc * 8, acc_phase);
                const long long c1 = trace ? clock64() : 0;
                dbg_a += c1 - c0;
                tc_fence_after();
#pragma unroll 1
                for (uint32_t c = 0; c < kBN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((w * 32u) << 16) + acc * kBN + c * 32, v);
                    const uint32_t q0 = qt * kBN + c * 32;
                    epilogue_filter(v, tau_s + q0, q0, row_ok, inv_row, app_keys, app_cnt, cap, overflow);
                }
                tc_fence_before();
                __syncwarp();
                if (trace) dbg_w += clock64() - c1;
                if (lane == 0) mbar_arrive_cluster((tempty_bar + acc * 8) & kPeerMask);   // leader's barrier
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    }
    if (trace && tid == 128) { dbg[8 * rank + 3] = dbg_a; dbg[8 * rank + 4] = dbg_w; }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // nobody touches the peer after this
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// One CTA (128 threads) per query: merge the appended candidates into the running top-M (both
// as rank keys), write the new top-M (sorted), raise tau[q] to the M-th score, clear the
// append list.  state/app sizes <= 1024 each -> at most 2048 keys, sorted in registers.
__global__ void __launch_bounds__(128, 4)
batch_prune_kernel(unsigned long long *__restrict__ state_keys, uint32_t *__restrict__ state_cnt, uint32_t m,
                   unsigned long long *__restrict__ app_keys, uint32_t *__restrict__ app_cnt, uint32_t cap,
                   float *__restrict__ tau, uint32_t nq)
{
    __shared__ uint64_t keys[kTopBuf];
    __shared__ float dummy[kTopBuf];
    const uint32_t q = blockIdx.x, t = threadIdx.x;
    if (q >= nq) return;
    const uint32_t ns = state_cnt[q];
    uint32_t na = app_cnt[static_cast<size_t>(q) * kCntStride];
    if (na > cap) na = cap;
    const uint32_t n = ns + na;
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    if (n2 < 128) n2 = 128;
    for (uint32_t i = t; i < n2; i += 128) {
        uint64_t k = 0;
        if (i < ns) k = state_keys[static_cast<size_t>(q) * m + i];
        else if (i < n) k = app_keys[static_cast<size_t>(q) * cap + (i - ns)];
        keys[i] = k;
        dummy[i] = 0.0f;
    }
    named_bar_sync(1, 128);
    bitonic_desc(keys, dummy, n2, t);
    const uint32_t keep = n < m ? n : m;
    for (uint32_t i = t; i < keep; i += 128) state_keys[static_cast<size_t>(q) * m + i] = keys[i];
    if (t == 0) {
        state_cnt[q] = keep;
        app_cnt[static_cast<size_t>(q) * kCntStride] = 0;
        if (keep >= m) tau[q] = key_score(keys[m - 1]);
    }
}

__global__ void to_half_rows_kernel(const float *__restrict__ src, uint32_t dim, __half *__restrict__ dst, uint32_t pitch,
                                    uint32_t n_valid, uint32_t n_pad)
{
    for (uint32_t r = blockIdx.x; r < n_pad; r += gridDim.x)
        for (uint32_t c = threadIdx.x; c < pitch; c += blockDim.x)
            dst[static_cast<size_t>(r) * pitch + c] = (r < n_valid && c < dim) ? __float2half_rn(src[static_cast<size_t>(r) * dim + c]) : __float2half_rn(0.0f);
}

// Exact re-score of the shortlist: one thread per (query, candidate) runs the reference's
// sequential f32 dot (src/rag_engine.rs:1776-1779) over the stored row and rewrites the rank key.
template <bool kHalf>
__global__ void batch_rescore_kernel(const void *__restrict__ rows, uint32_t pitch, uint32_t dim, uint32_t row_base,
                                     const float *__restrict__ q32, unsigned long long *__restrict__ state_keys,
                                     const uint32_t *__restrict__ state_cnt, uint32_t m, uint32_t nq)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = idx / m, i = idx - q * m;
    if (q >= nq || i >= state_cnt[q]) return;
    const unsigned long long key = state_keys[static_cast<size_t>(q) * m + i];
    const uint32_t row = key_row(key);
    const float *qv = q32 + static_cast<size_t>(q) * dim;
    float acc = 0.0f;
    if constexpr (kHalf) {
        const __half *r = static_cast<const __half *>(rows) + static_cast<size_t>(row - row_base) * pitch;
        for (uint32_t d = 0; d < dim; ++d) acc = add_rn(acc, mul_rn(qv[d], __half2float(r[d])));
    } else {
        const float *r = static_cast<const float *>(rows) + static_cast<size_t>(row - row_base) * pitch;
        for (uint32_t d = 0; d < dim; ++d) acc = add_rn(acc, mul_rn(qv[d], r[d]));
    }
    state_keys[static_cast<size_t>(q) * m + i] = make_key(acc, row);
}

__global__ void batch_init_kernel(float *tau, uint32_t *state_cnt, uint32_t *app_cnt, uint32_t nq, uint32_t nq_pad, uint32_t *overflow)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq_pad) {
        tau[i] = i < nq ? -INFINITY : INFINITY;     // padded queries never pass
        state_cnt[i] = 0;
        app_cnt[static_cast<size_t>(i) * kCntStride] = 0;
    }
    if (i == 0) *overflow = 0;
}

} // namespace

size_t batch_smem_bytes(uint32_t nq_pad) { return kStagesB * kStageB + 256 + nq_pad * sizeof(float) + 1024; }

cudaError_t batch_configure(int smem_optin)
{
    cudaError_t e = cudaFuncSetAttribute(batch_gemm_topm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(batch_gemm2_topm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
}

// 2-CTA variant: tiles are handed out in PAIRS of 128-row tiles; tile0 must be even
cudaError_t batch_gemm2_launch(const CUtensorMap *tmapA, const CUtensorMap *tmapQ128, int sm_count, uint32_t n_rows,
                               uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_pad, uint32_t pitch16,
                               const float *tau, unsigned long long *app_keys, uint32_t *app_cnt, uint32_t cap,
                               uint32_t *overflow, cudaStream_t st)
{
    if (tile1 <= tile0) return cudaSuccess;
    const uint32_t pair0 = tile0 / 2, pair1 = (tile1 + 1) / 2;
    uint32_t clusters = static_cast<uint32_t>(sm_count / 2);
    if (clusters > pair1 - pair0) clusters = pair1 - pair0;
    const size_t smem = kStages2 * kStage2 + 256 + nq_pad * sizeof(float) + 1024;
    unsigned long long *dbg = nullptr;
    if (getenv("RLR_DEBUG_BATCH_TRACE") && pair1 - pair0 > 1000) { cudaMalloc(&dbg, 128); cudaMemset(dbg, 0, 128); }
    batch_gemm2_topm_kernel<<<clusters * 2, kBatchThreads, smem, st>>>(
        *tmapA, *tmapQ128, n_rows, row_base, pair0, pair1, nq_pad / kBN, pitch16 / kBK, tau, app_keys, app_cnt, cap, overflow, dbg);
    if (dbg) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, 128, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        fprintf(stderr, "[batch2 trace pairs=%u clusters=%u] leader MMA thread: wait tmem_empty %llu, wait smem_full %llu, total %llu cycles; "
                        "epilogue warp 4 (leader/peer): wait tmem_full %llu/%llu, work %llu/%llu; producer wait smem_empty (leader/peer) %llu/%llu\n",
                pair1 - pair0, clusters, h[0], h[1], h[2], h[3], h[11], h[4], h[12], h[5], h[13]);
    }
    return cudaGetLastError();
}

cudaError_t batch_init_launch(float *tau, uint32_t *state_cnt, uint32_t *app_cnt, uint32_t nq, uint32_t nq_pad,
                              uint32_t *overflow, cudaStream_t st)
{
    batch_init_kernel<<<(nq_pad + 255) / 256, 256, 0, st>>>(tau, state_cnt, app_cnt, nq, nq_pad, overflow);
    return cudaGetLastError();
}

cudaError_t batch_queries_to_half_launch(const float *d_q, uint32_t dim, void *d_q16, uint32_t pitch16, uint32_t nq,
                                         uint32_t nq_pad, cudaStream_t st)
{
    to_half_rows_kernel<<<nq_pad < 592 ? nq_pad : 592, 256, 0, st>>>(d_q, dim, static_cast<__half *>(d_q16), pitch16, nq, nq_pad);
    return cudaGetLastError();
}

cudaError_t batch_gemm_launch(const CUtensorMap *tmapA, const CUtensorMap *tmapQ, int grid, uint32_t n_rows,
                              uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_pad, uint32_t pitch16,
                              const float *tau, unsigned long long *app_keys, uint32_t *app_cnt, uint32_t cap,
                              uint32_t *overflow, cudaStream_t st)
{
    if (tile1 <= tile0) return cudaSuccess;
    const uint32_t tiles = tile1 - tile0;
    if (static_cast<uint32_t>(grid) > tiles) grid = static_cast<int>(tiles);
    unsigned long long *dbg = nullptr;
    if (getenv("RLR_DEBUG_BATCH_TRACE") && tiles > 2000) { cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); }
    batch_gemm_topm_kernel<<<grid, kBatchThreads, batch_smem_bytes(nq_pad), st>>>(
        *tmapA, *tmapQ, n_rows, row_base, tile0, tile1, nq_pad / kBN, pitch16 / kBK, tau, app_keys, app_cnt, cap, overflow, dbg);
    if (dbg) {
        unsigned long long h[8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, 64, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        fprintf(stderr, "[batch trace tiles=%u] CTA0 MMA thread: wait tmem_empty %llu, wait smem_full %llu, total %llu cycles; "
                        "epilogue warp: wait tmem_full %llu, work %llu cycles\n", tiles, h[0], h[1], h[2], h[3], h[4]);
    }
    return cudaGetLastError();
}

cudaError_t batch_rescore_launch(const void *d_rows, int half, uint32_t pitch, uint32_t dim, uint32_t row_base,
                                 const float *d_q32, unsigned long long *state_keys, const uint32_t *state_cnt, uint32_t m,
                                 uint32_t nq, cudaStream_t st)
{
    const uint32_t total = nq * m;
    if (total == 0) return cudaSuccess;
    if (half) batch_rescore_kernel<true><<<(total + 127) / 128, 128, 0, st>>>(d_rows, pitch, dim, row_base, d_q32, state_keys, state_cnt, m, nq);
    else batch_rescore_kernel<false><<<(total + 127) / 128, 128, 0, st>>>(d_rows, pitch, dim, row_base, d_q32, state_keys, state_cnt, m, nq);
    return cudaGetLastError();
}

cudaError_t batch_prune_launch(unsigned long long *state_keys, uint32_t *state_cnt, uint32_t m, unsigned long long *app_keys,
                               uint32_t *app_cnt, uint32_t cap, float *tau, uint32_t nq, cudaStream_t st)
{
    batch_prune_kernel<<<nq, 128, 0, st>>>(state_keys, state_cnt, m, app_keys, app_cnt, cap, tau, nq);
    return cudaGetLastError();
}

} // namespace rlr
