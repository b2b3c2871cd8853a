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

#include <cuda_bf16.h>
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
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
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
// instruction descriptor: D = f32, A and B of one input type, both K-major, M = 128 (256 for the CTA pair), N = 256
//   bits [4,6) c_format = 1 (F32); [7,10) a_format, [10,13) b_format: kind::f16 -> 0 (F16) / 1 (BF16),
//   kind::tf32 -> 2 (TF32); bit 15 / 16 a/b major = 0 (K); bits [17,23) N >> 3; bits [24,29) M >> 4
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(kBN >> 3) << 17) | (static_cast<uint32_t>(kBM >> 4) << 24);
// operand precision of the contraction (host + kernels): what the 128-byte k-chunks hold
//   0: binary16 (64 per chunk, kind::f16)   1: bfloat16 (64 per chunk, kind::f16)   2: tf32 = the f32 store itself
//   (32 per chunk, kind::tf32: the tensor core reads the upper 19 bits of every f32, K = 8 per instruction)
__host__ __device__ constexpr uint32_t idesc_fmt(int prec) { return (static_cast<uint32_t>(prec) << 7) | (static_cast<uint32_t>(prec) << 10); }

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

template <bool kTf32>
__global__ void __launch_bounds__(kBatchThreads, 1)
batch_gemm_topm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapQ,
                       uint32_t n_rows, uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_tiles,
                       uint32_t n_k, uint32_t idesc, const float *__restrict__ tau, unsigned long long *__restrict__ app_keys,
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
                        tma_load_2d_nohint(a_dst, &tmapA, static_cast<int32_t>(kc * (kTf32 ? 32 : kBK)), static_cast<int32_t>(rt * kBM), full_bar + stage * 8);
                        tma_load_2d_nohint(a_dst + kABytes, &tmapQ, static_cast<int32_t>(kc * (kTf32 ? 32 : kBK)), static_cast<int32_t>(qt * kBN), full_bar + stage * 8);
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
                        for (uint32_t k = 0; k < 4; ++k) {             // 4 x 32 bytes (>>4 == 2) inside the swizzle span: K = 16 halves / 8 tf32
                            if constexpr (kTf32) tc_mma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                            else tc_mma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                        }
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


// ---- staged epilogue (2-CTA kernel) ----
// TMEM loads split from their wait so that the next 32 columns are in flight while the current
// ones are filtered.
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
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
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr uint32_t kStageCap = 128;      // staged survivors per epilogue warp (keys 1 KB + queries 512 B)
constexpr uint32_t kStageBytesPerWarp = kStageCap * 12 + 16;   // + the warp's record counter

// Survivors are first parked in a warp-private shared-memory buffer (no global traffic while the
// accumulator is being drained) and appended to their queries' global lists in batches: one
// atomicAdd per record, 32 of them in flight per warp, so a batch costs ONE L2 atomic round trip.
// (Appending per 32-column chunk cost a round trip per chunk and made the epilogue, not the tensor
// pipe, the bound of the early phases: 25.7k cycles per tile against 8.2k of MMA.)
__device__ __forceinline__ void sts_u64(uint32_t addr, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// Append the `cnt` staged records (keys at keys_s, their queries at qs_s: shared-space addresses) to
// the per-query global lists.  Records of one query sit next to each other, so one atomicAdd per
// (batch of 32, query); the up-to-32 atomics of a batch are in flight together.  Everything is
// passed by value: state that lives in a struct handed to a non-inlined function ends up in local
// memory, and the per-column bookkeeping then runs at local-memory latency (measured: 4.8k cycles
// per chunk with a survivor).
__device__ __noinline__ void flush_staged(uint32_t keys_s, uint32_t qs_s, uint32_t cnt, unsigned long long *__restrict__ app_keys,
                                          uint32_t *__restrict__ app_cnt, uint32_t cap, uint32_t *__restrict__ overflow)
{
    const uint32_t lane = threadIdx.x & 31u;
    __syncwarp();
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool live = j < cnt;
        const uint32_t q = live ? lds_u32(qs_s + j * 4) : 0xffffffffu;
        const uint32_t same = __match_any_sync(0xffffffffu, q);
        const uint32_t leader = __ffs(same) - 1;
        uint32_t base = 0;
        if (live && lane == leader) base = atomicAdd(app_cnt + static_cast<size_t>(q) * kCntStride, static_cast<uint32_t>(__popc(same)));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live) {
            const uint32_t slot = base + __popc(same & ((1u << lane) - 1u));
            if (slot < cap) app_keys[static_cast<size_t>(q) * cap + slot] = lds_u64(keys_s + j * 8);
            else *overflow = 1u;
        }
    }
    __syncwarp();
}

// v[i] for a run-time i without spilling the array to local memory: a 5-level select tree (31 SELs)
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], uint32_t i)
{
    uint32_t a[16], b[8], c[4];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (i & 16u) ? v[16 + j] : v[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (i & 8u) ? a[8 + j] : a[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (i & 4u) ? b[4 + j] : b[j];
    const uint32_t d0 = (i & 2u) ? c[2] : c[0], d1 = (i & 2u) ? c[3] : c[1];
    return (i & 1u) ? d1 : d0;
}

// 32 accumulator columns (queries q0 .. q0+31) of this lane's store row.  The number of staged
// records of the warp lives in shared memory at cnt_s (lanes with survivors claim their slots with
// one shared atomicAdd each).
__device__ __forceinline__ void filter_stage(const uint32_t (&v)[32], const float *tau32, uint32_t q0, bool row_ok, uint32_t inv_row,
                                             uint32_t lane, uint32_t keys_s, uint32_t qs_s, uint32_t cnt_s,
                                             unsigned long long *__restrict__ app_keys, uint32_t *__restrict__ app_cnt, uint32_t cap,
                                             uint32_t *__restrict__ overflow, bool trace, long long &t_slow, uint32_t &n_slow)
{
    // fast path: d = v - tau >= +0  <=>  v >= tau; AND the sign bits of the 32 differences
    uint32_t sgn = 0xffffffffu;
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        const float4 t = *reinterpret_cast<const float4 *>(tau32 + i4 * 4);       // broadcast LDS.128
        sgn &= __float_as_uint(sub_rn(__uint_as_float(v[i4 * 4 + 0]), t.x)) & __float_as_uint(sub_rn(__uint_as_float(v[i4 * 4 + 1]), t.y));
        sgn &= __float_as_uint(sub_rn(__uint_as_float(v[i4 * 4 + 2]), t.z)) & __float_as_uint(sub_rn(__uint_as_float(v[i4 * 4 + 3]), t.w));
    }
    const bool mine = row_ok && static_cast<int32_t>(sgn) >= 0;
    if (!__any_sync(0xffffffffu, mine)) return;
    // slow path (some lane has a survivor).  Kept short and branch-light: a first version walked the 32
    // columns with a warp-uniform branch + ballot each and cost ~4k cycles per call (instruction-cache
    // misses on rarely executed, jumpy code); a prefix-sum version with 32 predicated blocks still ~1.9k.
    const long long s0 = trace ? clock64() : 0;
    uint32_t mask = 0;
    if (mine) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(v[i]) >= tau32[i] ? 1u : 0u) << i;
    }
    const uint32_t total = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(__popc(mask)));
    uint32_t n_groups = 1;
    if (lds_u32(cnt_s) + total > kStageCap) {
        flush_staged(keys_s, qs_s, lds_u32(cnt_s), app_keys, app_cnt, cap, overflow);
        if (lane == 0) sts_u32(cnt_s, 0);
        __syncwarp();
        if (total > kStageCap) n_groups = 8;      // dense: 4 lanes (<= 128 records) at a time
    }
#pragma unroll 1
    for (uint32_t g = 0; g < n_groups; ++g) {
        uint32_t m = (n_groups == 1 || (lane >> 2) == g) ? mask : 0u;
        if (m) {
            uint32_t pos;
            const uint32_t n = __popc(m);
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(pos) : "r"(cnt_s), "r"(n) : "memory");
            while (m) {
                const uint32_t i = __ffs(m) - 1;
                m &= m - 1;
                sts_u64(keys_s + pos * 8, (static_cast<unsigned long long>(ord_f32(__uint_as_float(pick32(v, i)))) << 32) | inv_row);
                sts_u32(qs_s + pos * 4, q0 + i);
                ++pos;
            }
        }
        if (n_groups != 1) {
            __syncwarp();
            flush_staged(keys_s, qs_s, lds_u32(cnt_s), app_keys, app_cnt, cap, overflow);
            if (lane == 0) sts_u32(cnt_s, 0);
            __syncwarp();
        }
    }
    __syncwarp();
    if (trace) { t_slow += clock64() - s0; ++n_slow; }
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
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
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

constexpr int kBatch2Threads = 384;     // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue (two per TMEM lane quadrant)
constexpr uint32_t kEpiWarps2 = 8;

template <bool kTf32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBatch2Threads, 1)
batch_gemm2_topm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapQ,
                        uint32_t n_rows, uint32_t row_base, uint32_t pair0, uint32_t pair1, uint32_t nq_tiles,
                        uint32_t n_k, uint32_t idesc, const float *__restrict__ tau, unsigned long long *__restrict__ app_keys,
                        uint32_t *__restrict__ app_cnt, uint32_t cap, uint32_t *__restrict__ overflow, uint32_t dense,
                        unsigned long long *dbg /* dev-only cycle counters of cluster 0, may be null */)
{
    // Work items are (row pair, query tile) flattened, handed out round-robin to the clusters, so
    // that the early, short phases (4 / 28 row pairs) still spread over all 74 clusters.
    // dense != 0: the very first phase (tau = -inf, every row is kept, <= cap rows): no filter and no
    // atomics at all -- row r of the phase goes to slot r of every query's list.
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
    const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint32_t n_items = (pair1 - pair0) * nq_tiles;
    const uint32_t nq_pad = nq_tiles * kBN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmapA);
        tma_prefetch_desc(&tmapQ);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages2; ++s) { mbar_init(full_bar + s * 8, 1); mbar_init(empty_bar + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + a * 8, 1); mbar_init(tempty_bar + a * 8, 2 * kEpiWarps2); }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t *>(s_tmem))), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = tid; i < nq_pad; i += kBatch2Threads) tau_s[i] = tau[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                        // both CTAs' barriers + TMEM exist
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs; completion on the LEADER's full barrier) ----------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t it = cluster_id; it < n_items; it += n_clusters) {
                const uint32_t rp = pair0 + it / nq_tiles, qt = it % nq_tiles;
                    for (uint32_t kc = 0; kc < n_k; ++kc) {
                        const long long c0 = trace ? clock64() : 0;
                        mbar_wait(empty_bar + stage * 8, phase ^ 1);
                        if (trace) dbg_a += clock64() - c0;
                        const uint32_t lead_full = (full_bar + stage * 8) & kPeerMask;
                        if (rank == 0) mbar_arrive_expect_tx(full_bar + stage * 8, 2 * kStage2);   // both CTAs' bytes
                        const uint32_t a_dst = stages_addr + stage * kStage2;
                        tma_load_2d_2sm(a_dst, &tmapA, static_cast<int32_t>(kc * (kTf32 ? 32 : kBK)), static_cast<int32_t>((rp * 2 + rank) * kBM), lead_full);
                        tma_load_2d_2sm(a_dst + kABytes, &tmapQ, static_cast<int32_t>(kc * (kTf32 ? 32 : kBK)), static_cast<int32_t>(qt * kBN + rank * 128), lead_full);
                        if (++stage == kStages2) { stage = 0; phase ^= 1; }
                    }
            }
            if (trace) { dbg[8 * rank + 5] = dbg_a; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
                for (uint32_t it = cluster_id; it < n_items; it += n_clusters) {
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
                        for (uint32_t k = 0; k < 4; ++k) {             // 4 x 32 bytes of the swizzle span: K = 16 halves / 8 tf32 each
                            if constexpr (kTf32) tc_mma_tf32_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                            else tc_mma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                        }
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
        // ---------------- epilogue (each CTA: its own 128 rows; 2 warps per TMEM lane quadrant) ----------------
        const uint32_t w = warp - 4, quad = w & 3u, half = w >> 2;       // columns [half * 128, half * 128 + 128)
        const uint32_t keys_s = smem_u32(reinterpret_cast<uint8_t *>(tau_s + nq_pad) + w * kStageBytesPerWarp);
        const uint32_t qs_s = keys_s + kStageCap * 8;
        const uint32_t cnt_s = qs_s + kStageCap * 4;
        if (lane == 0) sts_u32(cnt_s, 0);
        __syncwarp();
        uint32_t n_slow = 0;
        long long t_slow = 0;
        const bool tr = trace && w == 0;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t it = cluster_id; it < n_items; it += n_clusters) {
            const uint32_t rp = pair0 + it / nq_tiles, qt = it % nq_tiles;
            const uint32_t row_local = (rp * 2 + rank) * kBM + quad * 32 + lane;
            const bool row_ok = row_local < n_rows;
            const uint32_t inv_row = ~(row_base + row_local);
            {
                const long long c0 = trace ? clock64() : 0;
                mbar_wait(tfull_bar + acc * 8, acc_phase);
                const long long c1 = trace ? clock64() : 0;
                dbg_a += c1 - c0;
                tc_fence_after();
                const uint32_t t_base = tmem_base + ((quad * 32u) << 16) + acc * kBN + half * 128;
                const uint32_t q_base = qt * kBN + half * 128;
                uint32_t va[32], vb[32];
                if (dense) {
                    const uint32_t slot = row_local - pair0 * 2 * kBM;            // < cap by construction of phase 0
#pragma unroll 1
                    for (uint32_t c = 0; c < 4; ++c) {
                        tmem_ld32_nowait(t_base + c * 32, va);
                        tmem_wait_ld();
                        if (row_ok) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)                          // a warp writes 32 consecutive slots of one query
                                app_keys[static_cast<size_t>(q_base + c * 32 + i) * cap + slot] =
                                    (static_cast<unsigned long long>(ord_f32(__uint_as_float(va[i]))) << 32) | inv_row;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster((tempty_bar + acc * 8) & kPeerMask);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                    continue;
                }
                tmem_ld32_nowait(t_base, va);
                tmem_wait_ld();
#pragma unroll 1
                for (uint32_t cc = 0; cc < 2; ++cc) {       // two copies of the filter code, not four
                    const uint32_t qa = q_base + cc * 64;
                    tmem_ld32_nowait(t_base + cc * 64 + 32, vb);
                    filter_stage(va, tau_s + qa, qa, row_ok, inv_row, lane, keys_s, qs_s, cnt_s, app_keys, app_cnt, cap, overflow, tr, t_slow, n_slow);
                    tmem_wait_ld();
                    if (cc == 0) {
                        tmem_ld32_nowait(t_base + 64, va);
                    } else {
                        // every column of this warp's half is in registers: hand the accumulator back first
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster((tempty_bar + acc * 8) & kPeerMask);   // leader's barrier
                    }
                    filter_stage(vb, tau_s + qa + 32, qa + 32, row_ok, inv_row, lane, keys_s, qs_s, cnt_s, app_keys, app_cnt, cap, overflow, tr, t_slow, n_slow);
                    if (cc == 0) tmem_wait_ld();
                }
                if (trace) dbg_w += clock64() - c1;
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
        flush_staged(keys_s, qs_s, lds_u32(cnt_s), app_keys, app_cnt, cap, overflow);
        if (tr && lane == 0) { dbg[8 * rank + 6] = 0; dbg[8 * rank + 7] = t_slow; dbg[16 + 2 * rank] = n_slow; dbg[17 + 2 * rank] = 0; }
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

// One CTA (128 threads) per query: fold the appended candidates into the running top-M (both as
// rank keys), write the new top-M (sorted), raise tau[q] to the M-th score, clear the append list.
// Sorting all n = M + appended keys (55 bitonic stages for 1024) made the five prunes of a batch
// cost 13 % of it.  Only the best M are needed, so: (1) radix-select the M-th largest key T exactly
// -- 64 rounds of "how many keys are >= T | bit", keys in registers, one barrier per round --
// (2) compact the keys >= T (keys are unique: exactly min(M, n) of them), (3) sort just those.
constexpr int kPruneKeysPerThread = kTopBuf / 128;     // 16
__global__ void __launch_bounds__(128, 4)
batch_prune_kernel(unsigned long long *__restrict__ state_keys, uint32_t *__restrict__ state_cnt, uint32_t m,
                   unsigned long long *__restrict__ app_keys, uint32_t *__restrict__ app_cnt, uint32_t cap,
                   float *__restrict__ tau, uint32_t nq)
{
    __shared__ uint64_t keys[kTopBuf];
    __shared__ float dummy[kTopBuf];
    __shared__ uint32_t s_part[2][4];
    __shared__ uint32_t s_sel;
    const uint32_t q = blockIdx.x, t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (q >= nq) return;
    const uint32_t ns = state_cnt[q];
    uint32_t na = app_cnt[static_cast<size_t>(q) * kCntStride];
    if (na > cap) na = cap;
    const uint32_t n = ns + na;
    const uint32_t keep = n < m ? n : m;
    // this thread's keys: positions t, t + 128, ...
    uint64_t mine[kPruneKeysPerThread];
#pragma unroll
    for (int j = 0; j < kPruneKeysPerThread; ++j) {
        const uint32_t i = t + j * 128;
        uint64_t k = 0;
        if (i < ns) k = state_keys[static_cast<size_t>(q) * m + i];
        else if (i < n) k = app_keys[static_cast<size_t>(q) * cap + (i - ns)];
        mine[j] = k;
    }
    if (t == 0) s_sel = 0;
    // (1) T = the largest value with count(keys >= T) >= keep  ==  the keep-th largest key (keys are unique)
    uint64_t T = 0;
    if (n > m) {
#pragma unroll 1
        for (int bit = 63; bit >= 0; --bit) {
            const uint64_t cand = T | (1ull << bit);
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < kPruneKeysPerThread; ++j) c += mine[j] >= cand ? 1u : 0u;
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0) s_part[bit & 1][warp] = c;
            named_bar_sync(1, 128);
            const uint32_t tot = s_part[bit & 1][0] + s_part[bit & 1][1] + s_part[bit & 1][2] + s_part[bit & 1][3];
            if (tot >= m) T = cand;
        }
    } else {
        T = 1;                                        // keep every (non-zero) key
        named_bar_sync(1, 128);
    }
    // (2) compact the selected keys (order is irrelevant: they are sorted next)
    uint32_t n2 = 128;
    while (n2 < keep) n2 <<= 1;
    for (uint32_t i = t; i < n2; i += 128) { keys[i] = 0; dummy[i] = 0.0f; }
    named_bar_sync(1, 128);
#pragma unroll
    for (int j = 0; j < kPruneKeysPerThread; ++j)
        if (mine[j] >= T) keys[atomicAdd(&s_sel, 1u)] = mine[j];
    named_bar_sync(1, 128);
    // (3) sort the <= M survivors
    bitonic_desc(keys, dummy, n2, t);
    for (uint32_t i = t; i < keep; i += 128) state_keys[static_cast<size_t>(q) * m + i] = keys[i];
    if (t == 0) {
        state_cnt[q] = keep;
        app_cnt[static_cast<size_t>(q) * kCntStride] = 0;
        if (keep >= m) tau[q] = key_score(keys[m - 1]);
    }
}

// Warp-per-query variant of the prune for the common shapes (M <= 128, M + cap <= 32 * KPL): the
// keys of one query live in one warp's registers, so the 64 select rounds need no block barrier
// (the CTA version spends most of its time in them), and 4 queries share a CTA.
template <int KPL>
__global__ void __launch_bounds__(128)
batch_prune_warp_kernel(unsigned long long *__restrict__ state_keys, uint32_t *__restrict__ state_cnt, uint32_t m,
                        unsigned long long *__restrict__ app_keys, uint32_t *__restrict__ app_cnt, uint32_t cap,
                        float *__restrict__ tau, uint32_t nq)
{
    __shared__ uint64_t s_keys[4][128];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * 4 + warp;
    if (q >= nq) return;
    const uint32_t ns = state_cnt[q];
    uint32_t na = app_cnt[static_cast<size_t>(q) * kCntStride];
    if (na > cap) na = cap;
    const uint32_t n = ns + na;
    const uint32_t keep = n < m ? n : m;
    uint64_t mine[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const uint32_t i = lane + j * 32;
        uint64_t k = 0;
        if (i < ns) k = state_keys[static_cast<size_t>(q) * m + i];
        else if (i < n) k = app_keys[static_cast<size_t>(q) * cap + (i - ns)];
        mine[j] = k;
    }
    uint64_t T = 1;                                   // n <= m: keep every (non-zero) key
    if (n > m) {
        // Select on the score half of the keys first (32-bit compares: the kernel is instruction bound,
        // 1.7 warps per scheduler); the row half only matters when equal scores straddle the cut.
        uint32_t mxh = 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) { const uint32_t h = static_cast<uint32_t>(mine[j] >> 32); mxh = h > mxh ? h : mxh; }
        mxh = __reduce_max_sync(0xffffffffu, mxh);
        uint32_t Th = 0;
#pragma unroll 1
        for (int bit = 31 - __clz(static_cast<int>(mxh | 1u)); bit >= 0; --bit) {   // bits above the top set bit are zero everywhere
            const uint32_t cand = Th | (1u << bit);
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < KPL; ++j) c += static_cast<uint32_t>(mine[j] >> 32) >= cand ? 1u : 0u;
            if (__reduce_add_sync(0xffffffffu, c) >= m) Th = cand;
        }
        uint32_t c_ge = 0;
#pragma unroll
        for (int j = 0; j < KPL; ++j) c_ge += static_cast<uint32_t>(mine[j] >> 32) >= Th ? 1u : 0u;
        c_ge = __reduce_add_sync(0xffffffffu, c_ge);
        T = static_cast<uint64_t>(Th) << 32;
        if (c_ge != m) {                               // equal scores at the cut: order by the row half
#pragma unroll 1
            for (int bit = 31; bit >= 0; --bit) {
                const uint64_t cand = T | (1ull << bit);
                uint32_t c = 0;
#pragma unroll
                for (int j = 0; j < KPL; ++j) c += mine[j] >= cand ? 1u : 0u;
                if (__reduce_add_sync(0xffffffffu, c) >= m) T = cand;
            }
        }
    }
    // compact the selected keys into this warp's 128 slots, then sort them (4 keys per lane)
    uint64_t *sk = s_keys[warp];
    for (uint32_t i = lane; i < 128; i += 32) sk[i] = 0;
    __syncwarp();
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) cnt += mine[j] >= T ? 1u : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += x;
    }
    uint32_t pos = incl - cnt;
#pragma unroll
    for (int j = 0; j < KPL; ++j)
        if (mine[j] >= T) sk[pos++] = mine[j];
    __syncwarp();
    for (uint32_t k = 2; k <= 128; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (uint32_t e = 0; e < 2; ++e) {
                const uint32_t i = lane + e * 32;                         // 64 compare-exchanges per stage
                const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const uint64_t a = sk[lo], b = sk[hi];
                const bool desc = (lo & k) == 0;
                if ((a < b) == desc) { sk[lo] = b; sk[hi] = a; }
            }
            __syncwarp();
        }
    for (uint32_t i = lane; i < keep; i += 32) state_keys[static_cast<size_t>(q) * m + i] = sk[i];
    if (lane == 0) {
        state_cnt[q] = keep;
        app_cnt[static_cast<size_t>(q) * kCntStride] = 0;
        if (keep >= m) tau[q] = key_score(sk[m - 1]);
    }
}

// f32 queries -> zero-padded operand rows of the contraction's input type (prec 0: binary16, 1: bfloat16,
// both round to nearest even; 2: f32 kept as it is -- kind::tf32 reads the upper 19 bits); also the NaN/Inf
// check of the (normalised) queries
template <int kPrec>
__global__ void to_operand_rows_kernel(const float *__restrict__ src, uint32_t dim, void *__restrict__ dst_v, uint32_t pitch,
                                       uint32_t n_valid, uint32_t n_pad, uint32_t *__restrict__ nonfinite)
{
    bool bad = false;
    for (uint32_t r = blockIdx.x; r < n_pad; r += gridDim.x)
        for (uint32_t c = threadIdx.x; c < pitch; c += blockDim.x) {
            float x = 0.0f;
            if (r < n_valid && c < dim) { x = src[static_cast<size_t>(r) * dim + c]; bad |= !is_finite_f32(x); }
            const size_t o = static_cast<size_t>(r) * pitch + c;
            if constexpr (kPrec == 0) static_cast<__half *>(dst_v)[o] = __float2half_rn(x);
            else if constexpr (kPrec == 1) static_cast<__nv_bfloat16 *>(dst_v)[o] = __float2bfloat16_rn(x);
            else static_cast<float *>(dst_v)[o] = x;
        }
    if (bad && nonfinite != nullptr) atomicOr(nonfinite, 1u);
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

__global__ void batch_set_cnt_kernel(uint32_t *app_cnt, uint32_t nq, uint32_t value)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) app_cnt[static_cast<size_t>(i) * kCntStride] = value;
}

__global__ void batch_init_kernel(float *tau, uint32_t *state_cnt, uint32_t *app_cnt, uint32_t nq, uint32_t nq_pad, uint32_t *overflow)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq_pad) {
        tau[i] = i < nq ? -INFINITY : INFINITY;     // padded queries never pass
        state_cnt[i] = 0;
        app_cnt[static_cast<size_t>(i) * kCntStride] = 0;
    }
    if (i == 0) { overflow[0] = 0; overflow[1] = 0; }     // [0] list overflow, [1] non-finite query seen
}

} // namespace

size_t batch_smem_bytes(uint32_t nq_pad) { return kStagesB * kStageB + 256 + nq_pad * sizeof(float) + 1024; }

cudaError_t batch_configure(int smem_optin)
{
    cudaError_t e = cudaFuncSetAttribute(batch_gemm_topm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(batch_gemm_topm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(batch_gemm2_topm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(batch_gemm2_topm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    return e;
}

// 2-CTA variant: tiles are handed out in PAIRS of 128-row tiles; tile0 must be even
cudaError_t batch_gemm2_launch(const CUtensorMap *tmapA, const CUtensorMap *tmapQ128, int sm_count, uint32_t n_rows,
                               uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_pad, uint32_t pitch_elems, int prec,
                               const float *tau, unsigned long long *app_keys, uint32_t *app_cnt, uint32_t cap,
                               uint32_t *overflow, int dense, cudaStream_t st)
{
    const uint32_t n_k = pitch_elems / (prec == 2 ? 32u : static_cast<uint32_t>(kBK));
    const uint32_t idesc = kIdesc2 | idesc_fmt(prec);
    if (tile1 <= tile0) return cudaSuccess;
    const uint32_t pair0 = tile0 / 2, pair1 = (tile1 + 1) / 2;
    uint32_t clusters = static_cast<uint32_t>(sm_count / 2);
    const uint32_t n_items = (pair1 - pair0) * (nq_pad / kBN);
    if (clusters > n_items) clusters = n_items;
    const size_t smem = kStages2 * kStage2 + 256 + nq_pad * sizeof(float) + kEpiWarps2 * kStageBytesPerWarp + 1024;
    unsigned long long *dbg = nullptr;
    if (getenv("RLR_DEBUG_BATCH_TRACE") && pair1 - pair0 > 1000) { cudaMalloc(&dbg, 256); cudaMemset(dbg, 0, 256); }
    if (prec == 2)
        batch_gemm2_topm_kernel<true><<<clusters * 2, kBatch2Threads, smem, st>>>(
            *tmapA, *tmapQ128, n_rows, row_base, pair0, pair1, nq_pad / kBN, n_k, idesc, tau, app_keys, app_cnt, cap, overflow, dense ? 1u : 0u, dbg);
    else
        batch_gemm2_topm_kernel<false><<<clusters * 2, kBatch2Threads, smem, st>>>(
            *tmapA, *tmapQ128, n_rows, row_base, pair0, pair1, nq_pad / kBN, n_k, idesc, tau, app_keys, app_cnt, cap, overflow, dense ? 1u : 0u, dbg);
    if (dbg) {
        unsigned long long h[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, 256, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        fprintf(stderr, "[batch2 trace] leader warp 4: flush %llu cycles, slow path %llu cycles over %llu chunks, %llu records\n", h[6], h[7], h[16], h[17]);
        fprintf(stderr, "[batch2 trace pairs=%u clusters=%u] leader MMA thread: wait tmem_empty %llu, wait smem_full %llu, total %llu cycles; "
                        "epilogue warp 4 (leader/peer): wait tmem_full %llu/%llu, work %llu/%llu; producer wait smem_empty (leader/peer) %llu/%llu\n",
                pair1 - pair0, clusters, h[0], h[1], h[2], h[3], h[11], h[4], h[12], h[5], h[13]);
    }
    return cudaGetLastError();
}

cudaError_t batch_set_cnt_launch(uint32_t *app_cnt, uint32_t nq, uint32_t value, cudaStream_t st)
{
    batch_set_cnt_kernel<<<(nq + 255) / 256, 256, 0, st>>>(app_cnt, nq, value);
    return cudaGetLastError();
}

cudaError_t batch_init_launch(float *tau, uint32_t *state_cnt, uint32_t *app_cnt, uint32_t nq, uint32_t nq_pad,
                              uint32_t *overflow, cudaStream_t st)
{
    batch_init_kernel<<<(nq_pad + 255) / 256, 256, 0, st>>>(tau, state_cnt, app_cnt, nq, nq_pad, overflow);
    return cudaGetLastError();
}

cudaError_t batch_queries_to_operand_launch(const float *d_q, uint32_t dim, void *d_qop, uint32_t pitch_elems, int prec,
                                            uint32_t nq, uint32_t nq_pad, uint32_t *d_nonfinite, cudaStream_t st)
{
    const uint32_t grid = nq_pad < 592 ? nq_pad : 592;
    if (prec == 0) to_operand_rows_kernel<0><<<grid, 256, 0, st>>>(d_q, dim, d_qop, pitch_elems, nq, nq_pad, d_nonfinite);
    else if (prec == 1) to_operand_rows_kernel<1><<<grid, 256, 0, st>>>(d_q, dim, d_qop, pitch_elems, nq, nq_pad, d_nonfinite);
    else to_operand_rows_kernel<2><<<grid, 256, 0, st>>>(d_q, dim, d_qop, pitch_elems, nq, nq_pad, d_nonfinite);
    return cudaGetLastError();
}

cudaError_t batch_gemm_launch(const CUtensorMap *tmapA, const CUtensorMap *tmapQ, int grid, uint32_t n_rows,
                              uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_pad, uint32_t pitch_elems, int prec,
                              const float *tau, unsigned long long *app_keys, uint32_t *app_cnt, uint32_t cap,
                              uint32_t *overflow, cudaStream_t st)
{
    const uint32_t n_k = pitch_elems / (prec == 2 ? 32u : static_cast<uint32_t>(kBK));
    const uint32_t idesc = kIdesc | idesc_fmt(prec);
    if (tile1 <= tile0) return cudaSuccess;
    const uint32_t tiles = tile1 - tile0;
    if (static_cast<uint32_t>(grid) > tiles) grid = static_cast<int>(tiles);
    unsigned long long *dbg = nullptr;
    if (getenv("RLR_DEBUG_BATCH_TRACE") && tiles > 2000) { cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); }
    if (prec == 2)
        batch_gemm_topm_kernel<true><<<grid, kBatchThreads, batch_smem_bytes(nq_pad), st>>>(
            *tmapA, *tmapQ, n_rows, row_base, tile0, tile1, nq_pad / kBN, n_k, idesc, tau, app_keys, app_cnt, cap, overflow, dbg);
    else
        batch_gemm_topm_kernel<false><<<grid, kBatchThreads, batch_smem_bytes(nq_pad), st>>>(
            *tmapA, *tmapQ, n_rows, row_base, tile0, tile1, nq_pad / kBN, n_k, idesc, tau, app_keys, app_cnt, cap, overflow, dbg);
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

// f32 store rows -> bfloat16 copy (round to nearest even), zero padded to dst_pitch
__global__ void rows_to_bf16_kernel(const float *__restrict__ src, uint32_t src_pitch, __nv_bfloat16 *__restrict__ dst,
                                    uint32_t dst_pitch, uint32_t dim, uint64_t n_rows)
{
    for (uint64_t r = blockIdx.x; r < n_rows; r += gridDim.x)
        for (uint32_t c = threadIdx.x; c < dst_pitch; c += blockDim.x)
            dst[r * dst_pitch + c] = __float2bfloat16_rn(c < dim ? src[r * src_pitch + c] : 0.0f);
}

cudaError_t to_bf16_launch(const float *d_src, uint32_t src_pitch, void *d_dst, uint32_t dst_pitch, uint32_t dim,
                           uint64_t n_rows, cudaStream_t stream)
{
    if (n_rows == 0) return cudaSuccess;
    const uint32_t grid = static_cast<uint32_t>(n_rows < 148 * 16 ? n_rows : 148 * 16);
    rows_to_bf16_kernel<<<grid, 256, 0, stream>>>(d_src, src_pitch, static_cast<__nv_bfloat16 *>(d_dst), dst_pitch, dim, n_rows);
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
    constexpr int kKPL = 40;                       // warp version holds up to 1280 keys per query in registers
    if (m <= 128 && m + cap <= 32 * kKPL && getenv("RLR_BATCH_PRUNE_CTA") == nullptr)
        batch_prune_warp_kernel<kKPL><<<(nq + 3) / 4, 128, 0, st>>>(state_keys, state_cnt, m, app_keys, app_cnt, cap, tau, nq);
    else
        batch_prune_kernel<<<nq, 128, 0, st>>>(state_keys, state_cnt, m, app_keys, app_cnt, cap, tau, nq);
    return cudaGetLastError();
}

} // namespace rlr
