// common.cuh -- device helpers shared by the sm_100a kernels of the retrieval path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rlr_b200.h"

namespace rlr {

// ---------------------------------------------------------------------------------
// Exact f32 arithmetic of the reference (src/rag_engine.rs:1776-1779): one rounding
// per multiply and one per add, never contracted into an FMA.  The __f*_rn
// intrinsics are never fused by nvcc regardless of -fmad.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ bool is_finite_f32(float x)
{
    return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u;
}

// ---------------------------------------------------------------------------------
// Rank keys.  A candidate's rank under the reference's stable descending sort of
// row-ordered input (src/rag_engine.rs:543) is (score desc, row asc).  Encoded as one
// u64 whose unsigned order is that rank order (larger key == ranks earlier):
//   high 32 bits: order-preserving image of the f32 score (-0.0 folded onto +0.0,
//                 because partial_cmp says they are Equal)
//   low  32 bits: ~row
// key == 0 is "no candidate" (it would need a NaN score with all bits set).
// ---------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord_from_bits(uint32_t b)
{
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t bits_from_ord(uint32_t o)
{
    return (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
}
__device__ __forceinline__ uint32_t ord_f32(float x)
{
    x = __fadd_rn(x, 0.0f); // -0.0 + 0.0 == +0.0 ; every other value unchanged
    return ord_from_bits(__float_as_uint(x));
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row)
{
    return (static_cast<uint64_t>(ord_f32(score)) << 32) | static_cast<uint64_t>(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~static_cast<uint32_t>(key); }
__device__ __forceinline__ float key_score(uint64_t key)
{
    return __uint_as_float(bits_from_ord(static_cast<uint32_t>(key >> 32)));
}

// ---------------------------------------------------------------------------------
// PTX wrappers: shared-memory addresses, mbarrier, TMA (cp.async.bulk.tensor).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// L2 eviction policy for data that is streamed exactly once per query.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 2-D tiled TMA load, global -> this CTA's shared memory, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int32_t x, int32_t y,
                                            uint32_t bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(tmap), "r"(x), "r"(y), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// non-blocking half of a named barrier: counts this warp's threads as arrived and carries on
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

// system-scope flag traffic for the cross-GPU mailbox (peer HBM over NVLink)
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns_common()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// warp-wide max of a u64 with two redux.sync (keys compare as unsigned integers)
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t key)
{
    const uint32_t hi = static_cast<uint32_t>(key >> 32);
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t lo = (hi == mh) ? static_cast<uint32_t>(key) : 0u;
    const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
    return (static_cast<uint64_t>(mh) << 32) | ml;
}

// ---------------------------------------------------------------------------------
// splitmix64-based counter hash for synthetic embeddings; the CPU
// twin lives in oracle/rlr_oracle.c and must produce the same bits.
// ---------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx)
{
    uint32_t u = static_cast<uint32_t>(splitmix64(seed ^ splitmix64(idx)) >> 40);
    return static_cast<float>(static_cast<int32_t>(u) - 8388608) * (1.0f / 8388608.0f);
}

} // namespace rlr
