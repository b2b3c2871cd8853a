// sort_regs.cuh -- register/shuffle bitonic sort used by the scan kernel's top-M bookkeeping.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace rlr {

namespace {

constexpr int kSortThreads = kScanRows;   // the R threads of one consumer group; `bar` = that group's named barrier
#define R kSortThreads

// ---------------------------------------------------------------------------------
// Bitonic sort (descending) of n = 2^k <= kTopBuf entries (u64 key + f32 payload) held in
// shared memory, by the R consumer threads.  With one warp per scheduler there is no other
// warp to hide shared-memory latency behind, so the network runs in REGISTERS: thread t
// owns the E = n/R consecutive elements [t*E, (t+1)*E); compare-exchanges at distance
// j < E are register-to-register, E <= j < 32E go through warp shuffles, and only the
// few j >= 32E sub-stages take a round trip through shared memory.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void ce_keep(uint64_t &k, float &v, uint64_t ok, float ov, bool keep_max)
{
    // branch-free select.  Keys are unique except for the zero padding, where taking either
    // copy is equivalent, so no equality test (which nvcc turns into a branch) is needed.
    const bool take = (ok > k) == keep_max;
    k = take ? ok : k;
    v = take ? ov : v;
}

template <int E, int J>
__device__ __forceinline__ void ce_in_regs(uint64_t (&k)[E], float (&v)[E], uint32_t g0, uint32_t K)
{
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if ((e & J) == 0) {
            const int f = e | J;
            const bool desc = ((g0 + e) & K) == 0;
            const bool swap = (k[e] < k[f]) == desc;   // equal keys are zero padding: either order
            const uint64_t ke = k[e], kf = k[f];
            const float ve = v[e], vf = v[f];
            k[e] = swap ? kf : ke; k[f] = swap ? ke : kf;
            v[e] = swap ? vf : ve; v[f] = swap ? ve : vf;
        }
    }
}

// one sub-stage (compile-time K, J) of the network over the thread's E registers
template <int E, uint32_t K, uint32_t J>
__device__ __forceinline__ void substage(uint64_t (&k)[E], float (&v)[E], uint64_t *skeys, float *sembs, uint32_t g0,
                                         uint32_t lane, uint32_t bar)
{
    if constexpr (J < static_cast<uint32_t>(E)) {
        ce_in_regs<E, static_cast<int>(J)>(k, v, g0, K);
    } else if constexpr (J < 32u * E) {
        constexpr uint32_t lane_mask = J / E;
        const bool is_lower = (lane & lane_mask) == 0;
        uint64_t ok[E];
        float ov[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {           // all shuffles first: E independent exchanges in flight
            ok[e] = __shfl_xor_sync(0xffffffffu, k[e], lane_mask);
            ov[e] = __shfl_xor_sync(0xffffffffu, v[e], lane_mask);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool desc = ((g0 + e) & K) == 0;
            ce_keep(k[e], v[e], ok[e], ov[e], desc == is_lower);
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) { skeys[g0 + e] = k[e]; sembs[g0 + e] = v[e]; }
        named_bar_sync(bar, R);
        uint64_t ok[E];
        float ov[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { ok[e] = skeys[(g0 + e) ^ J]; ov[e] = sembs[(g0 + e) ^ J]; }
        named_bar_sync(bar, R);
        const bool is_lower = (g0 & J) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool desc = ((g0 + e) & K) == 0;
            ce_keep(k[e], v[e], ok[e], ov[e], desc == is_lower);
        }
    }
}

template <int E, uint32_t K, uint32_t J>
__device__ __forceinline__ void stage_from(uint64_t (&k)[E], float (&v)[E], uint64_t *skeys, float *sembs, uint32_t g0,
                                           uint32_t lane, uint32_t bar)
{
    substage<E, K, J>(k, v, skeys, sembs, g0, lane, bar);
    if constexpr (J > 1) stage_from<E, K, J / 2>(k, v, skeys, sembs, g0, lane, bar);
}

template <int E, uint32_t K>
__device__ __forceinline__ void network_from(uint64_t (&k)[E], float (&v)[E], uint64_t *skeys, float *sembs, uint32_t g0,
                                             uint32_t lane, uint32_t bar)
{
    stage_from<E, K, K / 2>(k, v, skeys, sembs, g0, lane, bar);
    if constexpr (K < static_cast<uint32_t>(E) * R) network_from<E, K * 2>(k, v, skeys, sembs, g0, lane, bar);
}

template <int E>
__device__ __noinline__ void sort_desc_regs(uint64_t *skeys, float *sembs, uint32_t t, uint32_t bar)
{
    uint64_t k[E];
    float v[E];
    const uint32_t g0 = t * E;
    const uint32_t lane = t & 31;
#pragma unroll
    for (int e = 0; e < E; ++e) { k[e] = skeys[g0 + e]; v[e] = sembs[g0 + e]; }
    named_bar_sync(bar, R);                       // everyone has read its block before anyone overwrites
    network_from<E, 2>(k, v, skeys, sembs, g0, lane, bar);   // the whole network, unrolled at compile time
#pragma unroll
    for (int e = 0; e < E; ++e) { skeys[g0 + e] = k[e]; sembs[g0 + e] = v[e]; }
    named_bar_sync(bar, R);
}

// n must be a power of two <= kTopBuf; entries [n, max(n, R)) are zeroed here when n < R
__device__ __noinline__ void bitonic_desc(uint64_t *keys, float *embs, uint32_t n, uint32_t t, uint32_t bar = 1)
{
    if (n < static_cast<uint32_t>(R)) {
        for (uint32_t i = n + t; i < static_cast<uint32_t>(R); i += R) keys[i] = 0;
        named_bar_sync(bar, R);
        n = R;
    }
    switch (n / R) {
    case 1: sort_desc_regs<1>(keys, embs, t, bar); break;
    case 2: sort_desc_regs<2>(keys, embs, t, bar); break;
    case 4: sort_desc_regs<4>(keys, embs, t, bar); break;
    case 8: sort_desc_regs<8>(keys, embs, t, bar); break;
    default: sort_desc_regs<16>(keys, embs, t, bar); break;
    }
}


#undef R

} // namespace

} // namespace rlr
