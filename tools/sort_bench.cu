// dev tool: correctness + cycle cost of the register/shuffle bitonic sort (sort_regs.cuh)
#include <algorithm>
#include <cstdio>
#include <vector>
#include "../rust-local-rag_b200/csrc/sort_regs.cuh"

__global__ void __launch_bounds__(192, 1) k_sort(uint64_t *g_keys, float *g_v, uint32_t n, long long *cycles, int reps)
{
    __shared__ uint64_t keys[2048];
    __shared__ float v[2048];
    const uint32_t t = threadIdx.x;
    if (t >= 128) return;
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
        for (uint32_t i = t; i < n; i += 128) { keys[i] = g_keys[i]; v[i] = g_v[i]; }
        rlr::named_bar_sync(1, 128);
        const long long c0 = clock64();
        rlr::bitonic_desc(keys, v, n, t);
        total += clock64() - c0;
    }
    for (uint32_t i = t; i < n; i += 128) { g_keys[i] = keys[i]; g_v[i] = v[i]; }
    if (t == 0) *cycles = total / reps;
}

int main()
{
    for (uint32_t n : {32u, 128u, 256u, 512u, 1024u, 2048u}) {
        std::vector<uint64_t> h(n); std::vector<float> hv(n);
        uint64_t x = 88172645463325252ull;
        for (uint32_t i = 0; i < n; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x | 1; hv[i] = float(h[i] % 1000); }
        uint64_t *d; float *dv; long long *dc;
        cudaMalloc(&d, n * 8); cudaMalloc(&dv, n * 4); cudaMalloc(&dc, 8);
        cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dv, hv.data(), n * 4, cudaMemcpyHostToDevice);
        k_sort<<<1, 192>>>(d, dv, n, dc, 10);
        std::vector<uint64_t> o(n); std::vector<float> ov(n); long long c = 0;
        cudaMemcpy(o.data(), d, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(ov.data(), dv, n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        std::sort(h.begin(), h.end(), std::greater<uint64_t>());
        bool ok = cudaGetLastError() == cudaSuccess;
        for (uint32_t i = 0; i < n && ok; ++i) ok = (o[i] == h[i]) && (ov[i] == float(h[i] % 1000));
        printf("n=%4u  %s  %lld cycles/sort\n", n, ok ? "OK " : "BAD", c);
        cudaFree(d); cudaFree(dv); cudaFree(dc);
    }
    return 0;
}
