/* Dev tool: per-call latency of rlr_search_mmr measured AT the C ABI (no Python in the timed region), BASELINE
 * config 1 shape by default (10k x 768, top_k=5, diversity_factor=0.3).
 *   gcc -O2 -Iinclude tools/lat_bench.c -o /tmp/lat_bench -Lrust-local-rag_b200 -l:librlr_b200.so -Wl,-rpath,$PWD/rust-local-rag_b200 -lm
 *   /tmp/lat_bench [n_rows] [top_k] [diversity] [store flags]                                                          */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "rlr_b200.h"

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static int cmp(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : x > y;
}

int main(int argc, char **argv)
{
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 10000;
    const uint32_t k = argc > 2 ? (uint32_t)atoi(argv[2]) : 5;
    const float lam = argc > 3 ? (float)atof(argv[3]) : 0.3f;
    const uint32_t sflags = argc > 4 ? (uint32_t)strtoul(argv[4], 0, 0) : 0;
    enum { DIM = 768, NQ = 64, CALLS = 4000 };
    rlr_store *s = 0;
    if (rlr_store_create(0, DIM, n, 0, 0, 0, sflags, &s) != RLR_OK) { fprintf(stderr, "create: %s\n", rlr_last_error()); return 1; }
    if (rlr_store_fill_synthetic(s, 1, 7, 11, 256, 0.35f) != RLR_OK) { fprintf(stderr, "fill: %s\n", rlr_last_error()); return 1; }
    static float q[NQ][DIM];
    uint64_t x = 88172645463325252ull;
    for (int i = 0; i < NQ; ++i)
        for (int j = 0; j < DIM; ++j) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            q[i][j] = (float)((double)(x >> 11) / 9007199254740992.0 - 0.5);
        }
    const rlr_resolved_weights w = {0.7f, 0.3f, 0.7f, 0.3f};
    uint32_t rows[128], cnt = 0;
    float sc[128], em[128], lx[128];
    static double lat[CALLS];
    for (int i = 0; i < 50; ++i)
        if (rlr_search_mmr(s, q[i % NQ], DIM, 0, k, lam, &w, 0, 0, 0, rows, sc, em, lx, &cnt) != RLR_OK) { fprintf(stderr, "search: %s\n", rlr_last_error()); return 1; }
    for (int i = 0; i < CALLS; ++i) {
        const double t0 = now_us();
        rlr_search_mmr(s, q[i % NQ], DIM, 0, k, lam, &w, 0, 0, 0, rows, sc, em, lx, &cnt);
        lat[i] = now_us() - t0;
    }
    qsort(lat, CALLS, sizeof(double), cmp);
    printf("C ABI rlr_search_mmr: n=%llu dim=%d k=%u lam=%.2f store_flags=0x%x: p50 %.1f us, p90 %.1f us, p99 %.1f us, min %.1f us (%d calls, %u results)\n",
           (unsigned long long)n, DIM, k, lam, sflags, lat[CALLS / 2], lat[CALLS * 9 / 10], lat[CALLS * 99 / 100], lat[0], CALLS, cnt);
    rlr_store_destroy(s);
    return 0;
}
