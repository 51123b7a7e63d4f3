// INT32 ALU throughput probes: the denominator of the integer roofline (BASELINE.json asks for GCUPS as a
// fraction of the INT32 ALU peak, "measured on the box by microbenchmark").
// Each thread runs 8 independent dependent-chains so the pipes, not the latency, are the limit.
#pragma once
#include "launch.h"

namespace poyb200 {


// kind 0: a += b; b += a            (2 adds per chain step)
// kind 1: a = min(a, b); b = max(b, c); c = min(c, a) ... (3 min/max per chain step)
// kind 2: a = min(a + b, c); c += a (the min-plus primitive: 2 adds + 1 min per chain step)
template <int KIND>
__global__ void __launch_bounds__(256) int32_peak_kernel(int *out, int seed) {
    int a[PEAK_CHAINS], b[PEAK_CHAINS], c[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) {
        a[k] = seed + threadIdx.x + k;
        b[k] = seed * 3 + k + blockIdx.x;
        c[k] = seed ^ (k * 977 + threadIdx.x);
    }
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < PEAK_CHAINS; k++) {
                if (KIND == 0) {
                    asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k]));
                    asm volatile("add.s32 %0, %0, %1;" : "+r"(b[k]) : "r"(a[k]));
                } else if (KIND == 1) {
                    asm volatile("min.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b[k]));
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(b[k]) : "r"(c[k]));
                    asm volatile("min.s32 %0, %0, %1;" : "+r"(c[k]) : "r"(a[k]));
                } else {
                    int t;
                    asm volatile("add.s32 %0, %1, %2;" : "=r"(t) : "r"(a[k]), "r"(b[k]));
                    asm volatile("min.s32 %0, %1, %2;" : "=r"(a[k]) : "r"(t), "r"(c[k]));
                    asm volatile("add.s32 %0, %0, %1;" : "+r"(c[k]) : "r"(a[k]));
                }
            }
        }
    }
    int r = 0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) r ^= a[k] ^ b[k] ^ c[k];
    if (r == 0x7fffffff) out[0] = r;  // keeps the chains alive without real traffic
}

}  // namespace poyb200
