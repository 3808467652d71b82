// Micro-benchmark of the integer multiply pipes on sm_100a: issue rate of IMAD (lo), IMAD.HI, IMAD.WIDE and
// the carry-chained IMAD.WIDE.X forms that field.cuh's Montgomery product is made of, plus DFMA for reference.
// Prints lane-ops per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/imad_ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 16

template <int MODE>
__global__ void k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a + i;
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = (double)(a + i);
    double da = (double)a * 1e-9, db = (double)b * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL / 8; u++) {
            if (MODE == 0) {   // IMAD lo, 8 independent chains
#pragma unroll
                for (int i = 0; i < 8; i++) r[i] = r[i] * a + b;
            } else if (MODE == 1) {   // IMAD.HI
#pragma unroll
                for (int i = 0; i < 8; i++) r[i] = __umulhi(r[i], a) + b;
            } else if (MODE == 2) {   // IMAD.WIDE (64-bit accumulate), 4 independent chains on register pairs
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    uint64_t acc = ((uint64_t)r[i + 1] << 32) | r[i];
                    acc = (uint64_t)a * (uint32_t)acc + acc;
                    uint64_t acc2 = (uint64_t)b * (uint32_t)(acc >> 32) + acc;
                    r[i] = (uint32_t)acc2; r[i + 1] = (uint32_t)(acc2 >> 32);
                }
            } else if (MODE == 3) {   // carry chained mad.lo.cc / madc.hi.cc pairs (what fp_mul issues)
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                             "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                             "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5;"
                             "madc.lo.cc.u32 %6, %8, %8, %6; madc.hi.u32 %7, %8, %8, %7;"
                             : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                             : "r"(a), "r"(b), "r"(seed));
            } else if (MODE == 4) {   // DFMA
#pragma unroll
                for (int i = 0; i < 8; i++) d[i] = fma(d[i], da, db);
            } else if (MODE == 5) {   // IADD3 pipe
#pragma unroll
                for (int i = 0; i < 8; i++) r[i] = (r[i] + a) ^ b;
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i] + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, double ops_per_inner, int threads) {
    uint32_t *out; long long *cyc, h = 0;
    cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<MODE><<<148, threads>>>(out, 1, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, 2, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)ITERS * (UNROLL / 8) * ops_per_inner * threads;   // per SM (1 CTA per SM)
    printf("%-28s threads/SM=%4d  %.1f lane-ops/clk/SM  (%lld cycles, %.3f ms)\n", name, threads, ops / (double)h, h, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("IMAD lo (mul.lo+add)", 8, threads);
        run<1>("IMAD.HI", 8, threads);
        run<2>("IMAD.WIDE u64 acc", 8, threads);
        run<3>("mad.lo.cc/madc.hi.cc pairs", 4, threads);   // counted as 4 wide multiplies
        run<4>("DFMA", 8, threads);
        run<5>("IADD3/LOP3", 16, threads);
    }
    return 0;
}
