// Does the FP64 pipe add throughput to the G1 mixed addition on B200?  
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o dfma_ubench dfma_ubench.cu && ./dfma_ubench
// Part 1: Montgomery products per second of fp_mul (IMAD.WIDE pipe), fp_mul_dfma (FP64 pipe) and of threads that
//         alternate the two on independent chains.
// Part 2: mixed additions per second of the accumulation loop's body (madd-2008-s on an XYZZ running sum, points
//         gathered from an L2-resident table) when the products selected by MASK go through fp_mul_dfma:
//         bit 0 x2*zz, 1 y2*zzz, 2 p*pp, 3 zz*pp, 4 zzz*ppp, 5 x1*pp  (the two squarings and the fused Y3 stay on IMAD).
//         Every variant must produce the same sums (checked against MASK 0): the device check of fp_mul_dfma.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../rapidsnark_old_b200/csrc/field_dfma.cuh"
#include "../../rapidsnark_old_b200/csrc/curve.cuh"
using namespace b200;

template <int MODE>
__global__ void __launch_bounds__(128) k_mul(Fq *x, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a = x[t], b = x[t + 1], c = x[t + 2], d = x[t + 3];
    for (int i = 0; i < n; i++) {
        if (MODE == 0) { a = fp_mul(a, b); c = fp_mul(c, d); }
        if (MODE == 1) { a = fp_mul_dfma(a, b); c = fp_mul_dfma(c, d); }
        if (MODE == 2) { a = fp_mul(a, b); c = fp_mul_dfma(c, d); }
    }
    x[t] = fp_add(a, c);
}

template <int MASK>
DEVFN Fq pm(int bit, const Fq &x, const Fq &y) { return (MASK >> bit) & 1 ? fp_mul_dfma(x, y) : fp_mul(x, y); }

template <int MASK>
DEVFN void madd_mixed(Xyzz<Fq> &acc, const Affine<Fq> &q) {
    if (q.is_zero()) return;
    if (acc.zz.is_zero()) { acc.x = q.x; acc.y = q.y; acc.zz = Fq::one(); acc.zzz = Fq::one(); return; }
    Fq p = fp_sub(pm<MASK>(0, q.x, acc.zz), acc.x);
    Fq r = fp_sub(pm<MASK>(1, q.y, acc.zzz), acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = ec_dbl_affine(q); else acc = Xyzz<Fq>::zero();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = pm<MASK>(2, p, pp);
    Fq zz = pm<MASK>(3, acc.zz, pp);
    Fq zzz = pm<MASK>(4, acc.zzz, ppp);
    Fq qq = pm<MASK>(5, acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
    acc.y = fp_mul_sub_mul(r, fp_sub(qq, x3), acc.y, ppp);
    acc.x = x3; acc.zz = zz; acc.zzz = zzz;
}

template <int MASK, int MINB>
__global__ void __launch_bounds__(128, MINB) k_madd(const Affine<Fq> *tbl, u32 tmask, Xyzz<Fq> *out, int n) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    Xyzz<Fq> acc = Xyzz<Fq>::zero();
    u32 idx = (t * 2654435761u) & tmask;
    Affine<Fq> nxt = tbl[idx];
    for (int i = 0; i < n; i++) {
        Affine<Fq> p = nxt;
        idx = (idx * 1664525u + 1013904223u) & tmask;
        nxt = tbl[idx];
        madd_mixed<MASK>(acc, p);
    }
    out[t] = acc;
}

__global__ void k_make_points(Affine<Fq> *tbl, int n) {   // tbl[i] = (i + 1) * G, G = (1, 2)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<Fq> g;
    g.x = Fq::one();
    g.y = fp_add(Fq::one(), Fq::one());
    Xyzz<Fq> a = Xyzz<Fq>::from_affine(g);
    u32 k = (u32)i + 1;
    Xyzz<Fq> r = ec_mul(a, &k, 1);
    tbl[i] = ec_to_affine(r);
}

static float time_ms(void (*launch)(void *), void *arg) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(arg);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(arg);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

struct MulArg { Fq *x; int n, grid, mode; };
static void launch_mul(void *p) {
    MulArg *a = (MulArg *)p;
    if (a->mode == 0) k_mul<0><<<a->grid, 128>>>(a->x, a->n);
    if (a->mode == 1) k_mul<1><<<a->grid, 128>>>(a->x, a->n);
    if (a->mode == 2) k_mul<2><<<a->grid, 128>>>(a->x, a->n);
}

struct MaddArg { const Affine<Fq> *tbl; u32 tmask; Xyzz<Fq> *out; int n, grid, mask, minb; };
template <int MASK>
static void launch_madd_t(MaddArg *a) {
    if (a->minb == 4) k_madd<MASK, 4><<<a->grid, 128>>>(a->tbl, a->tmask, a->out, a->n);
    else if (a->minb == 3) k_madd<MASK, 3><<<a->grid, 128>>>(a->tbl, a->tmask, a->out, a->n);
    else k_madd<MASK, 2><<<a->grid, 128>>>(a->tbl, a->tmask, a->out, a->n);
}
static const int MASKS[] = {0, 2, 8, 10, 26, 42, 58, 27, 59, 63};
static void launch_madd(void *p) {
    MaddArg *a = (MaddArg *)p;
    switch (a->mask) {
        case 0: launch_madd_t<0>(a); break;
        case 2: launch_madd_t<2>(a); break;      // y2*zzz
        case 8: launch_madd_t<8>(a); break;      // zz*pp
        case 10: launch_madd_t<10>(a); break;    // y2*zzz, zz*pp
        case 26: launch_madd_t<26>(a); break;    // y2*zzz, zz*pp, zzz*ppp
        case 42: launch_madd_t<42>(a); break;    // y2*zzz, zz*pp, x1*pp
        case 58: launch_madd_t<58>(a); break;    // y2*zzz, zz*pp, zzz*ppp, x1*pp
        case 27: launch_madd_t<27>(a); break;    // x2*zz, y2*zzz, zz*pp, zzz*ppp
        case 59: launch_madd_t<59>(a); break;    // five
        case 63: launch_madd_t<63>(a); break;    // all six
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    // ---- part 1
    {
        const int grid = sms * 8, n = 2000;
        Fq *x;
        cudaMalloc(&x, (size_t)(grid * 128 + 8) * sizeof(Fq));
        std::vector<Fq> h(grid * 128 + 8);
        srand(1);
        for (auto &e : h) { for (int i = 0; i < 8; i++) e.v[i] = (u32)rand() * 2654435761u + rand(); e.v[7] &= 0x1fffffffu; }
        const char *names[3] = {"fp_mul (IMAD.WIDE) x2", "fp_mul_dfma (FP64) x2", "one of each, independent"};
        std::vector<Fq> res[3];
        for (int mode = 0; mode < 3; mode++) {
            cudaMemcpy(x, h.data(), h.size() * sizeof(Fq), cudaMemcpyHostToDevice);
            MulArg a{x, n, grid, mode};
            // time the second of two identical launches from the same start state
            launch_mul(&a);
            cudaDeviceSynchronize();
            cudaMemcpy(x, h.data(), h.size() * sizeof(Fq), cudaMemcpyHostToDevice);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            launch_mul(&a);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            res[mode].resize(grid * 128);
            cudaMemcpy(res[mode].data(), x, (size_t)grid * 128 * sizeof(Fq), cudaMemcpyDeviceToHost);
            double prods = 2.0 * n * grid * 128;
            printf("products  %-28s %8.3f ms  %7.1f G products/s  (%.2f per clk per SM at 1.965 GHz)\n", names[mode], ms,
                   prods / ms / 1e6, prods / (ms * 1e-3) / 1.965e9 / sms);
        }
        // mode 1 and 2 recompute mode 0's values only where the threads do not overlap; compare thread 0 mod 4 chains
        long bad = 0;
        for (int t = 0; t < grid * 128; t += 4)
            for (int m = 1; m < 3; m++)
                if (!(res[m][t] == res[0][t])) bad++;
        printf("products: %ld mismatches between the IMAD and the DFMA results\n", bad);
        cudaFree(x);
    }
    // ---- part 2
    {
        const int tn = 1 << 16;
        Affine<Fq> *tbl;
        cudaMalloc(&tbl, (size_t)tn * sizeof(Affine<Fq>));
        k_make_points<<<tn / 128, 128>>>(tbl, tn);
        cudaDeviceSynchronize();
        const int n = 400;
        for (int minb : {4, 3, 2}) {
            const int grid = sms * minb * 4;
            Xyzz<Fq> *out;
            cudaMalloc(&out, (size_t)grid * 128 * sizeof(Xyzz<Fq>));
            std::vector<Xyzz<Fq>> ref;
            for (int mask : MASKS) {
                MaddArg a{tbl, (u32)tn - 1, out, n, grid, mask, minb};
                float ms = time_ms(launch_madd, &a);
                std::vector<Xyzz<Fq>> got((size_t)grid * 128);
                cudaMemcpy(got.data(), out, got.size() * sizeof(Xyzz<Fq>), cudaMemcpyDeviceToHost);
                long bad = 0;
                if (mask == 0) ref = got;
                else for (size_t i = 0; i < got.size(); i++) if (memcmp(&got[i], &ref[i], sizeof(Xyzz<Fq>)) != 0) bad++;
                double adds = (double)n * grid * 128;
                printf("madd  CTAs/SM=%d  mask=%2d (%d on FP64)  %8.3f ms  %6.2f G adds/s  mismatches=%ld  err=%s\n", minb, mask,
                       __builtin_popcount(mask), ms, adds / ms / 1e6, bad, cudaGetErrorString(cudaGetLastError()));
            }
            cudaFree(out);
        }
        cudaFree(tbl);
    }
    return 0;
}
