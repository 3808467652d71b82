// Instruction mix of one mixed addition / one product under the build-time variants of field.cuh:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200_KARATSUBA=K -DB200_LAZY_PAIR=L -cubin -o v.cubin sass.cu
//   python sass_mix.py v.cubin "k_madd|k_mul|k_sqr"      (counts per pipe of the hot loop; no GPU needed)
#include "../../rapidsnark_old_b200/csrc/curve.cuh"
using namespace b200;
template <class F>
__global__ void k_madd(Xyzz<F> *acc, const Affine<F> *pts, int n) {
    Xyzz<F> a = acc[threadIdx.x];
    for (int i = 0; i < n; i++) { Affine<F> p = pts[i * 32 + threadIdx.x]; ec_madd(a, p); }
    acc[threadIdx.x] = a;
}
template __global__ void k_madd<Fq>(Xyzz<Fq> *, const Affine<Fq> *, int);
template __global__ void k_madd<Fq2>(Xyzz<Fq2> *, const Affine<Fq2> *, int);
__global__ void k_mul(Fq *x, int n) { Fq a = x[threadIdx.x], b = x[threadIdx.x + 32]; for (int i = 0; i < n; i++) { a = fp_mul(a, b); } x[threadIdx.x] = a; }
__global__ void k_sqr(Fq *x, int n) { Fq a = x[threadIdx.x]; for (int i = 0; i < n; i++) { a = fp_sqr(a); } x[threadIdx.x] = a; }
