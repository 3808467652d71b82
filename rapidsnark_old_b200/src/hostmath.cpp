// Host-side instantiation of the SAME field/curve templates the kernels use (bigint.cuh emulates the
// PTX carry flag when compiled by g++).  Exposed through the C-ABI as b200_host_* so that
//   * the host prover can do its O(1) group work (blinding, to-affine: groth16.cpp:209-253),
//   * ranks can fold gathered partial MSM results,
//   * the CPU test-suite can check the device arithmetic's algorithms bit-for-bit without a GPU.
#include <string.h>
#include "../../include/b200snark.h"
#include "../csrc/curve.cuh"

using namespace b200;

namespace {
template <class T> T ld(const void *p) { T t; memcpy(&t, p, sizeof(T)); return t; }
template <class T> void st(void *p, const T &t) { memcpy(p, &t, sizeof(T)); }
}

extern "C" {

void b200_host_fq_mul(void *r, const void *a, const void *b) { st(r, fp_mul(ld<Fq>(a), ld<Fq>(b))); }
void b200_host_fq_add(void *r, const void *a, const void *b) { st(r, fp_add(ld<Fq>(a), ld<Fq>(b))); }
void b200_host_fq_sub(void *r, const void *a, const void *b) { st(r, fp_sub(ld<Fq>(a), ld<Fq>(b))); }
void b200_host_fq_neg(void *r, const void *a) { st(r, fp_neg(ld<Fq>(a))); }
void b200_host_fq_inv(void *r, const void *a) { st(r, fp_inv(ld<Fq>(a))); }
void b200_host_fr_mul(void *r, const void *a, const void *b) { st(r, fp_mul(ld<Fr>(a), ld<Fr>(b))); }
void b200_host_fr_add(void *r, const void *a, const void *b) { st(r, fp_add(ld<Fr>(a), ld<Fr>(b))); }
void b200_host_fr_sub(void *r, const void *a, const void *b) { st(r, fp_sub(ld<Fr>(a), ld<Fr>(b))); }
void b200_host_fr_neg(void *r, const void *a) { st(r, fp_neg(ld<Fr>(a))); }
void b200_host_fr_inv(void *r, const void *a) { st(r, fp_inv(ld<Fr>(a))); }
void b200_host_fq2_mul(void *r, const void *a, const void *b) { st(r, fmul(ld<Fq2>(a), ld<Fq2>(b))); }
void b200_host_fq2_sqr(void *r, const void *a) { st(r, fsqr(ld<Fq2>(a))); }

void b200_host_g1_add(void *r, const void *a, const void *b) { G1Xyzz t = ld<G1Xyzz>(a); ec_add(t, ld<G1Xyzz>(b)); st(r, t); }
void b200_host_g1_madd(void *r, const void *a, const void *b) { G1Xyzz t = ld<G1Xyzz>(a); ec_madd(t, ld<G1Affine>(b)); st(r, t); }
void b200_host_g1_dbl(void *r, const void *a) { st(r, ec_dbl(ld<G1Xyzz>(a))); }
void b200_host_g1_neg(void *r, const void *a) { st(r, ec_neg(ld<G1Xyzz>(a))); }
void b200_host_g1_to_affine(void *r, const void *a) { st(r, ec_to_affine(ld<G1Xyzz>(a))); }
void b200_host_g1_mul(void *r, const void *base_affine, const void *scalar, uint32_t scalar_size) {
    u32 k[16] = {0};
    memcpy(k, scalar, scalar_size > 64 ? 64 : scalar_size);
    st(r, ec_mul(G1Xyzz::from_affine(ld<G1Affine>(base_affine)), k, (int)((scalar_size + 3) / 4)));
}

void b200_host_g2_add(void *r, const void *a, const void *b) { G2Xyzz t = ld<G2Xyzz>(a); ec_add(t, ld<G2Xyzz>(b)); st(r, t); }
void b200_host_g2_madd(void *r, const void *a, const void *b) { G2Xyzz t = ld<G2Xyzz>(a); ec_madd(t, ld<G2Affine>(b)); st(r, t); }
void b200_host_g2_dbl(void *r, const void *a) { st(r, ec_dbl(ld<G2Xyzz>(a))); }
void b200_host_g2_neg(void *r, const void *a) { st(r, ec_neg(ld<G2Xyzz>(a))); }
void b200_host_g2_to_affine(void *r, const void *a) { st(r, ec_to_affine(ld<G2Xyzz>(a))); }
void b200_host_g2_mul(void *r, const void *base_affine, const void *scalar, uint32_t scalar_size) {
    u32 k[16] = {0};
    memcpy(k, scalar, scalar_size > 64 ? 64 : scalar_size);
    st(r, ec_mul(G2Xyzz::from_affine(ld<G2Affine>(base_affine)), k, (int)((scalar_size + 3) / 4)));
}

}  // extern "C"
