// Host-side instantiation of the SAME field/curve templates the kernels use (bigint.cuh emulates the
// PTX carry flag when compiled by g++).  Exposed through the C-ABI as b200_host_* so that
//   * the host prover can do its O(1) group work (blinding, to-affine: groth16.cpp:209-253),
//   * ranks can fold gathered partial MSM results,
//   * the CPU test-suite can check the device arithmetic's algorithms bit-for-bit without a GPU.
#include <string.h>
#include "../../include/b200snark.h"
#include "../csrc/curve.cuh"
#include "../csrc/hostfield.hpp"
#include "../csrc/hostgroup.hpp"

using namespace b200;

namespace {
template <class T> T ld(const void *p) { T t; memcpy(&t, p, sizeof(T)); return t; }
template <class T> void st(void *p, const T &t) { memcpy(p, &t, sizeof(T)); }
}

namespace b200 {
// Host end of the MSM (called from msm.cuh): per bucket set  T = W + L * sum_t 2^t * plane_t,  then the window
// Horner of multiexp.cpp:137-141 when the sets are windows (one set = resident per-window tables, nothing to do).
template <class F>
static void msm_finish(const Xyzz<F> *planes, int nsets, int nplanes, int L, int nwin, int c, Xyzz<F> *out) {
    Xyzz<F> total = Xyzz<F>::zero();
    for (int w = nsets - 1; w >= 0; w--) {
        const Xyzz<F> *p = planes + (size_t)w * (nplanes + 1);
        Xyzz<F> t = Xyzz<F>::zero();
        for (int k = nplanes - 1; k >= 0; k--) { t = ec_dbl(t); ec_add(t, p[k]); }
        for (int l = L; l > 1; l >>= 1) t = ec_dbl(t);
        ec_add(t, p[nplanes]);
        if (nsets > 1) for (int k = 0; k < c; k++) total = ec_dbl(total);
        ec_add(total, t);
    }
    (void)nwin;
    *out = total;
}
void host_msm_finish_g1(const void *planes, int nsets, int nplanes, int L, int nwin, int c, void *out) {
    msm_finish<HFq>((const Xyzz<HFq> *)planes, nsets, nplanes, L, nwin, c, (Xyzz<HFq> *)out);
}
void host_msm_finish_g2(const void *planes, int nsets, int nplanes, int L, int nwin, int c, void *out) {
    msm_finish<HFq2>((const Xyzz<HFq2> *)planes, nsets, nplanes, L, nwin, c, (Xyzz<HFq2> *)out);
}

// Blinding + finalisation of src/groth16.cpp:209-253 with explicit r, s (32-byte little-endian, used
// un-reduced like the reference's 248-bit values): in = pih, pi_a, pib1 (G1 XYZZ), pi_b (G2 XYZZ), pi_c.
//
// Split in two so that the part which depends only on the verification key and r, s - three G1 and one G2 scalar
// multiplication, most of the host work of a proof - can run on a host thread WHILE the GPU computes the MSMs:
//   prepare : r*delta1 | s*delta1 | (rs)*delta1 | s*delta2                              (3 x 128 + 256 bytes, XYZZ)
//   finish  : the additions, s*pi_a + r*pib1 and the affine conversions, once the MSM results are there
// groth16_finalize = prepare + finish.  The proof's affine coordinates do not depend on the order of evaluation.
void groth16_blind_prepare(const void *delta1, const void *delta2, const uint8_t *r32, const uint8_t *s32, void *prep640) {
    HG1Affine d1;
    HG2Affine d2;
    memcpy(&d1, delta1, 64);
    memcpy(&d2, delta2, 128);
    HFr r, s;                                            // rs = r*s mod r                 (:242-243)
    memcpy(&r, r32, 32); memcpy(&s, s32, 32);
    HFr rs = hfp_to_mont(fmul(r, s));
    uint8_t rsb[32];
    memcpy(rsb, &rs, 32);
    HG1 rd1 = scalar_mul(d1, r32, 32), sd1 = scalar_mul(d1, s32, 32), rsd1 = scalar_mul(d1, rsb, 32);
    HG2 sd2 = scalar_mul(d2, s32, 32);
    uint8_t *o = (uint8_t *)prep640;
    memcpy(o, &rd1, 128); memcpy(o + 128, &sd1, 128); memcpy(o + 256, &rsd1, 128); memcpy(o + 384, &sd2, 256);
}

// The finish in three pieces, each needing only some of the MSM results, so that a caller can run them as the
// results arrive (b200_groth16_prove: pi_a and pib1 are ready long before pi_c and pih):
//   blind_ab : A = affine(pi_a + alpha1 + r*delta1), B1 = affine(pib1 + beta1 + s*delta1), T = s*A + r*B1
//   blind_b  : B = affine(pi_b + beta2 + s*delta2)
//   blind_c  : C = affine(pi_c + pih + T - (rs)*delta1)
void groth16_blind_ab(const void *pi_a128, const void *pib1_128, const void *alpha1, const void *beta1, const void *prep640,
                      const uint8_t *r32, const uint8_t *s32, void *outA64, void *outT128) {
    const uint8_t *pp = (const uint8_t *)prep640;
    HG1 pi_a, pib1, rd1, sd1;
    memcpy(&pi_a, pi_a128, 128); memcpy(&pib1, pib1_128, 128);
    memcpy(&rd1, pp, 128); memcpy(&sd1, pp + 128, 128);
    HG1Affine a1, b1;
    memcpy(&a1, alpha1, 64); memcpy(&b1, beta1, 64);
    ec_madd(pi_a, a1);                                   // pi_a += alpha1 + r*delta1      (:222-224)
    ec_add(pi_a, rd1);
    ec_madd(pib1, b1);                                   // pib1 += beta1 + s*delta1       (:230-232)
    ec_add(pib1, sd1);
    HG1Affine pa = ec_to_affine(pi_a), pb1 = ec_to_affine(pib1);
    HG1 t = double_scalar_mul(pa, s32, pb1, r32, 32);    // s*pi_a + r*pib1                (:236-240)
    memcpy(outA64, &pa, 64);
    memcpy(outT128, &t, 128);
}

void groth16_blind_b(const void *pi_b256, const void *beta2, const void *prep640, void *outB128) {
    HG2 pi_b, sd2;
    memcpy(&pi_b, pi_b256, 256);
    memcpy(&sd2, (const uint8_t *)prep640 + 384, 256);
    HG2Affine b2;
    memcpy(&b2, beta2, 128);
    ec_madd(pi_b, b2);                                   // pi_b += beta2 + s*delta2       (:226-228)
    ec_add(pi_b, sd2);
    HG2Affine B = ec_to_affine(pi_b);
    memcpy(outB128, &B, 128);
}

void groth16_blind_c(const void *pi_c128, const void *pih128, const void *T128, const void *prep640, void *outC64) {
    HG1 pi_c, pih, t, rsd1;
    memcpy(&pi_c, pi_c128, 128); memcpy(&pih, pih128, 128); memcpy(&t, T128, 128);
    memcpy(&rsd1, (const uint8_t *)prep640 + 256, 128);
    ec_add(pi_c, pih);                                   // pi_c += pih                    (:234)
    ec_add(pi_c, t);                                     // + s*pi_a + r*pib1              (:236-240)
    ec_add(pi_c, ec_neg(rsd1));                          // - (rs)*delta1                  (:245-246)
    HG1Affine C = ec_to_affine(pi_c);
    memcpy(outC64, &C, 64);
}

void groth16_finalize_prepared(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                               const void *prep640, const uint8_t *r32, const uint8_t *s32, void *out256) {
    const uint8_t *m = (const uint8_t *)msms768;
    uint8_t *o = (uint8_t *)out256;
    uint8_t T[128];
    groth16_blind_ab(m + 128, m + 256, alpha1, beta1, prep640, r32, s32, o, T);
    groth16_blind_b(m + 384, beta2, prep640, o + 64);
    groth16_blind_c(m + 640, m, T, prep640, o + 192);
}

void groth16_finalize(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                      const void *delta1, const void *delta2, const uint8_t *r32, const uint8_t *s32, void *out256) {
    uint8_t prep[640];
    groth16_blind_prepare(delta1, delta2, r32, s32, prep);
    groth16_finalize_prepared(msms768, alpha1, beta1, beta2, prep, r32, s32, out256);
}

// sum of n per-GPU partial records (pih, pi_a, pib1 | pi_b | pi_c), one call instead of 5 (n - 1) from the binding
void fold_partials(const void *parts768, int n, void *out768) {
    const uint8_t *p = (const uint8_t *)parts768;
    HG1 g1[4];
    HG2 g2;
    static const int off1[4] = {0, 128, 256, 640};
    for (int k = 0; k < 4; k++) memcpy(&g1[k], p + off1[k], 128);
    memcpy(&g2, p + 384, 256);
    for (int i = 1; i < n; i++) {
        const uint8_t *q = p + (size_t)768 * i;
        for (int k = 0; k < 4; k++) { HG1 t; memcpy(&t, q + off1[k], 128); ec_add(g1[k], t); }
        HG2 t2;
        memcpy(&t2, q + 384, 256);
        ec_add(g2, t2);
    }
    uint8_t *o = (uint8_t *)out768;
    for (int k = 0; k < 4; k++) memcpy(o + off1[k], &g1[k], 128);
    memcpy(o + 384, &g2, 256);
}

// canonical decimal string of a Montgomery-form Fq element (RawFq::toString, fr.cpp.ejs:202-213)
void fq_to_decimal(const void *mont32, char *out80) {
    HFq x;
    memcpy(&x, mont32, 32);
    x = hfp_from_mont(x);
    uint64_t v[4] = {x.v[0], x.v[1], x.v[2], x.v[3]};
    char tmp[80];
    int n = 0;
    while (v[0] | v[1] | v[2] | v[3]) {
        unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--) {
            unsigned __int128 cur = (rem << 64) | v[i];
            v[i] = (uint64_t)(cur / 10);
            rem = cur % 10;
        }
        tmp[n++] = (char)('0' + (int)rem);
    }
    if (n == 0) tmp[n++] = '0';
    for (int i = 0; i < n; i++) out80[i] = tmp[n - 1 - i];
    out80[n] = 0;
}
}  // namespace b200

extern "C" {

void b200_groth16_finalize(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                           const void *delta1, const void *delta2, const void *r32, const void *s32, void *out256) {
    b200::groth16_finalize(msms768, alpha1, beta1, beta2, delta1, delta2, (const uint8_t *)r32, (const uint8_t *)s32, out256);
}
void b200_groth16_blind_prepare(const void *delta1, const void *delta2, const void *r32, const void *s32, void *prep640) {
    b200::groth16_blind_prepare(delta1, delta2, (const uint8_t *)r32, (const uint8_t *)s32, prep640);
}
void b200_groth16_finalize_prepared(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                                    const void *prep640, const void *r32, const void *s32, void *out256) {
    b200::groth16_finalize_prepared(msms768, alpha1, beta1, beta2, prep640, (const uint8_t *)r32, (const uint8_t *)s32, out256);
}
void b200_host_fold_partials(const void *parts768, int n, void *out768) { b200::fold_partials(parts768, n, out768); }
void b200_fq_to_decimal(const void *mont32, char *out80) { b200::fq_to_decimal(mont32, out80); }

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
