/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * extern "C" wrappers (ref_* of oracle_api.h) over the reference's own, unmodified C++ templates,
 * included by path from /root/reference at build time (see oracle/Makefile).  Nothing from the
 * reference is copied into this repository; this file only *calls* it:
 *   Curve<F>::multiMulByScalar / add / dbl / copy / mulByScalar   ffiasm/c/curve.hpp:66-121
 *   FFT<RawFr>::fft / ifft / root                                 ffiasm/c/fft.hpp:22-28
 *   Groth16::makeProver / Prover::prove                           src/groth16.hpp:101-121
 * The output is oracle/_ref/libref_oracle.so (git-ignored; travels to the GPU box prebuilt).
 */
#include <stdint.h>
#include <string.h>
#include <map>
#include <memory>
#include <omp.h>

#include <alt_bn128.hpp>
#include "fft.hpp"
#include "groth16.hpp"

#define ORACLE_PREFIX(name) ref_##name
#include "oracle_api.h"

using namespace AltBn128;

/* The reference's logger writes ./MyLogFile.log from its constructor; the library build swaps in
 * silent definitions of the few Logger members src/groth16.cpp references (declared in the
 * reference's src/logger.hpp).  The ref_prover binary links the real src/logger.cpp. */
namespace CPlusPlusLogging {
Logger *Logger::m_Instance = 0;
Logger::Logger() {}
Logger::~Logger() {}
Logger *Logger::getInstance() throw() { if (!m_Instance) m_Instance = new Logger(); return m_Instance; }
void Logger::trace(const char *) throw() {}
void Logger::debug(const char *) throw() {}
void Logger::debug(std::ostringstream &) throw() {}
}

typedef Curve<F2Field<RawFq>> G2Curve;

static FFT<RawFr> *fft_for(uint64_t maxDomain)
{
    static std::map<uint64_t, FFT<RawFr> *> cache;
    auto it = cache.find(maxDomain);
    if (it != cache.end()) return it->second;
    FFT<RawFr> *f = new FFT<RawFr>(maxDomain);
    cache[maxDomain] = f;
    return f;
}

extern "C" {

int ref_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py --impl reference asks for all host cores explicitly */
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

void ref_g1_msm(const void *bases, const void *scalars, uint32_t scalarSize, uint32_t n, void *out)
{
    G1Point r;
    G1.multiMulByScalar(r, (G1PointAffine *)bases, (uint8_t *)scalars, scalarSize, n);
    memcpy(out, &r, sizeof(r));
}

void ref_g2_msm(const void *bases, const void *scalars, uint32_t scalarSize, uint32_t n, void *out)
{
    G2Point r;
    G2.multiMulByScalar(r, (G2PointAffine *)bases, (uint8_t *)scalars, scalarSize, n);
    memcpy(out, &r, sizeof(r));
}

void ref_g1_to_affine(const void *xyzz, void *out)
{
    G1Point p; memcpy(&p, xyzz, sizeof(p));
    G1PointAffine a; G1.copy(a, p);
    memcpy(out, &a, sizeof(a));
}

void ref_g2_to_affine(const void *xyzz, void *out)
{
    G2Point p; memcpy(&p, xyzz, sizeof(p));
    G2PointAffine a; G2.copy(a, p);
    memcpy(out, &a, sizeof(a));
}

#define WRAP_BIN(NAME, CURVE, PT, OP, T2)                                   \
    void NAME(void *r, const void *a, const void *b)                        \
    {                                                                       \
        PT pa; memcpy(&pa, a, sizeof(pa));                                  \
        T2 pb; memcpy(&pb, b, sizeof(pb));                                  \
        PT pr; CURVE.OP(pr, pa, pb);                                        \
        memcpy(r, &pr, sizeof(pr));                                         \
    }
WRAP_BIN(ref_g1_add, G1, G1Point, add, G1Point)
WRAP_BIN(ref_g1_madd, G1, G1Point, add, G1PointAffine)
WRAP_BIN(ref_g2_add, G2, G2Point, add, G2Point)
WRAP_BIN(ref_g2_madd, G2, G2Point, add, G2PointAffine)

void ref_g1_dbl(void *r, const void *a)
{
    G1Point pa; memcpy(&pa, a, sizeof(pa));
    G1Point pr; G1.dbl(pr, pa);
    memcpy(r, &pr, sizeof(pr));
}

void ref_g2_dbl(void *r, const void *a)
{
    G2Point pa; memcpy(&pa, a, sizeof(pa));
    G2Point pr; G2.dbl(pr, pa);
    memcpy(r, &pr, sizeof(pr));
}

void ref_g1_mul(void *r, const void *base, const void *scalar, uint32_t scalarSize)
{
    G1PointAffine b; memcpy(&b, base, sizeof(b));
    G1Point pr; G1.mulByScalar(pr, b, (uint8_t *)scalar, scalarSize);
    memcpy(r, &pr, sizeof(pr));
}

void ref_g2_mul(void *r, const void *base, const void *scalar, uint32_t scalarSize)
{
    G2PointAffine b; memcpy(&b, base, sizeof(b));
    G2Point pr; G2.mulByScalar(pr, b, (uint8_t *)scalar, scalarSize);
    memcpy(r, &pr, sizeof(pr));
}

void ref_fq2_mul(void *r, const void *a, const void *b)
{
    F2Element x, y, z; memcpy(&x, a, 64); memcpy(&y, b, 64);
    F2.mul(z, x, y);
    memcpy(r, &z, 64);
}

void ref_fr_fft(void *a, uint64_t n) { fft_for(n)->fft((FrElement *)a, n); }
void ref_fr_ifft(void *a, uint64_t n) { fft_for(n)->ifft((FrElement *)a, n); }

void ref_fr_root(uint32_t domainPow, uint64_t idx, void *out)
{
    FrElement &e = fft_for(1ULL << domainPow)->root(domainPow, idx);
    memcpy(out, &e, sizeof(e));
}

/* The next three call sequences follow src/groth16.cpp:48-253 phase by phase, using only the
 * reference's building blocks, so that the intermediate values (h scalars, the five MSM results)
 * that Prover::prove keeps private can be tapped.  ref_groth16_prove below runs the real
 * Prover::prove end to end; tests check the taps against it. */
void ref_h_scalars(uint32_t domainSize, uint64_t nCoefs, const void *coefs_section, const void *wtns_, void *h_out)
{
    typedef Groth16::Coef<Engine> CoefT;
    const CoefT *coefs = (const CoefT *)((const uint8_t *)coefs_section + 4);
    FrElement *wtns = (FrElement *)wtns_;
    RawFr &fr = Fr;
    FrElement *a = (FrElement *)h_out;                 /* the result is written in place, as :158-163 */
    FrElement *b = new FrElement[domainSize];
    FrElement *c = new FrElement[domainSize];
    for (uint32_t i = 0; i < domainSize; i++) { fr.copy(a[i], fr.zero()); fr.copy(b[i], fr.zero()); }
    for (uint64_t i = 0; i < nCoefs; i++) {            /* :66-84 without the locks (serial) */
        FrElement *ab = (coefs[i].m == 0) ? a : b;
        FrElement aux, cf;
        memcpy(&cf, &coefs[i].coef, sizeof(cf));
        fr.mul(aux, wtns[coefs[i].s], cf);
        fr.add(ab[coefs[i].c], ab[coefs[i].c], aux);
    }
    #pragma omp parallel for
    for (uint32_t i = 0; i < domainSize; i++) fr.mul(c[i], a[i], b[i]);
    FFT<RawFr> *fft = fft_for((uint64_t)domainSize * 2);        /* groth16.hpp:94 */
    uint32_t domainPower = fft->log2(domainSize);
    FrElement *abc[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {
        FrElement *x = abc[k];
        fft->ifft(x, domainSize);
        #pragma omp parallel for
        for (uint64_t i = 0; i < domainSize; i++) fr.mul(x[i], x[i], fft->root(domainPower + 1, i));
        fft->fft(x, domainSize);
    }
    #pragma omp parallel for
    for (uint64_t i = 0; i < domainSize; i++) {
        fr.mul(a[i], a[i], b[i]);
        fr.sub(a[i], a[i], c[i]);
        fr.fromMontgomery(a[i], a[i]);
    }
    delete[] b;
    delete[] c;
}

void ref_prove_msms(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                    const void *coefs_section, const void *pointsA, const void *pointsB1,
                    const void *pointsB2, const void *pointsC, const void *pointsH,
                    const void *wtns, void *out768)
{
    uint8_t *out = (uint8_t *)out768;
    FrElement *h = new FrElement[domainSize];
    ref_h_scalars(domainSize, nCoefs, coefs_section, wtns, h);
    uint32_t sW = 32;
    G1Point pih, pi_a, pib1, pi_c;
    G2Point pi_b;
    G1.multiMulByScalar(pih, (G1PointAffine *)pointsH, (uint8_t *)h, sW, domainSize);
    G1.multiMulByScalar(pi_a, (G1PointAffine *)pointsA, (uint8_t *)wtns, sW, nVars);
    G1.multiMulByScalar(pib1, (G1PointAffine *)pointsB1, (uint8_t *)wtns, sW, nVars);
    G2.multiMulByScalar(pi_b, (G2PointAffine *)pointsB2, (uint8_t *)wtns, sW, nVars);
    G1.multiMulByScalar(pi_c, (G1PointAffine *)pointsC, (uint8_t *)wtns + (uint64_t)(nPublic + 1) * sW, sW,
                        nVars - nPublic - 1);
    memcpy(out, &pih, 128);
    memcpy(out + 128, &pi_a, 128);
    memcpy(out + 256, &pib1, 128);
    memcpy(out + 384, &pi_b, 256);
    memcpy(out + 640, &pi_c, 128);
    delete[] h;
}

void ref_blind(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
               const void *delta1, const void *delta2, const void *r32, const void *s32, void *out_proof256)
{
    const uint8_t *in = (const uint8_t *)msms768;
    G1Point pih, pi_a, pib1, pi_c, p1;
    G2Point pi_b, p2;
    memcpy(&pih, in, 128); memcpy(&pi_a, in + 128, 128); memcpy(&pib1, in + 256, 128);
    memcpy(&pi_b, in + 384, 256); memcpy(&pi_c, in + 640, 128);
    G1PointAffine a1, b1, d1; G2PointAffine b2, d2;
    memcpy(&a1, alpha1, 64); memcpy(&b1, beta1, 64); memcpy(&d1, delta1, 64);
    memcpy(&b2, beta2, 128); memcpy(&d2, delta2, 128);
    FrElement r, s, rs;
    memcpy(&r, r32, 32); memcpy(&s, s32, 32);

    G1.add(pi_a, pi_a, a1);                                     /* groth16.cpp:222-224 */
    G1.mulByScalar(p1, d1, (uint8_t *)&r, sizeof(r));
    G1.add(pi_a, pi_a, p1);
    G2.add(pi_b, pi_b, b2);                                     /* :226-228 */
    G2.mulByScalar(p2, d2, (uint8_t *)&s, sizeof(s));
    G2.add(pi_b, pi_b, p2);
    G1.add(pib1, pib1, b1);                                     /* :230-232 */
    G1.mulByScalar(p1, d1, (uint8_t *)&s, sizeof(s));
    G1.add(pib1, pib1, p1);
    G1.add(pi_c, pi_c, pih);                                    /* :234 */
    G1.mulByScalar(p1, pi_a, (uint8_t *)&s, sizeof(s));         /* :236-237 */
    G1.add(pi_c, pi_c, p1);
    G1.mulByScalar(p1, pib1, (uint8_t *)&r, sizeof(r));         /* :239-240 */
    G1.add(pi_c, pi_c, p1);
    Fr.mul(rs, r, s);                                           /* :242-243 */
    Fr.toMontgomery(rs, rs);
    G1.mulByScalar(p1, d1, (uint8_t *)&rs, sizeof(rs));         /* :245-246 */
    G1.sub(pi_c, pi_c, p1);

    G1PointAffine A, C; G2PointAffine B;
    G1.copy(A, pi_a); G2.copy(B, pi_b); G1.copy(C, pi_c);
    uint8_t *out = (uint8_t *)out_proof256;
    memcpy(out, &A, 64); memcpy(out + 64, &B, 128); memcpy(out + 192, &C, 64);
}

/* The real thing: Groth16::makeProver + Prover::prove, untouched.  r,s come from randombytes_buf
 * (oracle/shim/sodium.h; deterministic under ORACLE_FIXED_RS).  out = A(64) B(128) C(64) affine. */
void ref_groth16_prove(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                       const void *alpha1, const void *beta1, const void *beta2, const void *delta1,
                       const void *delta2, const void *coefs_section, const void *pointsA,
                       const void *pointsB1, const void *pointsB2, const void *pointsC, const void *pointsH,
                       const void *wtns, void *out_proof256)
{
    auto prover = Groth16::makeProver<Engine>(nVars, nPublic, domainSize, nCoefs, (void *)alpha1, (void *)beta1,
                                              (void *)beta2, (void *)delta1, (void *)delta2, (void *)coefs_section,
                                              (void *)pointsA, (void *)pointsB1, (void *)pointsB2, (void *)pointsC,
                                              (void *)pointsH);
    auto proof = prover->prove((FrElement *)wtns);
    uint8_t *out = (uint8_t *)out_proof256;
    memcpy(out, &proof->A, 64); memcpy(out + 64, &proof->B, 128); memcpy(out + 192, &proof->C, 64);
}

/* The same two calls kept apart, so that a caller timing proofs can build the Prover once (the reference's
 * server mode does: src/fullprover.cpp:43-59 builds it at start-up, :155 proves per request).  makeProver's FFT
 * root table (fft.cpp:32-115) is then outside the per-proof time, exactly like this repo's b200_zkey_upload. */
void *ref_make_prover(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                      const void *alpha1, const void *beta1, const void *beta2, const void *delta1,
                      const void *delta2, const void *coefs_section, const void *pointsA,
                      const void *pointsB1, const void *pointsB2, const void *pointsC, const void *pointsH)
{
    auto prover = Groth16::makeProver<Engine>(nVars, nPublic, domainSize, nCoefs, (void *)alpha1, (void *)beta1,
                                              (void *)beta2, (void *)delta1, (void *)delta2, (void *)coefs_section,
                                              (void *)pointsA, (void *)pointsB1, (void *)pointsB2, (void *)pointsC,
                                              (void *)pointsH);
    return prover.release();
}

void ref_prover_prove(void *prover, const void *wtns, void *out_proof256)
{
    auto proof = ((Groth16::Prover<Engine> *)prover)->prove((FrElement *)wtns);
    uint8_t *out = (uint8_t *)out_proof256;
    memcpy(out, &proof->A, 64); memcpy(out + 64, &proof->B, 128); memcpy(out + 192, &proof->C, 64);
}

void ref_free_prover(void *prover) { delete (Groth16::Prover<Engine> *)prover; }

} /* extern "C" */
