/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement (orc_* of oracle_api.h) of the reference's CPU algorithm for the Groth16
 * hot path.  Each function names the reference code it follows (paths under /root/reference,
 * "ffiasm" = depends/ffiasm).  Pinned by tests/test_oracle.py against
 *   - the golden vectors of ffiasm/c/alt_bn128_test.cpp (multiExp2 KAT, Fq2 KAT, group order,
 *     algebraic multiExp, NTT round trip), and
 *   - oracle/_ref/libref_oracle.so = the reference's own templates compiled here.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "bn254_field.h"

#define ORACLE_PREFIX(name) orc_##name
#include "oracle_api.h"

/* ------------------------------------------------------------------ Fq wrappers */
typedef struct { uint64_t v[4]; } fq_t;
typedef fq_t fr_t;

static fq_t FQ_ONE, FQ_ZERO;
static fr_t FR_ONE, FR_ZERO;

static inline void fq_mul(fq_t *r, const fq_t *a, const fq_t *b) { Fq_rawMMul(r->v, a->v, b->v); }
static inline void fq_sqr(fq_t *r, const fq_t *a) { Fq_rawMSquare(r->v, a->v); }
static inline void fq_add(fq_t *r, const fq_t *a, const fq_t *b) { Fq_rawAdd(r->v, a->v, b->v); }
static inline void fq_sub(fq_t *r, const fq_t *a, const fq_t *b) { Fq_rawSub(r->v, a->v, b->v); }
static inline void fq_neg(fq_t *r, const fq_t *a) { Fq_rawNeg(r->v, a->v); }
static inline int fq_iszero(const fq_t *a) { return Fq_rawIsZero(a->v); }
static inline void fq_inv(fq_t *r, const fq_t *a) { Fq_rawInv(r->v, a->v); }   /* fr.cpp.ejs:215-227 */

/* ------------------------------------------------------------------ Fq2 = Fq[u]/(u^2+1), ffiasm/c/f2field.cpp */
typedef struct { fq_t a, b; } fq2_t;
static fq2_t FQ2_ONE, FQ2_ZERO;

static inline void fq2_add(fq2_t *r, const fq2_t *x, const fq2_t *y) { fq_add(&r->a, &x->a, &y->a); fq_add(&r->b, &x->b, &y->b); } /* :69-73 */
static inline void fq2_sub(fq2_t *r, const fq2_t *x, const fq2_t *y) { fq_sub(&r->a, &x->a, &y->a); fq_sub(&r->b, &x->b, &y->b); } /* :75-79 */
static inline void fq2_neg(fq2_t *r, const fq2_t *x) { fq_neg(&r->a, &x->a); fq_neg(&r->b, &x->b); }                              /* :81-85 */
static inline int fq2_iszero(const fq2_t *x) { return fq_iszero(&x->a) && fq_iszero(&x->b); }                                      /* :163-166 */

static void fq2_mul(fq2_t *r, const fq2_t *e1, const fq2_t *e2)          /* f2field.cpp:93-112, nr = -1 */
{
    fq_t aa, bb, bbr, sum1, sum2, t;
    fq_mul(&aa, &e1->a, &e2->a);
    fq_mul(&bb, &e1->b, &e2->b);
    fq_neg(&bbr, &bb);
    fq_add(&sum1, &e1->a, &e1->b);
    fq_add(&sum2, &e2->a, &e2->b);
    fq_mul(&t, &sum1, &sum2);
    fq_add(&r->a, &aa, &bbr);
    fq_sub(&t, &t, &aa);
    fq_sub(&r->b, &t, &bb);
}

static void fq2_sqr(fq2_t *r, const fq2_t *e1)                           /* f2field.cpp:114-126 (nr_is_negone branch) */
{
    fq_t ab, t1, t2;
    fq_mul(&ab, &e1->a, &e1->b);
    fq_add(&t1, &e1->a, &e1->b);
    fq_sub(&t2, &e1->a, &e1->b);
    fq_mul(&r->a, &t1, &t2);
    fq_add(&r->b, &ab, &ab);
}

static void fq2_inv(fq2_t *r, const fq2_t *e1)                           /* f2field.cpp:144-155 */
{
    fq_t t0, t1, t2, t3;
    fq_sqr(&t0, &e1->a);
    fq_sqr(&t1, &e1->b);
    fq_neg(&t2, &t1);
    fq_sub(&t2, &t0, &t2);
    fq_inv(&t3, &t2);
    fq_mul(&r->a, &e1->a, &t3);
    fq_mul(&r->b, &e1->b, &t3);
    fq_neg(&r->b, &r->b);
}

/* ------------------------------------------------------------------ G1, G2 */
#define CN(x) g1_##x
#define FE fq_t
#define FE_MUL fq_mul
#define FE_SQR fq_sqr
#define FE_ADD fq_add
#define FE_SUB fq_sub
#define FE_NEG fq_neg
#define FE_ISZERO fq_iszero
#define FE_INV fq_inv
#define FE_ONE (&FQ_ONE)
#define FE_ZERO (&FQ_ZERO)
#include "curve_tmpl.inc"
#undef CN
#undef FE
#undef FE_MUL
#undef FE_SQR
#undef FE_ADD
#undef FE_SUB
#undef FE_NEG
#undef FE_ISZERO
#undef FE_INV
#undef FE_ONE
#undef FE_ZERO

#define CN(x) g2_##x
#define FE fq2_t
#define FE_MUL fq2_mul
#define FE_SQR fq2_sqr
#define FE_ADD fq2_add
#define FE_SUB fq2_sub
#define FE_NEG fq2_neg
#define FE_ISZERO fq2_iszero
#define FE_INV fq2_inv
#define FE_ONE (&FQ2_ONE)
#define FE_ZERO (&FQ2_ZERO)
#include "curve_tmpl.inc"

__attribute__((constructor)) static void orc_init(void)
{
    memset(&FQ_ZERO, 0, sizeof(FQ_ZERO));
    memset(&FR_ZERO, 0, sizeof(FR_ZERO));
    memcpy(FQ_ONE.v, Fq_rawR_ptr(), 32);
    memcpy(FR_ONE.v, Fr_rawR_ptr(), 32);
    FQ2_ONE.a = FQ_ONE; FQ2_ONE.b = FQ_ZERO;
    FQ2_ZERO.a = FQ_ZERO; FQ2_ZERO.b = FQ_ZERO;
}

/* ------------------------------------------------------------------ FFT<RawFr>, ffiasm/c/fft.cpp */
typedef struct {
    uint32_t s;
    fr_t *roots;       /* roots[i] = w^i, w = primitive 2^s-th root of unity */
    fr_t *powTwoInv;   /* powTwoInv[i] = 2^-i */
} fft_t;

static inline void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { Fr_rawMMul(r->v, a->v, b->v); }

static void fr_pow(fr_t *r, const fr_t *base, const uint64_t e[4])
{
    fr_t acc = FR_ONE, b = *base;
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fr_mul(&acc, &acc, &b);
        fr_mul(&b, &b, &b);
    }
    *r = acc;
}

static uint32_t log2_u64(uint64_t n) { uint32_t r = 0; while (n != 1) { n >>= 1; r++; } return r; }   /* fft.cpp:10-18 */

/* fft.cpp:32-115.  The reference searches the smallest quadratic non-residue with GMP (-> 5 for
 * BN254 r) and sets roots[1] = nqr^((r-1)/2^s); s = min(2-adicity, log2(maxDomainSize)). */
static fft_t *fft_new(uint64_t maxDomainSize)
{
    uint32_t domainPow = log2_u64(maxDomainSize);
    const uint64_t *q = Fr_rawq_ptr();
    uint64_t qm1d2[4], e[4];
    /* (r-1)/2 */
    uint64_t qm1[4] = {q[0] - 1, q[1], q[2], q[3]};
    for (int i = 0; i < 4; i++) qm1d2[i] = (qm1[i] >> 1) | (i < 3 ? qm1[i + 1] << 63 : 0);
    /* nqr: smallest k >= 2 with k^((r-1)/2) != 1 */
    fr_t nqr, t, k;
    for (uint64_t cand = 2;; cand++) {
        uint64_t raw[4] = {cand, 0, 0, 0};
        Fr_rawToMontgomery(k.v, raw);
        fr_pow(&t, &k, qm1d2);
        if (!Fr_rawIsEq(t.v, FR_ONE.v)) { nqr = k; break; }
    }
    /* s and the odd-ish exponent (r-1)/2^s */
    uint32_t s = 1;
    memcpy(e, qm1d2, sizeof(e));
    while (!(e[0] & 1) && s < domainPow) {
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
        s++;
    }
    if (s < domainPow) return NULL;                  /* "Domain size too big for the curve", fft.cpp:70-72 */
    fft_t *f = (fft_t *)malloc(sizeof(fft_t));
    uint64_t nRoots = 1ULL << s;
    f->s = s;
    f->roots = (fr_t *)malloc(sizeof(fr_t) * nRoots);
    f->powTwoInv = (fr_t *)malloc(sizeof(fr_t) * (s + 1));
    f->roots[0] = FR_ONE;
    f->powTwoInv[0] = FR_ONE;
    if (nRoots > 1) {
        fr_pow(&f->roots[1], &nqr, e);
        uint64_t two[4] = {2, 0, 0, 0};
        fr_t two_m;
        Fr_rawToMontgomery(two_m.v, two);
        Fr_rawInv(f->powTwoInv[1].v, two_m.v);
    }
    #pragma omp parallel
    {
        int idThread = omp_get_thread_num();
        int nThreads = omp_get_num_threads();
        uint64_t increment = nRoots / (uint64_t)nThreads;
        uint64_t start = idThread == 0 ? 2 : (uint64_t)idThread * increment;
        uint64_t end = idThread == nThreads - 1 ? nRoots : (uint64_t)(idThread + 1) * increment;
        if (end > start) {
            uint64_t ee[4] = {start, 0, 0, 0};
            fr_pow(&f->roots[start], &f->roots[1], ee);
        }
        for (uint64_t i = start + 1; i < end; i++) fr_mul(&f->roots[i], &f->roots[i - 1], &f->roots[1]);
    }
    for (uint32_t i = 2; i <= s; i++) fr_mul(&f->powTwoInv[i], &f->powTwoInv[i - 1], &f->powTwoInv[1]);
    return f;
}

static fft_t *fft_for(uint64_t maxDomain)
{
    static fft_t *cache[64];
    uint32_t p = log2_u64(maxDomain);
    if (!cache[p]) cache[p] = fft_new(maxDomain);
    return cache[p];
}

static inline const fr_t *fft_root(const fft_t *f, uint32_t domainPow, uint64_t idx)   /* fft.hpp:28 */
{
    return &f->roots[idx << (f->s - domainPow)];
}

static inline uint64_t bit_reverse(uint64_t x, uint32_t domainPow)                     /* fft.cpp:20-27 */
{
    uint64_t r = 0;
    for (uint32_t i = 0; i < domainPow; i++) r |= ((x >> i) & 1) << (domainPow - 1 - i);
    return r;
}

static void fft_run(const fft_t *f, fr_t *a, uint64_t n)                               /* fft.cpp:159-195 */
{
    uint32_t domainPow = log2_u64(n);
    #pragma omp parallel for
    for (uint64_t i = 0; i < n; i++) {
        uint64_t r = bit_reverse(i, domainPow);
        if (i > r) { fr_t tmp = a[i]; a[i] = a[r]; a[r] = tmp; }
    }
    for (uint32_t s = 1; s <= domainPow; s++) {
        uint64_t m = 1ULL << s, mdiv2 = m >> 1;
        #pragma omp parallel for
        for (uint64_t i = 0; i < (n >> 1); i++) {
            fr_t t, u;
            uint64_t k = (i / mdiv2) * m;
            uint64_t j = i % mdiv2;
            fr_mul(&t, fft_root(f, s, j), &a[k + j + mdiv2]);
            u = a[k + j];
            Fr_rawAdd(a[k + j].v, t.v, u.v);
            Fr_rawSub(a[k + j + mdiv2].v, u.v, t.v);
        }
    }
}

static void ifft_run(const fft_t *f, fr_t *a, uint64_t n)                              /* fft.cpp:198-212 */
{
    fft_run(f, a, n);
    uint32_t domainPow = log2_u64(n);
    uint64_t nDiv2 = n >> 1;
    #pragma omp parallel for
    for (uint64_t i = 1; i < nDiv2; i++) {
        fr_t tmp = a[i];
        uint64_t r = n - i;
        fr_mul(&a[i], &a[r], &f->powTwoInv[domainPow]);
        fr_mul(&a[r], &tmp, &f->powTwoInv[domainPow]);
    }
    fr_mul(&a[0], &a[0], &f->powTwoInv[domainPow]);
    if (n > 1) fr_mul(&a[n >> 1], &a[n >> 1], &f->powTwoInv[domainPow]);
}

/* ------------------------------------------------------------------ exported API */
int orc_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

void orc_g1_msm(const void *bases, const void *scalars, uint32_t scalarSize, uint32_t n, void *out)
{
    g1_point r;
    g1_multiexp(&r, (const g1_affine *)bases, (const uint8_t *)scalars, scalarSize, n);
    memcpy(out, &r, sizeof(r));
}

void orc_g2_msm(const void *bases, const void *scalars, uint32_t scalarSize, uint32_t n, void *out)
{
    g2_point r;
    g2_multiexp(&r, (const g2_affine *)bases, (const uint8_t *)scalars, scalarSize, n);
    memcpy(out, &r, sizeof(r));
}

void orc_g1_to_affine(const void *xyzz, void *out) { g1_point p; memcpy(&p, xyzz, sizeof(p)); g1_affine a; g1_to_affine(&a, &p); memcpy(out, &a, sizeof(a)); }
void orc_g2_to_affine(const void *xyzz, void *out) { g2_point p; memcpy(&p, xyzz, sizeof(p)); g2_affine a; g2_to_affine(&a, &p); memcpy(out, &a, sizeof(a)); }

void orc_g1_add(void *r, const void *a, const void *b) { g1_point x, y, z; memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y)); g1_add(&z, &x, &y); memcpy(r, &z, sizeof(z)); }
void orc_g2_add(void *r, const void *a, const void *b) { g2_point x, y, z; memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y)); g2_add(&z, &x, &y); memcpy(r, &z, sizeof(z)); }
void orc_g1_madd(void *r, const void *a, const void *b) { g1_point x, z; g1_affine y; memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y)); g1_madd(&z, &x, &y); memcpy(r, &z, sizeof(z)); }
void orc_g2_madd(void *r, const void *a, const void *b) { g2_point x, z; g2_affine y; memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y)); g2_madd(&z, &x, &y); memcpy(r, &z, sizeof(z)); }
void orc_g1_dbl(void *r, const void *a) { g1_point x, z; memcpy(&x, a, sizeof(x)); g1_dbl(&z, &x); memcpy(r, &z, sizeof(z)); }
void orc_g2_dbl(void *r, const void *a) { g2_point x, z; memcpy(&x, a, sizeof(x)); g2_dbl(&z, &x); memcpy(r, &z, sizeof(z)); }

void orc_g1_mul(void *r, const void *base, const void *scalar, uint32_t scalarSize)
{
    g1_affine b; memcpy(&b, base, sizeof(b));
    g1_point z; g1_mul_scalar(&z, &b, (const uint8_t *)scalar, scalarSize);
    memcpy(r, &z, sizeof(z));
}

void orc_g2_mul(void *r, const void *base, const void *scalar, uint32_t scalarSize)
{
    g2_affine b; memcpy(&b, base, sizeof(b));
    g2_point z; g2_mul_scalar(&z, &b, (const uint8_t *)scalar, scalarSize);
    memcpy(r, &z, sizeof(z));
}

void orc_fq2_mul(void *r, const void *a, const void *b) { fq2_t x, y, z; memcpy(&x, a, 64); memcpy(&y, b, 64); fq2_mul(&z, &x, &y); memcpy(r, &z, 64); }

void orc_fr_fft(void *a, uint64_t n) { fft_run(fft_for(n), (fr_t *)a, n); }
void orc_fr_ifft(void *a, uint64_t n) { ifft_run(fft_for(n), (fr_t *)a, n); }
void orc_fr_root(uint32_t domainPow, uint64_t idx, void *out) { memcpy(out, fft_root(fft_for(1ULL << domainPow), domainPow, idx), 32); }

/* src/groth16.cpp:52-163 */
void orc_h_scalars(uint32_t domainSize, uint64_t nCoefs, const void *coefs_section, const void *wtns_, void *h_out)
{
    const uint8_t *rec = (const uint8_t *)coefs_section + 4;      /* groth16.cpp:38: skip the u32 count */
    const fr_t *wtns = (const fr_t *)wtns_;
    fr_t *a = (fr_t *)h_out;
    fr_t *b = (fr_t *)malloc(sizeof(fr_t) * domainSize);
    fr_t *c = (fr_t *)malloc(sizeof(fr_t) * domainSize);
    for (uint32_t i = 0; i < domainSize; i++) { a[i] = FR_ZERO; b[i] = FR_ZERO; }          /* :56-60 */
    for (uint64_t i = 0; i < nCoefs; i++, rec += 44) {                                     /* :66-84; Coef layout groth16.hpp:27-35 */
        uint32_t m, ci, si;
        fr_t coef, aux;
        memcpy(&m, rec, 4); memcpy(&ci, rec + 4, 4); memcpy(&si, rec + 8, 4); memcpy(&coef, rec + 12, 32);
        fr_t *ab = (m == 0) ? a : b;
        fr_mul(&aux, &wtns[si], &coef);
        Fr_rawAdd(ab[ci].v, ab[ci].v, aux.v);
    }
    #pragma omp parallel for
    for (uint32_t i = 0; i < domainSize; i++) fr_mul(&c[i], &a[i], &b[i]);                 /* :89-96 */
    fft_t *f = fft_for((uint64_t)domainSize * 2);                                          /* groth16.hpp:94 */
    uint32_t domainPower = log2_u64(domainSize);
    fr_t *abc[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {                                                          /* :101-155 */
        fr_t *x = abc[k];
        ifft_run(f, x, domainSize);
        #pragma omp parallel for
        for (uint64_t i = 0; i < domainSize; i++) fr_mul(&x[i], &x[i], fft_root(f, domainPower + 1, i));
        fft_run(f, x, domainSize);
    }
    #pragma omp parallel for
    for (uint64_t i = 0; i < domainSize; i++) {                                            /* :158-163 */
        fr_mul(&a[i], &a[i], &b[i]);
        Fr_rawSub(a[i].v, a[i].v, c[i].v);
        Fr_rawFromMontgomery(a[i].v, a[i].v);
    }
    free(b);
    free(c);
}

/* src/groth16.cpp:165-207 */
void orc_prove_msms(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                    const void *coefs_section, const void *pointsA, const void *pointsB1,
                    const void *pointsB2, const void *pointsC, const void *pointsH,
                    const void *wtns, void *out768)
{
    uint8_t *out = (uint8_t *)out768;
    uint8_t *h = (uint8_t *)malloc((size_t)domainSize * 32);
    orc_h_scalars(domainSize, nCoefs, coefs_section, wtns, h);
    orc_g1_msm(pointsH, h, 32, domainSize, out);
    orc_g1_msm(pointsA, wtns, 32, nVars, out + 128);
    orc_g1_msm(pointsB1, wtns, 32, nVars, out + 256);
    orc_g2_msm(pointsB2, wtns, 32, nVars, out + 384);
    orc_g1_msm(pointsC, (const uint8_t *)wtns + (uint64_t)(nPublic + 1) * 32, 32, nVars - nPublic - 1, out + 640);
    free(h);
}

/* src/groth16.cpp:209-253 with r,s supplied by the caller */
void orc_blind(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
               const void *delta1, const void *delta2, const void *r32, const void *s32, void *out_proof256)
{
    const uint8_t *in = (const uint8_t *)msms768;
    g1_point pih, pi_a, pib1, pi_c, p1, t1;
    g2_point pi_b, p2, t2;
    memcpy(&pih, in, 128); memcpy(&pi_a, in + 128, 128); memcpy(&pib1, in + 256, 128);
    memcpy(&pi_b, in + 384, 256); memcpy(&pi_c, in + 640, 128);
    g1_affine a1, b1, d1, tmpa; g2_affine b2, d2;
    memcpy(&a1, alpha1, 64); memcpy(&b1, beta1, 64); memcpy(&d1, delta1, 64);
    memcpy(&b2, beta2, 128); memcpy(&d2, delta2, 128);
    fr_t r, s, rs;
    memcpy(&r, r32, 32); memcpy(&s, s32, 32);

    g1_madd(&pi_a, &pi_a, &a1);
    g1_mul_scalar(&p1, &d1, (uint8_t *)&r, 32);
    g1_add(&pi_a, &pi_a, &p1);
    g2_madd(&pi_b, &pi_b, &b2);
    g2_mul_scalar(&p2, &d2, (uint8_t *)&s, 32);
    g2_add(&pi_b, &pi_b, &p2);
    g1_madd(&pib1, &pib1, &b1);
    g1_mul_scalar(&p1, &d1, (uint8_t *)&s, 32);
    g1_add(&pib1, &pib1, &p1);
    g1_add(&pi_c, &pi_c, &pih);
    /* the reference multiplies XYZZ points by scalars here (mulByScalar(Point,Point,...)); going
     * through the affine form first yields the same group element */
    g1_to_affine(&tmpa, &pi_a);
    g1_mul_scalar(&p1, &tmpa, (uint8_t *)&s, 32);
    g1_add(&pi_c, &pi_c, &p1);
    g1_to_affine(&tmpa, &pib1);
    g1_mul_scalar(&p1, &tmpa, (uint8_t *)&r, 32);
    g1_add(&pi_c, &pi_c, &p1);
    fr_mul(&rs, &r, &s);                         /* :242-243: mont_mul then toMontgomery = plain r*s mod r */
    Fr_rawToMontgomery(rs.v, rs.v);
    g1_mul_scalar(&p1, &d1, (uint8_t *)&rs, 32);
    g1_sub(&pi_c, &pi_c, &p1);
    (void)t1; (void)t2;

    g1_affine A, C; g2_affine B;
    g1_to_affine(&A, &pi_a); g2_to_affine(&B, &pi_b); g1_to_affine(&C, &pi_c);
    uint8_t *out = (uint8_t *)out_proof256;
    memcpy(out, &A, 64); memcpy(out + 64, &B, 128); memcpy(out + 192, &C, 64);
}
