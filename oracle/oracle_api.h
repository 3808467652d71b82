/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * One C API, two implementations that tests compare with each other and with the CUDA path:
 *
 *   ref_*  oracle/ref_wrap.cpp   -> oracle/_ref/libref_oracle.so
 *          thin extern "C" wrappers over the reference's OWN templates, compiled unmodified from
 *          /root/reference (ffiasm/c/curve.*, multiexp.*, fft.*, f2field.*, exp.hpp, naf.cpp,
 *          src/groth16.*) on top of the restated L0 field (oracle/port/bn254_field.c).
 *   orc_*  oracle/port/bn254_oracle.c -> oracle/port/liboracle_port.so
 *          plain-C restatement of the same algorithms, function by function, citing file:line.
 *
 * Byte conventions: field elements 32 B little-endian; Fq/Fq2 coordinates in
 * Montgomery form (R = 2^256); affine (0,0) = infinity; XYZZ point = {x,y,zz,zzz};
 * G1 affine 64 B, G1 XYZZ 128 B, G2 affine 128 B, G2 XYZZ 256 B; MSM scalars plain integers of
 * `scalarSize` bytes; NTT data in Montgomery form.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H
#include <stdint.h>

#ifndef ORACLE_PREFIX
#error "define ORACLE_PREFIX(name) before including oracle_api.h"
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* curve.hpp:118-121 -> multiexp.cpp:98-144 */
void ORACLE_PREFIX(g1_msm)(const void *bases, const void *scalars, uint32_t scalarSize, uint32_t n, void *out_xyzz);
void ORACLE_PREFIX(g2_msm)(const void *bases, const void *scalars, uint32_t scalarSize, uint32_t n, void *out_xyzz);
/* curve.cpp:562-574 */
void ORACLE_PREFIX(g1_to_affine)(const void *xyzz, void *out_affine);
void ORACLE_PREFIX(g2_to_affine)(const void *xyzz, void *out_affine);
/* curve.cpp:88-164 (add), :182-248 (mixed add), :337-394 (dbl) */
void ORACLE_PREFIX(g1_add)(void *r_xyzz, const void *a_xyzz, const void *b_xyzz);
void ORACLE_PREFIX(g1_madd)(void *r_xyzz, const void *a_xyzz, const void *b_affine);
void ORACLE_PREFIX(g1_dbl)(void *r_xyzz, const void *a_xyzz);
void ORACLE_PREFIX(g2_add)(void *r_xyzz, const void *a_xyzz, const void *b_xyzz);
void ORACLE_PREFIX(g2_madd)(void *r_xyzz, const void *a_xyzz, const void *b_affine);
void ORACLE_PREFIX(g2_dbl)(void *r_xyzz, const void *a_xyzz);
/* exp.hpp:6-28 + naf.cpp:55-74 */
void ORACLE_PREFIX(g1_mul)(void *r_xyzz, const void *base_affine, const void *scalar, uint32_t scalarSize);
void ORACLE_PREFIX(g2_mul)(void *r_xyzz, const void *base_affine, const void *scalar, uint32_t scalarSize);
/* fft.cpp:175-195 / :198-212, table built for maxDomain = n (ctor :32-115) */
void ORACLE_PREFIX(fr_fft)(void *a, uint64_t n);
void ORACLE_PREFIX(fr_ifft)(void *a, uint64_t n);
/* fft.hpp:28 root(domainPow, idx) of an FFT built for 2^domainPow */
void ORACLE_PREFIX(fr_root)(uint32_t domainPow, uint64_t idx, void *out);
/* groth16.cpp:52-163: a,b from coefs x wtns, c=a.b, 3x(ifft, twist, fft), h = fromMontgomery(a.b-c).
 * `coefs_section` is the zkey section-4 pointer (u32 count first, then 44-byte records). */
void ORACLE_PREFIX(h_scalars)(uint32_t domainSize, uint64_t nCoefs, const void *coefs_section,
                              const void *wtns, void *h_out);
/* groth16.cpp:165-207: the five pre-blinding MSMs.  out = pih(128) pi_a(128) pib1(128) pi_b(256) pi_c(128) */
void ORACLE_PREFIX(prove_msms)(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                               const void *coefs_section, const void *pointsA, const void *pointsB1,
                               const void *pointsB2, const void *pointsC, const void *pointsH,
                               const void *wtns, void *out768);
/* groth16.cpp:209-253: blinding with explicit r,s (32-byte LE each) -> affine A(64) B(128) C(64) */
void ORACLE_PREFIX(blind)(const void *msms768, const void *alpha1, const void *beta1, const void *beta2,
                          const void *delta1, const void *delta2, const void *r32, const void *s32,
                          void *out_proof256);
/* f2field.cpp:93-112 (Fq2 multiplication, 64-byte elements {a,b}) */
void ORACLE_PREFIX(fq2_mul)(void *r, const void *a, const void *b);
/* worker threads actually used (omp_get_max_threads for ref_, 1-or-omp for orc_) */
int ORACLE_PREFIX(threads)(void);

#ifdef __cplusplus
}
#endif
#endif
