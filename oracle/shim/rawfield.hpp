/*
 * ORACLE build shim (test infrastructure).
 *
 * The reference's C++ templates (ffiasm/c/curve.*, multiexp.*, fft.*, f2field.*, src/groth16.*)
 * are compiled UNMODIFIED from /root/reference against a field class with the surface of the
 * generated `Raw<Name>` (ffiasm/src/fr.hpp.ejs:79-131, behaviour per ffiasm/src/fr.cpp.ejs:170-272).
 * The generator (node + nasm) cannot run in this image, so that surface is restated here over the
 * plain-C raw routines of oracle/port/bn254_field.c.  One class template, two instantiations.
 */
#ifndef ORACLE_SHIM_RAWFIELD_HPP
#define ORACLE_SHIM_RAWFIELD_HPP

#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <gmp.h>
#include "../port/bn254_field.h"

struct OracleFieldOps {
    void (*copy)(uint64_t *, const uint64_t *);
    void (*swap)(uint64_t *, uint64_t *);
    void (*add)(uint64_t *, const uint64_t *, const uint64_t *);
    void (*sub)(uint64_t *, const uint64_t *, const uint64_t *);
    void (*neg)(uint64_t *, const uint64_t *);
    void (*mmul)(uint64_t *, const uint64_t *, const uint64_t *);
    void (*msquare)(uint64_t *, const uint64_t *);
    void (*mmul1)(uint64_t *, const uint64_t *, uint64_t);
    void (*toMont)(uint64_t *, const uint64_t *);
    void (*fromMont)(uint64_t *, const uint64_t *);
    int (*isEq)(const uint64_t *, const uint64_t *);
    int (*isZero)(const uint64_t *);
};

#define ORACLE_RAW_CLASS(CLS, N, BITS)                                                              \
class CLS {                                                                                         \
public:                                                                                             \
    const static int N64 = 4;                                                                       \
    const static int MaxBits = BITS;                                                                \
    struct Element { uint64_t v[4]; };                                                              \
private:                                                                                            \
    Element fZero, fOne, fNegOne;                                                                   \
    static void qMpz(mpz_t q) { mpz_init(q); mpz_import(q, 4, -1, 8, -1, 0, N##_rawq_ptr()); }      \
public:                                                                                             \
    CLS() { fromString(fZero, "0"); fromString(fOne, "1"); neg(fNegOne, fOne); }                    \
    ~CLS() {}                                                                                       \
    Element &zero() { return fZero; }                                                               \
    Element &one() { return fOne; }                                                                 \
    Element &negOne() { return fNegOne; }                                                           \
    void fromString(Element &r, std::string s) {                                                    \
        mpz_t mr, q; qMpz(q);                                                                       \
        mpz_init_set_str(mr, s.c_str(), 10);                                                        \
        mpz_fdiv_r(mr, mr, q);                                                                      \
        for (int i = 0; i < 4; i++) r.v[i] = 0;                                                     \
        mpz_export((void *)r.v, NULL, -1, 8, -1, 0, mr);                                            \
        N##_rawToMontgomery(r.v, r.v);                                                              \
        mpz_clear(mr); mpz_clear(q);                                                                \
    }                                                                                               \
    std::string toString(Element &a, uint32_t radix = 10) {                                         \
        Element tmp; mpz_t r;                                                                       \
        N##_rawFromMontgomery(tmp.v, a.v);                                                          \
        mpz_init(r);                                                                                \
        mpz_import(r, 4, -1, 8, -1, 0, (const void *)tmp.v);                                        \
        char *res = mpz_get_str(0, radix, r);                                                       \
        mpz_clear(r);                                                                               \
        std::string out(res); free(res); return out;                                                \
    }                                                                                               \
    void inline copy(Element &r, Element &a) { N##_rawCopy(r.v, a.v); }                             \
    void inline swap(Element &a, Element &b) { N##_rawSwap(a.v, b.v); }                             \
    void inline add(Element &r, Element &a, Element &b) { N##_rawAdd(r.v, a.v, b.v); }              \
    void inline sub(Element &r, Element &a, Element &b) { N##_rawSub(r.v, a.v, b.v); }              \
    void inline mul(Element &r, Element &a, Element &b) { N##_rawMMul(r.v, a.v, b.v); }             \
    void inline mul1(Element &r, Element &a, uint64_t b) { N##_rawMMul1(r.v, a.v, b); }             \
    void inline neg(Element &r, Element &a) { N##_rawNeg(r.v, a.v); }                               \
    void inline square(Element &r, Element &a) { N##_rawMSquare(r.v, a.v); }                        \
    void inv(Element &r, Element &a) {                                                              \
        mpz_t mr, q; qMpz(q); mpz_init(mr);                                                         \
        mpz_import(mr, 4, -1, 8, -1, 0, (const void *)a.v);                                         \
        mpz_invert(mr, mr, q);                                                                      \
        for (int i = 0; i < 4; i++) r.v[i] = 0;                                                     \
        mpz_export((void *)r.v, NULL, -1, 8, -1, 0, mr);                                            \
        N##_rawMMul(r.v, r.v, N##_rawR3_ptr());                                                     \
        mpz_clear(mr); mpz_clear(q);                                                                \
    }                                                                                               \
    void div(Element &r, Element &a, Element &b) { Element t; inv(t, b); mul(r, a, t); }            \
    void exp(Element &r, Element &base, uint8_t *scalar, unsigned int scalarSize) {                 \
        bool oneFound = false; Element cb; copy(cb, base);                                          \
        for (int i = (int)scalarSize * 8 - 1; i >= 0; i--) {                                        \
            bool bit = scalar[i >> 3] & (1 << (i & 7));                                             \
            if (!oneFound) { if (!bit) continue; copy(r, cb); oneFound = true; continue; }          \
            square(r, r);                                                                           \
            if (bit) mul(r, r, cb);                                                                 \
        }                                                                                           \
        if (!oneFound) copy(r, fOne);                                                               \
    }                                                                                               \
    void inline toMontgomery(Element &r, Element &a) { N##_rawToMontgomery(r.v, a.v); }             \
    void inline fromMontgomery(Element &r, Element &a) { N##_rawFromMontgomery(r.v, a.v); }         \
    int inline eq(Element &a, Element &b) { return N##_rawIsEq(a.v, b.v); }                         \
    int inline isZero(Element &a) { return N##_rawIsZero(a.v); }                                    \
    void toMpz(mpz_t r, Element &a) {                                                               \
        Element tmp; N##_rawFromMontgomery(tmp.v, a.v);                                             \
        mpz_import(r, 4, -1, 8, -1, 0, (const void *)tmp.v);                                        \
    }                                                                                               \
    void fromMpz(Element &r, mpz_t a) {                                                             \
        for (int i = 0; i < 4; i++) r.v[i] = 0;                                                     \
        mpz_export((void *)r.v, NULL, -1, 8, -1, 0, a);                                             \
        N##_rawToMontgomery(r.v, r.v);                                                              \
    }                                                                                               \
    void fromUI(Element &r, unsigned long int v) {                                                  \
        r.v[0] = v; r.v[1] = r.v[2] = r.v[3] = 0;                                                   \
        N##_rawToMontgomery(r.v, r.v);                                                              \
    }                                                                                               \
    static CLS field;                                                                               \
};

#endif
