/*
 * ORACLE (test infrastructure; see oracle/README.md).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
 *
 * C declarations of the restated L0 routines of the reference:
 * the generated `<Name>_raw*` entry points, ffiasm/src/fr.hpp.ejs:53-64, for
 * Fq (BN254 base field, tasksfile.js:10) and Fr (scalar field, tasksfile.js:11).
 */
#ifndef ORACLE_BN254_FIELD_H
#define ORACLE_BN254_FIELD_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_DECL_FIELD(N)                                                        \
    void N##_rawCopy(uint64_t *r, const uint64_t *a);                               \
    void N##_rawSwap(uint64_t *a, uint64_t *b);                                     \
    void N##_rawAdd(uint64_t *r, const uint64_t *a, const uint64_t *b);             \
    void N##_rawSub(uint64_t *r, const uint64_t *a, const uint64_t *b);             \
    void N##_rawNeg(uint64_t *r, const uint64_t *a);                                \
    void N##_rawMMul(uint64_t *r, const uint64_t *a, const uint64_t *b);            \
    void N##_rawMSquare(uint64_t *r, const uint64_t *a);                            \
    void N##_rawMMul1(uint64_t *r, const uint64_t *a, uint64_t b);                  \
    void N##_rawToMontgomery(uint64_t *r, const uint64_t *a);                       \
    void N##_rawFromMontgomery(uint64_t *r, const uint64_t *a);                     \
    int N##_rawIsEq(const uint64_t *a, const uint64_t *b);                          \
    int N##_rawIsZero(const uint64_t *a);                                           \
    void N##_rawInv(uint64_t *r, const uint64_t *a);                                \
    const uint64_t *N##_rawq_ptr(void);                                             \
    const uint64_t *N##_rawR_ptr(void);                                             \
    const uint64_t *N##_rawR2_ptr(void);                                            \
    const uint64_t *N##_rawR3_ptr(void);

ORACLE_DECL_FIELD(Fq)
ORACLE_DECL_FIELD(Fr)

/* field backend of rawMMul / rawMSquare: 1 = ADX assembly (bn254_mmul_adx.S), 0 = plain C; picked at load time */
extern int oracle_use_adx;
int oracle_field_backend(void);
void oracle_set_field_backend(int adx);

#ifdef __cplusplus
}
#endif
#endif
