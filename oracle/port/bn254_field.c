/*
 * ORACLE (test infrastructure).  Instantiates field_tmpl.inc for the two BN254 primes.
 * Constants: SURVEY.md Appendix B (q, R2, R3, np cross-checked against the reference's
 * pre-generated sample ffiasm/benchmark/fr.asm:7098-7103 for the Fq prime).
 */
#include "bn254_field.h"
#include <stdlib.h>

/* Field backend: 1 = ADX assembly (bn254_mmul_adx.S) when the CPU has ADX + BMI2 and ORACLE_NO_ADX is not set,
 * 0 = the plain-C CIOS below.  Both are canonical and bit-identical (tests/test_oracle.py). */
int oracle_use_adx = 0;
__attribute__((constructor)) static void oracle_pick_backend(void)
{
    __builtin_cpu_init();
    oracle_use_adx = __builtin_cpu_supports("adx") && __builtin_cpu_supports("bmi2") && !getenv("ORACLE_NO_ADX");
}
int oracle_field_backend(void) { return oracle_use_adx; }
void oracle_set_field_backend(int adx)
{
    __builtin_cpu_init();
    oracle_use_adx = adx && __builtin_cpu_supports("adx") && __builtin_cpu_supports("bmi2");
}

/* ---- Fq : 21888242871839275222246405745257275088696311157297823662689037894645226208583 */
static const uint64_t FQ_Q[4]  = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FQ_R[4]  = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
static const uint64_t FQ_R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
static const uint64_t FQ_R3[4] = {0xb1cd6dafda1530dfULL, 0x62f210e6a7283db6ULL, 0xef7f0b0c0ada0afbULL, 0x20fd6e902d592544ULL};
#define FNAME(x) Fq_##x
#define F_Q FQ_Q
#define F_R FQ_R
#define F_R2 FQ_R2
#define F_R3 FQ_R3
#define F_NP 0x87d20782e4866389ULL
#include "field_tmpl.inc"
#undef FNAME
#undef F_Q
#undef F_R
#undef F_R2
#undef F_R3
#undef F_NP

/* ---- Fr : 21888242871839275222246405745257275088548364400416034343698204186575808495617 */
static const uint64_t FR_Q[4]  = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_R[4]  = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
static const uint64_t FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
static const uint64_t FR_R3[4] = {0x5e94d8e1b4bf0040ULL, 0x2a489cbe1cfbb6b8ULL, 0x893cc664a19fcfedULL, 0x0cf8594b7fcc657cULL};
#define FNAME(x) Fr_##x
#define F_Q FR_Q
#define F_R FR_R
#define F_R2 FR_R2
#define F_R3 FR_R3
#define F_NP 0xc2e1f593efffffffULL
#include "field_tmpl.inc"
