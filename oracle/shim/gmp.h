/*
 * ORACLE build shim (test infrastructure): declaration-only stand-in for <gmp.h>.
 * The image has the GMP runtime (libgmp.so.10) but not its header, so the handful of
 * mpz entry points the reference's hot path and loaders call (src/main_prover.cpp,
 * src/zkey_utils.cpp, src/wtns_utils.cpp, ffiasm/c/fft.cpp:32-115, ffiasm/c/alt_bn128_test.cpp)
 * are declared here with GMP's public ABI names (__gmpz_*), and the binary links with
 * -l:libgmp.so.10.  No GMP code is reproduced.
 */
#ifndef ORACLE_SHIM_GMP_H
#define ORACLE_SHIM_GMP_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef unsigned long mp_bitcnt_t;
typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define mpz_init            __gmpz_init
#define mpz_clear           __gmpz_clear
#define mpz_import          __gmpz_import
#define mpz_export          __gmpz_export
#define mpz_get_str         __gmpz_get_str
#define mpz_set_str         __gmpz_set_str
#define mpz_init_set_str    __gmpz_init_set_str
#define mpz_init_set_ui     __gmpz_init_set_ui
#define mpz_init_set_si     __gmpz_init_set_si
#define mpz_set             __gmpz_set
#define mpz_set_ui          __gmpz_set_ui
#define mpz_set_si          __gmpz_set_si
#define mpz_get_si          __gmpz_get_si
#define mpz_fits_sint_p     __gmpz_fits_sint_p
#define mpz_cmp             __gmpz_cmp
#define mpz_cmp_ui          __gmpz_cmp_ui
#define mpz_add             __gmpz_add
#define mpz_add_ui          __gmpz_add_ui
#define mpz_sub             __gmpz_sub
#define mpz_mul             __gmpz_mul
#define mpz_mul_2exp        __gmpz_mul_2exp
#define mpz_fdiv_q          __gmpz_fdiv_q
#define mpz_fdiv_r          __gmpz_fdiv_r
#define mpz_fdiv_q_2exp     __gmpz_fdiv_q_2exp
#define mpz_powm            __gmpz_powm
#define mpz_invert          __gmpz_invert
#define mpz_tstbit          __gmpz_tstbit
#define mpz_sizeinbase      __gmpz_sizeinbase

void mpz_init(mpz_ptr);
void mpz_clear(mpz_ptr);
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *mpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
char *mpz_get_str(char *, int, mpz_srcptr);
int mpz_set_str(mpz_ptr, const char *, int);
int mpz_init_set_str(mpz_ptr, const char *, int);
void mpz_init_set_ui(mpz_ptr, unsigned long);
void mpz_init_set_si(mpz_ptr, long);
void mpz_set(mpz_ptr, mpz_srcptr);
void mpz_set_ui(mpz_ptr, unsigned long);
void mpz_set_si(mpz_ptr, long);
long mpz_get_si(mpz_srcptr);
int mpz_fits_sint_p(mpz_srcptr);
int mpz_cmp(mpz_srcptr, mpz_srcptr);
int __gmpz_cmp_ui(mpz_srcptr, unsigned long);
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
int mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
int mpz_tstbit(mpz_srcptr, mp_bitcnt_t);
size_t mpz_sizeinbase(mpz_srcptr, int);

#ifdef __cplusplus
}
#endif
#endif
