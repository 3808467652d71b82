// Fq2 = Fq[u]/(u^2 + 1) for the G2 tables (reference: ffiasm/c/f2field.cpp:69-172, non-residue -1 at
// ffiasm/c/alt_bn128.cpp:6).  Element = {a, b} = a + b*u, 64 bytes, both halves Montgomery.
#pragma once
#include "field.cuh"

namespace b200 {

template <class B>
struct alignas(16) Fq2T {
    B a, b;
    HD static Fq2T zero() { Fq2T r; r.a = B::zero(); r.b = B::zero(); return r; }
    HD static Fq2T one() { Fq2T r; r.a = B::one(); r.b = B::zero(); return r; }
    HD bool is_zero() const { return a.is_zero() && b.is_zero(); }
    HD bool operator==(const Fq2T &o) const { return a == o.a && b == o.b; }
    HD bool operator!=(const Fq2T &o) const { return !(*this == o); }
};
typedef Fq2T<Fq> Fq2;

// uniform operator names so curve.cuh is written once for Fq and Fq2
HD Fq fadd(const Fq &x, const Fq &y) { return fp_add(x, y); }
HD Fq fsub(const Fq &x, const Fq &y) { return fp_sub(x, y); }
HD Fq fmul(const Fq &x, const Fq &y) { return fp_mul(x, y); }
HD Fq fsqr(const Fq &x) { return fp_sqr(x); }
HD Fq fneg(const Fq &x) { return fp_neg(x); }
HD Fq fdbl(const Fq &x) { return fp_dbl(x); }
HD Fq finv(const Fq &x) { return fp_inv(x); }

HD Fr fadd(const Fr &x, const Fr &y) { return fp_add(x, y); }
HD Fr fsub(const Fr &x, const Fr &y) { return fp_sub(x, y); }
HD Fr fmul(const Fr &x, const Fr &y) { return fp_mul(x, y); }
HD Fr fsqr(const Fr &x) { return fp_sqr(x); }
HD Fr fneg(const Fr &x) { return fp_neg(x); }
HD Fr fdbl(const Fr &x) { return fp_dbl(x); }
HD Fr finv(const Fr &x) { return fp_inv(x); }

template <class B> HD Fq2T<B> fadd(const Fq2T<B> &x, const Fq2T<B> &y) { Fq2T<B> r; r.a = fadd(x.a, y.a); r.b = fadd(x.b, y.b); return r; }
template <class B> HD Fq2T<B> fsub(const Fq2T<B> &x, const Fq2T<B> &y) { Fq2T<B> r; r.a = fsub(x.a, y.a); r.b = fsub(x.b, y.b); return r; }
template <class B> HD Fq2T<B> fneg(const Fq2T<B> &x) { Fq2T<B> r; r.a = fneg(x.a); r.b = fneg(x.b); return r; }
template <class B> HD Fq2T<B> fdbl(const Fq2T<B> &x) { Fq2T<B> r; r.a = fdbl(x.a); r.b = fdbl(x.b); return r; }

// Karatsuba, 3 base-field products (f2field.cpp:93-112)
template <class B>
HD Fq2T<B> fmul(const Fq2T<B> &x, const Fq2T<B> &y) {
    B aa = fmul(x.a, y.a);
    B bb = fmul(x.b, y.b);
    B s = fmul(fadd(x.a, x.b), fadd(y.a, y.b));
    Fq2T<B> r;
    r.a = fsub(aa, bb);
    r.b = fsub(fsub(s, aa), bb);
    return r;
}

// Device representation: lazily reduced product - three plain 512-bit products, the Karatsuba combination on the
// unreduced values, then TWO Montgomery reductions instead of three (320 instead of 390 wide multiplies).
//   c0 = a0 b0 - a1 b1   (+ p * 2^256 when negative, so that 0 <= c0 < p * 2^256)
//   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 = a0 b1 + a1 b0 < 2 p^2 < p * 2^256
// The sums a0 + a1, b0 + b1 stay unreduced (< 2p < 2^255 fits the 8 limbs).  Results are canonical like fp_mul's.
HD Fq2T<Fq> fmul(const Fq2T<Fq> &x, const Fq2T<Fq> &y) {
    u32 t0[16], t1[16], t2[16], sa[8], sb[8];
    detail::mul8x8(t0, x.a.v, y.a.v);
    detail::mul8x8(t1, x.b.v, y.b.v);
    sa[0] = add_cc(x.a.v[0], x.b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sa[i] = addc_cc(x.a.v[i], x.b.v[i]);
    sa[7] = addc(x.a.v[7], x.b.v[7]);
    sb[0] = add_cc(y.a.v[0], y.b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sb[i] = addc_cc(y.a.v[i], y.b.v[i]);
    sb[7] = addc(y.a.v[7], y.b.v[7]);
    detail::mul8x8(t2, sa, sb);
    // t2 -= t0; t2 -= t1   (never negative)
    t2[0] = sub_cc(t2[0], t0[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) t2[i] = subc_cc(t2[i], t0[i]);
    t2[15] = subc(t2[15], t0[15]);
    t2[0] = sub_cc(t2[0], t1[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) t2[i] = subc_cc(t2[i], t1[i]);
    t2[15] = subc(t2[15], t1[15]);
    // t0 -= t1, add p * 2^256 back on borrow
    t0[0] = sub_cc(t0[0], t1[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) t0[i] = subc_cc(t0[i], t1[i]);
    u32 borrow = subc(0, 0);
    t0[8] = add_cc(t0[8], borrow & FqParams::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) t0[8 + i] = addc_cc(t0[8 + i], borrow & FqParams::mod(i));
    t0[15] = addc(t0[15], borrow & FqParams::mod(7));
    Fq2T<Fq> r;
    detail::mont_reduce16<FqParams>(r.a.v, t0);
    fp_reduce_once(r.a);
    detail::mont_reduce16<FqParams>(r.b.v, t2);
    fp_reduce_once(r.b);
    return r;
}

// x*y - z*w (the Y3 of the mixed addition) with one reduction per component instead of two
HD Fq fmul_sub_mul(const Fq &x, const Fq &y, const Fq &z, const Fq &w) { return fp_mul_sub_mul(x, y, z, w); }
HD Fr fmul_sub_mul(const Fr &x, const Fr &y, const Fr &z, const Fr &w) { return fp_mul_sub_mul(x, y, z, w); }
template <class B>
HD Fq2T<B> fmul_sub_mul(const Fq2T<B> &x, const Fq2T<B> &y, const Fq2T<B> &z, const Fq2T<B> &w) {
    return fsub(fmul(x, y), fmul(z, w));
}
#if B200_LAZY_PAIR
namespace detail {
// c0 = a0 b0 - a1 b1, c1 = a0 b1 + a1 b0 as unreduced 512-bit values: c0 may be negative and is returned as
// c0 mod 2^512 with its sign in the return value (all-ones = negative); 0 <= c1 < 2 p^2
HD u32 fq2_mul_wide(u32 *c0, u32 *c1, const Fq2T<Fq> &x, const Fq2T<Fq> &y) {
    u32 t1[16], sa[8], sb[8];
    mul8x8(c0, x.a.v, y.a.v);
    mul8x8(t1, x.b.v, y.b.v);
    sa[0] = add_cc(x.a.v[0], x.b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sa[i] = addc_cc(x.a.v[i], x.b.v[i]);
    sa[7] = addc(x.a.v[7], x.b.v[7]);
    sb[0] = add_cc(y.a.v[0], y.b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) sb[i] = addc_cc(y.a.v[i], y.b.v[i]);
    sb[7] = addc(y.a.v[7], y.b.v[7]);
    mul8x8(c1, sa, sb);
    c1[0] = sub_cc(c1[0], c0[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) c1[i] = subc_cc(c1[i], c0[i]);
    c1[15] = subc(c1[15], c0[15]);
    c1[0] = sub_cc(c1[0], t1[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) c1[i] = subc_cc(c1[i], t1[i]);
    c1[15] = subc(c1[15], t1[15]);
    c0[0] = sub_cc(c0[0], t1[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) c0[i] = subc_cc(c0[i], t1[i]);
    return subc(0, 0);
}
}  // namespace detail

#ifndef B200_FQ2_Y3_SEQ
#define B200_FQ2_Y3_SEQ 0
#endif
#if B200_FQ2_Y3_SEQ
namespace detail {
// 17-limb two's complement accumulators (limb 16 = sign extension): c += t / c -= t for a 16-limb t >= 0
HD void acc17_add(u32 *c, const u32 *t) {
    c[0] = add_cc(c[0], t[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) c[i] = addc_cc(c[i], t[i]);
    c[16] = addc(c[16], 0);
}
HD void acc17_sub(u32 *c, const u32 *t) {
    c[0] = sub_cc(c[0], t[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) c[i] = subc_cc(c[i], t[i]);
    c[16] = subc(c[16], 0);
}
HD void add8_noreduce(u32 *s, const u32 *a, const u32 *b) {     // a + b < 2p < 2^255
    s[0] = add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) s[i] = addc_cc(a[i], b[i]);
    s[7] = addc(a[7], b[7]);
}
// -2 p^2 < c < 2 p^2 in two's complement -> a valid Montgomery-reduction input congruent to it
HD void acc17_fix(u32 *c) {
    const u32 neg = c[16];                                      // 0 or all-ones
    c[8] = add_cc(c[8], neg & FqParams::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) c[8 + i] = addc_cc(c[8 + i], neg & FqParams::mod(i));
    c[15] = addc(c[15], neg & FqParams::mod(7));
}
}  // namespace detail

// Same value as below, one 512-bit product alive at a time: every product is added to / subtracted from the two
// signed accumulators as soon as it exists (the Karatsuba middle products first), so that at most c0, c1, one product
// and one pair of operand sums are alive - 66 registers instead of 96 + operands.  For the G2 accumulation kernel
// compiled for three CTAs per SM (168 registers), where the Y3 of the mixed addition is the register peak.
HD Fq2T<Fq> fmul_sub_mul(const Fq2T<Fq> &x, const Fq2T<Fq> &y, const Fq2T<Fq> &z, const Fq2T<Fq> &w) {
    u32 c0[17], c1[17], t[16];
    {
        u32 sa[8], sb[8];
        detail::add8_noreduce(sa, x.a.v, x.b.v);
        detail::add8_noreduce(sb, y.a.v, y.b.v);
        detail::mul8x8(c1, sa, sb);                   // c1 = (xa + xb)(ya + yb)
        c1[16] = 0;
    }
    detail::mul8x8(c0, x.a.v, y.a.v);                 // c0 = xa ya
    c0[16] = 0;
    detail::acc17_sub(c1, c0);
    detail::mul8x8(t, x.b.v, y.b.v);
    detail::acc17_sub(c0, t);                         // c0 = xa ya - xb yb
    detail::acc17_sub(c1, t);                         // c1 = xa yb + xb ya
    {
        u32 sa[8], sb[8];
        detail::add8_noreduce(sa, z.a.v, z.b.v);
        detail::add8_noreduce(sb, w.a.v, w.b.v);
        detail::mul8x8(t, sa, sb);
        detail::acc17_sub(c1, t);
    }
    detail::mul8x8(t, z.a.v, w.a.v);
    detail::acc17_sub(c0, t);
    detail::acc17_add(c1, t);
    detail::mul8x8(t, z.b.v, w.b.v);
    detail::acc17_add(c0, t);                         // c0 = (xa ya - xb yb) - (za wa - zb wb)
    detail::acc17_add(c1, t);                         // c1 = (xa yb + xb ya) - (za wb + zb wa)
    detail::acc17_fix(c0);
    detail::acc17_fix(c1);
    Fq2T<Fq> r;
    detail::mont_reduce16<FqParams>(r.a.v, c0);
    fp_reduce_once(r.a);
    detail::mont_reduce16<FqParams>(r.b.v, c1);
    fp_reduce_once(r.b);
    return r;
}
#else
// |c0|, |c1| of the difference stay below 2 p^2 < p * 2^255, so one conditional + p * 2^256 makes each a valid
// Montgomery-reduction input: 6 products and 2 reductions instead of 6 and 4
HD Fq2T<Fq> fmul_sub_mul(const Fq2T<Fq> &x, const Fq2T<Fq> &y, const Fq2T<Fq> &z, const Fq2T<Fq> &w) {
    u32 c0[16], c1[16], d0[16], d1[16];
    u32 neg0 = detail::fq2_mul_wide(c0, c1, x, y);
    u32 negd = detail::fq2_mul_wide(d0, d1, z, w);
    // c0 - d0 as a signed 512 + 1 bit value: sign word = neg0 - negd - borrow
    c0[0] = sub_cc(c0[0], d0[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) c0[i] = subc_cc(c0[i], d0[i]);
    u32 s0 = subc(neg0, negd);                       // 0 = non-negative, all-ones = negative (|value| < 2^511)
    c0[8] = add_cc(c0[8], s0 & FqParams::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) c0[8 + i] = addc_cc(c0[8 + i], s0 & FqParams::mod(i));
    c0[15] = addc(c0[15], s0 & FqParams::mod(7));
    detail::sub16_mod<FqParams>(c1, d1);
    Fq2T<Fq> r;
    detail::mont_reduce16<FqParams>(r.a.v, c0);
    fp_reduce_once(r.a);
    detail::mont_reduce16<FqParams>(r.b.v, c1);
    fp_reduce_once(r.b);
    return r;
}
#endif
#endif

// complex squaring, 2 base-field products (f2field.cpp:114-126)
template <class B>
HD Fq2T<B> fsqr(const Fq2T<B> &x) {
    B ab = fmul(x.a, x.b);
    Fq2T<B> r;
    r.a = fmul(fadd(x.a, x.b), fsub(x.a, x.b));
    r.b = fdbl(ab);
    return r;
}

// inverse through the norm a^2 + b^2 (f2field.cpp:144-155)
template <class B>
HD Fq2T<B> finv(const Fq2T<B> &x) {
    B n = finv(fadd(fsqr(x.a), fsqr(x.b)));
    Fq2T<B> r;
    r.a = fmul(x.a, n);
    r.b = fneg(fmul(x.b, n));
    return r;
}

// Products as called from the COLD group operations (full add, doubling: bucket folding / reduction, table
// building).  For Fq2 they are out-of-line on the device: a fully inlined G2 addition is ~60 KB of SASS per call
// site, and the cold kernels are latency-bound anyway; the hot mixed addition keeps using the inlined fmul/fsqr.
template <class F> HD F cmul(const F &x, const F &y) { return fmul(x, y); }
template <class F> HD F csqr(const F &x) { return fsqr(x); }
template <class B> HD_COLD Fq2T<B> cmul(const Fq2T<B> &x, const Fq2T<B> &y) { return fmul(x, y); }
template <class B> HD_COLD Fq2T<B> csqr(const Fq2T<B> &x) { return fsqr(x); }
template <class F> HD F cmul_sub_mul(const F &x, const F &y, const F &z, const F &w) { return fmul_sub_mul(x, y, z, w); }
template <class B> HD_COLD Fq2T<B> cmul_sub_mul(const Fq2T<B> &x, const Fq2T<B> &y, const Fq2T<B> &z, const Fq2T<B> &w) { return fmul_sub_mul(x, y, z, w); }

// Products of the HOT mixed addition.  The G2 loop is ~100 KB of straight-line SASS when everything is inlined and
// ncu attributes a quarter of its issue stalls to instruction fetch, so two out-of-line forms were built and
// measured on B200 (2^20 proof, G2 accumulation phase 7.8 ms inlined):
//   B200_G2_HOT_CALLS = 1  Fq2 products through the out-of-line cold copies (operands via the local-memory stack): 10.3 ms
//   B200_G2_HOT_CALLS = 2  base-field product / reduction as by-value calls (operands in registers, loop 45 KB):  8.5 ms
// Both lose to the inlined loop (0, the default); they stay as build variants for the record.
#ifndef B200_G2_HOT_CALLS
#define B200_G2_HOT_CALLS 0
#endif
#if B200_G2_HOT_CALLS == 2 && defined(__CUDA_ARCH__)
// Variant 2: the Fq2 products of the hot loop are assembled from two out-of-line base-field routines whose
// operands and results travel BY VALUE in registers (no stack traffic): the 512-bit product and the Montgomery
// reduction.  One copy of each in the kernel image instead of ~46 inlined ones per mixed addition.
struct W16 { u32 v[16]; };
static __device__ __noinline__ W16 fq_mul_wide_call(Fq a, Fq b) { W16 r; detail::mul8x8(r.v, a.v, b.v); return r; }
static __device__ __noinline__ Fq fq_reduce_call(W16 t) { Fq r; detail::mont_reduce16<FqParams>(r.v, t.v); fp_reduce_once(r); return r; }
DEVFN Fq fq_add_noreduce(const Fq &a, const Fq &b) {
    Fq r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    r.v[7] = addc(a.v[7], b.v[7]);
    return r;
}
DEVFN void w16_sub(W16 &t, const W16 &s) {        // t -= s, never negative
    t.v[0] = sub_cc(t.v[0], s.v[0]);
#pragma unroll
    for (int i = 1; i < 15; i++) t.v[i] = subc_cc(t.v[i], s.v[i]);
    t.v[15] = subc(t.v[15], s.v[15]);
}
DEVFN u32 w16_sub_sign(W16 &t, const W16 &s) {    // t -= s mod 2^512, returns all-ones when negative
    t.v[0] = sub_cc(t.v[0], s.v[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) t.v[i] = subc_cc(t.v[i], s.v[i]);
    return subc(0, 0);
}
DEVFN void w16_add_p_hi(W16 &t, u32 mask) {       // t += (mask & p) * 2^256
    t.v[8] = add_cc(t.v[8], mask & FqParams::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) t.v[8 + i] = addc_cc(t.v[8 + i], mask & FqParams::mod(i));
    t.v[15] = addc(t.v[15], mask & FqParams::mod(7));
}
// c0 = a0 b0 - a1 b1 (mod 2^512, sign returned), c1 = a0 b1 + a1 b0
DEVFN u32 fq2_mul_wide_calls(W16 &c0, W16 &c1, const Fq2 &x, const Fq2 &y) {
    c0 = fq_mul_wide_call(x.a, y.a);
    W16 t1 = fq_mul_wide_call(x.b, y.b);
    c1 = fq_mul_wide_call(fq_add_noreduce(x.a, x.b), fq_add_noreduce(y.a, y.b));
    w16_sub(c1, c0);
    w16_sub(c1, t1);
    return w16_sub_sign(c0, t1);
}
DEVFN Fq2 hmul(const Fq2 &x, const Fq2 &y) {
    W16 c0, c1;
    u32 neg = fq2_mul_wide_calls(c0, c1, x, y);
    w16_add_p_hi(c0, neg);
    Fq2 r;
    r.a = fq_reduce_call(c0);
    r.b = fq_reduce_call(c1);
    return r;
}
DEVFN Fq2 hsqr(const Fq2 &x) {
    Fq2 r;
    r.a = fq_reduce_call(fq_mul_wide_call(fadd(x.a, x.b), fsub(x.a, x.b)));
    r.b = fdbl(fq_reduce_call(fq_mul_wide_call(x.a, x.b)));
    return r;
}
DEVFN Fq2 hmul_sub_mul(const Fq2 &x, const Fq2 &y, const Fq2 &z, const Fq2 &w) {
    W16 c0, c1, d0, d1;
    u32 neg0 = fq2_mul_wide_calls(c0, c1, x, y);
    u32 negd = fq2_mul_wide_calls(d0, d1, z, w);
    c0.v[0] = sub_cc(c0.v[0], d0.v[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) c0.v[i] = subc_cc(c0.v[i], d0.v[i]);
    u32 s0 = subc(neg0, negd);
    w16_add_p_hi(c0, s0);
    u32 s1 = w16_sub_sign(c1, d1);
    w16_add_p_hi(c1, s1);
    Fq2 r;
    r.a = fq_reduce_call(c0);
    r.b = fq_reduce_call(c1);
    return r;
}
DEVFN Fq hmul(const Fq &x, const Fq &y) { return fmul(x, y); }
DEVFN Fq hsqr(const Fq &x) { return fsqr(x); }
DEVFN Fq hmul_sub_mul(const Fq &x, const Fq &y, const Fq &z, const Fq &w) { return fmul_sub_mul(x, y, z, w); }
#elif B200_G2_HOT_CALLS == 3 && defined(__CUDA_ARCH__)
// Variant 3: whole Fq2 products out of line with operands and result BY VALUE (registers, no stack): one copy of the
// 3-product/2-reduction body and one of the squaring, called 6 + 2 times per mixed addition; only the fused Y3
// (one call site) stays inlined.  Loop body ~1/4 of the inlined one (see tools/field_variants/sass_mix.py).
static __device__ __noinline__ Fq2 fq2_mul_call(Fq2 x, Fq2 y) { return fmul(x, y); }
static __device__ __noinline__ Fq2 fq2_sqr_call(Fq2 x) { return fsqr(x); }
DEVFN Fq2 hmul(const Fq2 &x, const Fq2 &y) { return fq2_mul_call(x, y); }
DEVFN Fq2 hsqr(const Fq2 &x) { return fq2_sqr_call(x); }
DEVFN Fq2 hmul_sub_mul(const Fq2 &x, const Fq2 &y, const Fq2 &z, const Fq2 &w) { return fmul_sub_mul(x, y, z, w); }
DEVFN Fq hmul(const Fq &x, const Fq &y) { return fmul(x, y); }
DEVFN Fq hsqr(const Fq &x) { return fsqr(x); }
DEVFN Fq hmul_sub_mul(const Fq &x, const Fq &y, const Fq &z, const Fq &w) { return fmul_sub_mul(x, y, z, w); }
#elif B200_G2_HOT_CALLS == 1
template <class F> HD F hmul(const F &x, const F &y) { return cmul(x, y); }
template <class F> HD F hsqr(const F &x) { return csqr(x); }
template <class F> HD F hmul_sub_mul(const F &x, const F &y, const F &z, const F &w) { return cmul_sub_mul(x, y, z, w); }
#else
template <class F> HD F hmul(const F &x, const F &y) { return fmul(x, y); }
template <class F> HD F hsqr(const F &x) { return fsqr(x); }
template <class F> HD F hmul_sub_mul(const F &x, const F &y, const F &z, const F &w) { return fmul_sub_mul(x, y, z, w); }
#endif

}  // namespace b200
