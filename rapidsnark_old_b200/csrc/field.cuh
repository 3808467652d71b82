// BN254 prime fields Fq / Fr as 8 x 32-bit little-endian limbs in Montgomery form, R = 2^256.
//
// Same bytes as the reference's 4 x 64-bit RawFq/RawFr elements, so zkey
// point tables and NTT data are consumed without conversion.  Semantics follow the reference's
// generated L0 routines: results are always fully reduced to [0,p)
//   mul  = <Name>_rawMMul   (ffiasm/src/montgomerybuilder.js:15-89, final canonical subtract :62-79)
//   add  = <Name>_rawAdd    (ffiasm/src/add.asm.ejs:210-236)
//   sub  = <Name>_rawSub    (ffiasm/src/sub.asm.ejs:288-305)
//   neg  = <Name>_rawNeg    (ffiasm/src/neg.asm.ejs:59-79)
// The multiplication is NOT the reference's 4-limb mulx/adcx CIOS: it is an 8-limb word-serial
// Montgomery product organised for the GPU's IMAD.WIDE pipe - two interleaved accumulators holding
// the even-aligned and odd-aligned 64-bit partial products, so no carry ripples between columns
// inside a round; the two are merged once at the end.
#pragma once
#include "bigint.cuh"

namespace b200 {

struct FqParams {
    static constexpr u32 INV = 0xe4866389u;  // -q^-1 mod 2^32
    HD static constexpr u32 mod(int i) {
        constexpr u32 t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    HD static constexpr u32 one(int i) {  // R mod q
        constexpr u32 t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    HD static constexpr u32 r2(int i) {  // R^2 mod q
        constexpr u32 t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return t[i];
    }
    HD static constexpr u32 r3(int i) {  // R^3 mod q
        constexpr u32 t[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
        return t[i];
    }
};

struct FrParams {
    static constexpr u32 INV = 0xefffffffu;  // -r^-1 mod 2^32
    HD static constexpr u32 mod(int i) {
        constexpr u32 t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    HD static constexpr u32 one(int i) {
        constexpr u32 t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    HD static constexpr u32 r2(int i) {
        constexpr u32 t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return t[i];
    }
    HD static constexpr u32 r3(int i) {
        constexpr u32 t[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return t[i];
    }
};

template <class P>
struct alignas(16) Fp {
    u32 v[8];
    typedef P Params;

    HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r;
    }
    HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
        return r;
    }
    HD static Fp r3() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::r3(i);
        return r;
    }
    HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
        return r;
    }
    HD bool is_zero() const {
        u32 t = v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) t |= v[i];
        return t == 0;
    }
    HD bool operator==(const Fp &o) const {
        u32 t = v[0] ^ o.v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) t |= v[i] ^ o.v[i];
        return t == 0;
    }
    HD bool operator!=(const Fp &o) const { return !(*this == o); }
};

// r = a - p if a >= p else a   (a < 2p)
template <class P>
HD void fp_reduce_once(Fp<P> &a) {
    u32 t[8];
    t[0] = sub_cc(a.v[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(a.v[i], P::mod(i));
    u32 borrow = subc(0, 0);  // 0 - 0 - borrow: 0xffffffff when a < p
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

template <class P>
HD Fp<P> fp_add(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    r.v[7] = addc(a.v[7], b.v[7]);  // p < 2^254: a+b < 2^255, no carry out
    fp_reduce_once(r);
    return r;
}

template <class P>
HD Fp<P> fp_sub(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
    u32 borrow = subc(0, 0);  // all-ones when a < b
    r.v[0] = add_cc(r.v[0], borrow & P::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], borrow & P::mod(i));
    r.v[7] = addc(r.v[7], borrow & P::mod(7));
    return r;
}

template <class P>
HD Fp<P> fp_neg(const Fp<P> &a) {
    Fp<P> r;
    r.v[0] = sub_cc(P::mod(0), a.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = subc_cc(P::mod(i), a.v[i]);
    r.v[7] = subc(P::mod(7), a.v[7]);
    bool z = a.is_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
    return r;
}

template <class P>
HD Fp<P> fp_dbl(const Fp<P> &a) {
    return fp_add(a, a);
}

namespace detail {

// acc[j], acc[j+1] = a[j] * bi for j = 0,2,4,6 (four independent 32x32->64 products)
HD void mul_pairs(u32 *acc, const u32 *a, u32 bi) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        acc[j] = mul_lo(a[j], bi);
        acc[j + 1] = mul_hi(a[j], bi);
    }
}

// acc += a_{0,2,4,6} * bi as one carry chain; carry out of acc[7] is left in the flag
HD void mad_pairs_cc(u32 *acc, const u32 *a, u32 bi) {
    acc[0] = mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        acc[j] = madc_lo_cc(a[j], bi, acc[j]);
        acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 1]);
    }
}

// acc = (acc >> 64) + a_{0,2,4,6} * bi, carry-in from the flag; top pair starts from zero
HD void mad_pairs_shift(u32 *acc, const u32 *a, u32 bi) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
        acc[j] = madc_lo_cc(a[j], bi, acc[j + 2]);
        acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 3]);
    }
    acc[6] = madc_lo_cc(a[6], bi, 0);
    acc[7] = madc_hi(a[6], bi, 0);
}

// One Montgomery round: T = (T + a*bi + m*p) / 2^32 with T = lo + 2^32 * hi.
//   lo  : accumulator whose 64-bit pairs sit at word positions (0,1),(2,3),(4,5),(6,7)
//   pend: on entry the previous round's `lo` (word 0 is zero, words 1..7 still to be shifted down by
//         one word); on exit the accumulator with pairs at positions (1,2),(3,4),(5,6),(7,8).
// After the round the roles swap: `pend` is word-aligned at position 0 for the next round.
template <class P>
HD void mont_round(u32 *lo, u32 *pend, const u32 *a, u32 bi, bool first) {
    u32 mod[8];
#pragma unroll
    for (int i = 0; i < 8; i++) mod[i] = P::mod(i);
    if (first) {
        mul_pairs(pend, a + 1, bi);
        mul_pairs(lo, a, bi);
    } else {
        lo[0] = add_cc(lo[0], pend[1]);   // stray word at position 0 of the shifted accumulator
        mad_pairs_shift(pend, a + 1, bi); // carry of position 0 enters position 1 = pend[0]
        mad_pairs_cc(lo, a, bi);
        pend[7] = addc(pend[7], 0);       // carry out of position 7 -> position 8
    }
    u32 m = lo[0] * P::INV;
    mad_pairs_cc(pend, mod + 1, m);       // top carry cannot occur: T < 2^(256+32+2)
    mad_pairs_cc(lo, mod, m);
    pend[7] = addc(pend[7], 0);
}

}  // namespace detail

// Montgomery product a*b*2^-256 mod p, fully reduced.
template <class P>
HD Fp<P> fp_mul(const Fp<P> &a, const Fp<P> &b) {
    u32 even[8], odd[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        detail::mont_round<P>(even, odd, a.v, b.v[i], i == 0);
        detail::mont_round<P>(odd, even, a.v, b.v[i + 1], false);
    }
    Fp<P> r;
    r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(even[i], odd[i + 1]);
    r.v[7] = addc(even[7], 0);
    fp_reduce_once(r);
    return r;
}

// ---- separated product / reduction (used by the lazily reduced Fq2 product in fq2.cuh) ----------------------
namespace detail {

// t[0..15] = a[0..7] * b[0..7], plain 512-bit product.  Same even/odd accumulator idea as mont_round: E holds the
// 64-bit partial products whose low word sits at an even position, O those at odd positions (O[k] = position
// k+1), so every product is one IMAD.WIDE with an aligned addend; the two are added once at the end.
HD void mul8x8(u32 *t, const u32 *a, const u32 *b) {
    u32 E[16], O[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 bi = b[i];
        // even limbs of a: first pair at position i; odd limbs: first pair at position i + 1
        u32 *Xe = (i & 1) ? O : E;
        const int se = (i & 1) ? i - 1 : i;
        u32 *Xo = (i & 1) ? E : O;
        const int so = (i & 1) ? i + 1 : i;
        Xe[se] = mad_lo_cc(a[0], bi, Xe[se]);
        Xe[se + 1] = madc_hi_cc(a[0], bi, Xe[se + 1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            Xe[se + j] = madc_lo_cc(a[j], bi, Xe[se + j]);
            Xe[se + j + 1] = madc_hi_cc(a[j], bi, Xe[se + j + 1]);
        }
        if (se + 8 < 16) Xe[se + 8] = addc(Xe[se + 8], 0);
        Xo[so] = mad_lo_cc(a[1], bi, Xo[so]);
        Xo[so + 1] = madc_hi_cc(a[1], bi, Xo[so + 1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            Xo[so + j] = madc_lo_cc(a[j + 1], bi, Xo[so + j]);
            Xo[so + j + 1] = madc_hi_cc(a[j + 1], bi, Xo[so + j + 1]);
        }
        if (so + 8 < 16) Xo[so + 8] = addc(Xo[so + 8], 0);
    }
    t[0] = E[0];
    t[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 15; k++) t[k] = addc_cc(E[k], O[k - 1]);
    t[15] = addc(E[15], O[14]);
}

// Montgomery reduction of a 16-limb value t < p * 2^256: r = t * 2^-256 mod p, r < 2p before the final
// conditional subtraction done by the caller.  Only the low half needs the m_i * p rounds; the high half is
// added at the end (t_hi < p, reduced low half <= p).
template <class P>
HD void mont_reduce16(u32 *r, const u32 *t) {
    u32 mod[8];
#pragma unroll
    for (int i = 0; i < 8; i++) mod[i] = P::mod(i);
    u32 ev[8], od[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ev[i] = t[i];
    {
        u32 m = ev[0] * P::INV;
        mul_pairs(od, mod + 1, m);
        mad_pairs_cc(ev, mod, m);
        od[7] = addc(od[7], 0);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
        u32 *lo = (i & 1) ? od : ev;      // roles swap every round (see mont_round)
        u32 *pend = (i & 1) ? ev : od;
        lo[0] = add_cc(lo[0], pend[1]);
        u32 m = lo[0] * P::INV;
        mad_pairs_shift(pend, mod + 1, m);
        mad_pairs_cc(lo, mod, m);
        pend[7] = addc(pend[7], 0);
    }
    r[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) r[i] = addc_cc(ev[i], od[i + 1]);
    r[7] = addc(ev[7], 0);
    r[0] = add_cc(r[0], t[8]);
#pragma unroll
    for (int i = 1; i < 7; i++) r[i] = addc_cc(r[i], t[8 + i]);
    r[7] = addc(r[7], t[15]);
}

}  // namespace detail

template <class P>
HD Fp<P> fp_sqr(const Fp<P> &a) {
    return fp_mul(a, a);
}

template <class P>
HD Fp<P> fp_to_mont(const Fp<P> &a) {
    return fp_mul(a, Fp<P>::r2());
}

template <class P>
HD Fp<P> fp_from_mont(const Fp<P> &a) {
    Fp<P> o = Fp<P>::zero();
    o.v[0] = 1;
    return fp_mul(a, o);
}

// a^e for a 256-bit little-endian exponent (host-side use: inversion by Fermat, root tables)
template <class P>
HD Fp<P> fp_pow(const Fp<P> &a, const u32 *e, int nwords = 8) {
    Fp<P> r = Fp<P>::one();
    for (int i = nwords * 32 - 1; i >= 0; i--) {
        r = fp_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
    }
    return r;
}

// a^-1 = a^(p-2); inv(0) = 0.  (Reference: GMP mpz_invert then *R^3, fr.cpp.ejs:215-227 - same value.)
template <class P>
HD Fp<P> fp_inv(const Fp<P> &a) {
    u32 e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::mod(i);
    e[0] -= 2;  // both moduli end in ...47 / ...01: no borrow
    return fp_pow(a, e);
}

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

}  // namespace b200
