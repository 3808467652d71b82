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

// Build-time variants (A/B measured on B200 at 2^20):
//   B200_KARATSUBA  0: the word-serial interleaved product (fp_mul_serial, 64 + 64 wide multiplies)      [default]
//                   1: 512-bit products through one Karatsuba level (48 wide multiplies) + separate reduction:
//                      12 % fewer multiplier-pipe cycles but 30 % more instructions and longer carry chains -
//                      measured 14 % SLOWER per proof (24.3 ms vs 21.4 ms), kept only as an experiment
//   B200_LAZY_PAIR  1: a*b - c*d shares one Montgomery reduction (Y3 of the mixed addition): -6 % per proof
#ifndef B200_KARATSUBA
#define B200_KARATSUBA 0
#endif
#ifndef B200_LAZY_PAIR
#define B200_LAZY_PAIR 1
#endif
//   B200_FAST_SQR   1: squaring = 36 wide multiplies (off-diagonal products once, doubled) + a separate Montgomery
//                      reduction (64) instead of the 128 of a general product: 2 of the 10 products of a G1 mixed
//                      addition are squarings (montgomerybuilder.js:134-141)
#ifndef B200_FAST_SQR
#define B200_FAST_SQR 1
#endif

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
    for (int j = 0; j < 8; j += 2) mul_wide(acc[j], acc[j + 1], a[j], bi);
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
HD Fp<P> fp_mul(const Fp<P> &a, const Fp<P> &b);

template <class P>
HD Fp<P> fp_mul_serial(const Fp<P> &a, const Fp<P> &b) {
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

// t[0..2N-1] = a[0..N-1] * b[0..N-1], plain product (N even).  Same even/odd accumulator idea as mont_round: E holds
// the 64-bit partial products whose low word sits at an even position, O those at odd positions (O[k] = position
// k+1), so every product is one IMAD.WIDE with an aligned addend; the two are added once at the end.  Row 0 has
// nothing to accumulate onto and is written as plain 64-bit products.
template <int N>
HD void mul_nxn(u32 *t, const u32 *a, const u32 *b) {
    u32 E[2 * N], O[2 * N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        mul_wide(E[j], E[j + 1], a[j], b[0]);
        mul_wide(O[j], O[j + 1], a[j + 1], b[0]);
    }
#pragma unroll
    for (int k = N; k < 2 * N; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 1; i < N; i++) {
        const u32 bi = b[i];
        // even limbs of a: first pair at position i; odd limbs: first pair at position i + 1
        u32 *Xe = (i & 1) ? O : E;
        const int se = (i & 1) ? i - 1 : i;
        u32 *Xo = (i & 1) ? E : O;
        const int so = (i & 1) ? i + 1 : i;
        Xe[se] = mad_lo_cc(a[0], bi, Xe[se]);
        Xe[se + 1] = madc_hi_cc(a[0], bi, Xe[se + 1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            Xe[se + j] = madc_lo_cc(a[j], bi, Xe[se + j]);
            Xe[se + j + 1] = madc_hi_cc(a[j], bi, Xe[se + j + 1]);
        }
        if (se + N < 2 * N) Xe[se + N] = addc(Xe[se + N], 0);
        Xo[so] = mad_lo_cc(a[1], bi, Xo[so]);
        Xo[so + 1] = madc_hi_cc(a[1], bi, Xo[so + 1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            Xo[so + j] = madc_lo_cc(a[j + 1], bi, Xo[so + j]);
            Xo[so + j + 1] = madc_hi_cc(a[j + 1], bi, Xo[so + j + 1]);
        }
        if (so + N < 2 * N) Xo[so + N] = addc(Xo[so + N], 0);
    }
    t[0] = E[0];
    t[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) t[k] = addc_cc(E[k], O[k - 1]);
    t[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2]);
}

#if B200_KARATSUBA
// t[0..15] = a[0..7] * b[0..7] for any 256-bit a, b: one level of Karatsuba over the 128-bit halves, 3 x 16 wide
// multiplies instead of 64 (the multiplier pipe is what bounds every kernel here; the extra additions run on the
// ALU pipe beside it).   a b = z0 + (zm - z0 - z2) 2^128 + z2 2^256,  zm = (a0 + a1)(b0 + b1) as a 9-limb value
HD void mul8x8(u32 *t, const u32 *a, const u32 *b) {
    u32 zm[9], sa[4], sb[4];
    mul_nxn<4>(t, a, b);
    mul_nxn<4>(t + 8, a + 4, b + 4);
    sa[0] = add_cc(a[0], a[4]);
    sa[1] = addc_cc(a[1], a[5]);
    sa[2] = addc_cc(a[2], a[6]);
    sa[3] = addc_cc(a[3], a[7]);
    const u32 ca = addc(0, 0);
    sb[0] = add_cc(b[0], b[4]);
    sb[1] = addc_cc(b[1], b[5]);
    sb[2] = addc_cc(b[2], b[6]);
    sb[3] = addc_cc(b[3], b[7]);
    const u32 cb = addc(0, 0);
    mul_nxn<4>(zm, sa, sb);
    const u32 ma = 0u - ca, mb = 0u - cb;          // the carries of the two sums: (sa + ca 2^128)(sb + cb 2^128)
    zm[4] = add_cc(zm[4], sb[0] & ma);
    zm[5] = addc_cc(zm[5], sb[1] & ma);
    zm[6] = addc_cc(zm[6], sb[2] & ma);
    zm[7] = addc_cc(zm[7], sb[3] & ma);
    zm[8] = addc(ca & cb, 0);
    zm[4] = add_cc(zm[4], sa[0] & mb);
    zm[5] = addc_cc(zm[5], sa[1] & mb);
    zm[6] = addc_cc(zm[6], sa[2] & mb);
    zm[7] = addc_cc(zm[7], sa[3] & mb);
    zm[8] = addc(zm[8], 0);
    zm[0] = sub_cc(zm[0], t[0]);                     // zm -= z0; zm -= z2: a0 b1 + a1 b0 >= 0
#pragma unroll
    for (int i = 1; i < 8; i++) zm[i] = subc_cc(zm[i], t[i]);
    zm[8] = subc(zm[8], 0);
    zm[0] = sub_cc(zm[0], t[8]);
#pragma unroll
    for (int i = 1; i < 8; i++) zm[i] = subc_cc(zm[i], t[8 + i]);
    zm[8] = subc(zm[8], 0);
    t[4] = add_cc(t[4], zm[0]);
#pragma unroll
    for (int i = 1; i < 9; i++) t[4 + i] = addc_cc(t[4 + i], zm[i]);
    t[13] = addc_cc(t[13], 0);
    t[14] = addc_cc(t[14], 0);
    t[15] = addc(t[15], 0);
}

// t = a^2 through the same 4 x 4 blocks: a0^2 + 2 a0 a1 2^128 + a1^2 2^256
HD void sqr8(u32 *t, const u32 *a) {
    u32 zm[8];
    mul_nxn<4>(t, a, a);
    mul_nxn<4>(t + 8, a + 4, a + 4);
    mul_nxn<4>(zm, a, a + 4);
    const u32 top = zm[7] >> 31;
#pragma unroll
    for (int i = 7; i > 0; i--) zm[i] = (zm[i] << 1) | (zm[i - 1] >> 31);
    zm[0] <<= 1;
    t[4] = add_cc(t[4], zm[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[4 + i] = addc_cc(t[4 + i], zm[i]);
    t[12] = addc_cc(t[12], top);
    t[13] = addc_cc(t[13], 0);
    t[14] = addc_cc(t[14], 0);
    t[15] = addc(t[15], 0);
}
#else
HD void mul8x8(u32 *t, const u32 *a, const u32 *b) { mul_nxn<8>(t, a, b); }
#if B200_FAST_SQR
// t[0..15] = a^2 with 36 wide multiplies instead of 64: the 28 products a_i a_j (i < j) once, doubled, plus the 8
// squares a_i^2 (the reference's generated squaring does the same on 4 x 64-bit limbs: montgomerybuilder.js:134-141).
// Same even/odd accumulators as mul_nxn: a product a_i a_j sits at word position i + j; those with i + j even are
// chained into E (E[k] = position k), the others into O (O[k] = position k + 1), so that within one row i the
// products j = i+1, i+3, .. (and j = i+2, i+4, ..) are adjacent, non-overlapping 64-bit values of ONE carry chain.
// The word above a chain's top only ever holds earlier end-of-chain carries (0, 1, 2), so one addc closes the chain.
HD void sqr8(u32 *t, const u32 *a) {
    u32 E[16], O[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < 7; i++) {
        {   // j = i+1, i+3, ..: odd positions 2i+1, 2i+3, .. -> O[2i], O[2i+2], ..
            const int s = 2 * i;
            O[s] = mad_lo_cc(a[i], a[i + 1], O[s]);
            O[s + 1] = madc_hi_cc(a[i], a[i + 1], O[s + 1]);
            int k = s + 2;
#pragma unroll
            for (int j = i + 3; j < 8; j += 2, k += 2) {
                O[k] = madc_lo_cc(a[i], a[j], O[k]);
                O[k + 1] = madc_hi_cc(a[i], a[j], O[k + 1]);
            }
            if (k < 16) O[k] = addc(O[k], 0);
        }
        if (i + 2 < 8) {   // j = i+2, i+4, ..: even positions 2i+2, 2i+4, .. -> E[2i+2], ..
            const int s = 2 * i + 2;
            E[s] = mad_lo_cc(a[i], a[i + 2], E[s]);
            E[s + 1] = madc_hi_cc(a[i], a[i + 2], E[s + 1]);
            int k = s + 2;
#pragma unroll
            for (int j = i + 4; j < 8; j += 2, k += 2) {
                E[k] = madc_lo_cc(a[i], a[j], E[k]);
                E[k + 1] = madc_hi_cc(a[i], a[j], E[k + 1]);
            }
            if (k < 16) E[k] = addc(E[k], 0);
        }
    }
    // off-diagonal sum = E + (O << 32); it is below 2^511, so doubling cannot overflow the 16 words
    t[0] = E[0];
    t[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 15; k++) t[k] = addc_cc(E[k], O[k - 1]);
    t[15] = addc(E[15], O[14]);
#pragma unroll
    for (int k = 15; k > 0; k--) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
    t[0] <<= 1;
    // + sum a_i^2 2^(64 i): one carry chain over all 16 words
    t[0] = mad_lo_cc(a[0], a[0], t[0]);
    t[1] = madc_hi_cc(a[0], a[0], t[1]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        t[2 * i] = madc_lo_cc(a[i], a[i], t[2 * i]);
        if (i < 7) t[2 * i + 1] = madc_hi_cc(a[i], a[i], t[2 * i + 1]);
        else t[2 * i + 1] = madc_hi(a[i], a[i], t[2 * i + 1]);
    }
}
#else
HD void sqr8(u32 *t, const u32 *a) { mul_nxn<8>(t, a, a); }
#endif
#endif

// t -= s over 16 limbs; when the difference is negative p * 2^256 is added back, so for |t - s| < p * 2^256 the
// result is a valid mont_reduce16 input congruent to t - s
template <class P>
HD void sub16_mod(u32 *t, const u32 *s) {
    t[0] = sub_cc(t[0], s[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) t[i] = subc_cc(t[i], s[i]);
    const u32 borrow = subc(0, 0);
    t[8] = add_cc(t[8], borrow & P::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) t[8 + i] = addc_cc(t[8 + i], borrow & P::mod(i));
    t[15] = addc(t[15], borrow & P::mod(7));
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
HD Fp<P> fp_mul(const Fp<P> &a, const Fp<P> &b) {
#if B200_KARATSUBA
    u32 t[16];
    detail::mul8x8(t, a.v, b.v);
    Fp<P> r;
    detail::mont_reduce16<P>(r.v, t);
    fp_reduce_once(r);
    return r;
#else
    return fp_mul_serial(a, b);
#endif
}

template <class P>
HD Fp<P> fp_sqr(const Fp<P> &a) {
#if B200_KARATSUBA || B200_FAST_SQR
    u32 t[16];
    detail::sqr8(t, a.v);
    Fp<P> r;
    detail::mont_reduce16<P>(r.v, t);
    fp_reduce_once(r);
    return r;
#else
    return fp_mul(a, a);
#endif
}

// a*b - c*d with ONE Montgomery reduction: the two 512-bit products are subtracted first (|a b - c d| < p^2 <
// p * 2^256).  Same canonical value as fp_sub(fp_mul(a, b), fp_mul(c, d)).
template <class P>
HD Fp<P> fp_mul_sub_mul(const Fp<P> &a, const Fp<P> &b, const Fp<P> &c, const Fp<P> &d) {
#if B200_LAZY_PAIR
    u32 t[16], s[16];
    detail::mul8x8(t, a.v, b.v);
    detail::mul8x8(s, c.v, d.v);
    detail::sub16_mod<P>(t, s);
    Fp<P> r;
    detail::mont_reduce16<P>(r.v, t);
    fp_reduce_once(r);
    return r;
#else
    return fp_sub(fp_mul(a, b), fp_mul(c, d));
#endif
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
