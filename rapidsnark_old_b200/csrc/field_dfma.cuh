// Montgomery product on the FP64 pipe: a*b*2^-256 mod p for canonical 8 x 32-bit operands, bit-identical to
// fp_mul (field.cuh), computed with DFMA instead of IMAD.WIDE.
//
// Why: on B200 the 32-bit multiplier pipe that fp_mul lives on issues 31 lane-ops/clk/SM and is 90 % busy in the
// accumulation kernel, while the FP64 pipe (59 DFMA lane-ops/clk/SM, tools/imad_ubench.cu) idles.  A product done
// here costs no IMAD slots at all, so a mixed addition that sends some of its ten products through this routine
// and the others through fp_mul runs on both pipes at once.
//
// How (the double-precision splitting trick of Emmart et al.): operands are cut into 52-bit limbs held as exact
// doubles; for limbs x, y < 2^52
//     h = fma_rz(x, y, 2^104)                 = 2^104 + 2^52 * floor(x*y / 2^52)       (exact: truncation, ulp 2^52)
//     l = fma_rz(x, y, (2^104 + 2^52) - h)    = 2^52  + (x*y mod 2^52)                  (exact)
// so the mantissa bits of h and l ARE the high and low halves of the 104-bit product.  The raw bit patterns are
// added as 64-bit integers into column accumulators (three-input adds), the exponent constants having been
// subtracted up front.  The Montgomery reduction runs in the same radix: four rounds of 52 bits and a last one of
// 48 bits (4 * 52 + 48 = 256), so R = 2^256 as everywhere else and stored values never change representation.
#pragma once
#include <string.h>
#include "field.cuh"

namespace b200 {
namespace dfma {

#if defined(__CUDA_ARCH__)
DEVFN double make_double(u32 hi, u32 lo) { return __hiloint2double((int)hi, (int)lo); }
DEVFN u64 bits(double d) { return (u64)__double_as_longlong(d); }
DEVFN double split_hi(double x, double y) { return __fma_rz(x, y, 0x1p104); }
DEVFN double split_lo(double x, double y, double h) { return __fma_rz(x, y, 0x1.0000000000001p104 - h); }
DEVFN u32 shr_pair(u32 lo, u32 hi, int s) { return __funnelshift_r(lo, hi, s); }
#else
inline double make_double(u32 hi, u32 lo) { u64 b = ((u64)hi << 32) | lo; double d; memcpy(&d, &b, 8); return d; }
inline u64 bits(double d) { u64 b; memcpy(&b, &d, 8); return b; }
// host model of the two roundings above on exact integers (x, y are integers below 2^53)
inline double split_hi(double x, double y) {
    unsigned __int128 p = (unsigned __int128)(u64)x * (u64)y;
    u64 hi = (u64)(p >> 52);
    return make_double(0x46700000u | (u32)(hi >> 32), (u32)hi);
}
inline double split_lo(double x, double y, double h) {
    unsigned __int128 p = (unsigned __int128)(u64)x * (u64)y;
    (void)h;
    u64 lo = (u64)p & ((1ull << 52) - 1);
    return make_double(0x43300000u | (u32)(lo >> 32), (u32)lo);
}
inline u32 shr_pair(u32 lo, u32 hi, int s) { return s == 0 ? lo : (u32)((((u64)hi << 32) | lo) >> s); }
#endif

// exact double of the 52-bit integer (hi20 : lo32)
HD double limb_to_double(u32 hi20, u32 lo) { return make_double(0x43300000u | hi20, lo) - 0x1p52; }

// canonical 8 x 32 words -> five 52-bit limbs (the last one holds bits 208..255) as exact doubles
HD void to_limbs(double *d, const u32 *w) {
    d[0] = limb_to_double(w[1] & 0xfffffu, w[0]);
    d[1] = limb_to_double(shr_pair(w[2], w[3], 20) & 0xfffffu, shr_pair(w[1], w[2], 20));
    d[2] = limb_to_double((w[4] >> 8) & 0xfffffu, shr_pair(w[3], w[4], 8));
    d[3] = limb_to_double(shr_pair(w[5], w[6], 28) & 0xfffffu, shr_pair(w[4], w[5], 28));
    d[4] = limb_to_double(w[7] >> 16, shr_pair(w[6], w[7], 16));
}

template <class P>
struct Consts;   // 52-bit limbs of the modulus and -p^-1 mod 2^52, as doubles
template <>
struct Consts<FqParams> {
    HD static double mod(int i) {
        const double t[5] = {0x8c16d87cfd47p0, 0x916871ca8d3c2p0, 0x181585d97816ap0, 0xa029b85045b68p0, 0x30644e72e131p0};
        return t[i];
    }
    HD static double np() { return 0x20782e4866389p0; }
};
template <>
struct Consts<FrParams> {
    HD static double mod(int i) {
        const double t[5] = {0x1f593f0000001p0, 0x4879b9709143ep0, 0x181585d2833e8p0, 0xa029b85045b68p0, 0x30644e72e131p0};
        return t[i];
    }
    HD static double np() { return 0x1f593efffffffp0; }
};

static const u64 BIAS_LO = 0x4330000000000000ull;   // bit pattern of 2^52  (exponent of every l)
static const u64 BIAS_HI = 0x4670000000000000ull;   // bit pattern of 2^104 (exponent of every h)
static const u64 M52 = (1ull << 52) - 1;

// col[k] += low half of x*y, col[k + 1] += high half (raw patterns; the biases are pre-subtracted by the caller)
HD void mac(u64 &c0, u64 &c1, double x, double y) {
    double h = split_hi(x, y);
    double l = split_lo(x, y, h);
    c0 += bits(l);
    c1 += bits(h);
}

}  // namespace dfma

// a * b * 2^-256 mod p, canonical in, canonical out; identical to fp_mul(a, b)
template <class P>
HD Fp<P> fp_mul_dfma(const Fp<P> &a, const Fp<P> &b) {
    using namespace dfma;
    double A[5], B[5];
    to_limbs(A, a.v);
    to_limbs(B, b.v);
    // column k (weight 2^(52 k)) collects the low halves of the limb products with i + j = k and the high halves of
    // those with i + j = k - 1; every pattern carries its exponent bias, subtracted here once per column.
    // Counts for the product: lo terms per column 1,2,3,4,5,4,3,2,1,0; hi terms 0,1,2,3,4,5,4,3,2,1.
    // Reduction round i (i < 4) adds q_i * p_j for j = 1..4 in full (lo into column i + j, hi into i + j + 1) and only
    // the HIGH half of q_i * p_0 (into column i + 1): the low half just cancels the column.  Round 4 adds all five.
    // Totals per column: lo terms 1,3,5,7,10,8,6,4,2,0 and hi terms 0,2,4,6,8,10,8,6,4,2.
    u64 col[10];
    {
        constexpr int n_lo[10] = {1, 3, 5, 7, 10, 8, 6, 4, 2, 0}, n_hi[10] = {0, 2, 4, 6, 8, 10, 8, 6, 4, 2};
#pragma unroll
        for (int k = 0; k < 10; k++) col[k] = 0ull - ((u64)n_lo[k] * BIAS_LO + (u64)n_hi[k] * BIAS_HI);
    }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) mac(col[i + j], col[i + j + 1], A[i], B[j]);

    const double np = Consts<P>::np();
#pragma unroll
    for (int i = 0; i < 5; i++) {
        // column i is final now (nothing below it is left): its low 52 (48) bits decide q_i
        const u64 v = col[i];
        const u64 t = v & M52;
        const double td = limb_to_double((u32)(t >> 32), (u32)t);
        double qh = split_hi(td, np);
        double ql = split_lo(td, np, qh);                       // 2^52 + (t * np mod 2^52)
        u64 qb = bits(ql);
        if (i == 4) qb &= ~(0xfull << 48);                      // last round clears 48 bits only: q mod 2^48
        const double q = make_double((u32)(qb >> 32), (u32)qb) - 0x1p52;
        if (i < 4) {
            // t + low(q p_0) = 0 or 2^52: the carry into the next column is (v >> 52) + (t != 0)
            col[i + 1] += (v >> 52) + (t != 0 ? 1u : 0u);
            col[i + 1] += bits(split_hi(q, Consts<P>::mod(0)));
#pragma unroll
            for (int j = 1; j < 5; j++) mac(col[i + j], col[i + j + 1], q, Consts<P>::mod(j));
        } else {
#pragma unroll
            for (int j = 0; j < 5; j++) mac(col[i + j], col[i + j + 1], q, Consts<P>::mod(j));
        }
    }
    // result = (col4 >> 48) + col5 2^4 + col6 2^56 + col7 2^108 + col8 2^160 + col9 2^212  (col4's low 48 bits are 0)
    // columns are < 2^58: make each hold 52 bits (column 4: the 4 bits above the 48 cleared ones)
    const u64 c = (col[4] >> 48) & 0xfull;
    u64 n5 = col[5] + (col[4] >> 52), n6, n7, n8, n9;
    n6 = col[6] + (n5 >> 52); n5 &= M52;
    n7 = col[7] + (n6 >> 52); n6 &= M52;
    n8 = col[8] + (n7 >> 52); n7 &= M52;
    n9 = col[9] + (n8 >> 52); n8 &= M52;
    // 64-bit words of the value; within each word the pieces occupy disjoint bits (n9 < 2^43: the result is < 2p)
    const u64 w0 = c | (n5 << 4) | (n6 << 56);           // bits 0..3 | 4..55 | 56..63 (low 8 bits of n6)
    const u64 w1 = (n6 >> 8) | (n7 << 44);               // 44 bits | low 20 bits of n7
    const u64 w2 = (n7 >> 20) | (n8 << 32);              // 32 bits | low 32 bits of n8
    const u64 w3 = (n8 >> 32) | (n9 << 20);              // 20 bits | n9
    Fp<P> r;
    r.v[0] = (u32)w0; r.v[1] = (u32)(w0 >> 32);
    r.v[2] = (u32)w1; r.v[3] = (u32)(w1 >> 32);
    r.v[4] = (u32)w2; r.v[5] = (u32)(w2 >> 32);
    r.v[6] = (u32)w3; r.v[7] = (u32)(w3 >> 32);
    fp_reduce_once(r);
    return r;
}

}  // namespace b200
