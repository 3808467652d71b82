// Host-side (x86-64) BN254 field arithmetic, 4 x 64-bit limbs, Montgomery R = 2^256 - the same bytes as
// the 8 x 32-bit device representation and as the reference's RawFq/RawFr elements.  Used for the serial
// O(1)-per-proof group work that stays on the CPU (window Horner of the MSM, blinding and to-affine of
// groth16.cpp:209-253), where one core beats one GPU thread by ~8x.  The function names match fq2.cuh /
// curve.cuh so Xyzz<HFq> and Xyzz<Fq2T<HFq>> instantiate the same group formulas as the kernels.
#pragma once
#include <stdint.h>
#include <string.h>
#include "field.cuh"

namespace b200 {

template <class P>
struct alignas(16) HFp {
    uint64_t v[4];
    typedef P Params;
    static uint64_t modw(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
    static uint64_t inv64() {   // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
        uint64_t p0 = modw(0), x = (uint64_t)(0u - P::INV);   // x = p^-1 mod 2^32
        x *= 2 - p0 * x;                                      // mod 2^64
        return 0 - x;
    }
    static HFp zero() { HFp r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
    static HFp from_words(uint32_t (*f)(int)) {
        HFp r;
        for (int i = 0; i < 4; i++) r.v[i] = (uint64_t)f(2 * i) | ((uint64_t)f(2 * i + 1) << 32);
        return r;
    }
    static HFp one() { HFp r; for (int i = 0; i < 4; i++) r.v[i] = (uint64_t)P::one(2 * i) | ((uint64_t)P::one(2 * i + 1) << 32); return r; }
    static HFp r2() { HFp r; for (int i = 0; i < 4; i++) r.v[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32); return r; }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const HFp &o) const { return ((v[0] ^ o.v[0]) | (v[1] ^ o.v[1]) | (v[2] ^ o.v[2]) | (v[3] ^ o.v[3])) == 0; }
    bool operator!=(const HFp &o) const { return !(*this == o); }
};

template <class P>
inline bool hfp_geq_mod(const uint64_t *a) {
    for (int i = 3; i >= 0; i--) {
        uint64_t m = HFp<P>::modw(i);
        if (a[i] != m) return a[i] > m;
    }
    return true;
}

template <class P>
inline void hfp_sub_mod(uint64_t *a) {
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)a[i] - HFp<P>::modw(i) - (uint64_t)br;
        a[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
}

template <class P>
inline HFp<P> fadd(const HFp<P> &a, const HFp<P> &b) {
    HFp<P> r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (hfp_geq_mod<P>(r.v)) hfp_sub_mod<P>(r.v);
    return r;
}

template <class P>
inline HFp<P> fsub(const HFp<P> &a, const HFp<P> &b) {
    HFp<P> r;
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)a.v[i] - b.v[i] - (uint64_t)br;
        r.v[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    if (br) {
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (unsigned __int128)r.v[i] + HFp<P>::modw(i);
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}

template <class P>
inline HFp<P> fneg(const HFp<P> &a) {
    if (a.is_zero()) return a;
    HFp<P> m;
    for (int i = 0; i < 4; i++) m.v[i] = HFp<P>::modw(i);
    return fsub(m, a);
}

template <class P>
inline HFp<P> fdbl(const HFp<P> &a) { return fadd(a, a); }

template <class P>
inline HFp<P> fmul(const HFp<P> &a, const HFp<P> &b) {
    static const uint64_t inv = HFp<P>::inv64();
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * inv;
        c = (unsigned __int128)m * HFp<P>::modw(0) + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (unsigned __int128)m * HFp<P>::modw(j) + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    HFp<P> r;
    for (int i = 0; i < 4; i++) r.v[i] = t[i];
    if (t[4] || hfp_geq_mod<P>(r.v)) hfp_sub_mod<P>(r.v);
    return r;
}

template <class P>
inline HFp<P> fsqr(const HFp<P> &a) { return fmul(a, a); }
template <class P>
inline HFp<P> fmul_sub_mul(const HFp<P> &a, const HFp<P> &b, const HFp<P> &c, const HFp<P> &d) { return fsub(fmul(a, b), fmul(c, d)); }

template <class P>
inline HFp<P> finv(const HFp<P> &a) {   // a^(p-2)
    uint64_t e[4];
    for (int i = 0; i < 4; i++) e[i] = HFp<P>::modw(i);
    e[0] -= 2;
    HFp<P> r = HFp<P>::one();
    for (int i = 255; i >= 0; i--) {
        r = fsqr(r);
        if ((e[i >> 6] >> (i & 63)) & 1) r = fmul(r, a);
    }
    return r;
}

template <class P>
inline HFp<P> hfp_to_mont(const HFp<P> &a) { return fmul(a, HFp<P>::r2()); }
template <class P>
inline HFp<P> hfp_from_mont(const HFp<P> &a) { HFp<P> o = HFp<P>::zero(); o.v[0] = 1; return fmul(a, o); }

typedef HFp<FqParams> HFq;
typedef HFp<FrParams> HFr;

}  // namespace b200
