// The reference's engine surface - AltBn128::Engine with f1, f2, fr, g1, g2 members, the RawFq / RawFr / F2Field /
// Curve class shapes and the global F1, F2, Fr, G1, G2 objects (depends/ffiasm/c/alt_bn128.hpp:8-60, curve.hpp:11-121,
// f2field.hpp:7-40, src/fr.hpp.ejs:79-131) - implemented over the C-ABI of include/b200snark.h.
//
// This is INTEGRATION.md's inner seam: a caller written against the reference's templates, such as the reference's own
// src/groth16.cpp, compiles UNCHANGED against this header (put this directory before depends/ffiasm/c on the include
// path) and then runs
//     Curve::multiMulByScalar   ->  b200_msm_g1 / b200_msm_g2      (curve.hpp:118-121)
//     FFT<RawFr>::fft / ifft    ->  b200_ntt_fr                    (fft.hpp:24-25, in fft.hpp next to this file)
// on the GPU, while the O(1) group and field operations around them (blinding, to-affine, strings) go through the
// library's host arithmetic.  Only the members the reference's callers use are provided.
#ifndef B200_ENGINE_ALT_BN128_HPP
#define B200_ENGINE_ALT_BN128_HPP
#include <stdint.h>
#include <string.h>
#include <sys/types.h>
#include <stdexcept>
#include <string>
#include "b200snark.h"

namespace b200engine {

// one library context per process for the seam (created on first use; the reference's objects are global too)
inline b200_ctx *context() {
    static b200_ctx *ctx = nullptr;
    if (!ctx) {
        int dev = 0;
        if (const char *e = getenv("B200_DEVICE")) dev = atoi(e);
        if (b200_init(dev, &ctx) != B200_OK) throw std::runtime_error(std::string("b200_init: ") + b200_last_error(nullptr));
    }
    return ctx;
}

inline std::string le32_to_decimal(const void *le32) {
    uint64_t v[4];
    memcpy(v, le32, 32);
    char tmp[80];
    int n = 0;
    while (v[0] | v[1] | v[2] | v[3]) {
        unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--) {
            unsigned __int128 cur = (rem << 64) | v[i];
            v[i] = (uint64_t)(cur / 10);
            rem = cur % 10;
        }
        tmp[n++] = (char)('0' + (int)rem);
    }
    if (n == 0) tmp[n++] = '0';
    std::string s(tmp, tmp + n);
    return std::string(s.rbegin(), s.rend());
}

}  // namespace b200engine

// ---- prime fields: same element bytes as the reference (4 x u64, Montgomery form where the reference's are)
template <int WHICH>   // 0 = Fq (base field), 1 = Fr (scalar field)
class B200RawField {
public:
    struct Element { uint64_t v[4]; };

private:
    Element fZero, fOne, fR2, fRaw1;
    static void mulf(void *r, const void *a, const void *b) { WHICH ? b200_host_fr_mul(r, a, b) : b200_host_fq_mul(r, a, b); }

public:
    B200RawField() {
        memset(&fZero, 0, sizeof fZero);
        memset(&fRaw1, 0, sizeof fRaw1);
        fRaw1.v[0] = 1;
        // R^2 mod p; one = toMontgomery(1)
        static const uint64_t r2q[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
        static const uint64_t r2r[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
        memcpy(&fR2, WHICH ? r2r : r2q, 32);
        mulf(&fOne, &fRaw1, &fR2);
    }
    Element &zero() { return fZero; }
    Element &one() { return fOne; }
    void copy(Element &r, const Element &a) { r = a; }
    void add(Element &r, const Element &a, const Element &b) { WHICH ? b200_host_fr_add(&r, &a, &b) : b200_host_fq_add(&r, &a, &b); }
    void sub(Element &r, const Element &a, const Element &b) { WHICH ? b200_host_fr_sub(&r, &a, &b) : b200_host_fq_sub(&r, &a, &b); }
    void neg(Element &r, const Element &a) { WHICH ? b200_host_fr_neg(&r, &a) : b200_host_fq_neg(&r, &a); }
    void mul(Element &r, const Element &a, const Element &b) { mulf(&r, &a, &b); }
    void square(Element &r, const Element &a) { mulf(&r, &a, &a); }
    void inv(Element &r, const Element &a) { WHICH ? b200_host_fr_inv(&r, &a) : b200_host_fq_inv(&r, &a); }
    void div(Element &r, const Element &a, const Element &b) { Element t; inv(t, b); mul(r, a, t); }
    void toMontgomery(Element &r, const Element &a) { mulf(&r, &a, &fR2); }
    void fromMontgomery(Element &r, const Element &a) { mulf(&r, &a, &fRaw1); }
    bool isZero(const Element &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
    bool eq(const Element &a, const Element &b) { return memcmp(&a, &b, 32) == 0; }
    // decimal string of a Montgomery-form element (RawFq::toString, fr.cpp.ejs:202-213)
    std::string toString(const Element &a, uint32_t radix = 10) {
        if (radix != 10) throw std::invalid_argument("toString: radix 10 only");
        Element n;
        fromMontgomery(n, a);
        return b200engine::le32_to_decimal(&n);
    }
};
typedef B200RawField<0> RawFq;
typedef B200RawField<1> RawFr;

// ---- Fq2 (f2field.hpp:7-40): layout only plus what Curve<F2Field>::toString needs
template <typename BaseField>
class F2Field {
public:
    struct Element { typename BaseField::Element a, b; };
    BaseField F;
    F2Field() {}
    explicit F2Field(const std::string &) {}             // the reference passes the non-residue "-1"
    void mul(Element &r, const Element &x, const Element &y) { b200_host_fq2_mul(&r, &x, &y); }
    void square(Element &r, const Element &x) { b200_host_fq2_sqr(&r, &x); }
    std::string toString(const Element &e, uint32_t radix = 10) { return "(" + F.toString(e.a, radix) + "," + F.toString(e.b, radix) + ")"; }
};

// ---- short Weierstrass group in the reference's XYZZ coordinates (curve.hpp:11-21): G1 over RawFq, G2 over F2Field
template <typename BaseField>
class Curve {
public:
    struct Point { typename BaseField::Element x, y, zz, zzz; };
    struct PointAffine { typename BaseField::Element x, y; };

private:
    static constexpr bool G2 = sizeof(typename BaseField::Element) == 64;
    BaseField *Fp;
    BaseField own;

public:
    Curve() : Fp(&own) {}
    Curve(BaseField &aF, const std::string &, const std::string &, const std::string &, const std::string &) : Fp(&aF) {}
    BaseField &F() { return *Fp; }

    void add(Point &r, Point &a, Point &b) { Point t; G2 ? b200_host_g2_add(&t, &a, &b) : b200_host_g1_add(&t, &a, &b); r = t; }
    void add(Point &r, Point &a, PointAffine &b) { Point t; G2 ? b200_host_g2_madd(&t, &a, &b) : b200_host_g1_madd(&t, &a, &b); r = t; }
    void add(Point &r, PointAffine &a, Point &b) { add(r, b, a); }
    void neg(Point &r, Point &a) { Point t; G2 ? b200_host_g2_neg(&t, &a) : b200_host_g1_neg(&t, &a); r = t; }
    void sub(Point &r, Point &a, Point &b) { Point t; neg(t, b); add(r, a, t); }
    void dbl(Point &r, Point &a) { Point t; G2 ? b200_host_g2_dbl(&t, &a) : b200_host_g1_dbl(&t, &a); r = t; }
    void copy(Point &r, Point &a) { r = a; }
    void copy(PointAffine &r, PointAffine &a) { r = a; }
    void copy(PointAffine &r, Point &a) { G2 ? b200_host_g2_to_affine(&r, &a) : b200_host_g1_to_affine(&r, &a); }
    bool isZero(Point &a) { static const Point z = Point(); return memcmp(&a.zz, &z.zz, sizeof a.zz) == 0; }
    bool isZero(PointAffine &a) { static const PointAffine z = PointAffine(); return memcmp(&a, &z, sizeof a) == 0; }
    // scalar: little-endian integer of scalarSize bytes (exp.hpp:6-28; the NAF there computes the same point)
    void mulByScalar(Point &r, PointAffine &base, uint8_t *scalar, unsigned int scalarSize) {
        G2 ? b200_host_g2_mul(&r, &base, scalar, scalarSize) : b200_host_g1_mul(&r, &base, scalar, scalarSize);
    }
    void mulByScalar(Point &r, Point &base, uint8_t *scalar, unsigned int scalarSize) {
        PointAffine a;
        copy(a, base);
        mulByScalar(r, a, scalar, scalarSize);
    }
    // THE hot call (curve.hpp:118-121 -> multiexp.cpp:98-144): on the GPU
    void multiMulByScalar(Point &r, PointAffine *bases, uint8_t *scalars, unsigned int scalarSize, unsigned int n, unsigned int nThreads = 0) {
        (void)nThreads;
        b200_ctx *ctx = b200engine::context();
        int rc = G2 ? b200_msm_g2(ctx, bases, scalars, scalarSize, n, &r) : b200_msm_g1(ctx, bases, scalars, scalarSize, n, &r);
        if (rc != B200_OK) throw std::runtime_error(std::string("b200_msm: ") + b200_last_error(ctx));
    }
    std::string toString(Point &p, uint32_t radix = 10) {
        PointAffine a;
        copy(a, p);
        return "(" + Fp->toString(a.x, radix) + "," + Fp->toString(a.y, radix) + ")";
    }
};

namespace AltBn128 {

typedef RawFq::Element F1Element;
typedef F2Field<RawFq>::Element F2Element;
typedef RawFr::Element FrElement;
typedef Curve<RawFq>::Point G1Point;
typedef Curve<RawFq>::PointAffine G1PointAffine;
typedef Curve<F2Field<RawFq>>::Point G2Point;
typedef Curve<F2Field<RawFq>>::PointAffine G2PointAffine;

// the reference defines these in alt_bn128.cpp; here they are header-only
inline RawFq F1;
inline F2Field<RawFq> F2("-1");
inline RawFr Fr;
inline Curve<RawFq> G1(F1, "0", "3", "1", "2");
inline Curve<F2Field<RawFq>> G2;

class Engine {
public:
    typedef RawFq F1;
    typedef F2Field<RawFq> F2;
    typedef RawFr Fr;
    typedef Curve<RawFq> G1;
    typedef Curve<F2Field<RawFq>> G2;

    F1 f1;
    F2 f2;
    Fr fr;
    G1 g1;
    G2 g2;

    Engine() : f1(), f2("-1"), fr(), g1(f1, "0", "3", "1", "2"), g2() {}

    typedef F1::Element F1Element;
    typedef F2::Element F2Element;
    typedef Fr::Element FrElement;
    typedef G1::Point G1Point;
    typedef G1::PointAffine G1PointAffine;
    typedef G2::Point G2Point;
    typedef G2::PointAffine G2PointAffine;

    inline static Engine *instance() { static Engine e; return &e; }
    static Engine &engine;
};
inline Engine &Engine::engine = *Engine::instance();

}  // namespace AltBn128
#endif
