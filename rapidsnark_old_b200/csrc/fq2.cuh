// Fq2 = Fq[u]/(u^2 + 1) for the G2 tables (reference: ffiasm/c/f2field.cpp:69-172, non-residue -1 at
// ffiasm/c/alt_bn128.cpp:6).  Element = {a, b} = a + b*u, 64 bytes, both halves Montgomery.
#pragma once
#include "field.cuh"

namespace b200 {

struct alignas(16) Fq2 {
    Fq a, b;
    HD static Fq2 zero() { Fq2 r; r.a = Fq::zero(); r.b = Fq::zero(); return r; }
    HD static Fq2 one() { Fq2 r; r.a = Fq::one(); r.b = Fq::zero(); return r; }
    HD bool is_zero() const { return a.is_zero() && b.is_zero(); }
    HD bool operator==(const Fq2 &o) const { return a == o.a && b == o.b; }
    HD bool operator!=(const Fq2 &o) const { return !(*this == o); }
};

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

HD Fq2 fadd(const Fq2 &x, const Fq2 &y) { Fq2 r; r.a = fp_add(x.a, y.a); r.b = fp_add(x.b, y.b); return r; }
HD Fq2 fsub(const Fq2 &x, const Fq2 &y) { Fq2 r; r.a = fp_sub(x.a, y.a); r.b = fp_sub(x.b, y.b); return r; }
HD Fq2 fneg(const Fq2 &x) { Fq2 r; r.a = fp_neg(x.a); r.b = fp_neg(x.b); return r; }
HD Fq2 fdbl(const Fq2 &x) { Fq2 r; r.a = fp_dbl(x.a); r.b = fp_dbl(x.b); return r; }

// Karatsuba, 3 base-field products (f2field.cpp:93-112)
HD Fq2 fmul(const Fq2 &x, const Fq2 &y) {
    Fq aa = fp_mul(x.a, y.a);
    Fq bb = fp_mul(x.b, y.b);
    Fq s = fp_mul(fp_add(x.a, x.b), fp_add(y.a, y.b));
    Fq2 r;
    r.a = fp_sub(aa, bb);
    r.b = fp_sub(fp_sub(s, aa), bb);
    return r;
}

// complex squaring, 2 base-field products (f2field.cpp:114-126)
HD Fq2 fsqr(const Fq2 &x) {
    Fq ab = fp_mul(x.a, x.b);
    Fq2 r;
    r.a = fp_mul(fp_add(x.a, x.b), fp_sub(x.a, x.b));
    r.b = fp_dbl(ab);
    return r;
}

// inverse through the norm a^2 + b^2 (f2field.cpp:144-155)
HD Fq2 finv(const Fq2 &x) {
    Fq n = fp_inv(fp_add(fp_sqr(x.a), fp_sqr(x.b)));
    Fq2 r;
    r.a = fp_mul(x.a, n);
    r.b = fp_neg(fp_mul(x.b, n));
    return r;
}

}  // namespace b200
