// Short-Weierstrass a = 0 curve points in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), written
// once for G1 (F = Fq) and G2 (F = Fq2).
//
// Reference: ffiasm/c/curve.hpp:11-21 (Point{x,y,zz,zzz}, PointAffine{x,y}), curve.cpp:88-164 (add),
// :182-248 (mixed add), :337-394 (dbl), :408-456 (dbl of an affine point), :529-537 (zero tests:
// XYZZ zero <=> zz == 0, affine zero <=> (0,0)), :562-574 (to affine).  The formulas are the EFD
// shortw-xyzz ones the reference names (add-2008-s, madd-2008-s, dbl-2008-s-1, mdbl-2008-s-1);
// representatives may differ from the reference's, the affine value never does.
#pragma once
#include "fq2.cuh"

namespace b200 {

template <class F>
struct alignas(16) Affine {
    F x, y;
    HD bool is_zero() const { return x.is_zero() && y.is_zero(); }
};

template <class F>
struct alignas(16) Xyzz {
    F x, y, zz, zzz;
    HD bool is_zero() const { return zz.is_zero(); }
    HD static Xyzz zero() {
        Xyzz r;
        r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    HD static Xyzz from_affine(const Affine<F> &p) {
        Xyzz r;
        if (p.is_zero()) return zero();
        r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
        return r;
    }
};

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef Xyzz<Fq> G1Xyzz;
typedef Xyzz<Fq2> G2Xyzz;

// 2 * (affine p), p != 0     (mdbl-2008-s-1: 2M + 3S... here 3M + 3S with the W*y product)
template <class F>
HD_COLD Xyzz<F> ec_dbl_affine(const Affine<F> &p) {
    Xyzz<F> r;
    F u = fdbl(p.y);
    F v = csqr(u);
    F w = cmul(u, v);
    F s = cmul(p.x, v);
    F xx = csqr(p.x);
    F m = fadd(fdbl(xx), xx);
    r.x = fsub(csqr(m), fdbl(s));
    r.y = cmul_sub_mul(m, fsub(s, r.x), w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * p   (dbl-2008-s-1, a = 0)
template <class F>
HD_COLD Xyzz<F> ec_dbl(const Xyzz<F> &p) {
    if (p.is_zero()) return p;
    Xyzz<F> r;
    F u = fdbl(p.y);
    F v = csqr(u);
    F w = cmul(u, v);
    F s = cmul(p.x, v);
    F xx = csqr(p.x);
    F m = fadd(fdbl(xx), xx);
    r.x = fsub(csqr(m), fdbl(s));
    r.y = cmul_sub_mul(m, fsub(s, r.x), w, p.y);
    r.zz = cmul(v, p.zz);
    r.zzz = cmul(w, p.zzz);
    return r;
}

// Accumulator held in registers.  ec_madd is written against this small accessor interface so the hot
// kernel can also keep the running bucket sum in shared memory (SmemAcc in msm.cuh) and hold only the
// temporaries of one mixed addition in registers.
template <class F>
struct RegAcc {
    Xyzz<F> p;
    HD F ld_x() const { return p.x; }
    HD F ld_y() const { return p.y; }
    HD F ld_zz() const { return p.zz; }
    HD F ld_zzz() const { return p.zzz; }
    HD void st_x(const F &v) { p.x = v; }
    HD void st_y(const F &v) { p.y = v; }
    HD void st_zz(const F &v) { p.zz = v; }
    HD void st_zzz(const F &v) { p.zzz = v; }
    HD void st_all(const Xyzz<F> &v) { p = v; }
    HD Xyzz<F> get() const { return p; }
};

// acc += q (q affine)   (madd-2008-s: 8M + 2S)
template <class Acc, class F>
HD void ec_madd_acc(Acc &acc, const Affine<F> &q) {
    if (q.is_zero()) return;
    F zz = acc.ld_zz();
    if (zz.is_zero()) {
        acc.st_x(q.x); acc.st_y(q.y); acc.st_zz(F::one()); acc.st_zzz(F::one());
        return;
    }
    F zzz = acc.ld_zzz();
    F x1 = acc.ld_x(), y1 = acc.ld_y();
    F p = fsub(hmul(q.x, zz), x1);
    F r = fsub(hmul(q.y, zzz), y1);
    if (p.is_zero()) {
        if (r.is_zero()) acc.st_all(ec_dbl_affine(q));   // same point
        else acc.st_all(Xyzz<F>::zero());                 // opposite points
        return;
    }
    F pp = hsqr(p);
    F ppp = hmul(p, pp);
    acc.st_zz(hmul(zz, pp));
    acc.st_zzz(hmul(zzz, ppp));
    F qq = hmul(x1, pp);
    F x3 = fsub(fsub(hsqr(r), ppp), fdbl(qq));
    acc.st_x(x3);
    acc.st_y(hmul_sub_mul(r, fsub(qq, x3), y1, ppp));
}

// Same addition with the affine operand behind an accessor too (ld_x / ld_y / is_zero / get): the experimental
// accumulation variant of msm.cuh stages the gathered points in shared memory and loads each coordinate where it is
// used, so neither the running sum nor the (current, prefetched) points occupy registers across the products.
template <class Acc, class Pt>
HD void ec_madd_acc_pt(Acc &acc, Pt &q) {
    typedef typename Pt::Field F;
    if (q.is_zero()) return;
    F zz = acc.ld_zz();
    if (zz.is_zero()) {
        acc.st_x(q.ld_x()); acc.st_y(q.ld_y()); acc.st_zz(F::one()); acc.st_zzz(F::one());
        return;
    }
    F p = fsub(hmul(q.ld_x(), zz), acc.ld_x());
    {
        F r = fsub(hmul(q.ld_y(), acc.ld_zzz()), acc.ld_y());
        if (p.is_zero()) {
            if (r.is_zero()) acc.st_all(ec_dbl_affine(q.get()));   // same point
            else acc.st_all(Xyzz<F>::zero());                       // opposite points
            return;
        }
        q.st_scratch(0, r);              // the point's own slot is dead from here on: park r (needed last) in it
    }
    F pp = hsqr(p);
    F ppp = hmul(p, pp);
    q.st_scratch(1, ppp);
    acc.st_zz(hmul(acc.ld_zz(), pp));
    acc.st_zzz(hmul(acc.ld_zzz(), ppp));
    F qq = hmul(acc.ld_x(), pp);
    F x3 = fsub(fsub(hsqr(q.ld_scratch(0)), q.ld_scratch(1)), fdbl(qq));
    acc.st_x(x3);
    acc.st_y(hmul_sub_mul(q.ld_scratch(0), fsub(qq, x3), acc.ld_y(), q.ld_scratch(1)));
}

template <class F>
HD void ec_madd(Xyzz<F> &acc, const Affine<F> &q) {
    RegAcc<F> a;
    a.p = acc;
    ec_madd_acc(a, q);
    acc = a.p;
}

// acc += q   (add-2008-s: 12M + 2S)
template <class F>
HD_COLD void ec_add(Xyzz<F> &acc, const Xyzz<F> &q) {
    if (q.is_zero()) return;
    if (acc.is_zero()) { acc = q; return; }
    F u1 = cmul(acc.x, q.zz);
    F u2 = cmul(q.x, acc.zz);
    F s1 = cmul(acc.y, q.zzz);
    F s2 = cmul(q.y, acc.zzz);
    F p = fsub(u2, u1);
    F r = fsub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = ec_dbl(acc);
        else acc = Xyzz<F>::zero();
        return;
    }
    F pp = csqr(p);
    F ppp = cmul(p, pp);
    F qq = cmul(u1, pp);
    F x3 = fsub(fsub(csqr(r), ppp), fdbl(qq));
    F y3 = cmul_sub_mul(r, fsub(qq, x3), s1, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = cmul(cmul(acc.zz, q.zz), pp);
    acc.zzz = cmul(cmul(acc.zzz, q.zzz), ppp);
}

template <class F>
HD Xyzz<F> ec_neg(const Xyzz<F> &p) {
    Xyzz<F> r = p;
    r.y = fneg(p.y);
    return r;
}

template <class F>
HD Affine<F> ec_neg(const Affine<F> &p) {
    Affine<F> r = p;
    r.y = fneg(p.y);
    return r;
}

// canonical affine value; infinity -> (0,0)   (curve.cpp:562-574)
template <class F>
HD Affine<F> ec_to_affine(const Xyzz<F> &p) {
    Affine<F> r;
    if (p.is_zero()) { r.x = F::zero(); r.y = F::zero(); return r; }
    F i = finv(fmul(p.zz, p.zzz));            // one inversion: 1/zz = i * zzz, 1/zzz = i * zz (same canonical values)
    r.x = fmul(p.x, fmul(i, p.zzz));
    r.y = fmul(p.y, fmul(i, p.zz));
    return r;
}

// k * p for a little-endian scalar of nwords 32-bit words (plain double-and-add; O(1) per proof, used
// by the host for blinding and by the device only for small bucket-index multipliers)
template <class F>
HD_COLD Xyzz<F> ec_mul(const Xyzz<F> &p, const u32 *k, int nwords) {
    Xyzz<F> r = Xyzz<F>::zero();
    int top = nwords * 32 - 1;
    while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        r = ec_dbl(r);
        if ((k[i >> 5] >> (i & 31)) & 1) ec_add(r, p);
    }
    return r;
}

}  // namespace b200
