// Bit-exactness of the build-time product variants of csrc/field.cuh / fq2.cuh (B200_KARATSUBA, B200_LAZY_PAIR) against
// the word-serial Montgomery product, on the HOST build of the very same source (carry flag emulated, bigint.cuh).
//   g++ -O2 -std=c++17 -DB200_KARATSUBA=K -DB200_LAZY_PAIR=L -x c++ host_test.cpp -o t && ./t [iterations]
// Run for all four (K, L) by tests/test_host_cpu.py.
#include "../../rapidsnark_old_b200/csrc/curve.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
using namespace b200;
template <class F> static F rnd(std::mt19937_64 &g, int mode) {
    F r;
    for (int i = 0; i < 8; i++) r.v[i] = (u32)g();
    if (mode == 1) for (int i = 0; i < 8; i++) r.v[i] = 0xffffffffu;
    if (mode == 2) for (int i = 0; i < 8; i++) r.v[i] = (g() & 1) ? 0xffffffffu : 0;
    if (mode == 3) for (int i = 0; i < 8; i++) r.v[i] = (g() & 3) ? (u32)g() : 0xffffffffu;
    r.v[7] &= 0x3fffffffu;
    F m = F::modulus();
    bool ge = true;
    for (int i = 7; i >= 0; i--) { if (r.v[i] != m.v[i]) { ge = r.v[i] > m.v[i]; break; } }
    if (ge) r = fp_sub(r, m);
    if (mode == 4) { r = m; r.v[0] -= 1 + (u32)(g() & 3); }
    if (mode == 5) { r = F::zero(); r.v[0] = (u32)(g() & 3); }
    return r;
}
template <class F> int run(const char *name, long iters) {
    std::mt19937_64 g(12345);
    long bad = 0;
    for (long it = 0; it < iters; it++) {
        int ma = it < 1000 ? (it % 6) : (g() % 12 < 6 ? 0 : g() % 6), mb = it < 1000 ? ((it / 6) % 6) : (g() % 12 < 6 ? 0 : g() % 6);
        F a = rnd<F>(g, ma), b = rnd<F>(g, mb), c = rnd<F>(g, g() % 6), d = rnd<F>(g, g() % 6);
        if (fp_mul_serial(a, b) != fp_mul(a, b)) bad++;
        if (fp_mul_serial(a, a) != fp_sqr(a)) bad++;
        if (fp_sub(fp_mul_serial(a, b), fp_mul_serial(c, d)) != fp_mul_sub_mul(a, b, c, d)) bad++;
    }
    printf("%s: bad=%ld\n", name, bad);
    return bad != 0;
}
static Fq2 fq2_mul_ref(const Fq2 &x, const Fq2 &y) {
    Fq aa = fp_mul_serial(x.a, y.a), bb = fp_mul_serial(x.b, y.b);
    Fq2 r; r.a = fp_sub(aa, bb); r.b = fp_add(fp_mul_serial(x.a, y.b), fp_mul_serial(x.b, y.a)); return r;
}
int run2(long iters) {
    std::mt19937_64 g(777);
    long bad = 0;
    for (long it = 0; it < iters; it++) {
        Fq2 x, y, z, w;
        x.a = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6); x.b = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6);
        y.a = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6); y.b = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6);
        z.a = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6); z.b = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6);
        w.a = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6); w.b = rnd<Fq>(g, g() % 12 < 6 ? 0 : g() % 6);
        if (it % 5 == 0) { z = x; w = y; }
        if (it % 7 == 0) { z = y; }
        Fq2 m = fq2_mul_ref(x, y);
        if (fmul(x, y) != m) bad++;
        if (fsqr(x) != fq2_mul_ref(x, x)) bad++;
        if (fmul_sub_mul(x, y, z, w) != fsub(m, fq2_mul_ref(z, w))) bad++;
    }
    printf("Fq2: bad=%ld\n", bad);
    return bad != 0;
}
// ec_madd_acc_pt (the accessor form used by the experimental shared-memory-staged accumulation kernel, which
// parks r and ppp in the consumed point's slot) against ec_madd, general and special cases, G1 and G2 fields
template <class F> struct HostPoint {
    typedef F Field;
    F slot[2];
    bool neg;
    F ld_x() const { return slot[0]; }
    F ld_y() const { return neg ? fneg(slot[1]) : slot[1]; }
    bool is_zero() const { return slot[0].is_zero() && slot[1].is_zero(); }
    Affine<F> get() const { Affine<F> a; a.x = ld_x(); a.y = ld_y(); return a; }
    void st_scratch(int i, const F &v) { slot[i] = v; }
    F ld_scratch(int i) const { return slot[i]; }
};
static Fq rndf(std::mt19937_64 &g, const Fq *) { return rnd<Fq>(g, g() % 12 < 8 ? 0 : g() % 6); }
static Fq2 rndf(std::mt19937_64 &g, const Fq2 *) { Fq2 r; r.a = rndf(g, (const Fq *)0); r.b = rndf(g, (const Fq *)0); return r; }
template <class F> int run_madd(const char *name, long iters) {
    std::mt19937_64 g(4242);
    long bad = 0;
    for (long it = 0; it < iters; it++) {
        Affine<F> q;
        q.x = rndf(g, (const F *)0); q.y = rndf(g, (const F *)0);
        Xyzz<F> a;
        a.x = rndf(g, (const F *)0); a.y = rndf(g, (const F *)0); a.zz = rndf(g, (const F *)0); a.zzz = rndf(g, (const F *)0);
        bool neg = g() & 1;
        switch (it % 8) {
            case 1: a = Xyzz<F>::zero(); break;                         // empty running sum
            case 2: q.x = F::zero(); q.y = F::zero(); break;            // point at infinity
            case 3: a = Xyzz<F>::from_affine(q); neg = false; break;    // same point: doubling
            case 4: a = Xyzz<F>::from_affine(q); a.y = fneg(a.y); neg = false; break;   // opposite points
            case 5: a = Xyzz<F>::from_affine(q); neg = true; break;     // opposite through the sign bit
            default: break;
        }
        Affine<F> qs = q;
        if (neg) qs.y = fneg(qs.y);
        Xyzz<F> want = a;
        ec_madd(want, qs);
        RegAcc<F> acc;
        acc.p = a;
        HostPoint<F> hp;
        hp.slot[0] = q.x; hp.slot[1] = q.y; hp.neg = neg;
        ec_madd_acc_pt(acc, hp);
        Xyzz<F> got = acc.get();
        if (got.x != want.x || got.y != want.y || got.zz != want.zz || got.zzz != want.zzz) bad++;
    }
    printf("%s: bad=%ld\n", name, bad);
    return bad != 0;
}

int main(int argc, char **argv) {
    long it = argc > 1 ? atol(argv[1]) : 1000000;
    return run<Fq>("Fq", it) | run<Fr>("Fr", it) | run2(it) | run_madd<Fq>("madd_pt G1", it / 4) | run_madd<Fq2>("madd_pt G2", it / 8);
}
