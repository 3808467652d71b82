// Host-side group helpers on the fast 4x64 field: Horner over MSM window sums, and the blinding /
// finalisation of a Groth16 proof (src/groth16.cpp:209-253).  g++ only.
#pragma once
#include "curve.cuh"
#include "hostfield.hpp"

namespace b200 {

typedef Fq2T<HFq> HFq2;
typedef Xyzz<HFq> HG1;
typedef Xyzz<HFq2> HG2;
typedef Affine<HFq> HG1Affine;
typedef Affine<HFq2> HG2Affine;

template <class F>
inline void horner(const Xyzz<F> *win, int nwin, int c, Xyzz<F> *out) {
    Xyzz<F> r = win[nwin - 1];
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) r = ec_dbl(r);
        ec_add(r, win[w]);
    }
    *out = r;
}

// k * P, k little-endian bytes
template <class F>
inline Xyzz<F> scalar_mul(const Affine<F> &p, const uint8_t *k, int nbytes) {
    Xyzz<F> base = Xyzz<F>::from_affine(p), r = Xyzz<F>::zero();
    int top = nbytes * 8 - 1;
    while (top >= 0 && !((k[top >> 3] >> (top & 7)) & 1)) top--;
    (void)base;
    for (int i = top; i >= 0; i--) {
        r = ec_dbl(r);
        if ((k[i >> 3] >> (i & 7)) & 1) ec_madd(r, p);      // mixed addition: the base is affine
    }
    return r;
}

// k1 * P + k2 * Q in ONE double-and-add pass (Shamir's trick: the doublings are shared, the addend of each step is
// P, Q or the precomputed P + Q): the proof's s * pi_a + r * pib1 (groth16.cpp:236-240), which sits on the critical
// path of every proof after the last MSM result has arrived
template <class F>
inline Xyzz<F> double_scalar_mul(const Affine<F> &p, const uint8_t *k1, const Affine<F> &q, const uint8_t *k2, int nbytes) {
    Xyzz<F> pq = Xyzz<F>::from_affine(p);
    ec_madd(pq, q);
    Affine<F> t[4];
    t[1] = p; t[2] = q; t[3] = ec_to_affine(pq);          // (0,0) when P = -Q: skipped by the mixed addition
    Xyzz<F> r = Xyzz<F>::zero();
    int top = nbytes * 8 - 1;
    while (top >= 0 && !(((k1[top >> 3] | k2[top >> 3]) >> (top & 7)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        r = ec_dbl(r);
        int idx = ((k1[i >> 3] >> (i & 7)) & 1) | (((k2[i >> 3] >> (i & 7)) & 1) << 1);
        if (idx) ec_madd(r, t[idx]);
    }
    return r;
}

}  // namespace b200
