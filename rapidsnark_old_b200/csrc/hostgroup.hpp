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
    for (int i = top; i >= 0; i--) {
        r = ec_dbl(r);
        if ((k[i >> 3] >> (i & 7)) & 1) ec_add(r, base);
    }
    return r;
}

}  // namespace b200
