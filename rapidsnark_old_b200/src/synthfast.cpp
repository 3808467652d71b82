// Scalar side of the synthetic chain circuit of rapidsnark_old_b200/synth.py, in C++ on the
// host's 4 x 64 Montgomery field: the Python big-integer version takes 8 s at 2^20 and two minutes at 2^24, which
// is setup time of every bench / profiling process on the GPU box.  Same definitions, same values (the CPU tests
// compare the two byte for byte); the point tables are still made by the fixed-base GPU kernel from these scalars.
//
//   wires      w_0 = 1, w_1 given, w_{j+2} = (w_j + 3 w_{j+1}) w_{j+1}                 (V = n - 6 wires)
//   rows       j < V - 2:  A_j = w_j + 3 w_{j+1},  B_j = w_{j+1},  C_j = w_{j+2};   rows V-2+i, i <= P: A = w_i
//   L_j(tau)   = (tau^n - 1) omega^j / (n (tau - omega^j))
//   A(tau)_s, B(tau)_s, C(tau)_s  = the column sums of the rows above weighted by L_j(tau)
//   K_s        = beta A_s + alpha B_s + C_s;  IC_s = K_s / gamma (s <= P),  C-table scalar_s = K_s / delta (s > P)
//   H table_i  = (tau^n + 1) x_i / (-n (tau - x_i)) * (tau^n - 1) / (-2 delta),  x_i = w_{2n} omega^i
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/b200snark.h"
#include "../csrc/hostfield.hpp"

using namespace b200;

namespace {

HFr ld_norm(const void *p) {   // 32-byte little-endian integer (< r) -> Montgomery
    HFr x;
    memcpy(&x, p, 32);
    return hfp_to_mont(x);
}
void st_norm(void *p, const HFr &x) {
    HFr y = hfp_from_mont(x);
    memcpy(p, &y, 32);
}
HFr from_u64(uint64_t v) {
    HFr x = HFr::zero();
    x.v[0] = v;
    return hfp_to_mont(x);
}
HFr pow_u64(HFr b, uint64_t e) {
    HFr r = HFr::one();
    while (e) {
        if (e & 1) r = fmul(r, b);
        b = fsqr(b);
        e >>= 1;
    }
    return r;
}
// w_{2^k} = 5^((r-1) / 2^28) squared 28 - k times (fft.cpp:52-83)
HFr root_of_unity(int k) {
    uint64_t e[4];
    for (int i = 0; i < 4; i++) e[i] = HFr::modw(i);
    e[0] -= 1;
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 28) | (i < 3 ? e[i + 1] << 36 : 0);
    HFr five = from_u64(5), w = HFr::one();
    for (int i = 255; i >= 0; i--) {
        w = fsqr(w);
        if ((e[i >> 6] >> (i & 63)) & 1) w = fmul(w, five);
    }
    for (int i = 0; i < 28 - k; i++) w = fsqr(w);
    return w;
}
// v[i] <- 1 / v[i] (no zeros), Montgomery's trick
void batch_inverse(std::vector<HFr> &v) {
    const size_t n = v.size();
    std::vector<HFr> pref(n + 1);
    pref[0] = HFr::one();
    for (size_t i = 0; i < n; i++) pref[i + 1] = fmul(pref[i], v[i]);
    HFr inv = finv(pref[n]);
    for (size_t i = n; i-- > 0;) {
        HFr o = fmul(pref[i], inv);
        inv = fmul(inv, v[i]);
        v[i] = o;
    }
}

}  // namespace

extern "C" int b200_synth_chain(uint32_t log_n, uint32_t n_public, const void *tau32, const void *alpha32,
                                const void *beta32, const void *gamma32, const void *delta32, const void *w1_32,
                                void *wtns, void *a_tau, void *b_tau, void *c_scalars, void *ic_scalars, void *h_tbl,
                                void *dlogs96) {
    if (log_n < 4 || log_n > 27 || !tau32 || !alpha32 || !beta32 || !gamma32 || !delta32 || !w1_32 || !wtns || !a_tau ||
        !b_tau || !c_scalars || !ic_scalars || !h_tbl || !dlogs96)
        return B200_ERR_ARG;
    const size_t n = (size_t)1 << log_n, V = n - 6, P = n_public, ncons = V - 2;
    if (P + 1 >= V) return B200_ERR_ARG;
    const HFr tau = ld_norm(tau32), alpha = ld_norm(alpha32), beta = ld_norm(beta32), gamma = ld_norm(gamma32),
              delta = ld_norm(delta32);
    const HFr three = from_u64(3), nn = from_u64(n);

    // wires
    std::vector<HFr> w(V);
    w[0] = HFr::one();
    w[1] = ld_norm(w1_32);
    for (size_t j = 0; j < ncons; j++) w[j + 2] = fmul(fadd(w[j], fmul(three, w[j + 1])), w[j + 1]);
    for (size_t i = 0; i < V; i++) st_norm((uint8_t *)wtns + 32 * i, w[i]);

    // Lagrange basis of the plain domain at tau
    const HFr omega = root_of_unity((int)log_n);
    const HFr tau_n = pow_u64(tau, n);
    const HFr zt = fsub(tau_n, HFr::one());
    std::vector<HFr> L(n), ws(n);
    {
        HFr wj = HFr::one();
        for (size_t j = 0; j < n; j++) {
            ws[j] = wj;
            L[j] = fmul(nn, fsub(tau, wj));
            wj = fmul(wj, omega);
        }
        batch_inverse(L);
        for (size_t j = 0; j < n; j++) L[j] = fmul(fmul(zt, ws[j]), L[j]);
    }

    // column sums A, B, C at tau
    std::vector<HFr> A(V, HFr::zero()), B(V, HFr::zero()), C(V, HFr::zero());
    for (size_t j = 0; j < ncons; j++) {
        A[j] = fadd(A[j], L[j]);
        A[j + 1] = fadd(A[j + 1], fmul(three, L[j]));
        B[j + 1] = fadd(B[j + 1], L[j]);
        C[j + 2] = fadd(C[j + 2], L[j]);
    }
    for (size_t i = 0; i <= P; i++) A[i] = fadd(A[i], L[ncons + i]);
    const HFr dinv = finv(delta), ginv = finv(gamma);
    HFr ea = HFr::zero(), eb = HFr::zero(), pub = HFr::zero();
    for (size_t i = 0; i < V; i++) {
        st_norm((uint8_t *)a_tau + 32 * i, A[i]);
        st_norm((uint8_t *)b_tau + 32 * i, B[i]);
        ea = fadd(ea, fmul(w[i], A[i]));
        eb = fadd(eb, fmul(w[i], B[i]));
        const HFr K = fadd(fadd(fmul(beta, A[i]), fmul(alpha, B[i])), C[i]);
        if (i <= P) {
            st_norm((uint8_t *)ic_scalars + 32 * i, fmul(K, ginv));
            pub = fadd(pub, fmul(w[i], K));
        } else {
            st_norm((uint8_t *)c_scalars + 32 * (i - P - 1), fmul(K, dinv));
        }
    }
    st_norm(dlogs96, ea);
    st_norm((uint8_t *)dlogs96 + 32, eb);
    st_norm((uint8_t *)dlogs96 + 64, pub);

    // H table on the coset g * omega^i
    {
        const HFr g = root_of_unity((int)log_n + 1);
        const HFr neg_n = fneg(nn);
        std::vector<HFr> &xs = ws, &den = L;   // reuse the storage
        HFr x = g;
        for (size_t i = 0; i < n; i++) {
            xs[i] = x;
            den[i] = fmul(neg_n, fsub(tau, x));
            x = fmul(x, omega);
        }
        batch_inverse(den);
        const HFr tn1 = fadd(tau_n, HFr::one());
        const HFr hfac = fmul(zt, finv(fmul(fneg(from_u64(2)), delta)));
        const HFr k = fmul(tn1, hfac);
        for (size_t i = 0; i < n; i++) st_norm((uint8_t *)h_tbl + 32 * i, fmul(fmul(k, xs[i]), den[i]));
    }
    return B200_OK;
}
