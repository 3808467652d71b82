// FFT<Field> with the reference's surface (depends/ffiasm/c/fft.hpp:8-31, fft.cpp:32-212): fft / ifft run on the GPU
// through b200_ntt_fr (natural order in and out, Montgomery data, in place), root(domainPow, idx) comes from a host
// table of w_{2^s}^i built like the reference's (w_{2^s} = 5^((r-1)/2^s), fft.cpp:52-102).  See alt_bn128.hpp here.
#ifndef B200_ENGINE_FFT_HPP
#define B200_ENGINE_FFT_HPP
#include <omp.h>     // the reference's fft.cpp brings it in; src/groth16.cpp relies on that for its omp_lock_t
#include <stdexcept>
#include <vector>
#include "alt_bn128.hpp"

template <typename Field>
class FFT {
    typedef typename Field::Element Element;
    Field f;
    uint32_t s;
    std::vector<Element> roots;

public:
    FFT(u_int64_t maxDomainSize, uint32_t nThreads = 0) {
        (void)nThreads;
        s = log2(maxDomainSize);
        if ((1ull << s) < maxDomainSize) s++;
        if (s > 28) throw std::range_error("Domain size too big for the curve");   // fft.cpp:70-72 (2-adicity of r)
        // w = 5^((r - 1) >> s): square-and-multiply over the bits of the exponent, host field arithmetic
        static const uint64_t rm1[4] = {0x43e1f593f0000000ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
        uint64_t e[4] = {rm1[0], rm1[1], rm1[2], rm1[3]};
        for (uint32_t k = 0; k < s; k++) {
            for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? (e[i + 1] << 63) : 0);
        }
        Element five, w = f.one();
        memset(&five, 0, sizeof five);
        five.v[0] = 5;
        f.toMontgomery(five, five);
        for (int i = 255; i >= 0; i--) {
            f.square(w, w);
            if ((e[i >> 6] >> (i & 63)) & 1) f.mul(w, w, five);
        }
        roots.resize((size_t)1 << s);
        roots[0] = f.one();
        for (size_t i = 1; i < roots.size(); i++) f.mul(roots[i], roots[i - 1], w);
    }
    u_int32_t log2(u_int64_t n) {
        u_int32_t r = 0;
        while (n > 1) { n >>= 1; r++; }
        return r;
    }
    void fft(Element *a, u_int64_t n) {
        b200_ctx *ctx = b200engine::context();
        if (b200_ntt_fr(ctx, a, n, 0) != B200_OK) throw std::runtime_error(std::string("b200_ntt_fr: ") + b200_last_error(ctx));
    }
    void ifft(Element *a, u_int64_t n) {
        b200_ctx *ctx = b200engine::context();
        if (b200_ntt_fr(ctx, a, n, 1) != B200_OK) throw std::runtime_error(std::string("b200_ntt_fr: ") + b200_last_error(ctx));
    }
    Element &root(u_int32_t domainPow, u_int64_t idx) { return roots[idx << (s - domainPow)]; }
};
#endif
