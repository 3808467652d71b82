#include "groth16.hpp"
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <string.h>
#include <sys/random.h>
#include <sstream>
#include <stdexcept>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "../../include/b200snark.h"

namespace AltBn128 {

const uint8_t kFrPrime[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                              0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

std::string f1ToString(const F1Element &e) {
    char buf[80];
    b200_fq_to_decimal(&e, buf);
    return std::string(buf);
}

std::string le32ToString(const void *le32) {
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

}  // namespace AltBn128

namespace Groth16 {

static void throwCtx(b200_ctx *ctx, const char *what) {
    throw std::runtime_error(std::string(what) + ": " + b200_last_error(ctx));
}

template <typename Engine>
Prover<Engine>::Prover(uint32_t _nVars, uint32_t _nPublic, uint32_t _domainSize, uint64_t nCoefs, void *_vk_alpha1,
                       void *_vk_beta1, void *_vk_beta2, void *_vk_delta1, void *_vk_delta2, void *coefs,
                       void *pointsA, void *pointsB1, void *pointsB2, void *pointsC, void *pointsH)
    : nVars(_nVars), nPublic(_nPublic), domainSize(_domainSize) {
    memcpy(&vk_alpha1, _vk_alpha1, sizeof vk_alpha1);
    memcpy(&vk_beta1, _vk_beta1, sizeof vk_beta1);
    memcpy(&vk_beta2, _vk_beta2, sizeof vk_beta2);
    memcpy(&vk_delta1, _vk_delta1, sizeof vk_delta1);
    memcpy(&vk_delta2, _vk_delta2, sizeof vk_delta2);
    memset(lastMsms, 0, sizeof lastMsms);
    // FFT ctor contract of the reference (groth16.hpp:94 builds FFT(2*domainSize); fft.cpp:70-72)
    if ((uint64_t)domainSize * 2 > (1ull << 28)) throw std::range_error("Domain size too big for the curve");

    int nGpus = 1, first = 0;
    if (const char *e = getenv("B200_GPUS")) nGpus = atoi(e) > 0 ? atoi(e) : 1;
    if (const char *e = getenv("B200_DEVICE")) first = atoi(e);
    int stride = 1;    // B200_DEVICE_STRIDE=0: all shards on one device (exercises the N-GPU host path on a 1-GPU box)
    if (const char *e = getenv("B200_DEVICE_STRIDE")) stride = atoi(e);
    // one host thread per GPU: context creation, section upload (staged through pinned buffers) and the device-side
    // CSR / table builds of the shards run side by side
    gpus.assign(nGpus, Gpu{nullptr, nullptr});
    std::vector<std::string> errs(nGpus);
    std::vector<int> codes(nGpus, B200_OK);
    const bool replicate = getenv("B200_REPLICATE_H") != nullptr;
    const bool timing = getenv("B200_TIMING") != nullptr;
    auto setup = [&](int g) {
        typedef std::chrono::steady_clock clk;
        auto t0 = clk::now();
        Gpu gp{nullptr, nullptr};
        if (b200_init(first + g * stride, &gp.ctx) != B200_OK) {
            errs[g] = std::string("b200_init: ") + b200_last_error(nullptr);
            codes[g] = B200_ERR_NO_GPU;
            return;
        }
        auto t1 = clk::now();
        // B200_PRECOMP=0|1 / B200_PRECOMP_C=<bits>: per-window tables on/off (on by default: they pay off from the
        // second proof on; the one-shot CLI turns them off itself)
        if (const char *e = getenv("B200_PRECOMP")) b200_set_option(gp.ctx, "precomp", atoi(e));
        if (const char *e = getenv("B200_PRECOMP_C")) b200_set_option(gp.ctx, "precomp_c", atoi(e));
        b200_zkey_desc d;
        d.n_vars = nVars; d.n_public = nPublic; d.domain_size = domainSize; d.n_coefs = nCoefs;
        // shards that run none of the three transform chains (polynomial i is built on GPU i % G) never read the coefficients
        d.coefs = (nGpus == 1 || replicate || g < 3) ? coefs : nullptr;
        d.points_a = pointsA; d.points_b1 = pointsB1; d.points_b2 = pointsB2;
        d.points_c = pointsC; d.points_h = pointsH;
        d.shard_index = (uint32_t)g; d.shard_count = (uint32_t)nGpus;
        d.shard_lo_num = d.shard_hi_num = d.shard_den = 0;
        int rc = b200_zkey_upload(gp.ctx, &d, &gp.zk);
        if (rc != B200_OK) {
            errs[g] = std::string("b200_zkey_upload: ") + b200_last_error(gp.ctx);
            codes[g] = rc;
            b200_free(gp.ctx);
            return;
        }
        gpus[g] = gp;
        if (timing) {
            auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "  gpu%d: context %.1f ms, zkey upload %.1f ms\n", g, ms(t0, t1), ms(t1, clk::now()));
        }
    };
    if (nGpus == 1) setup(0);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < nGpus; g++) th.emplace_back(setup, g);
        for (auto &t : th) t.join();
    }
    for (int g = 0; g < nGpus; g++) {
        if (codes[g] == B200_OK) continue;
        std::string msg = errs[g];
        int code = codes[g];
        for (auto &o : gpus) { if (o.zk) b200_zkey_free(o.zk); if (o.ctx) b200_free(o.ctx); }
        gpus.clear();
        if (code == B200_ERR_RANGE) throw std::range_error("Domain size too big for the curve");
        throw std::runtime_error(msg);
    }
}

template <typename Engine>
Prover<Engine>::~Prover() {
    for (auto &g : gpus) { b200_zkey_free(g.zk); b200_free(g.ctx); }
}

template <typename Engine>
void Prover<Engine>::setBlinding(const uint8_t r32[32], const uint8_t s32[32]) {
    fixedRS = true;
    memcpy(fixedR, r32, 32);
    memcpy(fixedS, s32, 32);
}

template <typename Engine>
std::unique_ptr<Proof<Engine>> Prover<Engine>::prove(typename Engine::FrElement *wtns) {
    const size_t G = gpus.size();
    std::vector<uint8_t> parts(768 * G);
    std::vector<int> rcs(G, 0);
    // r, s: 31 random bytes each, top byte zero (groth16.cpp:213-217).  Drawn first: the part of the blinding that
    // needs only the key and r, s (r*delta1, s*delta1, rs*delta1, s*delta2 - three G1 and one G2 scalar
    // multiplication, most of the host work of a proof) runs on a host thread while the GPUs compute the MSMs.
    uint8_t r[32] = {0}, s[32] = {0};
    if (fixedRS) {
        memcpy(r, fixedR, 32);
        memcpy(s, fixedS, 32);
    } else {
        if (getrandom(r, 31, 0) != 31 || getrandom(s, 31, 0) != 31) throw std::runtime_error("getrandom failed");
    }
    if (G == 1) {
        // one GPU: the fused call does the blinding on this thread while the GPU works, in the order the five
        // results arrive (include/b200snark.h b200_groth16_prove)
        b200_vkey vk{&vk_alpha1, &vk_beta1, &vk_beta2, &vk_delta1, &vk_delta2};
        uint8_t out[256];
        if (b200_groth16_prove(gpus[0].ctx, gpus[0].zk, wtns, 0, &vk, r, s, out, lastMsms) != B200_OK)
            throwCtx(gpus[0].ctx, "b200_groth16_prove");
        lastPhases.clear();
        float ms[16];
        int k = b200_last_phase_ms(gpus[0].ctx, ms, 16);
        for (int i = 0; i < k; i++) lastPhases.emplace_back(b200_phase_name(i), ms[i]);
        std::unique_ptr<Proof<Engine>> p(new Proof<Engine>());
        memcpy(&p->A, out, 64);
        memcpy(&p->B, out + 64, 128);
        memcpy(&p->C, out + 192, 64);
        return p;
    }
    uint8_t prep[640];
    std::thread blind([&]() { b200_groth16_blind_prepare(&vk_delta1, &vk_delta2, r, s, prep); });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joinBlind{blind};
    if (getenv("B200_REPLICATE_H")) {     // every GPU repeats the whole H pipeline (A/B, no exchange)
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++)
            th.emplace_back([&, g]() { rcs[g] = b200_prove_msms(gpus[g].ctx, gpus[g].zk, wtns, parts.data() + 768 * g); });
        for (auto &t : th) t.join();
    } else {
        // One host thread per GPU.  Stage 1 on every GPU (witness upload, its four witness MSMs, the transform
        // chains of the polynomials it owns: polynomial i on GPU i % G), then ONE thread orders the device-to-device
        // copies of the three transformed polynomials on the H streams, then stage 2 (combine, H MSM, collection).
        std::mutex mu;
        std::condition_variable cv;
        size_t arrived = 0;
        int exchange_rc = -1;                     // -1: not done yet
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++)
            th.emplace_back([&, g]() {
                uint32_t mask = 0;
                for (unsigned i = 0; i < 3; i++) if (i % G == g) mask |= 1u << i;
                void *bufs[3], *hs = nullptr;
                rcs[g] = b200_prove_begin(gpus[g].ctx, gpus[g].zk, wtns, 0, mask, bufs, &hs);
                {
                    std::unique_lock<std::mutex> lk(mu);
                    if (++arrived == G) {           // last one in: every stage 1 is enqueued
                        bool all_ok = true;
                        for (size_t k = 0; k < G; k++) all_ok = all_ok && rcs[k] == B200_OK;
                        if (all_ok) {
                            std::vector<b200_ctx *> cs(G);
                            std::vector<b200_zkey *> zs(G);
                            for (size_t k = 0; k < G; k++) { cs[k] = gpus[k].ctx; zs[k] = gpus[k].zk; }
                            exchange_rc = b200_exchange_polys(cs.data(), zs.data(), (int)G);
                        } else {
                            exchange_rc = B200_ERR_ARG;
                        }
                        cv.notify_all();
                    } else {
                        cv.wait(lk, [&] { return exchange_rc != -1; });
                    }
                }
                // a shard whose stage 1 went through is always finished (its streams and slots are drained), even
                // when another shard or the exchange failed - the error is reported below
                if (rcs[g] == B200_OK) {
                    int rc = b200_prove_finish(gpus[g].ctx, gpus[g].zk, parts.data() + 768 * g);
                    rcs[g] = rc != B200_OK ? rc : (exchange_rc == B200_OK ? B200_OK : exchange_rc);
                }
            });
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; g++)
        if (rcs[g] != B200_OK) throwCtx(gpus[g].ctx, "b200_prove_msms");
    // fold the per-GPU partial sums (one group addition per result and GPU)
    memcpy(lastMsms, parts.data(), 768);
    for (size_t g = 1; g < G; g++) {
        const uint8_t *p = parts.data() + 768 * g;
        b200_host_g1_add(lastMsms, lastMsms, p);
        b200_host_g1_add(lastMsms + 128, lastMsms + 128, p + 128);
        b200_host_g1_add(lastMsms + 256, lastMsms + 256, p + 256);
        b200_host_g2_add(lastMsms + 384, lastMsms + 384, p + 384);
        b200_host_g1_add(lastMsms + 640, lastMsms + 640, p + 640);
    }
    lastPhases.clear();
    float ms[16];
    int k = b200_last_phase_ms(gpus[0].ctx, ms, 16);
    for (int i = 0; i < k; i++) lastPhases.emplace_back(b200_phase_name(i), ms[i]);

    blind.join();
    uint8_t out[256];
    b200_groth16_finalize_prepared(lastMsms, &vk_alpha1, &vk_beta1, &vk_beta2, prep, r, s, out);
    std::unique_ptr<Proof<Engine>> p(new Proof<Engine>());
    memcpy(&p->A, out, 64);
    memcpy(&p->B, out + 64, 128);
    memcpy(&p->C, out + 192, 64);
    return p;
}

template <typename Engine>
std::string Proof<Engine>::toJsonStr() {
    using AltBn128::f1ToString;
    std::ostringstream ss;
    ss << "{ \"pi_a\":[\"" << f1ToString(A.x) << "\",\"" << f1ToString(A.y) << "\",\"1\"], ";
    ss << " \"pi_b\": [[\"" << f1ToString(B.x.a) << "\",\"" << f1ToString(B.x.b) << "\"],[\"" << f1ToString(B.y.a)
       << "\",\"" << f1ToString(B.y.b) << "\"], [\"1\",\"0\"]], ";
    ss << " \"pi_c\": [\"" << f1ToString(C.x) << "\",\"" << f1ToString(C.y) << "\",\"1\"], ";
    ss << " \"protocol\":\"groth16\" }";
    return ss.str();
}

template <typename Engine>
std::string Proof<Engine>::toJson() {
    // nlohmann::json dump of groth16.cpp:268-301: compact, object keys in alphabetical order
    using AltBn128::f1ToString;
    std::ostringstream ss;
    ss << "{\"pi_a\":[\"" << f1ToString(A.x) << "\",\"" << f1ToString(A.y) << "\",\"1\"],";
    ss << "\"pi_b\":[[\"" << f1ToString(B.x.a) << "\",\"" << f1ToString(B.x.b) << "\"],[\"" << f1ToString(B.y.a)
       << "\",\"" << f1ToString(B.y.b) << "\"],[\"1\",\"0\"]],";
    ss << "\"pi_c\":[\"" << f1ToString(C.x) << "\",\"" << f1ToString(C.y) << "\",\"1\"],";
    ss << "\"protocol\":\"groth16\"}";
    return ss.str();
}

template <typename Engine>
std::unique_ptr<Prover<Engine>> makeProver(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                                           void *vk_alpha1, void *vk_beta1, void *vk_beta2, void *vk_delta1,
                                           void *vk_delta2, void *coefs, void *pointsA, void *pointsB1,
                                           void *pointsB2, void *pointsC, void *pointsH) {
    return std::unique_ptr<Prover<Engine>>(new Prover<Engine>(nVars, nPublic, domainSize, nCoefs, vk_alpha1, vk_beta1,
                                                              vk_beta2, vk_delta1, vk_delta2, coefs, pointsA, pointsB1,
                                                              pointsB2, pointsC, pointsH));
}

template class Proof<AltBn128::Engine>;
template class Prover<AltBn128::Engine>;
template std::unique_ptr<Prover<AltBn128::Engine>> makeProver<AltBn128::Engine>(uint32_t, uint32_t, uint32_t, uint64_t,
                                                                                  void *, void *, void *, void *, void *,
                                                                                  void *, void *, void *, void *, void *,
                                                                                  void *);

}  // namespace Groth16
