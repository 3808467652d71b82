// Groth16 prover front end with the reference's API (src/groth16.hpp:11-123):
//   Groth16::makeProver<Engine>(nVars, nPublic, domainSize, nCoefs, vk_alpha1, vk_beta1, vk_beta2,
//                               vk_delta1, vk_delta2, coefs, pointsA, pointsB1, pointsB2, pointsC, pointsH)
//   Prover::prove(FrElement *wtns) -> unique_ptr<Proof>;  Proof::{A, B, C, toJson(), toJsonStr()}
// The hot path (groth16.cpp:52-207) runs on the GPU(s) behind include/b200snark.h; the O(1) blinding and
// affine conversion (groth16.cpp:209-253) run on the host.  Unlike the reference the Prover copies the
// tables to HBM in makeProver, so the zkey file may be closed afterwards.
#ifndef B200_GROTH16_HPP
#define B200_GROTH16_HPP
#include <memory>
#include <string>
#include <vector>
#include "alt_bn128.hpp"

struct b200_ctx;
struct b200_zkey;

namespace Groth16 {

template <typename Engine>
class Proof {
public:
    typename Engine::G1PointAffine A;
    typename Engine::G2PointAffine B;
    typename Engine::G1PointAffine C;
    std::string toJsonStr();   // the reference's hand-formatted variant (groth16.cpp:256-266)
    std::string toJson();      // what `proofFile << proof->toJson()` writes: compact, keys sorted
};

template <typename Engine>
class Prover {
    struct Gpu { b200_ctx *ctx; b200_zkey *zk; };
    std::vector<Gpu> gpus;     // one point-range shard per GPU
    uint32_t nVars, nPublic, domainSize;
    typename Engine::G1PointAffine vk_alpha1, vk_beta1, vk_delta1;
    typename Engine::G2PointAffine vk_beta2, vk_delta2;
    bool fixedRS = false;
    uint8_t fixedR[32], fixedS[32];

public:
    Prover(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs, void *vk_alpha1, void *vk_beta1,
           void *vk_beta2, void *vk_delta1, void *vk_delta2, void *coefs, void *pointsA, void *pointsB1,
           void *pointsB2, void *pointsC, void *pointsH);
    ~Prover();
    Prover(const Prover &) = delete;
    Prover &operator=(const Prover &) = delete;

    std::unique_ptr<Proof<Engine>> prove(typename Engine::FrElement *wtns);
    // test hook: deterministic blinding factors (the reference draws 31 random bytes each, groth16.cpp:213-217)
    void setBlinding(const uint8_t r32[32], const uint8_t s32[32]);
    // the five pre-blinding MSM results of the last prove() (the reference's LOG_DEBUG taps, groth16.cpp:174-207)
    uint8_t lastMsms[768];
    // per-phase device milliseconds of GPU 0 for the last prove()
    std::vector<std::pair<std::string, float>> lastPhases;
    unsigned gpuCount() const { return (unsigned)gpus.size(); }
};

template <typename Engine>
std::unique_ptr<Prover<Engine>> makeProver(uint32_t nVars, uint32_t nPublic, uint32_t domainSize, uint64_t nCoefs,
                                           void *vk_alpha1, void *vk_beta1, void *vk_beta2, void *vk_delta1,
                                           void *vk_delta2, void *coefs, void *pointsA, void *pointsB1,
                                           void *pointsB2, void *pointsC, void *pointsH);

extern template class Proof<AltBn128::Engine>;
extern template class Prover<AltBn128::Engine>;

}  // namespace Groth16
#endif
