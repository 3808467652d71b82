// The BN254 engine as the reference exposes it (depends/ffiasm/c/alt_bn128.hpp:8-60): AltBn128::Engine with its f1, f2,
// fr, g1, g2 members, the element / point types and the global F1, F2, Fr, G1, G2 objects come from engine/alt_bn128.hpp
// (implemented over the C-ABI: MSMs and transforms on the GPU, O(1) arithmetic on the host); this header adds the two
// small helpers the host prover uses.
#ifndef B200_ALT_BN128_HPP
#define B200_ALT_BN128_HPP
#include <stdint.h>
#include <string>
#include "engine/alt_bn128.hpp"

namespace AltBn128 {

// BN254 scalar field order r as 32 little-endian bytes (main_prover.cpp:36)
extern const uint8_t kFrPrime[32];

// decimal string of a Montgomery-form base-field element (RawFq::toString)
std::string f1ToString(const F1Element &e);
// decimal string of a NORMAL-form 32-byte little-endian integer (public signals, main_prover.cpp:85-93)
std::string le32ToString(const void *le32);

}  // namespace AltBn128
#endif
