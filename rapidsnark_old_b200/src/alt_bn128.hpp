// Plain data types of the BN254 engine as the reference exposes them (depends/ffiasm/c/alt_bn128.hpp:8-60,
// curve.hpp:11-21, f2field.hpp:7-10).  All arithmetic lives behind the C-ABI; these are layouts only.
#ifndef B200_ALT_BN128_HPP
#define B200_ALT_BN128_HPP
#include <stdint.h>
#include <string>

namespace AltBn128 {

struct FrElement { uint64_t v[4]; };
struct F1Element { uint64_t v[4]; };
struct F2Element { F1Element a, b; };
struct G1PointAffine { F1Element x, y; };
struct G2PointAffine { F2Element x, y; };
struct G1Point { F1Element x, y, zz, zzz; };
struct G2Point { F2Element x, y, zz, zzz; };

// BN254 scalar field order r as 32 little-endian bytes (main_prover.cpp:36)
extern const uint8_t kFrPrime[32];

// decimal string of a Montgomery-form base-field element (RawFq::toString)
std::string f1ToString(const F1Element &e);
// decimal string of a NORMAL-form 32-byte little-endian integer (public signals, main_prover.cpp:85-93)
std::string le32ToString(const void *le32);

struct Engine {
    typedef AltBn128::FrElement FrElement;
    typedef AltBn128::F1Element F1Element;
    typedef AltBn128::F2Element F2Element;
    typedef AltBn128::G1PointAffine G1PointAffine;
    typedef AltBn128::G2PointAffine G2PointAffine;
    typedef AltBn128::G1Point G1Point;
    typedef AltBn128::G2Point G2Point;
    static Engine engine;
};

}  // namespace AltBn128
#endif
