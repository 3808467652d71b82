// zkey header loader - ZKeyUtils::loadHeader of the reference (src/zkey_utils.hpp:9-34,
// src/zkey_utils.cpp:17-52).  The primes are kept as 32-byte little-endian arrays instead of GMP mpz_t.
#ifndef B200_ZKEY_UTILS_HPP
#define B200_ZKEY_UTILS_HPP
#include <memory>
#include <vector>
#include "binfile_utils.hpp"

namespace ZKeyUtils {

class Header {
public:
    uint32_t n8q;
    std::vector<uint8_t> qPrime;
    uint32_t n8r;
    std::vector<uint8_t> rPrime;
    uint32_t nVars;
    uint32_t nPublic;
    uint32_t domainSize;
    uint64_t nCoefs;
    void *vk_alpha1;
    void *vk_beta1;
    void *vk_beta2;
    void *vk_gamma2;
    void *vk_delta1;
    void *vk_delta2;
};

std::unique_ptr<Header> loadHeader(BinFileUtils::BinFile *f);

}  // namespace ZKeyUtils
#endif
