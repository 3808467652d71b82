#include "zkey_utils.hpp"
#include <stdexcept>

namespace ZKeyUtils {

// Sections read (reference: zkey_utils.cpp:17-52): 1 = protocol id (1 = groth16); 2 = field sizes and primes,
// circuit dimensions and the verification-key points; the coefficient count comes from the size of section 4.
std::unique_ptr<Header> loadHeader(BinFileUtils::BinFile *f) {
    {
        BinFileUtils::SectionReader proto(f, 1);
        if (proto.u32() != 1) throw std::invalid_argument("zkey file is not groth16");
        proto.finish();
    }
    auto hdr = std::make_unique<Header>();
    BinFileUtils::SectionReader rd(f, 2);
    hdr->n8q = rd.u32();
    hdr->qPrime = rd.bytes(hdr->n8q);
    hdr->n8r = rd.u32();
    hdr->rPrime = rd.bytes(hdr->n8r);
    hdr->nVars = rd.u32();
    hdr->nPublic = rd.u32();
    hdr->domainSize = rd.u32();
    const uint64_t g1 = (uint64_t)hdr->n8q * 2, g2 = (uint64_t)hdr->n8q * 4;   // affine point sizes
    hdr->vk_alpha1 = rd.raw(g1);
    hdr->vk_beta1 = rd.raw(g1);
    hdr->vk_beta2 = rd.raw(g2);
    hdr->vk_gamma2 = rd.raw(g2);
    hdr->vk_delta1 = rd.raw(g1);
    hdr->vk_delta2 = rd.raw(g2);
    rd.finish();
    hdr->nCoefs = f->getSectionSize(4) / (12 + hdr->n8r);   // 44-byte records; the leading u32 count is slack
    return hdr;
}

}  // namespace ZKeyUtils
