#include "wtns_utils.hpp"

namespace WtnsUtils {

// wtns section 1: u32 n8 | prime (n8 bytes, little-endian) | u32 nVars   (reference: wtns_utils.cpp:12-25)
std::unique_ptr<Header> loadHeader(BinFileUtils::BinFile *f) {
    BinFileUtils::SectionReader rd(f, 1);
    auto hdr = std::make_unique<Header>();
    hdr->n8 = rd.u32();
    hdr->prime = rd.bytes(hdr->n8);
    hdr->nVars = rd.u32();
    rd.finish();
    return hdr;
}

}  // namespace WtnsUtils
