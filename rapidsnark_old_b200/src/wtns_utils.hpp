// wtns header loader - WtnsUtils::loadHeader of the reference (src/wtns_utils.hpp:7-21, wtns_utils.cpp:12-25).
#ifndef B200_WTNS_UTILS_HPP
#define B200_WTNS_UTILS_HPP
#include <memory>
#include <vector>
#include "binfile_utils.hpp"

namespace WtnsUtils {

class Header {
public:
    uint32_t n8;
    std::vector<uint8_t> prime;
    uint32_t nVars;
};

std::unique_ptr<Header> loadHeader(BinFileUtils::BinFile *f);

}  // namespace WtnsUtils
#endif
