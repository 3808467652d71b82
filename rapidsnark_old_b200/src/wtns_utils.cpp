#include "wtns_utils.hpp"

namespace WtnsUtils {

std::unique_ptr<Header> loadHeader(BinFileUtils::BinFile *f) {
    std::unique_ptr<Header> h(new Header());
    f->startReadSection(1);
    h->n8 = f->readU32LE();
    const uint8_t *p = (const uint8_t *)f->read(h->n8);
    h->prime.assign(p, p + h->n8);
    h->nVars = f->readU32LE();
    f->endReadSection();
    return h;
}

}  // namespace WtnsUtils
