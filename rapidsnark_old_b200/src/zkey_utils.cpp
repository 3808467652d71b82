#include "zkey_utils.hpp"
#include <stdexcept>

namespace ZKeyUtils {

std::unique_ptr<Header> loadHeader(BinFileUtils::BinFile *f) {
    std::unique_ptr<Header> h(new Header());
    f->startReadSection(1);
    uint32_t protocol = f->readU32LE();
    if (protocol != 1) throw std::invalid_argument("zkey file is not groth16");
    f->endReadSection();

    f->startReadSection(2);
    h->n8q = f->readU32LE();
    const uint8_t *q = (const uint8_t *)f->read(h->n8q);
    h->qPrime.assign(q, q + h->n8q);
    h->n8r = f->readU32LE();
    const uint8_t *r = (const uint8_t *)f->read(h->n8r);
    h->rPrime.assign(r, r + h->n8r);
    h->nVars = f->readU32LE();
    h->nPublic = f->readU32LE();
    h->domainSize = f->readU32LE();
    h->vk_alpha1 = f->read(h->n8q * 2);
    h->vk_beta1 = f->read(h->n8q * 2);
    h->vk_beta2 = f->read(h->n8q * 4);
    h->vk_gamma2 = f->read(h->n8q * 4);
    h->vk_delta1 = f->read(h->n8q * 2);
    h->vk_delta2 = f->read(h->n8q * 4);
    f->endReadSection();

    h->nCoefs = f->getSectionSize(4) / (12 + h->n8r);
    return h;
}

}  // namespace ZKeyUtils
