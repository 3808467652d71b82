/* ORACLE build shim: stands in for the reference's generated build/fq.hpp (see ../rawfield.hpp). */
#ifndef ORACLE_SHIM_FQ_HPP
#define ORACLE_SHIM_FQ_HPP
#include "../rawfield.hpp"
#define Fq_N64 4
typedef uint64_t FqRawElement[4];
ORACLE_RAW_CLASS(RawFq, Fq, 254)
#endif
