/* ORACLE build shim: stands in for the reference's generated build/fr.hpp (see ../rawfield.hpp). */
#ifndef ORACLE_SHIM_FR_HPP
#define ORACLE_SHIM_FR_HPP
#include "../rawfield.hpp"
#define Fr_N64 4
typedef uint64_t FrRawElement[4];
ORACLE_RAW_CLASS(RawFr, Fr, 254)
#endif
