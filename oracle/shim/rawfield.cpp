/* ORACLE build shim: storage for the Raw<Name>::field singletons (ffiasm/src/fr.cpp.ejs:272). */
#include "build/fq.hpp"
#include "build/fr.hpp"
RawFq RawFq::field;
RawFr RawFr::field;
