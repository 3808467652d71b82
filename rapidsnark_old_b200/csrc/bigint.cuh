// 32-bit carry-chain primitives for 8x32-limb field arithmetic.
//
// Device build: thin wrappers over the PTX carry-flag instructions (add.cc / addc / mad.lo.cc /
// madc.hi.cc ...).  ptxas fuses a {mad.lo.cc, madc.hi.cc} pair on the same operands into one
// IMAD.WIDE.U32 with carry, which is what makes the 8x8 limb product cost 64 wide multiplies.
// Host build (g++, used by the host-side prover for the O(1) blinding work and by the CPU unit
// tests of these very routines): the same functions emulated with an explicit carry variable, so
// the algorithms in field.cuh are testable without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEVFN __device__ __forceinline__
#define HD_COLD __host__ __device__ __noinline__   // cold group operations: one out-of-line copy per kernel image
#else
#define HD inline __attribute__((always_inline))
#define DEVFN inline __attribute__((always_inline))
#define HD_COLD inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;

namespace b200 {

#if defined(__CUDA_ARCH__)

DEVFN u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEVFN u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEVFN u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEVFN u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEVFN u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEVFN u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEVFN u32 mul_lo(u32 a, u32 b) { return a * b; }
DEVFN u32 mul_hi(u32 a, u32 b) { return __umulhi(a, b); }
DEVFN void mul_wide(u32 &lo, u32 &hi, u32 a, u32 b) { u64 p = (u64)a * b; lo = (u32)p; hi = (u32)(p >> 32); }
DEVFN u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEVFN u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEVFN u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEVFN u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEVFN u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation of the PTX carry flag

static thread_local u32 g_cf = 0;
inline u32 add_cc(u32 a, u32 b) { u64 t = (u64)a + b; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 addc_cc(u32 a, u32 b) { u64 t = (u64)a + b + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
inline u32 addc(u32 a, u32 b) { return a + b + g_cf; }
inline u32 sub_cc(u32 a, u32 b) { u64 t = (u64)a - b; g_cf = (u32)((t >> 32) & 1); return (u32)t; }  // g_cf = borrow
inline u32 subc_cc(u32 a, u32 b) { u64 t = (u64)a - b - g_cf; g_cf = (u32)((t >> 32) & 1); return (u32)t; }
inline u32 subc(u32 a, u32 b) { return a - b - g_cf; }
inline u32 mul_lo(u32 a, u32 b) { return a * b; }
inline u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline void mul_wide(u32 &lo, u32 &hi, u32 a, u32 b) { u64 p = (u64)a * b; lo = (u32)p; hi = (u32)(p >> 32); }
inline u32 mad_lo_cc(u32 a, u32 b, u32 c) { return add_cc(mul_lo(a, b), c); }
inline u32 madc_lo_cc(u32 a, u32 b, u32 c) { return addc_cc(mul_lo(a, b), c); }
inline u32 mad_hi_cc(u32 a, u32 b, u32 c) { return add_cc(mul_hi(a, b), c); }
inline u32 madc_hi_cc(u32 a, u32 b, u32 c) { return addc_cc(mul_hi(a, b), c); }
inline u32 madc_hi(u32 a, u32 b, u32 c) { return mul_hi(a, b) + c + g_cf; }

#endif

}  // namespace b200
