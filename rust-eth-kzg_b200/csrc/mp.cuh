// Multi-precision building blocks for sm_100a: 32-bit limbs, IMAD carry chains.
//
// Every primitive below is ONE PTX instruction on the device (mad.lo.cc / madc.hi.cc / addc ...).
// The carry flag lives in the PTX condition-code register CC.CF; the statements are `asm volatile`
// so the front end keeps them in program order and never drops one link of a chain.
//
// For the no-GPU build container the same primitives have a host emulation with an explicit
// carry variable (EKZG_HOST_EMU): the Montgomery / curve code above this header is then
// bit-for-bit the program the device runs and can be checked on the CPU (tests/test_host_emu.py).
// The emulation is test scaffolding only -- no product entry point is compiled with it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EKZG_HD __host__ __device__ __forceinline__
#define EKZG_D __device__ __forceinline__
// big point-level routines are real calls on the device: ptxas compile time and code size stay sane,
// and the call overhead (operands through local memory) is <3% of the ~6k instructions of a point add
#define EKZG_HD_CALL static __host__ __device__ __noinline__
#else
#define EKZG_HD inline __attribute__((always_inline))
#define EKZG_D inline __attribute__((always_inline))
#define EKZG_HD_CALL static inline
#endif

namespace ekzg {

#if defined(__CUDA_ARCH__)

EKZG_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t d; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t d; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
EKZG_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
EKZG_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
EKZG_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
EKZG_D uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
EKZG_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
EKZG_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
EKZG_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

#else  // host emulation of the same instruction semantics (PTX ISA: CC.CF is carry for add, borrow for sub)

static thread_local uint32_t g_cf = 0;
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t emu_add3(uint32_t x, uint32_t y, uint32_t cin, bool setcc) {
    uint64_t s = (uint64_t)x + y + cin;
    if (setcc) g_cf = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_lo(a, b), c, 0, true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_lo(a, b), c, g_cf, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_hi(a, b), c, g_cf, true); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_lo(a, b), c, g_cf, false); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_hi(a, b), c, g_cf, false); }
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add3(a, b, g_cf, true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add3(a, b, g_cf, false); }
inline uint32_t emu_sub3(uint32_t x, uint32_t y, uint32_t bin, bool setcc) {
    uint64_t d = (uint64_t)x - y - bin;
    if (setcc) g_cf = (uint32_t)(d >> 63);
    return (uint32_t)d;
}
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub3(a, b, g_cf, true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub3(a, b, g_cf, false); }

#endif

}  // namespace ekzg
