// Shared-memory-operand Fp arithmetic for the two hot kernels (K4 fixed-base MSM, K5 G1 NTTs).
//
// Why: the Montgomery multiplier is ~620 instructions, so a kernel can afford ONE copy of it (32 KB instruction cache,
// DESIGN.md §4.1).  Round 1 called that copy with operands in registers: every call then cost 36 register moves (issued
// as IMAD.MOV on the very pipe the multiplier saturates) plus spills of the caller's live point coordinates around it
// (ncu: 1.0-1.3 G local-memory sectors per launch, 255 registers, 2-3 warps per scheduler).  Here the operands of a
// thread live in SHARED MEMORY -- eight 48-byte slots per thread, 48 KB per 128-thread CTA, four CTAs per SM -- and the
// point formulas are straight-line programs (tools/fpvm_asm.py -> fpvm_programs.inc) executed by one interpreter,
// fpvm_run.  A call passes three integers; the multiplier reads its operands with LDS.128 and writes the result with
// STS.128 (LSU pipe, otherwise idle), nothing is moved or spilled, and the kernels fit 128 registers = 4 warps per
// scheduler.
//
// Layout: element `slot` of thread t is three 16-byte chunks at  base + (slot*3 + q)*NT*16 + t*16, q = 0..2, so the 32
// lanes of a warp touch 512 contiguous bytes per LDS.128: conflict-free.
//
// The interpreter is a template over the memory so that tests/host_emu runs the very same code on the CPU.
#pragma once
#include "g1.cuh"

namespace ekzg {
namespace fpvm {

constexpr int NT = 128;                  // threads per CTA of every kernel that uses the VM
constexpr int NSLOT = 8;
constexpr uint32_t QS = NT * 16;         // bytes between the three chunks of one element
constexpr uint32_t SLOT_BYTES = 3 * QS;  // all threads' copies of one slot
constexpr uint32_t SMEM_BYTES = NSLOT * SLOT_BYTES;   // 48 KB per CTA
constexpr uint32_t NONE = 15;
enum : uint32_t { OP_MUL = 0, OP_SQR = 1, OP_MUL2 = 2, OP_LIN = 3 };
enum : uint32_t { LIN_ADD = 0, LIN_SUB, LIN_DBL, LIN_TRI, LIN_QUAD, LIN_OCT, LIN_NEG, LIN_COPY };

#include "fpvm_programs.inc"

template <class P>
EKZG_HD uint32_t fe_or_limbs(const Fe<P>& a) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < P::N; j++) x |= a.v[j];
    return x;
}

// one instruction; returns 1 if the result is zero
template <class Mem>
EKZG_HD uint32_t step(Mem& m, uint32_t ins) {
    const uint32_t op = ins >> 28, fl = (ins >> 24) & 15u, d = (ins >> 20) & 15u;
    const uint32_t f1 = (ins >> 16) & 15u, f2 = (ins >> 12) & 15u, f3 = (ins >> 8) & 15u, f4 = (ins >> 4) & 15u, f5 = ins & 15u;
    Fp r;
    if (op == OP_MUL) {            // (f1 [- f2]) * f3 [- f4], doubled if flag bit 0
        Fp a = m.ld(f1);
        if (f2 != NONE) { Fp t = m.ld(f2); fe_sub(a, a, t); }
        Fp b = m.ld(f3);
        fe_mul_inline(r, a, b);
        if (f4 != NONE) { Fp c = m.ld(f4); fe_sub(r, r, c); }
        if (fl & 1u) fe_dbl(r, r);
    } else if (op == OP_SQR) {     // (f1 [+ f2])^2 [- f3] [- f4] [- f5]
        Fp a = m.ld(f1);
        if (f2 != NONE) { Fp t = m.ld(f2); fe_add(a, a, t); }
        fe_sqr_inline(r, a);
        if (f3 != NONE) { Fp c = m.ld(f3); fe_sub(r, r, c); }
        if (f4 != NONE) { Fp c = m.ld(f4); fe_sub(r, r, c); }
        if (f5 != NONE) { Fp c = m.ld(f5); fe_sub(r, r, c); }
    } else if (op == OP_MUL2) {    // f1 * (f2 - f3) - f4 * f5, one reduction
        Fp b = m.ld(f2);
        { Fp t = m.ld(f3); fe_sub(b, b, t); }
        Fp e = m.ld(f5);
        fe_neg(e, e);
        Fp a = m.ld(f1), c = m.ld(f4);
        fe_mul2_inline(r, a, b, c, e);
    } else {                       // linear
        Fp a = m.ld(f1);
        if (fl == LIN_ADD) { Fp b = m.ld(f2); fe_add(r, a, b); }
        else if (fl == LIN_SUB) { Fp b = m.ld(f2); fe_sub(r, a, b); }
        else if (fl == LIN_DBL) fe_dbl(r, a);
        else if (fl == LIN_TRI) { fe_dbl(r, a); fe_add(r, r, a); }
        else if (fl == LIN_QUAD) { fe_dbl(r, a); fe_dbl(r, r); }
        else if (fl == LIN_OCT) { fe_dbl(r, a); fe_dbl(r, r); fe_dbl(r, r); }
        else if (fl == LIN_NEG) fe_neg(r, a);
        else r = a;
    }
    m.st(d, r);
    return fe_or_limbs(r) == 0 ? 1u : 0u;
}

#if defined(__CUDACC__)
// 32-bit shared-space address of this thread's chunk 0 of slot 0
struct Smem {
    uint32_t base;
    EKZG_D Fp ld(uint32_t slot) const {
        Fp r;
        const uint32_t a = base + slot * SLOT_BYTES;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a) : "memory");
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "r"(a), "n"(QS) : "memory");
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(r.v[8]), "=r"(r.v[9]), "=r"(r.v[10]), "=r"(r.v[11]) : "r"(a), "n"(2 * QS) : "memory");
        return r;
    }
    EKZG_D void st(uint32_t slot, const Fp& v) const {
        const uint32_t a = base + slot * SLOT_BYTES;
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.v[0]), "r"(v.v[1]), "r"(v.v[2]), "r"(v.v[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0+%1], {%2,%3,%4,%5};" ::"r"(a), "n"(QS), "r"(v.v[4]), "r"(v.v[5]), "r"(v.v[6]), "r"(v.v[7]) : "memory");
        asm volatile("st.shared.v4.u32 [%0+%1], {%2,%3,%4,%5};" ::"r"(a), "n"(2 * QS), "r"(v.v[8]), "r"(v.v[9]), "r"(v.v[10]), "r"(v.v[11]) : "memory");
    }
};
#endif

#if !defined(__CUDACC__)
// host-side twin of the shared-memory slots (tests/host_emu)
struct HostMem {
    Fp s[16];
    Fp ld(uint32_t slot) const { return s[slot]; }
    void st(uint32_t slot, const Fp& v) { s[slot] = v; }
};
static inline uint32_t host_run(HostMem& m, const uint32_t* prog, int pc, int n, int reps = 1) {
    uint32_t z = 0;
    for (int r = 0; r < reps; r++)
        for (int i = 0; i < n; i++) z |= step(m, prog[pc + i]) << i;
    return z;
}
#endif

}  // namespace fpvm
}  // namespace ekzg
