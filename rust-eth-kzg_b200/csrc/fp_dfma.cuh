// Fp Montgomery multiplication on the FP64 pipe: 8 x 48-bit limbs held in doubles, every 48x48-bit limb product
// split exactly into its high and low half by two DFMA.RZ and one DADD, halves summed per column in 64-bit integers.
//
// Why (tools/pipe_probe.cu on B200, profiles/r2_pipe_probe.json): the integer multiplier issues IMAD.WIDE.U32(.X)
// at 32 lanes/clk/SM (one warp instruction per 4 clocks per sub-partition) whether or not it carries, and
// DFMA runs in the same heavy pipe at 58-64 lanes/clk/SM (one per 2 clocks), overlapping with neither -- but a DFMA
// moves a 48x48-bit half product where the IMAD.WIDE moves a 32x32-bit whole one.  Per Fp multiplication:
//   integer : 12*(12+12+1)          = 300 IMAD.WIDE.U32.X * 4 clk = 1200 pipe clocks per warp
//   FP64    : 3*(64 + 64) + 24 conv = 408 DFMA/DADD       * 2 clk =  816
// and the ~390 integer adds/shifts/permutes of the FP64 variant go to the ALU pipe (128 lanes/clk/SM on sm_100,
// overlaps with the heavy pipe across warps).  Technique: Emmart, Zheng, Weems, "Faster modular exponentiation using
// double precision floating point arithmetic on the GPU" (ARITH 2018), adapted to radix 2^48 so that R = 2^384
// and the packed 12 x 32-bit Montgomery form of field.cuh is kept: tables, constants and callers are unchanged.
//
// Exactness: for integers x, y < 2^49 held in doubles,  hi = fma_rz(x, y, 2^100)  is 2^100 + floor(xy/2^48)*2^48
// (the ulp in [2^100, 2^101) is 2^48 and RZ truncates);  sub = (2^100 + 2^52) - hi  is exact (a multiple of 2^48
// below 2^99);  lo = fma_rz(x, y, sub) = (xy mod 2^48) + 2^52  is exact.  The IEEE bit patterns are then
// 0x463<<52 | floor(xy/2^48)  and  0x433<<52 | (xy mod 2^48): plain integers plus known offsets, which the column
// accumulators are pre-loaded with in negated form.
//
// Host build (tests/host_emu): same code, fma_rz through fesetround(FE_TOWARDZERO) + fma().
#pragma once
#include "constants.cuh"
#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cfenv>
#include <cstring>
#endif

namespace ekzg {

namespace fpd {

constexpr int L = 8;
constexpr uint64_t M48 = 0xffffffffffffull;
constexpr uint64_t OFF_HI = 0x4630000000000000ull;   // bits of 2^100
constexpr uint64_t OFF_LO = 0x4330000000000000ull;   // bits of 2^52
constexpr double C1 = 0x1p100, C2 = 0x1p100 + 0x1p52, T52 = 0x1p52;

EKZG_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    volatile double x = a, y = b, z = c;
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    volatile double r = std::fma(x, y, z);
    fesetround(old);
    return r;
#endif
}
EKZG_HD uint64_t dbits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
// the double whose value is the integer (hi16:lo32) < 2^48
EKZG_HD double from_u48(uint32_t lo, uint32_t hi16) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)(hi16 | 0x43300000u), (int)lo) - T52;
#else
    uint64_t u = ((uint64_t)(hi16 | 0x43300000u) << 32) | lo; double d; memcpy(&d, &u, 8); return d - T52;
#endif
}
EKZG_HD uint32_t funnel_r16(uint32_t lo, uint32_t hi) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, 16);
#else
    return (lo >> 16) | (hi << 16);
#endif
}

// limb j of p in radix 2^48, as a double
EKZG_HD constexpr uint64_t pl_u(int j) {
    const int q = j >> 1;
    return (j & 1) ? ((uint64_t)(FpParams::mod(3 * q + 1) >> 16) | ((uint64_t)FpParams::mod(3 * q + 2) << 16))
                   : ((uint64_t)FpParams::mod(3 * q) | ((uint64_t)(FpParams::mod(3 * q + 1) & 0xffffu) << 32));
}
EKZG_HD constexpr double pl(int j) { return (double)pl_u(j); }
// -p^-1 mod 2^48 (Hensel lifting of the inverse of the low 64 bits)
EKZG_HD constexpr uint64_t pinv48() {
    const uint64_t p0 = (uint64_t)FpParams::mod(0) | ((uint64_t)FpParams::mod(1) << 32);
    uint64_t y = 1;
    for (int i = 0; i < 6; i++) y *= 2 - p0 * y;
    return (0 - y) & M48;
}

// a (packed 12 x 32, < 2^384) -> 8 doubles
EKZG_HD void unpack(double* d, const uint32_t* a) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        d[2 * q] = from_u48(a[3 * q], a[3 * q + 1] & 0xffffu);
        d[2 * q + 1] = from_u48(funnel_r16(a[3 * q + 1], a[3 * q + 2]), a[3 * q + 2] >> 16);
    }
}

// col[k] += lo48(x*y), col[k+1] += floor(x*y / 2^48)   (plus the two known offsets)
EKZG_HD void prod(uint64_t* col, int k, double x, double y) {
    const double hi = fma_rz(x, y, C1);
    const double lo = fma_rz(x, y, C2 - hi);
    col[k] += dbits(lo);
    col[k + 1] += dbits(hi);
}

// One Montgomery row in radix 2^48: column i is complete; add m*p so that it becomes divisible by 2^48, push it up.
EKZG_HD void reduce_row(uint64_t* col, int i) {
    const uint64_t m = (col[i] * pinv48()) & M48;
    const double md = from_u48((uint32_t)m, (uint32_t)(m >> 32));
#pragma unroll
    for (int j = 0; j < L; j++) prod(col, i + j, md, pl(j));
    col[i + 1] += col[i] >> 48;
}

// columns 8..15 -> packed, fully reduced 12 x 32
EKZG_HD void finish(uint32_t* out, const uint64_t* col) {
    uint32_t r[12];
    uint64_t t = col[L];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint64_t e = t & M48;
        t = (t >> 48) + col[L + 2 * q + 1];
        const uint64_t o = (q == 3) ? t : (t & M48);     // result < 2p < 2^382: the top limb needs no mask
        if (q < 3) t = (t >> 48) + col[L + 2 * q + 2];
        r[3 * q] = (uint32_t)e;
        r[3 * q + 1] = (uint32_t)(e >> 32) | ((uint32_t)o << 16);
        r[3 * q + 2] = (uint32_t)(o >> 16);
    }
    fe_final_sub<FpParams>(r);
#pragma unroll
    for (int w = 0; w < 12; w++) out[w] = r[w];
}

// number of (i, j) in [0,8)^2 with i + j == k
EKZG_HD constexpr int npairs(int k) { return k < 0 || k > 14 ? 0 : (k < 8 ? k + 1 : 15 - k); }
// number of squaring products (i <= j) with i + j == k
EKZG_HD constexpr int nsq(int k) { return k < 0 || k > 14 ? 0 : npairs(k) / 2 + ((k & 1) ? 0 : 1); }
// negated offsets of everything column k will receive: nab products of the operand part, 64 of the m*p part
EKZG_HD constexpr uint64_t col_init(int n_lo, int n_hi) { return 0 - ((uint64_t)n_lo * OFF_LO + (uint64_t)n_hi * OFF_HI); }

}  // namespace fpd

// Rolled form of the multiplication: 4 iterations of 2 rows over a sliding window of 10 columns, so that the subroutine is
// ~4 KB of code instead of ~14 KB (fully unrolled, the three FP64 routines exceed the 32 KB instruction cache: ncu showed
// stall_no_instruction 1.2 per issue and a 77 % hit rate).  Offsets are removed when a column retires: column k < 8 has
// received (NP+1)*(k+1) low halves and (NP+1)*k high halves by then.   NP = 1: a*b, NP = 2: a*b + c*d.
template <int NP>
EKZG_HD void mont_rolled(uint32_t* out, double* a, const double* b, double* c, const double* d) {
    using namespace fpd;
    uint64_t col[2 * L];   // [0..9] sliding window; after the loop [0..7] are columns 8..15
#pragma unroll
    for (int k = 0; k < 10; k++) col[k] = 0;
    uint64_t off = (uint64_t)(NP + 1) * OFF_LO;
#pragma unroll 1
    for (int it = 0; it < L / 2; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int j = 0; j < L; j++) {
                prod(col + r, j, a[r], b[j]);
                if (NP == 2) prod(col + r, j, c[r], d[j]);
            }
            const uint64_t m = (col[r] * pinv48()) & M48;
            const double md = from_u48((uint32_t)m, (uint32_t)(m >> 32));
#pragma unroll
            for (int j = 0; j < L; j++) prod(col + r, j, md, pl(j));
            col[r + 1] += (col[r] - off) >> 48;
            off += (uint64_t)(NP + 1) * (OFF_LO + OFF_HI);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) col[k] = col[k + 2];
        col[8] = 0;
        col[9] = 0;
#pragma unroll
        for (int k = 0; k < L - 2; k++) {
            a[k] = a[k + 2];
            if (NP == 2) c[k] = c[k + 2];
        }
    }
    // columns 8..15: (NP+1)*(15-k) low halves, (NP+1)*(16-k) high halves
#pragma unroll
    for (int k = 0; k < L; k++) col[L + k] = col[k] - (uint64_t)(NP + 1) * ((uint64_t)(7 - k) * OFF_LO + (uint64_t)(8 - k) * OFF_HI);
    finish(out, col);
}

EKZG_HD void fp_mul_dfma_inline(Fe<FpParams>& out, const Fe<FpParams>& a_, const Fe<FpParams>& b_) {
    using namespace fpd;
    double a[L], b[L];
    unpack(a, a_.v);
    unpack(b, b_.v);
    mont_rolled<1>(out.v, a, b, a, b);
}

// out = a*b + c*d, one reduction
EKZG_HD void fp_mul2_dfma_inline(Fe<FpParams>& out, const Fe<FpParams>& a_, const Fe<FpParams>& b_, const Fe<FpParams>& c_, const Fe<FpParams>& d_) {
    using namespace fpd;
    double a[L], b[L], c[L], d[L];
    unpack(a, a_.v);
    unpack(b, b_.v);
    unpack(c, c_.v);
    unpack(d, d_.v);
    mont_rolled<2>(out.v, a, b, c, d);
}

// out = a^2: cross products once, against the doubled limb
EKZG_HD void fp_sqr_dfma_inline(Fe<FpParams>& out, const Fe<FpParams>& a_) {
    using namespace fpd;
    double a[L], d[L];
    unpack(a, a_.v);
#pragma unroll
    for (int k = 0; k < L; k++) d[k] = a[k] + a[k];
    uint64_t col[2 * L];
#pragma unroll
    for (int k = 0; k < 2 * L; k++) col[k] = col_init(nsq(k) + npairs(k), nsq(k - 1) + npairs(k - 1));
#pragma unroll
    for (int i = 0; i < L; i++) {
        prod(col, 2 * i, a[i], a[i]);
#pragma unroll
        for (int j = i + 1; j < L; j++) prod(col, i + j, a[i], d[j]);
        reduce_row(col, i);
    }
    finish(out.v, col);
}

}  // namespace ekzg
