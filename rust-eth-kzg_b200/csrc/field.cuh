// BLS12-381 prime fields on 32-bit limbs: Fp (12 limbs) and Fr (8 limbs), Montgomery form,
// one element per thread, all limbs in registers.
//
// Replaces the blst field arithmetic the reference reaches through blstrs
// (crates/cryptography/bls12_381/src/lib.rs:23-42 type aliases).
//
// mont_mul keeps the running sum as TWO staggered accumulators  T = X + Y * 2^32 : all products
// a[j]*b_i with even j land in X (lo at limb j, hi at limb j+1), those with odd j in Y, so each
// row is two uninterrupted mad.lo.cc / madc.hi.cc chains with no carry fix-ups in the middle.
// After the Montgomery step X[0] == 0 and the division by 2^32 is a role swap of X and Y.
// Cost: N*(2N+1) wide multiply-adds (ptxas fuses every mad.lo.cc / madc.hi.cc pair into one IMAD.WIDE.U32.X):
// Fp 12*25 = 300, Fr 8*17 = 136.
#pragma once
#include "constants.cuh"

namespace ekzg {

template <class P>
struct alignas(16) Fe {
    uint32_t v[P::N];
};
using Fp = Fe<FpParams>;
using Fr = Fe<FrParams>;

// r = (r >= p) ? r - p : r     (r < 2p on entry)
template <class P>
EKZG_HD void fe_final_sub(uint32_t* r) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = sub_cc(r[0], P::mod(0));
#pragma unroll
    for (int j = 1; j < N; j++) t[j] = subc_cc(r[j], P::mod(j));
    uint32_t borrow = subc(0u, 0u);  // 0xffffffff if r < p
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = borrow ? r[j] : t[j];
}

template <class P>
EKZG_HD void fe_mul_inline(Fe<P>& out, const Fe<P>& a_, const Fe<P>& b_) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    const uint32_t* a = a_.v;
    const uint32_t* b = b_.v;
    uint32_t ev[N], od[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t* x = (i & 1) ? od : ev;  // offset-0 accumulator of this row
        uint32_t* y = (i & 1) ? ev : od;  // offset-1 accumulator of this row
        const uint32_t bi = b[i];
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                x[j] = mul_lo(a[j], bi);
                x[j + 1] = mul_hi(a[j], bi);
                y[j] = mul_lo(a[j + 1], bi);
                y[j + 1] = mul_hi(a[j + 1], bi);
            }
        } else {
            // T/2^32 of the previous row:  new X = old Y (+ old X[1] at limb 0), new Y = old X >> 64
            x[0] = add_cc(x[0], y[1]);
#pragma unroll
            for (int j = 1; j < N - 1; j += 2) {
                y[j - 1] = madc_lo_cc(a[j], bi, y[j + 1]);
                y[j] = madc_hi_cc(a[j], bi, y[j + 2]);
            }
            y[N - 2] = madc_lo_cc(a[N - 1], bi, 0u);
            y[N - 1] = madc_hi(a[N - 1], bi, 0u);
            x[0] = mad_lo_cc(a[0], bi, x[0]);
            x[1] = madc_hi_cc(a[0], bi, x[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                x[j] = madc_lo_cc(a[j], bi, x[j]);
                x[j + 1] = madc_hi_cc(a[j], bi, x[j + 1]);
            }
            y[N - 1] = addc(y[N - 1], 0u);
        }
        const uint32_t m = mul_lo(x[0], P::M0);
        y[0] = mad_lo_cc(P::mod(1), m, y[0]);
        y[1] = madc_hi_cc(P::mod(1), m, y[1]);
#pragma unroll
        for (int j = 3; j < N; j += 2) {
            y[j - 1] = madc_lo_cc(P::mod(j), m, y[j - 1]);
            y[j] = madc_hi_cc(P::mod(j), m, y[j]);
        }
        x[0] = mad_lo_cc(P::mod(0), m, x[0]);  // == 0 now
        x[1] = madc_hi_cc(P::mod(0), m, x[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            x[j] = madc_lo_cc(P::mod(j), m, x[j]);
            x[j + 1] = madc_hi_cc(P::mod(j), m, x[j + 1]);
        }
        y[N - 1] = addc(y[N - 1], 0u);
    }
    // N is even: the last row used x = od, y = ev.  result = x/2^32 + y
    uint32_t* x = od;
    uint32_t* y = ev;
    uint32_t r[N];
    r[0] = add_cc(x[1], y[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = addc_cc(x[j + 1], y[j]);
    r[N - 1] = addc(y[N - 1], 0u);
    fe_final_sub<P>(r);
#pragma unroll
    for (int j = 0; j < N; j++) out.v[j] = r[j];
}

// out = a*b + c*d with ONE Montgomery reduction (444 instead of 600 multiply-adds for Fp): every row adds both
// partial products before its reduction step.  The running sum stays below 3p*2^32 + 3p, so the same headroom
// argument as for fe_sqr_inline applies (Fp only); the result is below (2p^2 + pR)/R < 2p, one final subtraction.
// Point formulas use it for  y3 = r*(v - x3) - y1*j  with the second product's factor negated beforehand.
template <class P>
EKZG_HD void fe_mul2_inline(Fe<P>& out, const Fe<P>& a_, const Fe<P>& b_, const Fe<P>& c_, const Fe<P>& d_) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    const uint32_t* a = a_.v;
    const uint32_t* b = b_.v;
    const uint32_t* c = c_.v;
    const uint32_t* d = d_.v;
    uint32_t ev[N], od[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t* x = (i & 1) ? od : ev;
        uint32_t* y = (i & 1) ? ev : od;
        const uint32_t bi = b[i], di = d[i];
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                x[j] = mul_lo(a[j], bi);
                x[j + 1] = mul_hi(a[j], bi);
                y[j] = mul_lo(a[j + 1], bi);
                y[j + 1] = mul_hi(a[j + 1], bi);
            }
        } else {
            x[0] = add_cc(x[0], y[1]);
#pragma unroll
            for (int j = 1; j < N - 1; j += 2) {
                y[j - 1] = madc_lo_cc(a[j], bi, y[j + 1]);
                y[j] = madc_hi_cc(a[j], bi, y[j + 2]);
            }
            y[N - 2] = madc_lo_cc(a[N - 1], bi, 0u);
            y[N - 1] = madc_hi(a[N - 1], bi, 0u);
            x[0] = mad_lo_cc(a[0], bi, x[0]);
            x[1] = madc_hi_cc(a[0], bi, x[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                x[j] = madc_lo_cc(a[j], bi, x[j]);
                x[j + 1] = madc_hi_cc(a[j], bi, x[j + 1]);
            }
            y[N - 1] = addc(y[N - 1], 0u);
        }
        // second partial product, then the Montgomery step: both add in place (no shift)
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const uint32_t m = pass ? mul_lo(x[0], P::M0) : di;
            auto f = [&](int k) -> uint32_t { return pass ? P::mod(k) : c[k]; };
            y[0] = mad_lo_cc(f(1), m, y[0]);
            y[1] = madc_hi_cc(f(1), m, y[1]);
#pragma unroll
            for (int j = 3; j < N; j += 2) {
                y[j - 1] = madc_lo_cc(f(j), m, y[j - 1]);
                y[j] = madc_hi_cc(f(j), m, y[j]);
            }
            x[0] = mad_lo_cc(f(0), m, x[0]);
            x[1] = madc_hi_cc(f(0), m, x[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                x[j] = madc_lo_cc(f(j), m, x[j]);
                x[j + 1] = madc_hi_cc(f(j), m, x[j + 1]);
            }
            y[N - 1] = addc(y[N - 1], 0u);
        }
    }
    uint32_t* x = od;
    uint32_t* y = ev;
    uint32_t r[N];
    r[0] = add_cc(x[1], y[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = addc_cc(x[j + 1], y[j]);
    r[N - 1] = addc(y[N - 1], 0u);
    fe_final_sub<P>(r);
#pragma unroll
    for (int j = 0; j < N; j++) out.v[j] = r[j];
}

// Squaring: row i multiplies a_i into  S_i = a_i*2^(32i) + 2*(a >> 32(i+1)) << 32(i+1)  instead of into all of a,
// so every cross product a_i*a_k (k > i) is computed once, already doubled: sum_i a_i * S_i = a^2.
// The limbs of S_i are a_i, then (a_(i+1) << 1), then the limbs of 2a -- plain register values, no fix-ups.
// Same staggered accumulators and Montgomery rows as fe_mul_inline; rows just start at limb i, the skipped
// slots only pass the carry along.  78 + 156 multiply-adds instead of 300 for Fp.
// Row 0 multiplies by ~2a, so the running sum reaches 3p * 2^32 before the shift: that must stay below the
// accumulators' 2^(32N+32), i.e. p < 2^(32N)/3 -- true for Fp (p ~ 0.10 * 2^384), NOT for Fr (r ~ 0.45 * 2^256),
// which therefore keeps squaring through fe_mul.
template <class P>
EKZG_HD void fe_sqr_inline(Fe<P>& out, const Fe<P>& a_) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    const uint32_t* a = a_.v;
    uint32_t d[N], e[N];   // d = limbs of 2a (2a < 2^(32N)), e[k] = a[k] << 1
#pragma unroll
    for (int k = 0; k < N; k++) {
        e[k] = a[k] << 1;
        d[k] = k ? (e[k] | (a[k - 1] >> 31)) : e[k];
    }
    uint32_t ev[N], od[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t* x = (i & 1) ? od : ev;
        uint32_t* y = (i & 1) ? ev : od;
        const uint32_t bi = a[i];
        // limb k of S_i (k >= i)
        auto s = [&](int k) -> uint32_t { return k == i ? a[k] : (k == i + 1 ? e[k] : d[k]); };
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                x[j] = mul_lo(s(j), bi);
                x[j + 1] = mul_hi(s(j), bi);
                y[j] = mul_lo(s(j + 1), bi);
                y[j + 1] = mul_hi(s(j + 1), bi);
            }
        } else {
            x[0] = add_cc(x[0], y[1]);
#pragma unroll
            for (int j = 1; j < N - 1; j += 2) {
                if (j >= i) {
                    y[j - 1] = madc_lo_cc(s(j), bi, y[j + 1]);
                    y[j] = madc_hi_cc(s(j), bi, y[j + 2]);
                } else {
                    y[j - 1] = addc_cc(y[j + 1], 0u);
                    y[j] = addc_cc(y[j + 2], 0u);
                }
            }
            y[N - 2] = madc_lo_cc(s(N - 1), bi, 0u);
            y[N - 1] = madc_hi(s(N - 1), bi, 0u);
            // even limbs k >= i (k = 0 only belongs to row 0)
            constexpr int dummy = 0; (void)dummy;
            const int k0 = (i + 1) & ~1;  // first even k >= i
            if (k0 < N) {
                x[k0] = mad_lo_cc(s(k0), bi, x[k0]);
                x[k0 + 1] = madc_hi_cc(s(k0), bi, x[k0 + 1]);
#pragma unroll
                for (int j = 2; j < N; j += 2) {
                    if (j > k0) {
                        x[j] = madc_lo_cc(s(j), bi, x[j]);
                        x[j + 1] = madc_hi_cc(s(j), bi, x[j + 1]);
                    }
                }
                y[N - 1] = addc(y[N - 1], 0u);
            }
        }
        const uint32_t m = mul_lo(x[0], P::M0);
        y[0] = mad_lo_cc(P::mod(1), m, y[0]);
        y[1] = madc_hi_cc(P::mod(1), m, y[1]);
#pragma unroll
        for (int j = 3; j < N; j += 2) {
            y[j - 1] = madc_lo_cc(P::mod(j), m, y[j - 1]);
            y[j] = madc_hi_cc(P::mod(j), m, y[j]);
        }
        x[0] = mad_lo_cc(P::mod(0), m, x[0]);
        x[1] = madc_hi_cc(P::mod(0), m, x[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            x[j] = madc_lo_cc(P::mod(j), m, x[j]);
            x[j + 1] = madc_hi_cc(P::mod(j), m, x[j + 1]);
        }
        y[N - 1] = addc(y[N - 1], 0u);
    }
    uint32_t* x = od;
    uint32_t* y = ev;
    uint32_t r[N];
    r[0] = add_cc(x[1], y[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = addc_cc(x[j + 1], y[j]);
    r[N - 1] = addc(y[N - 1], 0u);
    fe_final_sub<P>(r);
#pragma unroll
    for (int j = 0; j < N; j++) out.v[j] = r[j];
}

// Alternative Fp multipliers were built and measured on B200 -- 14 x 28-bit limbs with carry-free IMAD.WIDE column sums,
// 8 x 48-bit limbs in doubles with the DFMA.RZ high/low split, one Karatsuba level on the plain product -- and all three lose
// to the interleaved carry chains above at the 2-3 warps per sub-partition these kernels run at
// (profiles/r1_v4_dfma_experiment.md; the code of the last two is in the history up to commit 5b26d49).
EKZG_HD void fp_mul_impl(Fp& r, const Fp& a, const Fp& b) { fe_mul_inline(r, a, b); }
EKZG_HD void fp_sqr_impl(Fp& r, const Fp& a) { fe_sqr_inline(r, a); }
EKZG_HD void fp_mul2_impl(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) { fe_mul2_inline(r, a, b, c, d); }

// The ~620-instruction Fp multiplication is ONE subroutine per kernel image on the device: operands and result
// travel in registers (by-value aggregates; ptxas keeps them out of memory), so a point operation is a short
// sequence of calls and the hot code of every kernel fits the 32 KB L1.5 instruction cache.  Fully inlined,
// a Jacobian doubling alone is 70 KB of straight-line code and warps that are not in lock step starve on
// instruction fetch (ncu: stall_no_instruction 21 cycles per issue in the first persistent G1-NTT kernel).
#if defined(__CUDA_ARCH__) && !defined(EKZG_FP_MUL_INLINE)
static __device__ __noinline__ Fp fp_mul_call(Fp a, Fp b) {
    Fp r;
    fp_mul_impl(r, a, b);
    return r;
}
#endif

#if defined(__CUDA_ARCH__) && !defined(EKZG_FP_MUL_INLINE)
static __device__ __noinline__ Fp fp_mul2_call(Fp a, Fp b, Fp c, Fp d) {
    Fp r;
    fp_mul2_impl(r, a, b, c, d);
    return r;
}
#endif
// out = a*b + c*d  (Fp only, see fe_mul2_inline)
EKZG_HD void fp_mul2_add(Fp& out, const Fp& a, const Fp& b, const Fp& c, const Fp& d);

template <class P>
EKZG_HD void fe_mul(Fe<P>& out, const Fe<P>& a, const Fe<P>& b) {
    fe_mul_inline(out, a, b);
}
EKZG_HD void fe_mul(Fp& out, const Fp& a, const Fp& b) {
#if defined(__CUDA_ARCH__) && !defined(EKZG_FP_MUL_INLINE)
    out = fp_mul_call(a, b);
#else
    fp_mul_impl(out, a, b);
#endif
}

#if defined(__CUDA_ARCH__) && !defined(EKZG_FP_MUL_INLINE)
static __device__ __noinline__ Fp fp_sqr_call(Fp a) {
    Fp r;
    fp_sqr_impl(r, a);
    return r;
}
#endif

template <class P>
EKZG_HD void fe_sqr(Fe<P>& out, const Fe<P>& a) {
    fe_mul_inline(out, a, a);
}
EKZG_HD void fe_sqr(Fp& out, const Fp& a) {
#if defined(__CUDA_ARCH__) && !defined(EKZG_FP_MUL_INLINE)
    out = fp_sqr_call(a);
#else
    fp_sqr_impl(out, a);
#endif
}

template <class P>
EKZG_HD void fe_add(Fe<P>& out, const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t r[N];
    r[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = addc_cc(a.v[j], b.v[j]);
    r[N - 1] = addc(a.v[N - 1], b.v[N - 1]);  // 2p < 2^(32N): no carry out
    fe_final_sub<P>(r);
#pragma unroll
    for (int j = 0; j < N; j++) out.v[j] = r[j];
}

template <class P>
EKZG_HD void fe_sub(Fe<P>& out, const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t r[N];
    r[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int j = 1; j < N; j++) r[j] = subc_cc(a.v[j], b.v[j]);
    uint32_t mask = subc(0u, 0u);  // all ones if a < b
    r[0] = add_cc(r[0], P::mod(0) & mask);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = addc_cc(r[j], P::mod(j) & mask);
    r[N - 1] = addc(r[N - 1], P::mod(N - 1) & mask);
#pragma unroll
    for (int j = 0; j < N; j++) out.v[j] = r[j];
}

EKZG_HD void fp_mul2_add(Fp& out, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
#if defined(__CUDA_ARCH__) && !defined(EKZG_FP_MUL_INLINE)
    out = fp_mul2_call(a, b, c, d);
#else
    fp_mul2_impl(out, a, b, c, d);
#endif
}

template <class P>
EKZG_HD bool fe_is_zero(const Fe<P>& a) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < P::N; j++) x |= a.v[j];
    return x == 0;
}

template <class P>
EKZG_HD bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < P::N; j++) x |= a.v[j] ^ b.v[j];
    return x == 0;
}

template <class P>
EKZG_HD void fe_neg(Fe<P>& out, const Fe<P>& a) {
    constexpr int N = P::N;
    uint32_t nz = 0;
#pragma unroll
    for (int j = 0; j < N; j++) nz |= a.v[j];
    uint32_t mask = nz ? 0xffffffffu : 0u;
    uint32_t r[N];
    r[0] = sub_cc(P::mod(0) & mask, a.v[0]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) r[j] = subc_cc(P::mod(j) & mask, a.v[j]);
    r[N - 1] = subc(P::mod(N - 1) & mask, a.v[N - 1]);
#pragma unroll
    for (int j = 0; j < N; j++) out.v[j] = r[j];
}

// out = neg ? -a : a
template <class P>
EKZG_HD void fe_cneg(Fe<P>& out, const Fe<P>& a, bool neg) {
    Fe<P> n;
    fe_neg(n, a);
#pragma unroll
    for (int j = 0; j < P::N; j++) out.v[j] = neg ? n.v[j] : a.v[j];
}

template <class P>
EKZG_HD void fe_dbl(Fe<P>& out, const Fe<P>& a) {
    fe_add(out, a, a);
}

template <class P>
EKZG_HD void fe_set_zero(Fe<P>& a) {
#pragma unroll
    for (int j = 0; j < P::N; j++) a.v[j] = 0;
}
template <class P>
EKZG_HD void fe_set_one(Fe<P>& a) {
#pragma unroll
    for (int j = 0; j < P::N; j++) a.v[j] = P::one(j);
}
template <class P>
EKZG_HD Fe<P> fe_const_r2() {
    Fe<P> a;
#pragma unroll
    for (int j = 0; j < P::N; j++) a.v[j] = P::r2(j);
    return a;
}
// plain integer (< modulus) -> Montgomery form
template <class P>
EKZG_HD void fe_to_mont(Fe<P>& out, const Fe<P>& a) {
    fe_mul(out, a, fe_const_r2<P>());
}
// Montgomery form -> plain integer in [0, modulus)
template <class P>
EKZG_HD void fe_from_mont(Fe<P>& out, const Fe<P>& a) {
    Fe<P> one;
    fe_set_zero(one);
    one.v[0] = 1;
    fe_mul(out, a, one);
}

// plain-integer compare a >= modulus (for canonicity checks on the wire format)
template <class P>
EKZG_HD bool fe_plain_ge_mod(const Fe<P>& a) {
    constexpr int N = P::N;
    sub_cc(a.v[0], P::mod(0));
#pragma unroll
    for (int j = 1; j < N; j++) subc_cc(a.v[j], P::mod(j));
    uint32_t borrow = subc(0u, 0u);
    return borrow == 0;
}
// plain-integer compare a > (modulus-1)/2   (sign bit of the compressed encoding)
template <class P>
EKZG_HD bool fe_plain_gt_half(const Fe<P>& a) {
    constexpr int N = P::N;
    sub_cc(P::half(0), a.v[0]);
#pragma unroll
    for (int j = 1; j < N; j++) subc_cc(P::half(j), a.v[j]);
    uint32_t borrow = subc(0u, 0u);
    return borrow != 0;
}

// a^e for a plain little-endian exponent given by a constexpr limb function; square-and-multiply
// with a 4-bit fixed window (only used off the hot loop: inversion, square roots).
template <class P, class ExpFn>
EKZG_HD void fe_pow_limbs(Fe<P>& out, const Fe<P>& a, ExpFn e, int nbits) {
    Fe<P> tbl[16];
    fe_set_one(tbl[0]);
    tbl[1] = a;
    for (int i = 2; i < 16; i++) fe_mul(tbl[i], tbl[i - 1], a);
    Fe<P> acc;
    fe_set_one(acc);
    int top = ((nbits + 3) / 4) * 4;
    for (int pos = top - 4; pos >= 0; pos -= 4) {
        for (int s = 0; s < 4; s++) fe_sqr(acc, acc);
        uint32_t d = (e(pos >> 5) >> (pos & 31)) & 15u;
        Fe<P> t = tbl[d];
        fe_mul(acc, acc, t);
    }
    out = acc;
}

struct FpExpInv {  // p - 2
    EKZG_HD uint32_t operator()(int i) const { return FpParams::exp_inv(i); }
};
struct FrExpInv {  // r - 2
    EKZG_HD uint32_t operator()(int i) const { return FrParams::exp_inv(i); }
};
struct FpExpSqrt {  // (p+1)/4
    EKZG_HD uint32_t operator()(int i) const { return FpParams::exp_sqrt(i); }
};

EKZG_HD_CALL void fp_inv(Fp& out, const Fp& a) { fe_pow_limbs(out, a, FpExpInv(), 381); }
EKZG_HD_CALL void fr_inv(Fr& out, const Fr& a) { fe_pow_limbs(out, a, FrExpInv(), 255); }
EKZG_HD_CALL void fp_sqrt_candidate(Fp& out, const Fp& a) { fe_pow_limbs(out, a, FpExpSqrt(), 379); }

}  // namespace ekzg
