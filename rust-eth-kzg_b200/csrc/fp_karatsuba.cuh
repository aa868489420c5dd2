// Fp Montgomery multiplication with one level of Karatsuba on the 12 x 12-limb product: 3 * 36 + 144 + 12 = 264 wide
// multiply-adds instead of 300.  The price is ~250 extra additions, which go to the ALU pipe (128 lanes/clk/SM on sm_100,
// profiles/r1_v4_pipe_probe.json) while the multiplier -- the pipe K4 is bound by -- does 12 % less per product.
//
//   a = a0 + a1*B^6, b = b0 + b1*B^6 (B = 2^32):  a*b = z0 + z1*B^6 + z2*B^12,  z0 = a0*b0, z2 = a1*b1,
//   z1 = (a0+a1)(b0+b1) - z0 - z2, the carries of the two sums handled as masked additions.
// All partial products keep the discipline of field.cuh: every mad.lo.cc / madc.hi.cc pair adds into an EVEN-aligned pair
// of accumulator limbs, so ptxas fuses it into one IMAD.WIDE.U32.X; products landing on odd limb positions go to a second
// accumulator that is stored shifted by one limb.
// The 24-limb product is then reduced row by row (12 rows of m*p), again on the two accumulators.
// Same packed Montgomery form (R = 2^384), inputs < p, output < p: a drop-in for fe_mul_inline (Fp only).
#pragma once
#include "constants.cuh"

namespace ekzg {

namespace fpk {

// r[0..11] = x[0..5] * y[0..5]
EKZG_HD void mul6(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint32_t ev[12], od[12];   // ev: products on even limb positions; od[k] = limb k+1 of those on odd positions
#pragma unroll
    for (int k = 6; k < 12; k++) { ev[k] = 0; od[k] = 0; }
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
        ev[j] = mul_lo(x[j], y[0]);
        ev[j + 1] = mul_hi(x[j], y[0]);
        od[j] = mul_lo(x[j + 1], y[0]);
        od[j + 1] = mul_hi(x[j + 1], y[0]);
    }
#pragma unroll
    for (int i = 1; i < 6; i++) {
        const uint32_t yi = y[i];
        if (i & 1) {
            // odd j -> even positions i+j: ev[i+1 .. i+6]
            ev[i + 1] = mad_lo_cc(x[1], yi, ev[i + 1]);
            ev[i + 2] = madc_hi_cc(x[1], yi, ev[i + 2]);
            ev[i + 3] = madc_lo_cc(x[3], yi, ev[i + 3]);
            ev[i + 4] = madc_hi_cc(x[3], yi, ev[i + 4]);
            ev[i + 5] = madc_lo_cc(x[5], yi, ev[i + 5]);
            if (i + 7 < 12) {
                ev[i + 6] = madc_hi_cc(x[5], yi, ev[i + 6]);
                ev[i + 7] = addc(ev[i + 7], 0u);
            } else {
                ev[i + 6] = madc_hi(x[5], yi, ev[i + 6]);
            }
            // even j -> odd positions i+j: od[i-1 .. i+4]
            od[i - 1] = mad_lo_cc(x[0], yi, od[i - 1]);
            od[i] = madc_hi_cc(x[0], yi, od[i]);
            od[i + 1] = madc_lo_cc(x[2], yi, od[i + 1]);
            od[i + 2] = madc_hi_cc(x[2], yi, od[i + 2]);
            od[i + 3] = madc_lo_cc(x[4], yi, od[i + 3]);
            od[i + 4] = madc_hi_cc(x[4], yi, od[i + 4]);
            od[i + 5] = addc(od[i + 5], 0u);
        } else {
            // even j -> even positions: ev[i .. i+5]
            ev[i] = mad_lo_cc(x[0], yi, ev[i]);
            ev[i + 1] = madc_hi_cc(x[0], yi, ev[i + 1]);
            ev[i + 2] = madc_lo_cc(x[2], yi, ev[i + 2]);
            ev[i + 3] = madc_hi_cc(x[2], yi, ev[i + 3]);
            ev[i + 4] = madc_lo_cc(x[4], yi, ev[i + 4]);
            ev[i + 5] = madc_hi_cc(x[4], yi, ev[i + 5]);
            ev[i + 6] = addc(ev[i + 6], 0u);
            // odd j -> odd positions i+j: od[i .. i+5]
            od[i] = mad_lo_cc(x[1], yi, od[i]);
            od[i + 1] = madc_hi_cc(x[1], yi, od[i + 1]);
            od[i + 2] = madc_lo_cc(x[3], yi, od[i + 2]);
            od[i + 3] = madc_hi_cc(x[3], yi, od[i + 3]);
            od[i + 4] = madc_lo_cc(x[5], yi, od[i + 4]);
            od[i + 5] = madc_hi_cc(x[5], yi, od[i + 5]);
            od[i + 6] = addc(od[i + 6], 0u);
        }
    }
    r[0] = ev[0];
    r[1] = add_cc(ev[1], od[0]);
#pragma unroll
    for (int k = 2; k < 11; k++) r[k] = addc_cc(ev[k], od[k - 1]);
    r[11] = addc(ev[11], od[10]);
}

// t[0..23] = a * b  (12 x 12 limbs, one Karatsuba level)
EKZG_HD void mul12_karatsuba(uint32_t* t, const uint32_t* a, const uint32_t* b) {
    uint32_t sa[6], sb[6];
    sa[0] = add_cc(a[0], a[6]);
#pragma unroll
    for (int j = 1; j < 6; j++) sa[j] = addc_cc(a[j], a[j + 6]);
    const uint32_t ca = addc(0u, 0u);
    sb[0] = add_cc(b[0], b[6]);
#pragma unroll
    for (int j = 1; j < 6; j++) sb[j] = addc_cc(b[j], b[j + 6]);
    const uint32_t cb = addc(0u, 0u);
    uint32_t z0[12], z2[12], m[13];
    mul6(z0, a, b);
    mul6(z2, a + 6, b + 6);
    mul6(m, sa, sb);
    // (ca*B^6 + sa)(cb*B^6 + sb) = sa*sb + (ca ? sb : 0)*B^6 + (cb ? sa : 0)*B^6 + ca*cb*B^12
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    m[6] = add_cc(m[6], sb[0] & ma);
#pragma unroll
    for (int j = 1; j < 6; j++) m[6 + j] = addc_cc(m[6 + j], sb[j] & ma);
    m[12] = addc(ca & cb, 0u);
    m[6] = add_cc(m[6], sa[0] & mb);
#pragma unroll
    for (int j = 1; j < 6; j++) m[6 + j] = addc_cc(m[6 + j], sa[j] & mb);
    m[12] = addc(m[12], 0u);
    // z1 = m - z0 - z2  (13 limbs, never negative)
    m[0] = sub_cc(m[0], z0[0]);
#pragma unroll
    for (int j = 1; j < 12; j++) m[j] = subc_cc(m[j], z0[j]);
    m[12] = subc(m[12], 0u);
    m[0] = sub_cc(m[0], z2[0]);
#pragma unroll
    for (int j = 1; j < 12; j++) m[j] = subc_cc(m[j], z2[j]);
    m[12] = subc(m[12], 0u);
    // t = z0 + z1*B^6 + z2*B^12
#pragma unroll
    for (int j = 0; j < 6; j++) t[j] = z0[j];
    t[6] = add_cc(z0[6], m[0]);
#pragma unroll
    for (int j = 1; j < 6; j++) t[6 + j] = addc_cc(z0[6 + j], m[j]);
#pragma unroll
    for (int j = 0; j < 7; j++) t[12 + j] = addc_cc(z2[j], m[6 + j]);
#pragma unroll
    for (int j = 7; j < 11; j++) t[12 + j] = addc_cc(z2[j], 0u);
    t[23] = addc(z2[11], 0u);
}

// out[0..11] = t / B^12 mod p, fully reduced  (t < p*B^12: any sum of a few products of values < p).
// The running total is t + ev + od*B.  t is kept apart from the accumulators on purpose: a chain's last carry then always
// lands in a limb that holds 0 or 1, never in a full 32-bit limb of t where it could ripple further.
EKZG_HD void mont_reduce24(uint32_t* out, const uint32_t* t) {
    using P = FpParams;
    uint32_t ev[24], od[24];
#pragma unroll
    for (int k = 0; k < 24; k++) { ev[k] = 0; od[k] = 0; }
    uint32_t cin = 0;          // carry of the (zeroed) limbs below row i into limb i, at most 3
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint32_t lo = t[i] + ev[i] + (i ? od[i - 1] : 0u) + cin;
        const uint32_t mm = mul_lo(lo, P::M0);
        if (i & 1) {
            // odd j -> even positions i+j: ev[i+1 .. i+12]
            ev[i + 1] = mad_lo_cc(P::mod(1), mm, ev[i + 1]);
            ev[i + 2] = madc_hi_cc(P::mod(1), mm, ev[i + 2]);
#pragma unroll
            for (int j = 3; j < 12; j += 2) {
                ev[i + j] = madc_lo_cc(P::mod(j), mm, ev[i + j]);
                if (i + j + 1 < 23) ev[i + j + 1] = madc_hi_cc(P::mod(j), mm, ev[i + j + 1]);
                else ev[i + j + 1] = madc_hi(P::mod(j), mm, ev[i + j + 1]);
            }
            if (i + 13 < 24) ev[i + 13] = addc(ev[i + 13], 0u);
            // even j -> odd positions i+j: od[i-1 .. i+10]
            od[i - 1] = mad_lo_cc(P::mod(0), mm, od[i - 1]);
            od[i] = madc_hi_cc(P::mod(0), mm, od[i]);
#pragma unroll
            for (int j = 2; j < 12; j += 2) {
                od[i + j - 1] = madc_lo_cc(P::mod(j), mm, od[i + j - 1]);
                od[i + j] = madc_hi_cc(P::mod(j), mm, od[i + j]);
            }
            od[i + 11] = addc(od[i + 11], 0u);
        } else {
            // even j -> even positions: ev[i .. i+11]
            ev[i] = mad_lo_cc(P::mod(0), mm, ev[i]);
            ev[i + 1] = madc_hi_cc(P::mod(0), mm, ev[i + 1]);
#pragma unroll
            for (int j = 2; j < 12; j += 2) {
                ev[i + j] = madc_lo_cc(P::mod(j), mm, ev[i + j]);
                ev[i + j + 1] = madc_hi_cc(P::mod(j), mm, ev[i + j + 1]);
            }
            ev[i + 12] = addc(ev[i + 12], 0u);
            // odd j -> odd positions i+j: od[i .. i+11]
            od[i] = mad_lo_cc(P::mod(1), mm, od[i]);
            od[i + 1] = madc_hi_cc(P::mod(1), mm, od[i + 1]);
#pragma unroll
            for (int j = 3; j < 12; j += 2) {
                od[i + j - 1] = madc_lo_cc(P::mod(j), mm, od[i + j - 1]);
                od[i + j] = madc_hi_cc(P::mod(j), mm, od[i + j]);
            }
            od[i + 12] = addc(od[i + 12], 0u);
        }
        // limb i of the total is now 0 mod B; its carry goes into limb i+1
        uint32_t s = add_cc(t[i], ev[i]);
        uint32_t c = addc(0u, 0u);
        s = add_cc(s, i ? od[i - 1] : 0u);
        c = addc(c, 0u);
        s = add_cc(s, cin);
        cin = addc(c, 0u);
        (void)s;
    }
    uint32_t r[12];
    r[0] = add_cc(ev[12], od[11]);
#pragma unroll
    for (int k = 1; k < 11; k++) r[k] = addc_cc(ev[12 + k], od[11 + k]);
    r[11] = addc(ev[23], od[22]);
    r[0] = add_cc(r[0], t[12]);
#pragma unroll
    for (int k = 1; k < 11; k++) r[k] = addc_cc(r[k], t[12 + k]);
    r[11] = addc(r[11], t[23]);
    r[0] = add_cc(r[0], cin);
#pragma unroll
    for (int k = 1; k < 11; k++) r[k] = addc_cc(r[k], 0u);
    r[11] = addc(r[11], 0u);
    fe_final_sub<P>(r);
#pragma unroll
    for (int k = 0; k < 12; k++) out[k] = r[k];
}

}  // namespace fpk

EKZG_HD void fp_mulk_inline(Fe<FpParams>& out, const Fe<FpParams>& a, const Fe<FpParams>& b) {
    uint32_t t[24];
    fpk::mul12_karatsuba(t, a.v, b.v);
    fpk::mont_reduce24(out.v, t);
}

}  // namespace ekzg
