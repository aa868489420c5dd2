// Fp inversion by 2-adic division steps (Bernstein & Yang, "Fast constant-time gcd computation and modular inversion",
// TCHES 2019; the half-delta variant with 30 steps per batch on signed 30-bit limbs).
//
// Why it exists: the batched-affine MSM kernel (K4a in kzg_kernels.cu; reference batch_addition.rs:142-232 and
// batch_inversion.rs) shares ONE inversion between the ~19 independent point additions of a thread's round.  On this
// machine the binding resource is the wide integer multiply pipe; a Fermat inversion costs ~460 Montgomery products there
// (more than the 19 additions it serves), whereas the division steps are shifts, masks and adds on the ALU pipe (~11 k simple
// instructions per inversion) plus ~3.9 k wide multiplies for the 30 matrix applications -- about 13 product-equivalents.
//
// The routine is plain C++ on 32/64-bit integers (no PTX), so tests/host_emu runs the same code on the CPU.
// Iteration bound: for the half-delta variant a modulus below 2^381 needs at most floor((45907*381 + 26313)/19929) = 878
// division steps; 30 batches of 30 = 900 are run (fewer when g reaches zero early -- further steps would be no-ops).
#pragma once
#include "field.cuh"

namespace ekzg {
namespace gcdinv {

constexpr int L = 13;                       // signed 30-bit limbs: 13 * 30 = 390 bits
constexpr int32_t M30 = (1 << 30) - 1;
constexpr uint32_t P_INV30 = 0x30003u;      // p^-1 mod 2^30

EKZG_HD constexpr int32_t p30(int i) {
    switch (i) {
        case 0: return 0x3fffaaab; case 1: return 0x27fbffff; case 2: return 0x153ffffb; case 3: return 0x2affffac;
        case 4: return 0x30f6241e; case 5: return 0x034a83da; case 6: return 0x112bf673; case 7: return 0x12e13ce1;
        case 8: return 0x2cd76477; case 9: return 0x1ed90d2e; case 10: return 0x29a4b1ba; case 11: return 0x3a8e5ff9;
        case 12: return 0x001a0111; default: return 0;
    }
}
// R^3 mod p, R = 2^384 (plain limbs): mont_mul(x^-1 of a Montgomery-form x, R^3) = x^-1 in Montgomery form
EKZG_HD constexpr uint32_t r3(int i) {
    switch (i) {
        case 0: return 0xd94ca1e0u; case 1: return 0xed48ac6bu; case 2: return 0x03a7adf8u; case 3: return 0x315f831eu;
        case 4: return 0x615e29ddu; case 5: return 0x9a53352au; case 6: return 0x921e1761u; case 7: return 0x34c04e5eu;
        case 8: return 0x65724728u; case 9: return 0x2512d435u; case 10: return 0x91755d4du; case 11: return 0x0aa63460u;
        default: return 0u;
    }
}

struct S30 { int32_t v[L]; };
struct Mat { int32_t u, v, q, r; };         // 2^30 times the transition matrix of 30 division steps

// 30 division steps on the low words of f (odd) and g; zeta = -(delta + 1/2)
EKZG_HD int32_t divsteps_30(int32_t zeta, uint32_t f, uint32_t g, Mat& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 5
    for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(zeta >> 31);          // all ones: delta > 0
        const uint32_t c2 = 0u - (g & 1u);             // all ones: g odd
        const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;   // -f, -u, -v if delta > 0
        g += x & c2; q += y & c2; r += z & c2;         // g odd: g +- f
        c1 &= c2;                                      // swap case: delta > 0 and g odd
        zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;     // delta <- 1 - delta, else delta + 1
        f += g & c1; u += q & c1; v += r & c1;         // f <- old g
        g >>= 1; u <<= 1; v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}

// 32 x 32 -> 64 signed product (one IMAD.WIDE on the device; written so that the compiler sees two 32-bit operands)
EKZG_HD int64_t mul32(int32_t a, int32_t b) { return (int64_t)a * (int64_t)b; }

// (f, g) <- t * (f, g) / 2^30   (exact)
EKZG_HD void update_fg(S30& f, S30& g, const Mat& t) {
    int64_t cf = mul32(t.u, f.v[0]) + mul32(t.v, g.v[0]);
    int64_t cg = mul32(t.q, f.v[0]) + mul32(t.r, g.v[0]);
    cf >>= 30; cg >>= 30;
#pragma unroll
    for (int i = 1; i < L; i++) {
        const int32_t fi = f.v[i], gi = g.v[i];
        cf += mul32(t.u, fi) + mul32(t.v, gi);
        cg += mul32(t.q, fi) + mul32(t.r, gi);
        f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
    }
    f.v[L - 1] = (int32_t)cf;
    g.v[L - 1] = (int32_t)cg;
}

// (d, e) <- t * (d, e) / 2^30 mod p, both kept in (-2p, p)
EKZG_HD void update_de(S30& d, S30& e, const Mat& t) {
    const int32_t sd = d.v[L - 1] >> 31, se = e.v[L - 1] >> 31;
    int32_t md = (t.u & sd) + (t.v & se);
    int32_t me = (t.q & sd) + (t.r & se);
    int64_t cd = mul32(t.u, d.v[0]) + mul32(t.v, e.v[0]);
    int64_t ce = mul32(t.q, d.v[0]) + mul32(t.r, e.v[0]);
    md -= (int32_t)((P_INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);   // makes the low 30 bits of cd + p*md vanish
    me -= (int32_t)((P_INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += mul32(p30(0), md);
    ce += mul32(p30(0), me);
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < L; i++) {
        const int32_t di = d.v[i], ei = e.v[i];
        cd += mul32(t.u, di) + mul32(t.v, ei) + mul32(p30(i), md);
        ce += mul32(t.q, di) + mul32(t.r, ei) + mul32(p30(i), me);
        d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
    }
    d.v[L - 1] = (int32_t)cd;
    e.v[L - 1] = (int32_t)ce;
}

// r in (-2p, p) -> sign(f) * r in [0, p)
EKZG_HD void normalize(S30& r, int32_t sign) {
    int32_t add = r.v[L - 1] >> 31;
    const int32_t neg = sign >> 31;
#pragma unroll
    for (int i = 0; i < L; i++) {
        r.v[i] += p30(i) & add;
        r.v[i] = (r.v[i] ^ neg) - neg;
    }
#pragma unroll
    for (int i = 0; i < L - 1; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
    add = r.v[L - 1] >> 31;
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] += p30(i) & add;
#pragma unroll
    for (int i = 0; i < L - 1; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
}

EKZG_HD void to_s30(S30& r, const Fp& a) {
#pragma unroll
    for (int i = 0; i < L; i++) {
        const int bit = 30 * i, w = bit >> 5, s = bit & 31;
        uint32_t x = a.v[w] >> s;
        if (s > 2 && w + 1 < 12) x |= a.v[w + 1] << (32 - s);
        r.v[i] = (int32_t)(x & (uint32_t)M30);
    }
}
EKZG_HD void from_s30(Fp& r, const S30& a) {   // a in [0, p), limbs in [0, 2^30)
#pragma unroll
    for (int w = 0; w < 12; w++) {
        const int bit = 32 * w, i = bit / 30, s = bit - 30 * i;   // limb i contributes its bits s..29 at word bit 0
        uint32_t x = (uint32_t)a.v[i] >> s;
        x |= (uint32_t)a.v[i + 1] << (30 - s);
        if (30 - s + 30 < 32 && i + 2 < L) x |= (uint32_t)a.v[i + 2] << (60 - s);
        r.v[w] = x;
    }
}

}  // namespace gcdinv

// out = a^-1 for a in Montgomery form, result in Montgomery form; a = 0 gives 0.
EKZG_HD_CALL void fp_inv_gcd(Fp& out, const Fp& a) {
    using namespace gcdinv;
    S30 f, g, d, e;
#pragma unroll
    for (int i = 0; i < L; i++) { f.v[i] = p30(i); d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    to_s30(g, a);
    int32_t zeta = -1;
#pragma unroll 1
    for (int it = 0; it < 30; it++) {
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < L; i++) nz |= g.v[i];
        if (nz == 0) break;
        Mat t;
        zeta = divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        update_de(d, e, t);
        update_fg(f, g, t);
    }
    normalize(d, f.v[L - 1]);        // f = +-1 when a != 0 (p is prime); d * a = f
    Fp raw, r3c;
    from_s30(raw, d);
#pragma unroll
    for (int i = 0; i < 12; i++) r3c.v[i] = r3(i);
    fe_mul(out, raw, r3c);           // (aR)^-1 * R^3 / R = a^-1 R
}

}  // namespace ekzg
