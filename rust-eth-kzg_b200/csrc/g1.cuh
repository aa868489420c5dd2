// BLS12-381 G1 (y^2 = x^3 + 4 over Fp) point arithmetic, one point per thread.
//
// Replaces blst's P1 arithmetic that the reference uses through blstrs `G1Projective`/`G1Affine`
// (crates/cryptography/bls12_381/src/lib.rs:23-42) and the reference's own batched-affine adder
// (crates/cryptography/bls12_381/src/batch_addition.rs:14-39).  Only the final compressed bytes are
// pinned by the consensus vectors, so the coordinate systems are chosen for the GPU:
//   * XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) for accumulating table entries: mixed add 8M+2S, of which the two
//     products of  y3 = R*(Q - X3) - Y1*PPP  share one Montgomery reduction (fp_mul2_add);
//   * Jacobian for doubling-heavy scalar multiplication in the G1 NTT: doubling 2M+5S.  The Jacobian formulas keep
//     to the two subroutines fp_mul / fp_sqr: a third one (fp_mul2_add) pushed K5's hot code past the 32 KB
//     instruction cache and cost more than the saved reduction (measured: 34.8 -> 35.4 ms).
// All adders are COMPLETE: identity operands, P+P and P+(-P) are detected and handled, because a
// constant blob makes every proof the identity (SURVEY.md §7 "Identity handling is mandatory").
#pragma once
#include "field.cuh"

namespace ekzg {

struct G1Affine {  // identity encoded as (0, 0), which is not on the curve
    Fp x, y;
};
struct G1Xyzz {  // identity <=> zz == 0
    Fp x, y, zz, zzz;
};
struct G1Jac {  // identity <=> z == 0
    Fp x, y, z;
};

// word k of a point (x first): what the vector loads / stores of fr_ntt.cuh and g1_ntt_units.cuh address the members through
EKZG_HD uint32_t& limb_word(G1Affine& p, int k) { return k < 12 ? p.x.v[k] : p.y.v[k - 12]; }
EKZG_HD const uint32_t& limb_word(const G1Affine& p, int k) { return k < 12 ? p.x.v[k] : p.y.v[k - 12]; }
EKZG_HD uint32_t& limb_word(G1Jac& p, int k) { return k < 12 ? p.x.v[k] : k < 24 ? p.y.v[k - 12] : p.z.v[k - 24]; }
EKZG_HD const uint32_t& limb_word(const G1Jac& p, int k) { return k < 12 ? p.x.v[k] : k < 24 ? p.y.v[k - 12] : p.z.v[k - 24]; }
EKZG_HD uint32_t& limb_word(G1Xyzz& p, int k) { return k < 12 ? p.x.v[k] : k < 24 ? p.y.v[k - 12] : k < 36 ? p.zz.v[k - 24] : p.zzz.v[k - 36]; }
EKZG_HD const uint32_t& limb_word(const G1Xyzz& p, int k) { return k < 12 ? p.x.v[k] : k < 24 ? p.y.v[k - 12] : k < 36 ? p.zz.v[k - 24] : p.zzz.v[k - 36]; }

EKZG_HD bool g1a_is_inf(const G1Affine& p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
EKZG_HD void g1a_set_inf(G1Affine& p) { fe_set_zero(p.x); fe_set_zero(p.y); }
EKZG_HD bool xyzz_is_inf(const G1Xyzz& p) { return fe_is_zero(p.zz); }
EKZG_HD void xyzz_set_inf(G1Xyzz& p) { fe_set_zero(p.x); fe_set_zero(p.y); fe_set_zero(p.zz); fe_set_zero(p.zzz); }
EKZG_HD bool jac_is_inf(const G1Jac& p) { return fe_is_zero(p.z); }
EKZG_HD void jac_set_inf(G1Jac& p) { fe_set_zero(p.x); fe_set_zero(p.y); fe_set_zero(p.z); }

EKZG_HD void xyzz_from_affine(G1Xyzz& r, const G1Affine& p) {
    if (g1a_is_inf(p)) { xyzz_set_inf(r); return; }
    r.x = p.x; r.y = p.y; fe_set_one(r.zz); fe_set_one(r.zzz);
}
EKZG_HD void jac_from_affine(G1Jac& r, const G1Affine& p) {
    if (g1a_is_inf(p)) { jac_set_inf(r); return; }
    r.x = p.x; r.y = p.y; fe_set_one(r.z);
}

// 2*P for affine P != identity  (mdbl-2008-s-1, a = 0)
EKZG_HD_CALL void xyzz_dbl_affine(G1Xyzz& r, const G1Affine& p) {
    Fp u, v, w, s, m, t;
    fe_dbl(u, p.y);
    fe_sqr(v, u);
    fe_mul(w, u, v);
    fe_mul(s, p.x, v);
    fe_sqr(t, p.x);
    fe_dbl(m, t); fe_add(m, m, t);
    fe_sqr(r.x, m); fe_sub(r.x, r.x, s); fe_sub(r.x, r.x, s);
    fe_sub(t, s, r.x);
    fe_neg(u, p.y);
    fp_mul2_add(r.y, m, t, w, u);   // M*(S - X3) - W*Y1
    r.zz = v; r.zzz = w;
}

// acc += (neg ? -P : P), P affine  (madd-2008-s: 8M + 2S)
EKZG_HD void xyzz_madd(G1Xyzz& acc, const G1Affine& p_in, bool neg) {
    if (g1a_is_inf(p_in)) return;
    G1Affine p;
    p.x = p_in.x;
    fe_cneg(p.y, p_in.y, neg);
    if (xyzz_is_inf(acc)) { acc.x = p.x; acc.y = p.y; fe_set_one(acc.zz); fe_set_one(acc.zzz); return; }
    Fp pp, rr, t, ppp, q;
    fe_mul(pp, p.x, acc.zz);   // U2
    fe_mul(rr, p.y, acc.zzz);  // S2
    fe_sub(pp, pp, acc.x);     // P = U2 - X1
    fe_sub(rr, rr, acc.y);     // R = S2 - Y1
    if (fe_is_zero(pp)) {
        if (fe_is_zero(rr)) xyzz_dbl_affine(acc, p); else xyzz_set_inf(acc);
        return;
    }
    fe_sqr(t, pp);             // PP
    fe_mul(ppp, pp, t);        // PPP
    fe_mul(q, acc.x, t);       // Q
    fe_mul(acc.zz, acc.zz, t);
    fe_mul(acc.zzz, acc.zzz, ppp);
    fe_sqr(t, rr);
    fe_sub(t, t, ppp); fe_sub(t, t, q); fe_sub(t, t, q);  // X3
    fe_sub(q, q, t);
    fe_neg(ppp, ppp);
    fp_mul2_add(acc.y, rr, q, acc.y, ppp);   // R*(Q - X3) - Y1*PPP, one reduction
    acc.x = t;
}

// 2*P, general XYZZ  (dbl-2008-s-1, a = 0)
EKZG_HD_CALL void xyzz_dbl(G1Xyzz& r, const G1Xyzz& p) {
    if (xyzz_is_inf(p)) { xyzz_set_inf(r); return; }
    Fp u, v, w, s, m, t;
    fe_dbl(u, p.y);
    fe_sqr(v, u);
    fe_mul(w, u, v);
    fe_mul(s, p.x, v);
    fe_sqr(t, p.x);
    fe_dbl(m, t); fe_add(m, m, t);
    Fp x3;
    fe_sqr(x3, m); fe_sub(x3, x3, s); fe_sub(x3, x3, s);
    fe_sub(t, s, x3);
    fe_neg(u, p.y);
    fe_mul(r.zz, v, p.zz);
    fe_mul(r.zzz, w, p.zzz);
    fp_mul2_add(r.y, m, t, w, u);   // M*(S - X3) - W*Y1
    r.x = x3;
}

// acc += q, both XYZZ  (add-2008-s: 12M + 2S)
EKZG_HD_CALL void xyzz_add(G1Xyzz& acc, const G1Xyzz& q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(acc)) { acc = q; return; }
    Fp u1, u2, s1, s2, pp, ppp, t;
    fe_mul(u1, acc.x, q.zz);
    fe_mul(u2, q.x, acc.zz);
    fe_mul(s1, acc.y, q.zzz);
    fe_mul(s2, q.y, acc.zzz);
    fe_sub(u2, u2, u1);  // P
    fe_sub(s2, s2, s1);  // R
    if (fe_is_zero(u2)) {
        if (fe_is_zero(s2)) { G1Xyzz d; xyzz_dbl(d, acc); acc = d; } else xyzz_set_inf(acc);
        return;
    }
    fe_sqr(pp, u2);
    fe_mul(ppp, u2, pp);
    fe_mul(u1, u1, pp);  // Q
    fe_mul(acc.zz, acc.zz, q.zz); fe_mul(acc.zz, acc.zz, pp);
    fe_mul(acc.zzz, acc.zzz, q.zzz); fe_mul(acc.zzz, acc.zzz, ppp);
    fe_sqr(t, s2);
    fe_sub(t, t, ppp); fe_sub(t, t, u1); fe_sub(t, t, u1);  // X3
    fe_sub(u1, u1, t);
    fe_neg(ppp, ppp);
    fp_mul2_add(acc.y, s2, u1, s1, ppp);   // R*(Q - X3) - S1*PPP
    acc.x = t;
}

// XYZZ -> Jacobian without inversion: Z = ZZ*ZZZ (= z^5), X' = X*ZZ*ZZZ^2, Y' = Y*ZZ^3*ZZZ^2
EKZG_HD void jac_from_xyzz(G1Jac& r, const G1Xyzz& p) {
    if (xyzz_is_inf(p)) { jac_set_inf(r); return; }
    Fp a, b, c;
    fe_sqr(a, p.zzz);         // ZZZ^2
    fe_mul(a, a, p.zz);       // ZZ*ZZZ^2
    fe_sqr(b, p.zz);          // ZZ^2
    fe_mul(c, a, b);          // ZZ^3*ZZZ^2
    fe_mul(r.x, p.x, a);
    fe_mul(r.y, p.y, c);
    fe_mul(r.z, p.zz, p.zzz);
}

// 2*P Jacobian (dbl-2009-l, a = 0): 2M + 5S.  Identity stays identity (Z3 = 2*Y*0).
// (the _inl forms are the same formulas force-inlined: the fixed-scalar ladder of K5 keeps its accumulator in registers across
// them, whereas a real call passes the point through local memory -- 72 words per doubling, ncu: 18 % of K5's time)
EKZG_HD void jac_dbl_inl(G1Jac& r, const G1Jac& p) {
    Fp a, b, c, d, e, f;
    fe_sqr(a, p.x);
    fe_sqr(b, p.y);
    fe_sqr(c, b);
    fe_add(d, p.x, b); fe_sqr(d, d); fe_sub(d, d, a); fe_sub(d, d, c); fe_dbl(d, d);
    fe_dbl(e, a); fe_add(e, e, a);
    fe_sqr(f, e);
    fe_mul(r.z, p.y, p.z); fe_dbl(r.z, r.z);
    fe_sub(f, f, d); fe_sub(f, f, d);  // X3
    fe_sub(d, d, f);
    fe_mul(d, e, d);
    fe_dbl(c, c); fe_dbl(c, c); fe_dbl(c, c);
    fe_sub(r.y, d, c);
    r.x = f;
}
EKZG_HD_CALL void jac_dbl(G1Jac& r, const G1Jac& p) { jac_dbl_inl(r, p); }

// acc += q, both Jacobian  (add-2007-bl: 11M + 5S)
EKZG_HD_CALL void jac_add(G1Jac& acc, const G1Jac& q) {
    if (jac_is_inf(q)) return;
    if (jac_is_inf(acc)) { acc = q; return; }
    Fp z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v;
    fe_sqr(z1z1, acc.z);
    fe_sqr(z2z2, q.z);
    fe_mul(u1, acc.x, z2z2);
    fe_mul(u2, q.x, z1z1);
    fe_mul(s1, acc.y, q.z); fe_mul(s1, s1, z2z2);
    fe_mul(s2, q.y, acc.z); fe_mul(s2, s2, z1z1);
    fe_sub(h, u2, u1);
    fe_sub(rr, s2, s1);
    if (fe_is_zero(h)) {
        if (fe_is_zero(rr)) { G1Jac d; jac_dbl(d, acc); acc = d; } else jac_set_inf(acc);
        return;
    }
    fe_dbl(rr, rr);
    fe_dbl(i, h); fe_sqr(i, i);
    fe_mul(j, h, i);
    fe_mul(v, u1, i);
    // Z3 = ((Z1+Z2)^2 - Z1Z1 - Z2Z2) * H
    fe_add(u2, acc.z, q.z); fe_sqr(u2, u2); fe_sub(u2, u2, z1z1); fe_sub(u2, u2, z2z2);
    fe_mul(acc.z, u2, h);
    fe_sqr(u2, rr); fe_sub(u2, u2, j); fe_sub(u2, u2, v); fe_sub(u2, u2, v);  // X3
    fe_sub(v, v, u2); fe_mul(v, rr, v);
    fe_mul(s1, s1, j); fe_dbl(s1, s1);
    fe_sub(acc.y, v, s1);
    acc.x = u2;
}

EKZG_HD void jac_neg(G1Jac& r, const G1Jac& p) { r.x = p.x; fe_neg(r.y, p.y); r.z = p.z; }
EKZG_HD void jac_cneg(G1Jac& r, const G1Jac& p, bool neg) { r.x = p.x; fe_cneg(r.y, p.y, neg); r.z = p.z; }

// acc += (neg ? -P : P), P affine  (madd-2007-bl: 7M + 4S)
EKZG_HD void jac_madd_inl(G1Jac& acc, const G1Affine& p_in, bool neg) {
    if (g1a_is_inf(p_in)) return;
    G1Affine p;
    p.x = p_in.x;
    fe_cneg(p.y, p_in.y, neg);
    if (jac_is_inf(acc)) { acc.x = p.x; acc.y = p.y; fe_set_one(acc.z); return; }
    Fp z1z1, u2, s2, h, hh, i, j, rr, v;
    fe_sqr(z1z1, acc.z);
    fe_mul(u2, p.x, z1z1);
    fe_mul(s2, p.y, acc.z); fe_mul(s2, s2, z1z1);
    fe_sub(h, u2, acc.x);
    fe_sub(rr, s2, acc.y);
    if (fe_is_zero(h)) {
        if (fe_is_zero(rr)) { G1Jac d; jac_dbl(d, acc); acc = d; } else jac_set_inf(acc);
        return;
    }
    fe_dbl(rr, rr);
    fe_sqr(hh, h);
    fe_dbl(i, hh); fe_dbl(i, i);
    fe_mul(j, h, i);
    fe_mul(v, acc.x, i);
    fe_add(u2, acc.z, h); fe_sqr(u2, u2); fe_sub(u2, u2, z1z1); fe_sub(acc.z, u2, hh);
    fe_sqr(u2, rr); fe_sub(u2, u2, j); fe_sub(u2, u2, v); fe_sub(u2, u2, v);  // X3
    fe_sub(v, v, u2); fe_mul(v, rr, v);
    fe_mul(s2, acc.y, j); fe_dbl(s2, s2);
    fe_sub(acc.y, v, s2);
    acc.x = u2;
}
EKZG_HD_CALL void jac_madd(G1Jac& acc, const G1Affine& p_in, bool neg) { jac_madd_inl(acc, p_in, neg); }

// phi(P) = (beta*x, y, z): multiplication by lambda (GLV endomorphism)
EKZG_HD void jac_endo(G1Jac& r, const G1Jac& p) {
    Fp beta;
#pragma unroll
    for (int j = 0; j < 12; j++) beta.v[j] = FpParams::beta(j);
    fe_mul(r.x, p.x, beta);
    r.y = p.y; r.z = p.z;
}

// Jacobian -> affine given zinv = 1/Z (Montgomery form)
EKZG_HD void jac_to_affine_with_inv(G1Affine& r, const G1Jac& p, const Fp& zinv) {
    Fp zi2, zi3;
    fe_sqr(zi2, zinv);
    fe_mul(zi3, zi2, zinv);
    fe_mul(r.x, p.x, zi2);
    fe_mul(r.y, p.y, zi3);
}

// serialize an affine point (Montgomery coordinates) to the 48-byte compressed wire format
// (crates/serialization/src/lib.rs:84-86 -> blstrs to_compressed):
// big-endian x; byte0 bit7 = compressed, bit6 = identity, bit5 = y > (p-1)/2.
EKZG_HD_CALL void g1a_compress(uint8_t* out, const G1Affine& p) {
    if (g1a_is_inf(p)) {
        out[0] = 0xc0;
        for (int i = 1; i < 48; i++) out[i] = 0;
        return;
    }
    Fp x, y;
    fe_from_mont(x, p.x);
    fe_from_mont(y, p.y);
    bool big = fe_plain_gt_half(y);
#pragma unroll
    for (int j = 0; j < 12; j++) {
        uint32_t w = x.v[11 - j];
        out[4 * j + 0] = (uint8_t)(w >> 24);
        out[4 * j + 1] = (uint8_t)(w >> 16);
        out[4 * j + 2] = (uint8_t)(w >> 8);
        out[4 * j + 3] = (uint8_t)w;
    }
    out[0] |= big ? 0xa0 : 0x80;
}

// curve membership y^2 == x^3 + 4 (Montgomery coordinates)
EKZG_HD bool g1a_on_curve(const G1Affine& p) {
    Fp l, r, b;
    fe_sqr(l, p.y);
    fe_sqr(r, p.x); fe_mul(r, r, p.x);
#pragma unroll
    for (int j = 0; j < 12; j++) b.v[j] = FpParams::b4(j);
    fe_add(r, r, b);
    return fe_eq(l, r);
}

// parse 48 compressed bytes. returns 0 ok, 1 malformed / not on curve.  No subgroup check here.
EKZG_HD_CALL int g1a_decompress(G1Affine& r, const uint8_t* in) {
    uint8_t b0 = in[0];
    if (!(b0 & 0x80)) return 1;  // uncompressed form not accepted for 48-byte input
    bool inf = b0 & 0x40, sign = b0 & 0x20;
    Fp x;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        uint32_t w = ((uint32_t)in[4 * j] << 24) | ((uint32_t)in[4 * j + 1] << 16) | ((uint32_t)in[4 * j + 2] << 8) | in[4 * j + 3];
        if (j == 0) w &= 0x1fffffffu;
        x.v[11 - j] = w;
    }
    if (inf) {
        if (sign || !fe_is_zero(x)) return 1;
        g1a_set_inf(r);
        return 0;
    }
    if (fe_plain_ge_mod(x)) return 1;
    Fp xm, y2, y, b, chk;
    fe_to_mont(xm, x);
    fe_sqr(y2, xm); fe_mul(y2, y2, xm);
#pragma unroll
    for (int j = 0; j < 12; j++) b.v[j] = FpParams::b4(j);
    fe_add(y2, y2, b);
    fp_sqrt_candidate(y, y2);
    fe_sqr(chk, y);
    if (!fe_eq(chk, y2)) return 1;
    Fp yp;
    fe_from_mont(yp, y);
    if (fe_plain_gt_half(yp) != sign) fe_neg(y, y);
    r.x = xm; r.y = y;
    return 0;
}

}  // namespace ekzg
