// Scalar multiplication building blocks.
//
//  * booth_digit: signed window digit of a 256-bit little-endian scalar, the closed form of the
//    reference's get_booth_index (crates/cryptography/bls12_381/src/booth_encoding.rs:4-46;
//    SURVEY.md Appendix A.4):  v = ((s << 1) >> (t*w)) & (2^(w+1)-1),  d = ((v+1)>>1) - (v>>w)*2^w.
//  * jac_mul_glv16: k*P for a FIXED scalar k given as two 33-digit signed radix-16 strings
//    (k = k1 + k2*lambda, tools/gen_device_constants.py).  This is what a butterfly of the G1 NTT
//    does (reference: `*b * twiddle`, crates/cryptography/polynomial/src/fft.rs:164-177): the GLV
//    endomorphism halves the doublings (132 instead of 255) and the schedule is the same for every
//    twiddle, so lanes with different twiddles do not diverge.
//  * jac_mul_u256: generic double-and-add with a 4-bit window (setup only).
#pragma once
#include "g1.cuh"

namespace ekzg {

// bits [lo, lo+cnt) of an 8-limb little-endian integer, cnt <= 25, lo may be -1 (bit -1 := 0)
EKZG_HD uint32_t u256_bits(const uint32_t* s, int lo, int cnt) {
    // value of (s << 1) >> (lo+1)
    int sh = lo + 1;  // >= 0, position in (s<<1)
    // (s<<1) bit i = s bit i-1
    uint64_t acc = 0;
    int limb = sh >> 5;
    int off = sh & 31;
    // gather up to 3 limbs of (s<<1): limb L of (s<<1) = (s[L] << 1) | (s[L-1] >> 31)
    uint32_t l0 = 0, l1 = 0;
    {
        int L = limb;
        uint32_t cur = (L < 8) ? s[L] : 0u;
        uint32_t prv = (L >= 1 && L <= 8) ? s[L - 1] : 0u;
        l0 = (cur << 1) | (prv >> 31);
        L = limb + 1;
        cur = (L < 8) ? s[L] : 0u;
        prv = (L >= 1 && L <= 8) ? s[L - 1] : 0u;
        l1 = (cur << 1) | (prv >> 31);
    }
    acc = ((uint64_t)l1 << 32) | l0;
    return (uint32_t)(acc >> off) & ((1u << cnt) - 1u);
}

// signed Booth digit of window t (width w <= 24) of plain scalar s; result in [-2^(w-1), 2^(w-1)]
EKZG_HD int booth_digit(const uint32_t* s, int t, int w) {
    uint32_t v = u256_bits(s, t * w - 1, w + 1);
    return (int)((v + 1) >> 1) - (int)((v >> w) << w);
}

// k*P, k = k1 + k2*lambda given as signed radix-16 digits d[0..32] (k1) and d[33..65] (k2).
EKZG_HD_CALL void jac_mul_glv16(G1Jac& out, const G1Jac& p, const int8_t* d) {
    G1Jac tbl[8];  // tbl[i] = (i+1)*P
    tbl[0] = p;
    jac_dbl(tbl[1], p);
    tbl[2] = tbl[1]; jac_add(tbl[2], p);
    jac_dbl(tbl[3], tbl[1]);
    tbl[4] = tbl[3]; jac_add(tbl[4], p);
    jac_dbl(tbl[5], tbl[2]);
    tbl[6] = tbl[5]; jac_add(tbl[6], p);
    jac_dbl(tbl[7], tbl[3]);
    G1Jac acc;
    jac_set_inf(acc);
    for (int i = 32; i >= 0; i--) {
        if (i != 32) {
            for (int s = 0; s < 4; s++) jac_dbl(acc, acc);
        }
        int d1 = d[i], d2 = d[33 + i];
        if (d1 != 0) {
            int a = d1 < 0 ? -d1 : d1;
            G1Jac t;
            jac_cneg(t, tbl[a - 1], d1 < 0);
            jac_add(acc, t);
        }
        if (d2 != 0) {
            int a = d2 < 0 ? -d2 : d2;
            G1Jac t;
            jac_endo(t, tbl[a - 1]);
            jac_cneg(t, t, d2 < 0);
            jac_add(acc, t);
        }
    }
    out = acc;
}

// one GLV half on its own: sum_i d[i] 16^i * Q for the 33 signed radix-16 digits of a half (Q = P for the low half, phi(P) for the
// high one).  k_scalar_mul_split gives the two halves of a multiplication to two lanes: 128 doublings + 33 additions on the critical
// path instead of 128 + 66.
EKZG_HD_CALL void jac_mul_half16(G1Jac& out, const G1Jac& q, const int8_t* d /*33*/) {
    G1Jac tbl[8];  // tbl[i] = (i+1)*Q
    tbl[0] = q;
    jac_dbl(tbl[1], q);
    tbl[2] = tbl[1]; jac_add(tbl[2], q);
    jac_dbl(tbl[3], tbl[1]);
    tbl[4] = tbl[3]; jac_add(tbl[4], q);
    jac_dbl(tbl[5], tbl[2]);
    tbl[6] = tbl[5]; jac_add(tbl[6], q);
    jac_dbl(tbl[7], tbl[3]);
    G1Jac acc;
    jac_set_inf(acc);
    for (int i = 32; i >= 0; i--) {
        if (i != 32) {
            for (int s = 0; s < 4; s++) jac_dbl(acc, acc);
        }
        const int d1 = d[i];
        if (d1 != 0) {
            const int a = d1 < 0 ? -d1 : d1;
            G1Jac t;
            jac_cneg(t, tbl[a - 1], d1 < 0);
            jac_add(acc, t);
        }
    }
    out = acc;
}

// GLV split of a VARIABLE scalar k < r:  k = k1 + k2*lambda with k2 = floor(k / lambda), k1 = k mod lambda.  Because
// r = lambda^2 + lambda + 1, both halves are non-negative and below 2^128.  The quotient comes from the 129-bit reciprocal
// floor(2^256 / lambda) (at most 2 too small, fixed up by subtraction), then both halves are recoded into the signed radix-16
// digits jac_mul_glv16 takes: 128 doublings + 66 additions instead of the 252 + 63 of jac_mul_u256.
// Used by the verifiers' random-linear-combination scalar multiplications (reference: g1_lincomb -> blst Pippenger,
// crates/cryptography/bls12_381/src/lincomb.rs:7-30; one multiplication per thread here).
// the two halves on their own: k = k1 + k2*lambda, k1 and k2 as 5 plain limbs each (both below 2^128: limb 4 is zero on return)
EKZG_HD void glv_split_halves(uint32_t* k1 /*5*/, uint32_t* q /*5*/, const uint32_t* k /*8 limbs, < r*/) {
    const uint32_t lam[4] = {GLV_LAMBDA[0], GLV_LAMBDA[1], GLV_LAMBDA[2], GLV_LAMBDA[3]};
    const uint32_t rec[5] = {0xf6cfee30u, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x1u};   // floor(2^256 / lambda)
    // q = (k * rec) >> 256
    uint32_t prod[13];
    for (int i = 0; i < 13; i++) prod[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 5; j++) {
            uint64_t t = (uint64_t)k[i] * rec[j] + prod[i + j] + c;
            prod[i + j] = (uint32_t)t;
            c = t >> 32;
        }
        prod[i + 5] = (uint32_t)c;
    }
    for (int i = 0; i < 5; i++) q[i] = prod[8 + i];
    // k1 = k - q*lambda, exact below 3*lambda < 2^130: 5 limbs are enough
    uint32_t ql[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) {
            uint64_t t = (uint64_t)q[i] * lam[j] + ql[i + j] + c;
            ql[i + j] = (uint32_t)t;
            c = t >> 32;
        }
        ql[i + 4] = (uint32_t)c;
    }
    {
        uint64_t b = 0;
        for (int i = 0; i < 5; i++) {
            uint64_t t = (uint64_t)k[i] - ql[i] - b;
            k1[i] = (uint32_t)t;
            b = (t >> 32) & 1;
        }
    }
    for (int it = 0; it < 3; it++) {
        // k1 >= lambda ?
        bool ge = k1[4] != 0;
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--) {
                if (k1[i] != lam[i]) { ge = k1[i] > lam[i]; break; }
            }
        }
        if (!ge) break;
        uint64_t b = 0;
        for (int i = 0; i < 5; i++) {
            uint64_t t = (uint64_t)k1[i] - (i < 4 ? lam[i] : 0u) - b;
            k1[i] = (uint32_t)t;
            b = (t >> 32) & 1;
        }
        uint64_t c = 1;
        for (int i = 0; i < 5; i++) { uint64_t t = (uint64_t)q[i] + c; q[i] = (uint32_t)t; c = t >> 32; }
    }
}
EKZG_HD void glv_split_digits(int8_t* d /*66*/, const uint32_t* k /*8 limbs, < r*/) {
    uint32_t k1[5], q[5];
    glv_split_halves(k1, q, k);
    // signed radix-16 digits in [-7, 8], least significant first, 33 per half
    for (int h = 0; h < 2; h++) {
        const uint32_t* v = h ? q : k1;
        int carry = 0;
        for (int i = 0; i < 32; i++) {
            int x = (int)((v[i >> 3] >> ((i & 7) * 4)) & 15u) + carry;
            carry = x > 8;
            d[33 * h + i] = (int8_t)(carry ? x - 16 : x);
        }
        d[33 * h + 32] = (int8_t)carry;
    }
}
EKZG_HD_CALL void jac_mul_fr_glv(G1Jac& out, const G1Jac& p, const uint32_t* k /*8 limbs, < r*/) {
    int8_t d[66];
    glv_split_digits(d, k);
    jac_mul_glv16(out, p, d);
}

// r = a + b for Jacobian a, affine b, a != +-b, neither the identity (madd-2004-hmv: 8M + 3S, Z3 = Z1*H);
// zr = H = Z3/Z1 is handed back for the common-Z table below.
EKZG_HD_CALL void jac_madd_zr(G1Jac& r, const G1Jac& a, const G1Affine& b, Fp& zr) {
    Fp t1, t2, t3, t4;
    fe_sqr(t1, a.z);
    fe_mul(t2, t1, a.z);
    fe_mul(t1, t1, b.x);
    fe_mul(t2, t2, b.y);
    fe_sub(t1, t1, a.x);       // H
    fe_sub(t2, t2, a.y);       // R
    zr = t1;
    fe_mul(r.z, a.z, t1);
    fe_sqr(t3, t1);            // HH
    fe_mul(t4, t3, t1);        // HHH
    fe_mul(t3, t3, a.x);       // V
    fe_dbl(t1, t3);
    Fp x3;
    fe_sqr(x3, t2);
    fe_sub(x3, x3, t1);
    fe_sub(x3, x3, t4);
    fe_sub(t3, t3, x3);
    fe_mul(t3, t3, t2);
    fe_mul(t4, t4, a.y);
    fe_sub(r.y, t3, t4);
    r.x = x3;
}

// k*P for a FIXED scalar given as an op list (tools/gen_device_constants.py: width-5 NAF of the two GLV
// halves, merged).  This is the `*b * twiddle` of the reference's G1 butterfly (polynomial/src/fft.rs:164-177).
//   1. odd multiples P, 3P, .., 15P by repeated mixed addition of 2P on the curve isomorphic by 2P's Z, then
//      rescaled to ONE common Z (the z-ratios of the additions are the H values), so the eight entries are
//      affine points of an isomorphic curve y^2 = x^3 + 4*Zg^6 -- no inversion;
//   2. the ladder runs on that curve with mixed additions (11 instead of 16 multiplications each; the
//      a = 0 formulas never see the curve constant); phi(x, y) = (beta*x, y) holds there as well;
//   3. Z *= Zg maps the result back.
// P must be a non-identity point of the prime-order subgroup (callers skip the identity).
constexpr int MULOPS_STRIDE = 64;
EKZG_HD_CALL void jac_mul_ops(G1Jac& out, const G1Jac& p, const uint16_t* ops) {
    G1Affine tbl[8];   // (2i+1)P, common Z
    Fp bx[8];          // beta * x of the same
    Fp zr[8];
    Fp zg;             // common Z (before the factor Z(2P))
    G1Jac d;
    jac_dbl(d, p);
    {
        G1Jac cur;
        Fp dz2, dz3;
        fe_sqr(dz2, d.z);
        fe_mul(dz3, dz2, d.z);
        fe_mul(cur.x, p.x, dz2);
        fe_mul(cur.y, p.y, dz3);
        cur.z = p.z;
        G1Affine da;
        da.x = d.x; da.y = d.y;
        tbl[0].x = cur.x; tbl[0].y = cur.y;
        for (int i = 1; i < 8; i++) {
            G1Jac nxt;
            jac_madd_zr(nxt, cur, da, zr[i]);
            cur = nxt;
            tbl[i].x = cur.x; tbl[i].y = cur.y;
        }
        zg = cur.z;
        Fp zs = zr[7];
        for (int i = 6; i >= 0; i--) {
            Fp z2, z3;
            fe_sqr(z2, zs);
            fe_mul(z3, z2, zs);
            fe_mul(tbl[i].x, tbl[i].x, z2);
            fe_mul(tbl[i].y, tbl[i].y, z3);
            if (i) fe_mul(zs, zs, zr[i]);
        }
        Fp beta;
#pragma unroll
        for (int j = 0; j < 12; j++) beta.v[j] = FpParams::beta(j);
        for (int i = 0; i < 8; i++) fe_mul(bx[i], tbl[i].x, beta);
    }
    G1Jac acc;
    jac_set_inf(acc);
    const int n = ops[0];
    for (int c = 1; c <= n; c++) {
        const uint32_t op = ops[c];
#if defined(__CUDA_ARCH__) && !defined(EKZG_K5_CALL_POINT_OPS)
        for (int s = op >> 8; s > 0; s--) jac_dbl_inl(acc, acc);
#else
        for (int s = op >> 8; s > 0; s--) jac_dbl(acc, acc);
#endif
        if (op & 0x20) {
            const int idx = op & 7;
            G1Affine e;
            e.x = (op & 0x10) ? bx[idx] : tbl[idx].x;
            e.y = tbl[idx].y;
#if defined(__CUDA_ARCH__) && !defined(EKZG_K5_CALL_POINT_OPS)
            jac_madd_inl(acc, e, (op & 8) != 0);
#else
            jac_madd(acc, e, (op & 8) != 0);
#endif
        }
    }
    if (!jac_is_inf(acc)) {
        fe_mul(acc.z, acc.z, zg);
        fe_mul(acc.z, acc.z, d.z);
    }
    out = acc;
}

// k*P for a plain 256-bit little-endian scalar (unsigned 4-bit windows); setup paths only
EKZG_HD_CALL void jac_mul_u256(G1Jac& out, const G1Jac& p, const uint32_t* k) {
    G1Jac tbl[15];
    tbl[0] = p;
    for (int i = 1; i < 15; i++) { tbl[i] = tbl[i - 1]; jac_add(tbl[i], p); }
    G1Jac acc;
    jac_set_inf(acc);
    for (int pos = 252; pos >= 0; pos -= 4) {
        for (int s = 0; s < 4; s++) jac_dbl(acc, acc);
        uint32_t dg = (k[pos >> 5] >> (pos & 31)) & 15u;
        if (dg) jac_add(acc, tbl[dg - 1]);
    }
    out = acc;
}

// |x| * P for the BLS parameter |x| = 0xd201000000010000 (bit 63 set, five more bits): 63 doublings + 5 additions
EKZG_HD_CALL void jac_mul_bls_x_abs(G1Jac& out, const G1Jac& p) {
    const uint64_t x = 0xd201000000010000ull;
    G1Jac acc = p;
    for (int bit = 62; bit >= 0; bit--) {
        jac_dbl(acc, acc);
        if ((x >> bit) & 1) jac_add(acc, p);
    }
    out = acc;
}

// Prime-order subgroup membership of a curve point (the check blstrs' from_compressed performs for
// crates/serialization/src/lib.rs:69-81): phi acts on G1 as multiplication by lambda = x^2 - 1, and
//   P in G1  <=>  x^2 * P == phi(P) + P
// (Scott, "A note on group membership tests for G1, G2 and GT on BLS pairing-friendly curves", eprint 2021/1130, in
// the form phi^2(P) = -x^2 P with phi^2 + phi + 1 = 0).  Two multiplications by the 64-bit |x| instead of one by the
// 255-bit r.  a must be on the curve and not the identity.
EKZG_HD_CALL bool g1a_in_subgroup(const G1Affine& a) {
    G1Jac p, t;
    jac_from_affine(p, a);
    jac_mul_bls_x_abs(t, p);
    jac_mul_bls_x_abs(t, t);          // x^2 P
    G1Jac s;
    jac_endo(s, p);
    jac_madd(s, a, false);            // phi(P) + P
    if (jac_is_inf(t) || jac_is_inf(s)) return jac_is_inf(t) && jac_is_inf(s);
    Fp z1z1, z2z2, l, r;
    fe_sqr(z1z1, t.z);
    fe_sqr(z2z2, s.z);
    fe_mul(l, t.x, z2z2);
    fe_mul(r, s.x, z1z1);
    if (!fe_eq(l, r)) return false;
    fe_mul(z1z1, z1z1, t.z);
    fe_mul(z2z2, z2z2, s.z);
    fe_mul(l, t.y, z2z2);
    fe_mul(r, s.y, z1z1);
    return fe_eq(l, r);
}

// op lists of the 128th roots of unity, row e = omega_128^e (host copy; the kernels keep one in __constant__)
static const uint16_t TWIDDLE_OPS_HOST[128][MULOPS_STRIDE] =
#include "twiddle_ops.inc"
    ;

}  // namespace ekzg
