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

}  // namespace ekzg
