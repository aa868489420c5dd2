/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Prime-field arithmetic for BLS12-381 Fp (6x64-bit limbs) and Fr (4x64-bit limbs) in Montgomery
 * form, written with unsigned __int128.  In the reference this arithmetic is third-party:
 * blst >=0.3.16 via blstrs 0.7.1 (crates/cryptography/bls12_381/Cargo.toml:17-23, type aliases
 * crates/cryptography/bls12_381/src/lib.rs:23-42).  It is restated here from the published
 * algorithm (CIOS Montgomery multiplication, Fermat inversion); results are pinned by the
 * consensus vectors through tests/test_oracle_vectors.py.
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include "constants.h"

typedef unsigned __int128 u128;

#define ALWAYS_INLINE static inline __attribute__((always_inline))

/* ---- generic n-limb helpers (n is a compile-time constant at every call site) ---- */
ALWAYS_INLINE int limbs_ge(const uint64_t *a, const uint64_t *b, const int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
ALWAYS_INLINE uint64_t limbs_add(uint64_t *r, const uint64_t *a, const uint64_t *b, const int n) {
    u128 c = 0;
    for (int i = 0; i < n; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
ALWAYS_INLINE uint64_t limbs_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, const int n) {
    uint64_t borrow = 0;
    for (int i = 0; i < n; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
ALWAYS_INLINE int limbs_is_zero(const uint64_t *a, const int n) {
    uint64_t x = 0; for (int i = 0; i < n; i++) x |= a[i]; return x == 0;
}
ALWAYS_INLINE void mont_add(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, const int n) {
    uint64_t t[6]; uint64_t c = limbs_add(t, a, b, n);
    if (c || limbs_ge(t, m, n)) limbs_sub(t, t, m, n);
    memcpy(r, t, 8 * n);
}
ALWAYS_INLINE void mont_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, const int n) {
    uint64_t t[6]; uint64_t bw = limbs_sub(t, a, b, n);
    if (bw) limbs_add(t, t, m, n);
    memcpy(r, t, 8 * n);
}
ALWAYS_INLINE void mont_neg(uint64_t *r, const uint64_t *a, const uint64_t *m, const int n) {
    if (limbs_is_zero(a, n)) { memset(r, 0, 8 * n); return; }
    limbs_sub(r, m, a, n);
}
/* CIOS Montgomery product r = a*b/2^(64n) mod m */
ALWAYS_INLINE void mont_mul(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, const uint64_t inv, const int n) {
    uint64_t t[8] = {0};
    for (int i = 0; i < n; i++) {
        u128 c = 0;
        for (int j = 0; j < n; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[n]; t[n] = (uint64_t)c; t[n + 1] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * inv;
        c = ((u128)q * m[0] + t[0]) >> 64;
        for (int j = 1; j < n; j++) { c += (u128)q * m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[n]; t[n - 1] = (uint64_t)c; t[n] = t[n + 1] + (uint64_t)(c >> 64);
    }
    if (t[n] || limbs_ge(t, m, n)) limbs_sub(t, t, m, n);
    memcpy(r, t, 8 * n);
}
/* r = a^e (e plain little-endian limbs, elimbs of them); a, r in Montgomery form */
ALWAYS_INLINE void mont_pow(uint64_t *r, const uint64_t *a, const uint64_t *e, int elimbs, const uint64_t *one,
                            const uint64_t *m, const uint64_t inv, const int n) {
    uint64_t acc[6], base[6];
    memcpy(acc, one, 8 * n); memcpy(base, a, 8 * n);
    int top = elimbs * 64 - 1;
    while (top >= 0 && !((e[top / 64] >> (top % 64)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        mont_mul(acc, acc, acc, m, inv, n);
        if ((e[i / 64] >> (i % 64)) & 1) mont_mul(acc, acc, base, m, inv, n);
    }
    memcpy(r, acc, 8 * n);
}

/* ---------------------------------------------------------------- Fp */
typedef struct { uint64_t l[6]; } fp_t;
ALWAYS_INLINE void fp_add(fp_t *r, const fp_t *a, const fp_t *b) { mont_add(r->l, a->l, b->l, FP_MOD, 6); }
ALWAYS_INLINE void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) { mont_sub(r->l, a->l, b->l, FP_MOD, 6); }
ALWAYS_INLINE void fp_neg(fp_t *r, const fp_t *a) { mont_neg(r->l, a->l, FP_MOD, 6); }
ALWAYS_INLINE void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) { mont_mul(r->l, a->l, b->l, FP_MOD, FP_INV, 6); }
ALWAYS_INLINE void fp_sqr(fp_t *r, const fp_t *a) { mont_mul(r->l, a->l, a->l, FP_MOD, FP_INV, 6); }
ALWAYS_INLINE int fp_is_zero(const fp_t *a) { return limbs_is_zero(a->l, 6); }
ALWAYS_INLINE int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a->l, b->l, 48) == 0; }
ALWAYS_INLINE void fp_set_one(fp_t *r) { memcpy(r->l, FP_ONE, 48); }
ALWAYS_INLINE void fp_set_zero(fp_t *r) { memset(r->l, 0, 48); }
ALWAYS_INLINE void fp_dbl(fp_t *r, const fp_t *a) { fp_add(r, a, a); }
static void fp_inv(fp_t *r, const fp_t *a) { mont_pow(r->l, a->l, FP_EXP_INV, 6, FP_ONE, FP_MOD, FP_INV, 6); }
static void fp_pow(fp_t *r, const fp_t *a, const uint64_t *e) { mont_pow(r->l, a->l, e, 6, FP_ONE, FP_MOD, FP_INV, 6); }
/* plain integer (little-endian limbs) <-> Montgomery */
ALWAYS_INLINE void fp_from_plain(fp_t *r, const uint64_t *x) { fp_t t; memcpy(t.l, x, 48); fp_t r2; memcpy(r2.l, FP_R2, 48); fp_mul(r, &t, &r2); }
ALWAYS_INLINE void fp_to_plain(uint64_t *x, const fp_t *a) { fp_t one = {{1, 0, 0, 0, 0, 0}}; fp_t t; fp_mul(&t, a, &one); memcpy(x, t.l, 48); }
/* 48-byte big-endian; returns 0 if value >= p */
static int fp_from_be(fp_t *r, const uint8_t *b) {
    uint64_t x[6];
    for (int i = 0; i < 6; i++) { uint64_t v = 0; for (int j = 0; j < 8; j++) v = (v << 8) | b[(5 - i) * 8 + j]; x[i] = v; }
    if (limbs_ge(x, FP_MOD, 6)) return 0;
    fp_from_plain(r, x); return 1;
}
static void fp_to_be(uint8_t *b, const fp_t *a) {
    uint64_t x[6]; fp_to_plain(x, a);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 8; j++) b[(5 - i) * 8 + j] = (uint8_t)(x[i] >> (56 - 8 * j));
}
/* y > (p-1)/2 on the canonical integer */
static int fp_is_lex_largest(const fp_t *a) {
    uint64_t x[6]; fp_to_plain(x, a);
    for (int i = 5; i >= 0; i--) { if (x[i] > FP_HALF[i]) return 1; if (x[i] < FP_HALF[i]) return 0; }
    return 0;
}
/* square root for p = 3 mod 4; returns 0 if a is not a square */
static int fp_sqrt(fp_t *r, const fp_t *a) {
    fp_t s, c; fp_pow(&s, a, FP_EXP_SQRT); fp_sqr(&c, &s);
    if (!fp_eq(&c, a)) return 0;
    *r = s; return 1;
}

/* ---------------------------------------------------------------- Fr */
typedef struct { uint64_t l[4]; } fr_t;
ALWAYS_INLINE void fr_add(fr_t *r, const fr_t *a, const fr_t *b) { mont_add(r->l, a->l, b->l, FR_MOD, 4); }
ALWAYS_INLINE void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) { mont_sub(r->l, a->l, b->l, FR_MOD, 4); }
ALWAYS_INLINE void fr_neg(fr_t *r, const fr_t *a) { mont_neg(r->l, a->l, FR_MOD, 4); }
ALWAYS_INLINE void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { mont_mul(r->l, a->l, b->l, FR_MOD, FR_INV, 4); }
ALWAYS_INLINE int fr_is_zero(const fr_t *a) { return limbs_is_zero(a->l, 4); }
ALWAYS_INLINE int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a->l, b->l, 32) == 0; }
ALWAYS_INLINE void fr_set_one(fr_t *r) { memcpy(r->l, FR_ONE, 32); }
ALWAYS_INLINE void fr_set_zero(fr_t *r) { memset(r->l, 0, 32); }
static void fr_inv(fr_t *r, const fr_t *a) { mont_pow(r->l, a->l, FR_EXP_INV, 4, FR_ONE, FR_MOD, FR_INV, 4); }
static void fr_pow_u64(fr_t *r, const fr_t *a, uint64_t e) { mont_pow(r->l, a->l, &e, 1, FR_ONE, FR_MOD, FR_INV, 4); }
ALWAYS_INLINE void fr_from_plain(fr_t *r, const uint64_t *x) { fr_t t; memcpy(t.l, x, 32); fr_t r2; memcpy(r2.l, FR_R2, 32); fr_mul(r, &t, &r2); }
ALWAYS_INLINE void fr_to_plain(uint64_t *x, const fr_t *a) { fr_t one = {{1, 0, 0, 0}}; fr_t t; fr_mul(&t, a, &one); memcpy(x, t.l, 32); }
ALWAYS_INLINE void fr_from_u64(fr_t *r, uint64_t v) { uint64_t x[4] = {v, 0, 0, 0}; fr_from_plain(r, x); }
static void be32_to_limbs(uint64_t *x, const uint8_t *b) {
    for (int i = 0; i < 4; i++) { uint64_t v = 0; for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j]; x[i] = v; }
}
/* 32-byte big-endian, canonical only (serialization/src/lib.rs:47-63); returns 0 if >= r */
static int fr_from_be(fr_t *r, const uint8_t *b) {
    uint64_t x[4]; be32_to_limbs(x, b);
    if (limbs_ge(x, FR_MOD, 4)) return 0;
    fr_from_plain(r, x); return 1;
}
/* 32-byte big-endian reduced mod r (bls12_381/src/lib.rs:128-140 reduce_bytes_to_scalar_bias) */
static void fr_from_be_reduce(fr_t *r, const uint8_t *b) {
    uint64_t x[4]; be32_to_limbs(x, b);
    /* 2^256 < 3r, so at most two subtractions... r ~ 0.45*2^256: up to 2 */
    while (limbs_ge(x, FR_MOD, 4)) limbs_sub(x, x, FR_MOD, 4);
    fr_from_plain(r, x);
}
static void fr_to_be(uint8_t *b, const fr_t *a) {
    uint64_t x[4]; fr_to_plain(x, a);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(x[i] >> (56 - 8 * j));
}
