/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * G1 arithmetic for BLS12-381 (y^2 = x^3 + 4) in Jacobian coordinates, ZCash-format point
 * (de)compression and the subgroup check.  Third-party in the reference (blst through blstrs:
 * G1Projective / G1Affine operator overloads, from_compressed / to_compressed at
 * crates/serialization/src/lib.rs:69-99,132-156); restated here from the standard formulas
 * (EFD dbl-2009-l, add-2007-bl, madd-2007-bl).
 */
#pragma once
#include "field.h"

typedef struct { fp_t x, y; int inf; } g1a_t;  /* affine */
typedef struct { fp_t x, y, z; } g1_t;         /* Jacobian, z == 0 <=> identity */

static void g1_set_inf(g1_t *r) { fp_set_one(&r->x); fp_set_one(&r->y); fp_set_zero(&r->z); }
static int g1_is_inf(const g1_t *a) { return fp_is_zero(&a->z); }
static void g1_from_affine(g1_t *r, const g1a_t *a) {
    if (a->inf) { g1_set_inf(r); return; }
    r->x = a->x; r->y = a->y; fp_set_one(&r->z);
}
static void g1_neg(g1_t *r, const g1_t *a) { r->x = a->x; fp_neg(&r->y, &a->y); r->z = a->z; }
static void g1a_neg(g1a_t *r, const g1a_t *a) { r->x = a->x; fp_neg(&r->y, &a->y); r->inf = a->inf; }

static void g1_dbl(g1_t *r, const g1_t *p) {
    if (g1_is_inf(p)) { g1_set_inf(r); return; }
    fp_t A, B, C, D, E, F, t;
    fp_sqr(&A, &p->x); fp_sqr(&B, &p->y); fp_sqr(&C, &B);
    fp_add(&t, &p->x, &B); fp_sqr(&t, &t); fp_sub(&t, &t, &A); fp_sub(&t, &t, &C); fp_dbl(&D, &t);
    fp_dbl(&E, &A); fp_add(&E, &E, &A);
    fp_sqr(&F, &E);
    fp_t z3; fp_mul(&z3, &p->y, &p->z); fp_dbl(&z3, &z3);
    fp_t x3; fp_sub(&x3, &F, &D); fp_sub(&x3, &x3, &D);
    fp_t y3; fp_sub(&t, &D, &x3); fp_mul(&y3, &E, &t);
    fp_dbl(&C, &C); fp_dbl(&C, &C); fp_dbl(&C, &C); fp_sub(&y3, &y3, &C);
    r->x = x3; r->y = y3; r->z = z3;
}

static void g1_add(g1_t *r, const g1_t *p, const g1_t *q) {
    if (g1_is_inf(p)) { *r = *q; return; }
    if (g1_is_inf(q)) { *r = *p; return; }
    fp_t z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
    fp_sqr(&z1z1, &p->z); fp_sqr(&z2z2, &q->z);
    fp_mul(&u1, &p->x, &z2z2); fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&t, &q->z, &z2z2); fp_mul(&s1, &p->y, &t);
    fp_mul(&t, &p->z, &z1z1); fp_mul(&s2, &q->y, &t);
    fp_sub(&h, &u2, &u1); fp_sub(&rr, &s2, &s1);
    if (fp_is_zero(&h)) {
        if (fp_is_zero(&rr)) { g1_dbl(r, p); return; }
        g1_set_inf(r); return;
    }
    fp_t hh, hhh, v;
    fp_sqr(&hh, &h); fp_mul(&hhh, &h, &hh); fp_mul(&v, &u1, &hh);
    fp_t x3, y3, z3;
    fp_sqr(&x3, &rr); fp_sub(&x3, &x3, &hhh); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
    fp_sub(&t, &v, &x3); fp_mul(&y3, &rr, &t); fp_mul(&t, &s1, &hhh); fp_sub(&y3, &y3, &t);
    fp_mul(&z3, &p->z, &q->z); fp_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

static void g1_add_affine(g1_t *r, const g1_t *p, const g1a_t *q) {
    if (q->inf) { *r = *p; return; }
    if (g1_is_inf(p)) { g1_from_affine(r, q); return; }
    fp_t z1z1, u2, s2, h, rr, t;
    fp_sqr(&z1z1, &p->z);
    fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&t, &p->z, &z1z1); fp_mul(&s2, &q->y, &t);
    fp_sub(&h, &u2, &p->x); fp_sub(&rr, &s2, &p->y);
    if (fp_is_zero(&h)) {
        if (fp_is_zero(&rr)) { g1_dbl(r, p); return; }
        g1_set_inf(r); return;
    }
    fp_t hh, hhh, v, x3, y3, z3;
    fp_sqr(&hh, &h); fp_mul(&hhh, &h, &hh); fp_mul(&v, &p->x, &hh);
    fp_sqr(&x3, &rr); fp_sub(&x3, &x3, &hhh); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
    fp_sub(&t, &v, &x3); fp_mul(&y3, &rr, &t); fp_mul(&t, &p->y, &hhh); fp_sub(&y3, &y3, &t);
    fp_mul(&z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

static void g1_sub(g1_t *r, const g1_t *p, const g1_t *q) { g1_t n; g1_neg(&n, q); g1_add(r, p, &n); }

static void g1_to_affine(g1a_t *r, const g1_t *p) {
    if (g1_is_inf(p)) { r->inf = 1; fp_set_zero(&r->x); fp_set_zero(&r->y); return; }
    fp_t zi, zi2, zi3;
    fp_inv(&zi, &p->z); fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
    fp_mul(&r->x, &p->x, &zi2); fp_mul(&r->y, &p->y, &zi3); r->inf = 0;
}

/* shared-inversion normalisation, identity safe (bls12_381/src/lib.rs:56-104) */
static void g1_batch_to_affine(g1a_t *out, const g1_t *in, int n) {
    fp_t *pre = (fp_t *)__builtin_alloca(sizeof(fp_t) * (n + 1));
    fp_t acc; fp_set_one(&acc);
    for (int i = 0; i < n; i++) { pre[i] = acc; if (!g1_is_inf(&in[i])) fp_mul(&acc, &acc, &in[i].z); }
    fp_t inv; fp_inv(&inv, &acc);
    for (int i = n - 1; i >= 0; i--) {
        if (g1_is_inf(&in[i])) { out[i].inf = 1; fp_set_zero(&out[i].x); fp_set_zero(&out[i].y); continue; }
        fp_t zi, zi2, zi3; fp_mul(&zi, &inv, &pre[i]); fp_mul(&inv, &inv, &in[i].z);
        fp_sqr(&zi2, &zi); fp_mul(&zi3, &zi2, &zi);
        fp_mul(&out[i].x, &in[i].x, &zi2); fp_mul(&out[i].y, &in[i].y, &zi3); out[i].inf = 0;
    }
}

/* k: plain little-endian limbs, nbits significant; left-to-right double-and-add */
static void g1_mul_limbs(g1_t *r, const g1_t *p, const uint64_t *k, int nbits) {
    g1_t acc; g1_set_inf(&acc);
    for (int i = nbits - 1; i >= 0; i--) {
        g1_dbl(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) g1_add(&acc, &acc, p);
    }
    *r = acc;
}
static void g1_mul_fr(g1_t *r, const g1_t *p, const fr_t *s) {
    uint64_t k[4]; fr_to_plain(k, s); g1_mul_limbs(r, p, k, 255);
}

static int g1a_on_curve(const g1a_t *a) {
    if (a->inf) return 1;
    fp_t l, rr, four; fp_sqr(&l, &a->y); fp_sqr(&rr, &a->x); fp_mul(&rr, &rr, &a->x);
    uint64_t f[6] = {4, 0, 0, 0, 0, 0}; fp_from_plain(&four, f); fp_add(&rr, &rr, &four);
    return fp_eq(&l, &rr);
}
/* [r]P == O */
static int g1a_in_subgroup(const g1a_t *a) {
    if (a->inf) return 1;
    g1_t p, q; g1_from_affine(&p, a); g1_mul_limbs(&q, &p, FR_MOD, 255);
    return g1_is_inf(&q);
}

/* ZCash compressed encoding: bit7 = compressed, bit6 = infinity, bit5 = y lexicographically largest */
static void g1a_compress(uint8_t out[48], const g1a_t *a) {
    if (a->inf) { memset(out, 0, 48); out[0] = 0xC0; return; }
    fp_to_be(out, &a->x); out[0] |= 0x80;
    if (fp_is_lex_largest(&a->y)) out[0] |= 0x20;
}
static void g1_compress(uint8_t out[48], const g1_t *p) { g1a_t a; g1_to_affine(&a, p); g1a_compress(out, &a); }

/* returns 1 on success, 0 on any encoding / curve / subgroup failure */
static int g1a_decompress(g1a_t *r, const uint8_t in[48], int check_subgroup) {
    int c = in[0] >> 7, inf = (in[0] >> 6) & 1, sign = (in[0] >> 5) & 1;
    if (!c) return 0;
    uint8_t b[48]; memcpy(b, in, 48); b[0] &= 0x1F;
    if (inf) {
        if (sign) return 0;
        for (int i = 0; i < 48; i++) if (b[i]) return 0;
        r->inf = 1; fp_set_zero(&r->x); fp_set_zero(&r->y); return 1;
    }
    fp_t x, y2, y, four;
    if (!fp_from_be(&x, b)) return 0;
    uint64_t f[6] = {4, 0, 0, 0, 0, 0}; fp_from_plain(&four, f);
    fp_sqr(&y2, &x); fp_mul(&y2, &y2, &x); fp_add(&y2, &y2, &four);
    if (!fp_sqrt(&y, &y2)) return 0;
    if (fp_is_lex_largest(&y) != sign) fp_neg(&y, &y);
    r->x = x; r->y = y; r->inf = 0;
    if (check_subgroup && !g1a_in_subgroup(r)) return 0;
    return 1;
}
