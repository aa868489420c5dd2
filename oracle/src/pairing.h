/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Optimal-ate pairing check on BLS12-381, written for obviousness rather than speed:
 * tower Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - (1+u)), Fp12 = Fp6[w]/(w^2 - v); affine G2
 * Miller loop over |x| = 0xd201000000010000 with lines embedded as (c0.a0, c0.a1, c1.a1);
 * final exponentiation as one plain square-and-multiply by (p^12-1)/r.
 * Third-party in the reference: blstrs::Bls12::multi_miller_loop + final_exponentiation
 * (crates/cryptography/bls12_381/src/lib.rs:45-50).  Pinned through the verify_* consensus vectors.
 */
#pragma once
#include "g1.h"

typedef struct { fp_t c0, c1; } fp2_t;
typedef struct { fp2_t a0, a1, a2; } fp6_t;
typedef struct { fp6_t c0, c1; } fp12_t;

static void fp2_add(fp2_t *r, const fp2_t *a, const fp2_t *b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static void fp2_sub(fp2_t *r, const fp2_t *a, const fp2_t *b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static void fp2_neg(fp2_t *r, const fp2_t *a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static void fp2_mul(fp2_t *r, const fp2_t *a, const fp2_t *b) {
    fp_t t0, t1, t2, t3;
    fp_mul(&t0, &a->c0, &b->c0); fp_mul(&t1, &a->c1, &b->c1);
    fp_mul(&t2, &a->c0, &b->c1); fp_mul(&t3, &a->c1, &b->c0);
    fp_sub(&r->c0, &t0, &t1); fp_add(&r->c1, &t2, &t3);
}
static void fp2_sqr(fp2_t *r, const fp2_t *a) { fp2_mul(r, a, a); }
static void fp2_mul_fp(fp2_t *r, const fp2_t *a, const fp_t *b) { fp_mul(&r->c0, &a->c0, b); fp_mul(&r->c1, &a->c1, b); }
static int fp2_is_zero(const fp2_t *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static int fp2_eq(const fp2_t *a, const fp2_t *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static void fp2_set_zero(fp2_t *r) { fp_set_zero(&r->c0); fp_set_zero(&r->c1); }
static void fp2_set_one(fp2_t *r) { fp_set_one(&r->c0); fp_set_zero(&r->c1); }
static void fp2_inv(fp2_t *r, const fp2_t *a) {
    fp_t n, t; fp_sqr(&n, &a->c0); fp_sqr(&t, &a->c1); fp_add(&n, &n, &t); fp_inv(&n, &n);
    fp_mul(&r->c0, &a->c0, &n); fp_mul(&t, &a->c1, &n); fp_neg(&r->c1, &t);
}
/* multiply by xi = 1 + u */
static void fp2_mul_xi(fp2_t *r, const fp2_t *a) {
    fp_t t0, t1; fp_sub(&t0, &a->c0, &a->c1); fp_add(&t1, &a->c0, &a->c1); r->c0 = t0; r->c1 = t1;
}
static void fp2_pow(fp2_t *r, const fp2_t *a, const uint64_t *e, int nlimbs) {
    fp2_t acc, base = *a; fp2_set_one(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        fp2_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fp2_mul(&acc, &acc, &base);
    }
    *r = acc;
}
/* sqrt in Fp2 for p = 3 mod 4 (Adj & Rodriguez-Henriquez, alg. 9); returns 0 if non-residue */
static int fp2_sqrt(fp2_t *r, const fp2_t *a) {
    if (fp2_is_zero(a)) { fp2_set_zero(r); return 1; }
    fp2_t a1, alpha, x0, t, neg1;
    fp2_pow(&a1, a, FP_EXP_P34, 6);
    fp2_sqr(&t, &a1); fp2_mul(&alpha, &t, a);
    fp2_mul(&x0, &a1, a);
    fp2_set_one(&neg1); fp2_neg(&neg1, &neg1);
    fp2_t x;
    if (fp2_eq(&alpha, &neg1)) {
        /* x = u * x0 */
        fp_neg(&x.c0, &x0.c1); x.c1 = x0.c0;
    } else {
        fp2_t b, one; fp2_set_one(&one); fp2_add(&b, &alpha, &one);
        fp2_pow(&b, &b, FP_EXP_P12, 6);
        fp2_mul(&x, &b, &x0);
    }
    fp2_sqr(&t, &x);
    if (!fp2_eq(&t, a)) return 0;
    *r = x; return 1;
}

/* ---- Fp6 ---- */
static void fp6_add(fp6_t *r, const fp6_t *a, const fp6_t *b) { fp2_add(&r->a0, &a->a0, &b->a0); fp2_add(&r->a1, &a->a1, &b->a1); fp2_add(&r->a2, &a->a2, &b->a2); }
static void fp6_sub(fp6_t *r, const fp6_t *a, const fp6_t *b) { fp2_sub(&r->a0, &a->a0, &b->a0); fp2_sub(&r->a1, &a->a1, &b->a1); fp2_sub(&r->a2, &a->a2, &b->a2); }
/* schoolbook: v^3 = xi */
static void fp6_mul(fp6_t *r, const fp6_t *a, const fp6_t *b) {
    fp2_t t00, t01, t02, t10, t11, t12, t20, t21, t22, s, x;
    fp2_mul(&t00, &a->a0, &b->a0); fp2_mul(&t01, &a->a0, &b->a1); fp2_mul(&t02, &a->a0, &b->a2);
    fp2_mul(&t10, &a->a1, &b->a0); fp2_mul(&t11, &a->a1, &b->a1); fp2_mul(&t12, &a->a1, &b->a2);
    fp2_mul(&t20, &a->a2, &b->a0); fp2_mul(&t21, &a->a2, &b->a1); fp2_mul(&t22, &a->a2, &b->a2);
    fp6_t o;
    fp2_add(&s, &t12, &t21); fp2_mul_xi(&x, &s); fp2_add(&o.a0, &t00, &x);
    fp2_mul_xi(&x, &t22); fp2_add(&s, &t01, &t10); fp2_add(&o.a1, &s, &x);
    fp2_add(&s, &t02, &t11); fp2_add(&o.a2, &s, &t20);
    *r = o;
}
/* multiply by v */
static void fp6_mul_v(fp6_t *r, const fp6_t *a) {
    fp6_t o; fp2_mul_xi(&o.a0, &a->a2); o.a1 = a->a0; o.a2 = a->a1; *r = o;
}
static void fp6_set_zero(fp6_t *r) { fp2_set_zero(&r->a0); fp2_set_zero(&r->a1); fp2_set_zero(&r->a2); }

/* ---- Fp12 ---- */
static void fp12_set_one(fp12_t *r) { fp6_set_zero(&r->c0); fp6_set_zero(&r->c1); fp2_set_one(&r->c0.a0); }
static void fp12_mul(fp12_t *r, const fp12_t *a, const fp12_t *b) {
    fp6_t t0, t1, t2, t3, v;
    fp6_mul(&t0, &a->c0, &b->c0); fp6_mul(&t1, &a->c1, &b->c1);
    fp6_mul(&t2, &a->c0, &b->c1); fp6_mul(&t3, &a->c1, &b->c0);
    fp6_mul_v(&v, &t1);
    fp12_t o; fp6_add(&o.c0, &t0, &v); fp6_add(&o.c1, &t2, &t3); *r = o;
}
static int fp12_is_one(const fp12_t *a) {
    fp12_t one; fp12_set_one(&one);
    return memcmp(a, &one, sizeof(fp12_t)) == 0;
}

/* ---- G2 (affine, on the twist y^2 = x^3 + 4(1+u)) ---- */
typedef struct { fp2_t x, y; int inf; } g2a_t;

static void g2a_neg(g2a_t *r, const g2a_t *a) { r->x = a->x; fp2_neg(&r->y, &a->y); r->inf = a->inf; }
static void g2a_add(g2a_t *r, const g2a_t *p, const g2a_t *q) {
    if (p->inf) { *r = *q; return; }
    if (q->inf) { *r = *p; return; }
    fp2_t lam, num, den, t;
    if (fp2_eq(&p->x, &q->x)) {
        fp2_add(&t, &p->y, &q->y);
        if (fp2_is_zero(&t)) { r->inf = 1; fp2_set_zero(&r->x); fp2_set_zero(&r->y); return; }
        fp2_sqr(&num, &p->x); fp2_add(&t, &num, &num); fp2_add(&num, &t, &num);
        fp2_add(&den, &p->y, &p->y);
    } else {
        fp2_sub(&num, &q->y, &p->y); fp2_sub(&den, &q->x, &p->x);
    }
    fp2_inv(&den, &den); fp2_mul(&lam, &num, &den);
    fp2_t x3, y3;
    fp2_sqr(&x3, &lam); fp2_sub(&x3, &x3, &p->x); fp2_sub(&x3, &x3, &q->x);
    fp2_sub(&t, &p->x, &x3); fp2_mul(&y3, &lam, &t); fp2_sub(&y3, &y3, &p->y);
    r->x = x3; r->y = y3; r->inf = 0;
}
static void g2a_mul_fr(g2a_t *r, const g2a_t *p, const fr_t *s) {
    uint64_t k[4]; fr_to_plain(k, s);
    g2a_t acc; acc.inf = 1; fp2_set_zero(&acc.x); fp2_set_zero(&acc.y);
    for (int i = 254; i >= 0; i--) {
        g2a_add(&acc, &acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) g2a_add(&acc, &acc, p);
    }
    *r = acc;
}
/* 96-byte compressed: x.c1 (with flag bits) || x.c0; sign = y lexicographically largest (c1 first) */
static int fp2_is_lex_largest(const fp2_t *y) {
    if (!fp_is_zero(&y->c1)) return fp_is_lex_largest(&y->c1);
    return fp_is_lex_largest(&y->c0);
}
static int g2a_decompress(g2a_t *r, const uint8_t in[96]) {
    int c = in[0] >> 7, inf = (in[0] >> 6) & 1, sign = (in[0] >> 5) & 1;
    if (!c) return 0;
    uint8_t b[48]; memcpy(b, in, 48); b[0] &= 0x1F;
    if (inf) { r->inf = 1; fp2_set_zero(&r->x); fp2_set_zero(&r->y); return 1; }
    fp2_t x, y2, y, bb;
    if (!fp_from_be(&x.c1, b)) return 0;
    if (!fp_from_be(&x.c0, in + 48)) return 0;
    uint64_t f[6] = {4, 0, 0, 0, 0, 0}; fp_from_plain(&bb.c0, f); bb.c1 = bb.c0;
    fp2_sqr(&y2, &x); fp2_mul(&y2, &y2, &x); fp2_add(&y2, &y2, &bb);
    if (!fp2_sqrt(&y, &y2)) return 0;
    if (fp2_is_lex_largest(&y) != sign) fp2_neg(&y, &y);
    r->x = x; r->y = y; r->inf = 0; return 1;
}

/* f <- f * line, line = (l0 + l1 v) + (l2 v) w   [sparse positions c0.a0, c0.a1, c1.a1] */
static void fp12_mul_line(fp12_t *f, const fp2_t *l0, const fp2_t *l1, const fp2_t *l2) {
    fp12_t l; fp6_set_zero(&l.c0); fp6_set_zero(&l.c1);
    l.c0.a0 = *l0; l.c0.a1 = *l1; l.c1.a1 = *l2;
    fp12_mul(f, f, &l);
}

#define BLS_X_ABS 0xd201000000010000ULL

/* product of Miller loops f_{|x|,Q_i}(P_i); pairs with an identity on either side are skipped
 * (as blstrs does).  The conjugation for negative x is omitted: it does not affect "== 1". */
static void miller_loop_multi(fp12_t *out, const g1a_t *ps, const g2a_t *qs, int n) {
    fp12_t f; fp12_set_one(&f);
    g2a_t T[8];
    for (int k = 0; k < n; k++) T[k] = qs[k];
    for (int bit = 62; bit >= 0; bit--) {
        fp12_mul(&f, &f, &f);
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1 && !((BLS_X_ABS >> bit) & 1)) break;
            for (int k = 0; k < n; k++) {
                if (ps[k].inf || qs[k].inf) continue;
                /* slope of tangent at T (pass 0) or chord T,Q (pass 1) */
                fp2_t lam, num, den, t;
                if (pass == 0) {
                    fp2_sqr(&num, &T[k].x); fp2_add(&t, &num, &num); fp2_add(&num, &t, &num);
                    fp2_add(&den, &T[k].y, &T[k].y);
                } else {
                    fp2_sub(&num, &qs[k].y, &T[k].y); fp2_sub(&den, &qs[k].x, &T[k].x);
                }
                fp2_inv(&den, &den); fp2_mul(&lam, &num, &den);
                /* line(P) * w^3 = (lam*xT - yT) - lam*xP * v + yP * v w */
                fp2_t l0, l1, l2;
                fp2_mul(&l0, &lam, &T[k].x); fp2_sub(&l0, &l0, &T[k].y);
                fp2_mul_fp(&l1, &lam, &ps[k].x); fp2_neg(&l1, &l1);
                fp2_set_zero(&l2); l2.c0 = ps[k].y;
                fp12_mul_line(&f, &l0, &l1, &l2);
                /* advance T */
                fp2_t x3, y3; const fp2_t *x2 = pass == 0 ? &T[k].x : &qs[k].x;
                fp2_sqr(&x3, &lam); fp2_sub(&x3, &x3, &T[k].x); fp2_sub(&x3, &x3, x2);
                fp2_sub(&t, &T[k].x, &x3); fp2_mul(&y3, &lam, &t); fp2_sub(&y3, &y3, &T[k].y);
                T[k].x = x3; T[k].y = y3;
            }
        }
    }
    *out = f;
}

static void final_exponentiation(fp12_t *r, const fp12_t *a) {
    fp12_t acc, base = *a; fp12_set_one(&acc);
    int top = FINAL_EXP_LIMBS * 64 - 1;
    while (!((FINAL_EXP[top / 64] >> (top % 64)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        fp12_mul(&acc, &acc, &acc);
        if ((FINAL_EXP[i / 64] >> (i % 64)) & 1) fp12_mul(&acc, &acc, &base);
    }
    *r = acc;
}

/* prod e(P_i, Q_i) == 1 */
static int pairing_check(const g1a_t *ps, const g2a_t *qs, int n) {
    fp12_t f; miller_loop_multi(&f, ps, qs, n); final_exponentiation(&f, &f);
    return fp12_is_one(&f);
}
