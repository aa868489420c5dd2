// Host-side optimal-ate pairing check on BLS12-381 for the verifiers: prod_i e(P_i, Q_i) == 1 with the Q_i
// taken from the three FIXED G2 points of the verification keys, so all G2 work (the line coefficients of the
// Miller loop) is done once at start-up and a check costs two sparse Miller loops + one final exponentiation.
//   reference: multi_pairings (crates/cryptography/bls12_381/src/lib.rs:45-50 -> blstrs Bls12::multi_miller_loop +
//   final_exponentiation) with G2Prepared inputs (kzg_multi_open/src/fk20/verifier.rs:100-106, 251-254;
//   kzg_single_open/src/verifier.rs:33-58).  One pairing check per verify call: latency-bound scalar work that stays on
//   the CPU (SURVEY.md §2.2 "(host) 2-pairing check").
// Tower: Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3-(1+u)), Fp12 = Fp6[w]/(w^2-v).  Fp: 6x64-bit Montgomery.
#include "host_pairing.h"
#include <immintrin.h>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "host_consts.h"

namespace ekzg {
namespace host {

typedef unsigned __int128 u128;

struct Fp { uint64_t v[6]; };

static inline bool fp_is_zero(const Fp& a) { uint64_t x = 0; for (int i = 0; i < 6; i++) x |= a.v[i]; return x == 0; }
static inline bool fp_eq(const Fp& a, const Fp& b) { return memcmp(a.v, b.v, 48) == 0; }
static inline bool ge_p(const uint64_t* t) {
    for (int i = 5; i >= 0; i--) { if (t[i] > FP_P[i]) return true; if (t[i] < FP_P[i]) return false; }
    return true;
}
static inline void sub_p(uint64_t* t) {
    uint64_t br = 0;
    for (int i = 0; i < 6; i++) { u128 d = (u128)t[i] - FP_P[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
// branch-free: the tower arithmetic of one pairing check does ~100 k of these against ~24 k multiplications
static inline void fp_add(Fp& r, const Fp& a, const Fp& b) {
    unsigned long long t[6], s[6];
    unsigned char c = 0, br = 0;
    for (int i = 0; i < 6; i++) c = _addcarry_u64(c, a.v[i], b.v[i], &t[i]);       // a, b < p < 2^381: no carry out
    for (int i = 0; i < 6; i++) br = _subborrow_u64(br, t[i], FP_P[i], &s[i]);
    for (int i = 0; i < 6; i++) r.v[i] = br ? t[i] : s[i];                         // borrow <=> t < p
}
static inline void fp_sub(Fp& r, const Fp& a, const Fp& b) {
    unsigned long long t[6];
    unsigned char br = 0, c = 0;
    for (int i = 0; i < 6; i++) br = _subborrow_u64(br, a.v[i], b.v[i], &t[i]);
    const uint64_t mask = 0 - (uint64_t)br;                                        // a < b: add p back
    for (int i = 0; i < 6; i++) { c = _addcarry_u64(c, t[i], FP_P[i] & mask, &t[i]); r.v[i] = t[i]; }
}
static inline void fp_neg(Fp& r, const Fp& a) {
    Fp z; memset(&z, 0, sizeof z);
    const uint64_t nz = fp_is_zero(a) ? 0 : ~0ull;                                 // -0 = 0, not p
    unsigned long long t[6];
    unsigned char br = 0;
    for (int i = 0; i < 6; i++) { br = _subborrow_u64(br, FP_P[i] & nz, a.v[i], &t[i]); r.v[i] = t[i]; }
}
// CIOS with the two carry chains of a row interleaved; p < 2^381 leaves three spare bits in the top word, so a row never carries
// out of t5 (the "no-carry" variant) and one conditional subtraction finishes
#define EKZG_MAC(lo, hi, a, b, c, d) { const u128 x_ = (u128)(a) * (b) + (c) + (d); lo = (uint64_t)x_; hi = (uint64_t)(x_ >> 64); }
static inline void fp_mul(Fp& r, const Fp& a, const Fp& b) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
#pragma GCC unroll 6
    for (int i = 0; i < 6; i++) {
        const uint64_t bi = b.v[i];
        uint64_t c, c2, lo;
        EKZG_MAC(t0, c, a.v[0], bi, t0, 0);
        const uint64_t m = t0 * FP_M0;
        EKZG_MAC(lo, c2, m, FP_P[0], t0, 0);
        EKZG_MAC(t1, c, a.v[1], bi, t1, c); EKZG_MAC(t0, c2, m, FP_P[1], t1, c2);
        EKZG_MAC(t2, c, a.v[2], bi, t2, c); EKZG_MAC(t1, c2, m, FP_P[2], t2, c2);
        EKZG_MAC(t3, c, a.v[3], bi, t3, c); EKZG_MAC(t2, c2, m, FP_P[3], t3, c2);
        EKZG_MAC(t4, c, a.v[4], bi, t4, c); EKZG_MAC(t3, c2, m, FP_P[4], t4, c2);
        EKZG_MAC(t5, c, a.v[5], bi, t5, c); EKZG_MAC(t4, c2, m, FP_P[5], t5, c2);
        t5 = c + c2;
        (void)lo;
    }
    unsigned long long t[6] = {t0, t1, t2, t3, t4, t5}, s[6];
    unsigned char br = 0;
    for (int i = 0; i < 6; i++) br = _subborrow_u64(br, t[i], FP_P[i], &s[i]);
    for (int i = 0; i < 6; i++) r.v[i] = br ? t[i] : s[i];
}
static inline void fp_sqr(Fp& r, const Fp& a) { fp_mul(r, a, a); }
static Fp fp_from_plain(const uint64_t* x) { Fp a, r2; memcpy(a.v, x, 48); memcpy(r2.v, FP_R2, 48); Fp o; fp_mul(o, a, r2); return o; }
static Fp fp_one() { Fp o; memcpy(o.v, FP_ONE, 48); return o; }
static void fp_inv(Fp& r, const Fp& a) {
    Fp acc = fp_one(), base = a;
    for (int i = 0; i < 384; i++) {
        if ((FP_EXP_INV[i / 64] >> (i % 64)) & 1) fp_mul(acc, acc, base);
        fp_sqr(base, base);
    }
    r = acc;
}

struct Fp2 { Fp c0, c1; };
static inline void f2_add(Fp2& r, const Fp2& a, const Fp2& b) { fp_add(r.c0, a.c0, b.c0); fp_add(r.c1, a.c1, b.c1); }
static inline void f2_sub(Fp2& r, const Fp2& a, const Fp2& b) { fp_sub(r.c0, a.c0, b.c0); fp_sub(r.c1, a.c1, b.c1); }
static inline void f2_neg(Fp2& r, const Fp2& a) { fp_neg(r.c0, a.c0); fp_neg(r.c1, a.c1); }
static inline void f2_mul(Fp2& r, const Fp2& a, const Fp2& b) {  // Karatsuba, 3 Fp mul
    Fp t0, t1, s0, s1, m;
    fp_mul(t0, a.c0, b.c0); fp_mul(t1, a.c1, b.c1);
    fp_add(s0, a.c0, a.c1); fp_add(s1, b.c0, b.c1); fp_mul(m, s0, s1);
    fp_sub(m, m, t0); fp_sub(m, m, t1);
    fp_sub(r.c0, t0, t1); r.c1 = m;
}
static inline void f2_sqr(Fp2& r, const Fp2& a) {  // (a0+a1)(a0-a1), 2 a0 a1
    Fp s, d, m;
    fp_add(s, a.c0, a.c1); fp_sub(d, a.c0, a.c1); fp_mul(m, a.c0, a.c1);
    fp_mul(r.c0, s, d); fp_add(r.c1, m, m);
}
static inline void f2_mul_fp(Fp2& r, const Fp2& a, const Fp& b) { fp_mul(r.c0, a.c0, b); fp_mul(r.c1, a.c1, b); }
static inline void f2_mul_xi(Fp2& r, const Fp2& a) { Fp t0, t1; fp_sub(t0, a.c0, a.c1); fp_add(t1, a.c0, a.c1); r.c0 = t0; r.c1 = t1; }
static inline bool f2_is_zero(const Fp2& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
static void f2_inv(Fp2& r, const Fp2& a) {
    Fp n, t; fp_sqr(n, a.c0); fp_sqr(t, a.c1); fp_add(n, n, t); fp_inv(n, n);
    fp_mul(r.c0, a.c0, n); fp_mul(t, a.c1, n); fp_neg(r.c1, t);
}
static Fp2 f2_zero() { Fp2 z; memset(&z, 0, sizeof z); return z; }

struct Fp6 { Fp2 a0, a1, a2; };
static inline void f6_add(Fp6& r, const Fp6& a, const Fp6& b) { f2_add(r.a0, a.a0, b.a0); f2_add(r.a1, a.a1, b.a1); f2_add(r.a2, a.a2, b.a2); }
static inline void f6_sub(Fp6& r, const Fp6& a, const Fp6& b) { f2_sub(r.a0, a.a0, b.a0); f2_sub(r.a1, a.a1, b.a1); f2_sub(r.a2, a.a2, b.a2); }
static inline void f6_neg(Fp6& r, const Fp6& a) { f2_neg(r.a0, a.a0); f2_neg(r.a1, a.a1); f2_neg(r.a2, a.a2); }
static void f6_mul(Fp6& r, const Fp6& a, const Fp6& b) {  // Karatsuba (6 Fp2 mul), v^3 = xi
    Fp2 v0, v1, v2, t, s1, s2, o0, o1, o2;
    f2_mul(v0, a.a0, b.a0); f2_mul(v1, a.a1, b.a1); f2_mul(v2, a.a2, b.a2);
    // o0 = v0 + xi*((a1+a2)(b1+b2) - v1 - v2)
    f2_add(s1, a.a1, a.a2); f2_add(s2, b.a1, b.a2); f2_mul(t, s1, s2); f2_sub(t, t, v1); f2_sub(t, t, v2); f2_mul_xi(t, t); f2_add(o0, v0, t);
    // o1 = (a0+a1)(b0+b1) - v0 - v1 + xi*v2
    f2_add(s1, a.a0, a.a1); f2_add(s2, b.a0, b.a1); f2_mul(t, s1, s2); f2_sub(t, t, v0); f2_sub(t, t, v1); f2_mul_xi(s1, v2); f2_add(o1, t, s1);
    // o2 = (a0+a2)(b0+b2) - v0 - v2 + v1
    f2_add(s1, a.a0, a.a2); f2_add(s2, b.a0, b.a2); f2_mul(t, s1, s2); f2_sub(t, t, v0); f2_sub(t, t, v2); f2_add(o2, t, v1);
    r.a0 = o0; r.a1 = o1; r.a2 = o2;
}
static inline void f6_mul_v(Fp6& r, const Fp6& a) { Fp6 o; f2_mul_xi(o.a0, a.a2); o.a1 = a.a0; o.a2 = a.a1; r = o; }
static void f6_inv(Fp6& r, const Fp6& a) {
    // standard: c0 = a0^2 - xi a1 a2, c1 = xi a2^2 - a0 a1, c2 = a1^2 - a0 a2, t = a0 c0 + xi(a2 c1 + a1 c2)
    Fp2 c0, c1, c2, t, u;
    f2_sqr(c0, a.a0); f2_mul(t, a.a1, a.a2); f2_mul_xi(t, t); f2_sub(c0, c0, t);
    f2_sqr(c1, a.a2); f2_mul_xi(c1, c1); f2_mul(t, a.a0, a.a1); f2_sub(c1, c1, t);
    f2_sqr(c2, a.a1); f2_mul(t, a.a0, a.a2); f2_sub(c2, c2, t);
    f2_mul(t, a.a2, c1); f2_mul(u, a.a1, c2); f2_add(t, t, u); f2_mul_xi(t, t); f2_mul(u, a.a0, c0); f2_add(t, t, u);
    f2_inv(t, t);
    f2_mul(r.a0, c0, t); f2_mul(r.a1, c1, t); f2_mul(r.a2, c2, t);
}

struct Fp12 { Fp6 c0, c1; };
static Fp12 f12_one() { Fp12 o; memset(&o, 0, sizeof o); o.c0.a0.c0 = fp_one(); return o; }
static void f12_mul(Fp12& r, const Fp12& a, const Fp12& b) {  // Karatsuba (3 Fp6 mul), w^2 = v
    Fp6 t0, t1, s0, s1, m, v;
    f6_mul(t0, a.c0, b.c0); f6_mul(t1, a.c1, b.c1);
    f6_add(s0, a.c0, a.c1); f6_add(s1, b.c0, b.c1); f6_mul(m, s0, s1);
    f6_sub(m, m, t0); f6_sub(m, m, t1);
    f6_mul_v(v, t1);
    f6_add(r.c0, t0, v); r.c1 = m;
}
static void f12_sqr(Fp12& r, const Fp12& a) {  // complex squaring: 2 Fp6 mul
    Fp6 ab, s, t, v;
    f6_mul(ab, a.c0, a.c1);
    f6_add(s, a.c0, a.c1); f6_mul_v(v, a.c1); f6_add(t, a.c0, v);
    f6_mul(s, s, t);               // (c0+c1)(c0+v c1) = c0^2 + v c1^2 + c0c1 + v c0c1
    f6_sub(s, s, ab); f6_mul_v(v, ab); f6_sub(s, s, v);
    r.c0 = s; f6_add(r.c1, ab, ab);
}
// a * (b0 + b1 v) in Fp6: five Fp2 multiplications
static void f6_mul_sparse2(Fp6& r, const Fp6& a, const Fp2& b0, const Fp2& b1) {
    Fp2 v0, v1, t, s, o0, o1, o2;
    f2_mul(v0, a.a0, b0); f2_mul(v1, a.a1, b1);
    f2_mul(t, a.a2, b1); f2_mul_xi(t, t); f2_add(o0, v0, t);
    f2_add(s, a.a0, a.a1); f2_add(t, b0, b1); f2_mul(t, s, t); f2_sub(t, t, v0); f2_sub(o1, t, v1);
    f2_mul(t, a.a2, b0); f2_add(o2, t, v1);
    r.a0 = o0; r.a1 = o1; r.a2 = o2;
}
// f * ((A + B v) + (C v) w) with C in Fp: the value of a Miller-loop line at a G1 point.  36 Fp multiplications instead of the 54
// of a general Fp12 product.
static void f12_mul_line(Fp12& f, const Fp2& A, const Fp2& B, const Fp& C) {
    Fp6 t0, t1, s, m, v;
    f6_mul_sparse2(t0, f.c0, A, B);
    Fp2 x;
    f2_mul_fp(x, f.c1.a2, C); f2_mul_xi(t1.a0, x); f2_mul_fp(t1.a1, f.c1.a0, C); f2_mul_fp(t1.a2, f.c1.a1, C);   // c1 * (C v)
    Fp2 bc = B;
    fp_add(bc.c0, bc.c0, C);
    f6_add(s, f.c0, f.c1);
    f6_mul_sparse2(m, s, A, bc);
    f6_sub(m, m, t0); f6_sub(m, m, t1);
    f6_mul_v(v, t1);
    f6_add(f.c0, t0, v); f.c1 = m;
}
// squaring in the cyclotomic subgroup (Granger-Scott: three Fp4 squarings, 18 Fp multiplications instead of 36).  Only valid after
// the easy part of the final exponentiation.
static inline void fp4_square(Fp2& c0, Fp2& c1, const Fp2& a, const Fp2& b) {
    Fp2 t0, t1, t2;
    f2_sqr(t0, a); f2_sqr(t1, b);
    f2_mul_xi(t2, t1); f2_add(c0, t2, t0);
    f2_add(t2, a, b); f2_sqr(t2, t2); f2_sub(t2, t2, t0); f2_sub(c1, t2, t1);
}
static void f12_cyc_sqr(Fp12& r, const Fp12& f) {
    Fp2 z0 = f.c0.a0, z4 = f.c0.a1, z3 = f.c0.a2, z2 = f.c1.a0, z1 = f.c1.a1, z5 = f.c1.a2;
    Fp2 t0, t1, t2, t3;
    fp4_square(t0, t1, z0, z1);
    f2_sub(z0, t0, z0); f2_add(z0, z0, z0); f2_add(z0, z0, t0);     // 3 t0 - 2 z0
    f2_add(z1, t1, z1); f2_add(z1, z1, z1); f2_add(z1, z1, t1);     // 3 t1 + 2 z1
    fp4_square(t0, t1, z2, z3);
    fp4_square(t2, t3, z4, z5);
    f2_sub(z4, t0, z4); f2_add(z4, z4, z4); f2_add(z4, z4, t0);
    f2_add(z5, t1, z5); f2_add(z5, z5, z5); f2_add(z5, z5, t1);
    f2_mul_xi(t0, t3);
    f2_add(z2, t0, z2); f2_add(z2, z2, z2); f2_add(z2, z2, t0);
    f2_sub(z3, t2, z3); f2_add(z3, z3, z3); f2_add(z3, z3, t2);
    r.c0.a0 = z0; r.c0.a1 = z4; r.c0.a2 = z3; r.c1.a0 = z2; r.c1.a1 = z1; r.c1.a2 = z5;
}
static void f12_conj(Fp12& r, const Fp12& a) { r.c0 = a.c0; f6_neg(r.c1, a.c1); }
static void f12_inv(Fp12& r, const Fp12& a) {
    Fp6 t0, t1;
    f6_mul(t0, a.c0, a.c0); f6_mul(t1, a.c1, a.c1); f6_mul_v(t1, t1); f6_sub(t0, t0, t1);  // c0^2 - v c1^2
    f6_inv(t0, t0);
    f6_mul(r.c0, a.c0, t0); f6_mul(t1, a.c1, t0); f6_neg(r.c1, t1);
}
static bool f12_is_one(const Fp12& a) { Fp12 o = f12_one(); return memcmp(&a, &o, sizeof o) == 0; }

struct G2Aff { Fp2 x, y; };
struct Line { Fp2 lam, c; };  // line through T: value at P=(xP,yP) is (c - lam*xP v) + (yP v) w

struct State {
    Fp gamma[6];                 // xi^(i (p^2-1)/6), in Fp
    Fp2 gamma1[6];               // xi^(i (p-1)/6), in Fp2
    std::vector<Line> lines[6];  // G2Sel index
    bool ok = false;
};
static State g_state;
static std::once_flag g_once;

static void f12_frob2(Fp12& r, const Fp12& a) {
    const Fp* g = g_state.gamma;
    r.c0.a0 = a.c0.a0;
    f2_mul_fp(r.c0.a1, a.c0.a1, g[2]);
    f2_mul_fp(r.c0.a2, a.c0.a2, g[4]);
    f2_mul_fp(r.c1.a0, a.c1.a0, g[1]);
    f2_mul_fp(r.c1.a1, a.c1.a1, g[3]);
    f2_mul_fp(r.c1.a2, a.c1.a2, g[5]);
}

// a^p: the coefficient of w^k (w^6 = xi) is conjugated and multiplied by xi^(k (p-1)/6)
static void f12_frob1(Fp12& r, const Fp12& a) {
    const Fp2* g = g_state.gamma1;
    auto term = [&](Fp2& out, const Fp2& in, int k) {
        Fp2 c = in;
        fp_neg(c.c1, c.c1);
        if (k == 0) out = c; else f2_mul(out, c, g[k]);
    };
    Fp12 o;
    term(o.c0.a0, a.c0.a0, 0);
    term(o.c0.a1, a.c0.a1, 2);
    term(o.c0.a2, a.c0.a2, 4);
    term(o.c1.a0, a.c1.a0, 1);
    term(o.c1.a1, a.c1.a1, 3);
    term(o.c1.a2, a.c1.a2, 5);
    r = o;
}

// a^x for a in the cyclotomic subgroup, x = -0xd201000000010000 (negative: conjugate at the end)
static void f12_exp_x(Fp12& r, const Fp12& a) {
    Fp12 acc = a;
    for (int bit = 62; bit >= 0; bit--) {
        f12_cyc_sqr(acc, acc);
        if ((BLS_X_ABS >> bit) & 1) f12_mul(acc, acc, a);
    }
    f12_conj(r, acc);
}

static void prepare(std::vector<Line>& out, const G2Aff& q) {
    G2Aff t = q;
    out.clear();
    for (int bit = 62; bit >= 0; bit--) {
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1 && !((BLS_X_ABS >> bit) & 1)) break;
            Fp2 num, den, lam, tmp;
            if (pass == 0) {  // tangent at T
                f2_sqr(num, t.x); f2_add(tmp, num, num); f2_add(num, tmp, num);
                f2_add(den, t.y, t.y);
            } else {          // chord through T and Q
                f2_sub(num, q.y, t.y); f2_sub(den, q.x, t.x);
            }
            f2_inv(den, den); f2_mul(lam, num, den);
            Line l;
            l.lam = lam;
            f2_mul(l.c, lam, t.x); f2_sub(l.c, l.c, t.y);
            out.push_back(l);
            const Fp2& x2 = pass == 0 ? t.x : q.x;
            Fp2 x3, y3;
            f2_sqr(x3, lam); f2_sub(x3, x3, t.x); f2_sub(x3, x3, x2);
            f2_sub(tmp, t.x, x3); f2_mul(y3, lam, tmp); f2_sub(y3, y3, t.y);
            t.x = x3; t.y = y3;
        }
    }
}

static G2Aff g2_const(const uint64_t* x0, const uint64_t* x1, const uint64_t* y0, const uint64_t* y1, bool neg) {
    G2Aff q;
    q.x.c0 = fp_from_plain(x0); q.x.c1 = fp_from_plain(x1); q.y.c0 = fp_from_plain(y0); q.y.c1 = fp_from_plain(y1);
    if (neg) f2_neg(q.y, q.y);
    return q;
}

static bool on_twist(const G2Aff& q) {  // y^2 == x^3 + 4(1+u)
    Fp2 l, r, b;
    f2_sqr(l, q.y); f2_sqr(r, q.x); f2_mul(r, r, q.x);
    uint64_t four[6] = {4, 0, 0, 0, 0, 0};
    b.c0 = fp_from_plain(four); b.c1 = b.c0;
    f2_add(r, r, b);
    return fp_eq(l.c0, r.c0) && fp_eq(l.c1, r.c1);
}

static void init_state() {
    State& s = g_state;
    s.gamma[0] = fp_one();
    const uint64_t* gs[5] = {FROB2_GAMMA_1, FROB2_GAMMA_2, FROB2_GAMMA_3, FROB2_GAMMA_4, FROB2_GAMMA_5};
    for (int i = 0; i < 5; i++) memcpy(s.gamma[i + 1].v, gs[i], 48);
    {   // gamma1[1] = xi^((p-1)/6) by square-and-multiply, xi = 1 + u; the exponent is FP_P (p = 1 mod 6) divided by 6
        uint64_t e[6];
        u128 rem = 0;
        for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | FP_P[i]; e[i] = (uint64_t)(cur / 6); rem = cur % 6; }
        Fp2 xi, acc;
        xi.c0 = fp_one(); xi.c1 = fp_one();
        acc.c0 = fp_one(); memset(&acc.c1, 0, sizeof acc.c1);
        for (int i = 383; i >= 0; i--) {
            f2_sqr(acc, acc);
            if ((e[i / 64] >> (i % 64)) & 1) f2_mul(acc, acc, xi);
        }
        s.gamma1[0].c0 = fp_one(); memset(&s.gamma1[0].c1, 0, sizeof(Fp));
        s.gamma1[1] = acc;
        for (int i = 2; i < 6; i++) f2_mul(s.gamma1[i], s.gamma1[i - 1], acc);
    }
    bool ok = true;
    {   // self-check of the Frobenius constants: gamma1[k] * conj(gamma1[k]) == gamma[k]  (xi^(k(p-1)/6 * (p+1)) = xi^(k(p^2-1)/6))
        for (int k = 1; k < 6; k++) {
            Fp2 c = s.gamma1[k], prod;
            fp_neg(c.c1, c.c1);
            f2_mul(prod, s.gamma1[k], c);
            ok = ok && fp_eq(prod.c0, s.gamma[k]) && fp_is_zero(prod.c1);
        }
    }
    for (int neg = 0; neg < 2; neg++) {
        G2Aff gen = g2_const(G2_GEN_X0, G2_GEN_X1, G2_GEN_Y0, G2_GEN_Y1, neg);
        G2Aff tau = g2_const(G2_TAU_X0, G2_TAU_X1, G2_TAU_Y0, G2_TAU_Y1, neg);
        G2Aff t64 = g2_const(G2_TAU64_X0, G2_TAU64_X1, G2_TAU64_Y0, G2_TAU64_Y1, neg);
        ok = ok && on_twist(gen) && on_twist(tau) && on_twist(t64);
        prepare(s.lines[0 + 3 * neg], gen);
        prepare(s.lines[1 + 3 * neg], tau);
        prepare(s.lines[2 + 3 * neg], t64);
    }
    s.ok = ok;
}

namespace {
struct Pt { Fp x, y; const std::vector<Line>* ls; };   // affine, Montgomery form
bool pairing_product_is_one(const std::vector<Pt>& pts);
}

// The line tables of a caller-supplied setup's [1]_2, [tau]_2, [tau^64]_2 and their negations (G2Sel index), built once by
// g2_keys_from_compressed; the embedded ceremony's tables live in g_state.
struct G2Keys { std::vector<Line> lines[6]; };
static inline const std::vector<Line>& lines_of(const G2Keys* keys, G2Sel sel) { return keys ? keys->lines[(int)sel] : g_state.lines[(int)sel]; }

bool pairing_check(const PairingInput* in, int n) {
    std::call_once(g_once, init_state);
    if (!g_state.ok) return false;
    std::vector<Pt> pts;
    for (int i = 0; i < n; i++) {
        if (in[i].g1_is_identity) continue;  // e(O, Q) = 1 (blstrs skips identity pairs as well)
        Pt p;
        p.x = fp_from_plain(in[i].g1_x); p.y = fp_from_plain(in[i].g1_y);
        p.ls = &g_state.lines[(int)in[i].g2];
        pts.push_back(p);
    }
    return pairing_product_is_one(pts);
}

bool pairing_check_jac(const PairingInputJac* in, int n, const G2Keys* keys) {
    std::call_once(g_once, init_state);
    if (!g_state.ok) return false;
    std::vector<Pt> pts;
    for (int i = 0; i < n; i++) {
        if (in[i].g1_is_identity) continue;
        Fp X, Y, Z, zi, zi2, zi3;
        memcpy(X.v, in[i].x, 48); memcpy(Y.v, in[i].y, 48); memcpy(Z.v, in[i].z, 48);
        if (fp_is_zero(Z)) continue;
        fp_inv(zi, Z); fp_sqr(zi2, zi); fp_mul(zi3, zi2, zi);
        Pt p;
        fp_mul(p.x, X, zi2); fp_mul(p.y, Y, zi3);
        p.ls = &lines_of(keys, in[i].g2);
        pts.push_back(p);
    }
    return pairing_product_is_one(pts);
}

namespace {
bool pairing_product_is_one(const std::vector<Pt>& pts) {
    Fp12 f = f12_one();
    size_t idx = 0;
    for (int bit = 62; bit >= 0; bit--) {
        f12_sqr(f, f);
        int steps = ((BLS_X_ABS >> bit) & 1) ? 2 : 1;
        for (int sidx = 0; sidx < steps; sidx++, idx++) {
            for (const Pt& p : pts) {
                const Line& l = (*p.ls)[idx];
                Fp2 b;
                f2_mul_fp(b, l.lam, p.x); f2_neg(b, b);
                f12_mul_line(f, l.c, b, p.y);
            }
        }
    }
    // final exponentiation: easy part (p^6-1)(p^2+1), then 3x the hard part (p^4-p^2+1)/r through
    //   3 (p^4 - p^2 + 1)/r = (x-1)^2 (x+p) (x^2 + p^2 - 1) + 3      (Hayashida-Hayasaka-Teruya; checked in
    // tools/gen_host_constants.py) -- five exponentiations by the 64-bit x instead of a 1268-bit square-and-multiply.
    // gcd(3, r) = 1, so the result is one exactly when the pairing product is.  After the easy part the value lies in
    // the cyclotomic subgroup, where inversion is conjugation.
    Fp12 t, u;
    f12_conj(t, f); f12_inv(u, f); f12_mul(t, t, u);
    f12_frob2(u, t); f12_mul(t, u, t);
    Fp12 a, b, c, e1, e2;
    f12_exp_x(e1, t); f12_conj(e2, t); f12_mul(a, e1, e2);          // t^(x-1)
    f12_exp_x(e1, a); f12_conj(e2, a); f12_mul(a, e1, e2);          // ^(x-1)
    f12_exp_x(e1, a); f12_frob1(e2, a); f12_mul(b, e1, e2);         // ^(x+p)
    f12_exp_x(e1, b); f12_exp_x(e1, e1); f12_frob2(e2, b); f12_mul(c, e1, e2);
    f12_conj(e2, b); f12_mul(c, c, e2);                             // ^(x^2+p^2-1)
    Fp12 acc;
    f12_cyc_sqr(acc, t); f12_mul(acc, acc, t); f12_mul(acc, acc, c);    // * t^3
    return f12_is_one(acc);
}
}  // namespace

// The shortcuts above against the general routines, on values from a real Miller loop: the line product against a full Fp12
// product with the sparse element written out, the cyclotomic squaring against the general one after the easy part.
bool pairing_selftest() {
    std::call_once(g_once, init_state);
    if (!g_state.ok) return false;
    Fp12 f = f12_one();
    Fp px = fp_from_plain(FP_P), py = fp_one();   // any field elements will do (px = p mod p = 0 is avoided below)
    fp_add(px, py, py); fp_add(py, px, py);       // px = 2, py = 3 (Montgomery form)
    const std::vector<Line>& ls = g_state.lines[1];
    for (size_t i = 0; i < 24 && i < ls.size(); i++) {
        Fp2 b;
        f2_mul_fp(b, ls[i].lam, px); f2_neg(b, b);
        Fp12 lf, want;
        memset(&lf, 0, sizeof lf);
        lf.c0.a0 = ls[i].c; lf.c0.a1 = b; lf.c1.a1.c0 = py;
        f12_sqr(f, f);
        f12_mul(want, f, lf);
        f12_mul_line(f, ls[i].c, b, py);
        if (memcmp(&f, &want, sizeof f) != 0) return false;
    }
    Fp12 t, u;
    f12_conj(t, f); f12_inv(u, f); f12_mul(t, t, u);
    f12_frob2(u, t); f12_mul(t, u, t);            // now in the cyclotomic subgroup
    for (int i = 0; i < 8; i++) {
        Fp12 a, b;
        f12_sqr(a, t); f12_cyc_sqr(b, t);
        if (memcmp(&a, &b, sizeof a) != 0) return false;
        f12_mul(t, a, f12_is_one(a) ? a : t);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Caller-supplied trusted setup: the 65 G2 points of `g2_monomial` in the ZCash compressed form (96 bytes: x.c1 || x.c0 big-endian,
// flag bits compressed / infinity / y-is-the-larger-root in the top three bits), as deserialize_g2_points hands them to blstrs'
// G2Affine::from_compressed[_unchecked] (crates/serialization/src/trusted_setup.rs, crates/trusted_setup/src/lib.rs:40-61):
// field elements canonical, point on the twist, and -- with the subgroup check on -- [r]Q = O for every one of them.
static bool fp_from_be48(Fp& out, const uint8_t* b, bool mask_flags) {
    uint64_t t[6];
    for (int i = 0; i < 6; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | b[(5 - i) * 8 + k];
        t[i] = w;
    }
    if (mask_flags) t[5] &= 0x1fffffffffffffffULL;
    if (ge_p(t)) return false;
    out = fp_from_plain(t);
    return true;
}
static void fp_to_plain(uint64_t* out, const Fp& a) {
    Fp one_plain; memset(&one_plain, 0, sizeof one_plain); one_plain.v[0] = 1;
    Fp o; fp_mul(o, a, one_plain);
    memcpy(out, o.v, 48);
}
// a > (p-1)/2 as a plain integer  <=>  2a > p - 1  <=>  2a >= p + 1 (p odd)  <=>  2a > p
static bool fp_is_larger_half(const Fp& a) {
    uint64_t t[7], c = 0;
    uint64_t pl[6]; fp_to_plain(pl, a);
    for (int i = 0; i < 6; i++) { t[i] = (pl[i] << 1) | c; c = pl[i] >> 63; }
    t[6] = c;
    if (t[6]) return true;
    for (int i = 5; i >= 0; i--) { if (t[i] > FP_P[i]) return true; if (t[i] < FP_P[i]) return false; }
    return false;
}
// square root in Fp (p = 3 mod 4): a^((p+1)/4), checked
static bool fp_sqrt(Fp& r, const Fp& a) {
    uint64_t e[6], c = 1;
    for (int i = 0; i < 6; i++) { const uint64_t s = FP_P[i] + c; c = (s < c) ? 1 : 0; e[i] = s; }   // p + 1 (no overflow: p < 2^381)
    for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i < 5 ? e[i + 1] << 62 : 0);
    Fp acc = fp_one();
    for (int i = 383; i >= 0; i--) {
        fp_sqr(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(acc, acc, a);
    }
    Fp chk; fp_sqr(chk, acc);
    r = acc;
    return fp_eq(chk, a);
}
// square root in Fp2 = Fp[u]/(u^2+1): x0^2 = (a0 +- sqrt(a0^2 + a1^2))/2, x1 = a1 / (2 x0); checked by squaring
static bool f2_sqrt(Fp2& r, const Fp2& a) {
    Fp2 x = f2_zero();
    if (fp_is_zero(a.c1)) {
        Fp s;
        if (fp_sqrt(s, a.c0)) { x.c0 = s; }
        else { Fp na; fp_neg(na, a.c0); if (!fp_sqrt(s, na)) return false; x.c1 = s; }
    } else {
        Fp n, t, s, d, two_inv, x0;
        fp_sqr(n, a.c0); fp_sqr(t, a.c1); fp_add(n, n, t);
        if (!fp_sqrt(s, n)) return false;                       // the norm of a square is a square
        Fp two = fp_one(); fp_add(two, two, two); fp_inv(two_inv, two);
        fp_add(d, a.c0, s); fp_mul(d, d, two_inv);
        if (!fp_sqrt(x0, d)) { fp_sub(d, a.c0, s); fp_mul(d, d, two_inv); if (!fp_sqrt(x0, d)) return false; }
        Fp den; fp_add(den, x0, x0); fp_inv(den, den);
        x.c0 = x0; fp_mul(x.c1, a.c1, den);
    }
    Fp2 chk; f2_sqr(chk, x);
    if (!fp_eq(chk.c0, a.c0) || !fp_eq(chk.c1, a.c1)) return false;
    r = x;
    return true;
}

// 0 ok, 1 malformed encoding, 2 not on the curve; *inf set for the point at infinity
static int g2_decompress(G2Aff& q, bool* inf, const uint8_t* b) {
    *inf = false;
    if (!(b[0] & 0x80)) return 1;                               // the setup file holds compressed points only
    if (b[0] & 0x40) {
        if (b[0] & 0x3f) return 1;
        for (int i = 1; i < 96; i++) if (b[i]) return 1;
        *inf = true;
        return 0;
    }
    if (!fp_from_be48(q.x.c1, b, true) || !fp_from_be48(q.x.c0, b + 48, false)) return 1;
    Fp2 rhs, bb;
    f2_sqr(rhs, q.x); f2_mul(rhs, rhs, q.x);
    uint64_t four[6] = {4, 0, 0, 0, 0, 0};
    bb.c0 = fp_from_plain(four); bb.c1 = bb.c0;
    f2_add(rhs, rhs, bb);
    if (!f2_sqrt(q.y, rhs)) return 2;
    const bool larger = fp_is_zero(q.y.c1) ? fp_is_larger_half(q.y.c0) : fp_is_larger_half(q.y.c1);
    if (larger != ((b[0] & 0x20) != 0)) f2_neg(q.y, q.y);
    return 0;
}

// [r]Q == O, by a plain double-and-add in Jacobian coordinates over Fp2 (a = 0)
struct G2Jac { Fp2 x, y, z; };
static void g2_dbl(G2Jac& r, const G2Jac& p) {
    Fp2 a, b, c, d, e, f, t;
    f2_sqr(a, p.x); f2_sqr(b, p.y); f2_sqr(c, b);
    f2_add(d, p.x, b); f2_sqr(d, d); f2_sub(d, d, a); f2_sub(d, d, c); f2_add(d, d, d);
    f2_add(e, a, a); f2_add(e, e, a);
    f2_sqr(f, e);
    G2Jac o;
    f2_mul(o.z, p.y, p.z); f2_add(o.z, o.z, o.z);
    f2_sub(o.x, f, d); f2_sub(o.x, o.x, d);
    f2_sub(t, d, o.x); f2_mul(o.y, e, t);
    f2_add(c, c, c); f2_add(c, c, c); f2_add(c, c, c);
    f2_sub(o.y, o.y, c);
    r = o;
}
static void g2_add_affine(G2Jac& r, const G2Jac& p, const G2Aff& q) {
    if (f2_is_zero(p.z)) { r.x = q.x; r.y = q.y; r.z = f2_zero(); r.z.c0 = fp_one(); return; }
    Fp2 z2, u2, s2, h, rr, h2, h3, v, t;
    f2_sqr(z2, p.z); f2_mul(u2, q.x, z2); f2_mul(s2, q.y, z2); f2_mul(s2, s2, p.z);
    f2_sub(h, u2, p.x); f2_sub(rr, s2, p.y);
    if (f2_is_zero(h)) {
        if (f2_is_zero(rr)) { g2_dbl(r, p); return; }
        r.x = f2_zero(); r.y = f2_zero(); r.y.c0 = fp_one(); r.z = f2_zero();   // P + (-P)
        return;
    }
    f2_sqr(h2, h); f2_mul(h3, h2, h); f2_mul(v, p.x, h2);
    G2Jac o;
    f2_sqr(o.x, rr); f2_sub(o.x, o.x, h3); f2_sub(o.x, o.x, v); f2_sub(o.x, o.x, v);
    f2_sub(t, v, o.x); f2_mul(o.y, rr, t); f2_mul(t, p.y, h3); f2_sub(o.y, o.y, t);
    f2_mul(o.z, p.z, h);
    r = o;
}
static bool g2_in_subgroup(const G2Aff& q) {
    static const uint64_t R[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    G2Jac acc; acc.x = f2_zero(); acc.y = f2_zero(); acc.y.c0 = fp_one(); acc.z = f2_zero();
    for (int i = 254; i >= 0; i--) {
        if (!f2_is_zero(acc.z)) g2_dbl(acc, acc);
        if ((R[i / 64] >> (i % 64)) & 1) g2_add_affine(acc, acc, q);
    }
    return f2_is_zero(acc.z);
}

G2Keys* g2_keys_from_compressed(const uint8_t* g2, int count, bool subgroup_check, std::string* err) {
    std::call_once(g_once, init_state);
    if (!g_state.ok) { *err = "host pairing self-check failed"; return nullptr; }
    if (count != 65) { *err = "trusted setup: g2_monomial must hold 65 points, got " + std::to_string(count); return nullptr; }
    G2Aff used[3];
    for (int i = 0; i < count; i++) {
        G2Aff q; bool inf = false;
        const int rc = g2_decompress(q, &inf, g2 + (size_t)96 * i);
        if (rc) { *err = "trusted setup: g2_monomial[" + std::to_string(i) + (rc == 1 ? "] is not a valid compressed G2 encoding" : "] is not on the curve"); return nullptr; }
        if (!inf && subgroup_check && !g2_in_subgroup(q)) { *err = "trusted setup: g2_monomial[" + std::to_string(i) + "] is outside the prime-order subgroup"; return nullptr; }
        const int slot = i == 0 ? 0 : i == 1 ? 1 : i == 64 ? 2 : -1;
        if (slot >= 0) {
            if (inf) { *err = "trusted setup: g2_monomial[" + std::to_string(i) + "] is the point at infinity"; return nullptr; }
            used[slot] = q;
        }
    }
    G2Keys* k = new G2Keys();
    for (int neg = 0; neg < 2; neg++)
        for (int sidx = 0; sidx < 3; sidx++) {
            G2Aff q = used[sidx];
            if (neg) f2_neg(q.y, q.y);
            prepare(k->lines[sidx + 3 * neg], q);
        }
    return k;
}
void g2_keys_free(G2Keys* k) { delete k; }

// test hook: decompress one G2 point; out = x.c0, x.c1, y.c0, y.c1 as plain little-endian 64-bit limbs.  0 ok, 1 malformed,
// 2 off the curve, 3 outside the subgroup, 4 infinity
int g2_decompress_plain(const uint8_t* in96, uint64_t* out24) {
    std::call_once(g_once, init_state);
    G2Aff q; bool inf = false;
    const int rc = g2_decompress(q, &inf, in96);
    if (rc) return rc;
    if (inf) return 4;
    fp_to_plain(out24, q.x.c0); fp_to_plain(out24 + 6, q.x.c1); fp_to_plain(out24 + 12, q.y.c0); fp_to_plain(out24 + 18, q.y.c1);
    return g2_in_subgroup(q) ? 0 : 3;
}

}  // namespace host
}  // namespace ekzg
