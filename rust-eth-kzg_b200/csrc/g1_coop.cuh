// Warp-cooperative Fp / G1 arithmetic for K5's latency mode: ONE field element on FOUR lanes (3 x 32-bit limbs each), eight points
// per warp.  A dependent product costs 965 clocks in this layout against 1823 for the one-element-per-thread multiplier
// (tools/coop_probe.cu, profiles/r2_coop_probe.md), and a small batch is bound by exactly that latency: 7 fixed-scalar
// multiplications of ~1500 dependent products.  Throughput is 40 % lower, so only batches that leave most of the machine idle run here.
//
//   * Values are kept in [0, 2p) ("semi-reduced").  The Montgomery product of two such values is again below 2p (R = 2^384 > 4p), so
//     the multiplier never compares across lanes; additions and subtractions bring their result back below 2p with one conditional
//     subtraction / addition of 2p.  Conversion back to one-thread form subtracts p once more if needed: bit-identical results.
//   * Carries and borrows between the lanes of a group are resolved carry-lookahead style: every lane votes "generate" and
//     "propagate" (__ballot_sync), the carries INTO the lanes are (X + G) ^ X ^ G with X = G | P on the group's four bits.
//   * The fixed-scalar ladder below is straight-line: no exceptional-case branches.  tests/test_k5_coop_ops.py proves on integers
//     that none of the 126 op lists it runs ever adds equal or opposite points or doubles the identity (prime-order input).
//   reference: the `*b * twiddle` of the G1 butterfly, polynomial/src/fft.rs:164-177.
#pragma once
#include "g1_mul.cuh"

#ifdef __CUDACC__
namespace ekzg {
namespace coop {

constexpr unsigned FULL = 0xffffffffu;

struct CFp { uint32_t v[3]; };
struct CJac { CFp x, y, z; };

// 2p, limb i
__device__ constexpr uint32_t mod2(int i) { return (FpParams::mod(i) << 1) | (i ? FpParams::mod(i - 1) >> 31 : 0u); }

struct Ctx {
    unsigned gl;      // lane within the group: limbs 3 gl .. 3 gl + 2
    unsigned shift;   // first lane of the group within the warp
    CFp p, p2;        // this lane's limbs of p and 2p
};

#define EKZG_COOP_PICK(f, k) (gl == 0 ? f(k) : gl == 1 ? f(3 + (k)) : gl == 2 ? f(6 + (k)) : f(9 + (k)))
__device__ __forceinline__ Ctx make_ctx() {
    Ctx c;
    const unsigned lane = threadIdx.x & 31u, gl = lane & 3u;
    c.gl = gl;
    c.shift = lane & 28u;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        c.p.v[k] = EKZG_COOP_PICK(FpParams::mod, k);
        c.p2.v[k] = EKZG_COOP_PICK(mod2, k);
    }
    return c;
}
__device__ __forceinline__ CFp const_fp(const Ctx& c, uint32_t (*f)(int)) {
    CFp r;
    const unsigned gl = c.gl;
#pragma unroll
    for (int k = 0; k < 3; k++) r.v[k] = EKZG_COOP_PICK(f, k);
    return r;
}
__device__ __forceinline__ uint32_t one_limb(int i) { return FpParams::one(i); }
__device__ __forceinline__ uint32_t beta_limb(int i) { return FpParams::beta(i); }
// 3/2 mod p in Montgomery form (3 * 2^-1 * 2^384 mod p): the doubling below multiplies by it instead of tripling and halving
__device__ __forceinline__ uint32_t three_halves_limb(int i) {
    switch (i) { case 0: return 0x0004aaa6u; case 1: return 0xd40e0000u; case 2: return 0x4d680003u; case 3: return 0x52980012u; case 4: return 0x82528a06u; case 5: return 0x5b547b32u; case 6: return 0xaeb8f988u; case 7: return 0x8179debau; case 8: return 0x51dc8c38u; case 9: return 0xe47cd408u; case 10: return 0xdb01638fu; case 11: return 0x13f10530u; default: return 0u; }
}

// this lane's three limbs of a one-thread element (every lane of the group holds the same `a`)
__device__ __forceinline__ CFp from_fp(const Ctx& c, const Fp& a) {
    CFp r;
    const unsigned gl = c.gl;
#pragma unroll
    for (int k = 0; k < 3; k++) r.v[k] = gl == 0 ? a.v[k] : gl == 1 ? a.v[3 + k] : gl == 2 ? a.v[6 + k] : a.v[9 + k];
    return r;
}
// the whole element, canonical (< p), in every lane of the group
__device__ __forceinline__ Fp to_fp(const Ctx&, const CFp& a) {
    Fp r;
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
        for (int k = 0; k < 3; k++) r.v[3 * l + k] = __shfl_sync(FULL, a.v[k], l, 4);
    fe_final_sub<FpParams>(r.v);   // [0, 2p) -> [0, p)
    return r;
}

// carry (or borrow) INTO this lane from the lower lanes of its group; *out = what leaves the group's top lane
__device__ __forceinline__ uint32_t resolve(const Ctx& c, bool generate, bool propagate, uint32_t* out) {
    const uint32_t G = (__ballot_sync(FULL, generate) >> c.shift) & 15u, P = (__ballot_sync(FULL, propagate) >> c.shift) & 15u;
    const uint32_t X = G | P, C = (X + G) ^ X ^ G;
    *out = (C >> 4) & 1u;
    return (C >> c.gl) & 1u;
}

// r = a + b over the whole group (no reduction); returns the carry out of the top lane
__device__ __forceinline__ uint32_t add_raw(const Ctx& c, CFp& r, const CFp& a, const CFp& b) {
    uint32_t s0, s1, s2, g;
    asm("add.cc.u32 %0, %4, %7;\n\t addc.cc.u32 %1, %5, %8;\n\t addc.cc.u32 %2, %6, %9;\n\t addc.u32 %3, 0, 0;"
        : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(g) : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]));
    uint32_t out;
    const uint32_t cin = resolve(c, g != 0, (s0 & s1 & s2) == 0xffffffffu, &out);
    asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;" : "+r"(s0), "+r"(s1), "+r"(s2) : "r"(cin));
    r.v[0] = s0; r.v[1] = s1; r.v[2] = s2;
    return out;
}
// r = a - b over the whole group; returns the borrow out of the top lane (1 <=> a < b)
__device__ __forceinline__ uint32_t sub_raw(const Ctx& c, CFp& r, const CFp& a, const CFp& b) {
    uint32_t d0, d1, d2, g;
    asm("sub.cc.u32 %0, %4, %7;\n\t subc.cc.u32 %1, %5, %8;\n\t subc.cc.u32 %2, %6, %9;\n\t subc.u32 %3, 0, 0;"
        : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(g) : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]));
    uint32_t out;
    const uint32_t bin = resolve(c, g != 0, (d0 | d1 | d2) == 0u, &out);
    asm("sub.cc.u32 %0, %0, %3;\n\t subc.cc.u32 %1, %1, 0;\n\t subc.u32 %2, %2, 0;" : "+r"(d0), "+r"(d1), "+r"(d2) : "r"(bin));
    r.v[0] = d0; r.v[1] = d1; r.v[2] = d2;
    return out;
}

// a, b in [0, 2p) -> a + b in [0, 2p)
__device__ __forceinline__ void cadd(const Ctx& c, CFp& r, const CFp& a, const CFp& b) {
    CFp s, d;
    add_raw(c, s, a, b);                       // < 4p < 2^384: nothing leaves the group
    const uint32_t below = sub_raw(c, d, s, c.p2);
#pragma unroll
    for (int k = 0; k < 3; k++) r.v[k] = below ? s.v[k] : d.v[k];
}
// a, b in [0, 2p) -> a - b in [0, 2p)
__device__ __forceinline__ void csub(const Ctx& c, CFp& r, const CFp& a, const CFp& b) {
    CFp d, q;
    const uint32_t below = sub_raw(c, d, a, b);
#pragma unroll
    for (int k = 0; k < 3; k++) q.v[k] = below ? c.p2.v[k] : 0u;
    add_raw(c, r, d, q);                       // the carry out of the group cancels the borrow
}
__device__ __forceinline__ void cdbl(const Ctx& c, CFp& r, const CFp& a) { cadd(c, r, a, a); }
__device__ __forceinline__ void cneg(const Ctx& c, CFp& r, const CFp& a) {
    CFp z;
    z.v[0] = z.v[1] = z.v[2] = 0u;
    csub(c, r, z, a);
}

// (t0, t1, t2, h0, h1) += a(3 limbs) * s: two carry chains, every mad.lo.cc / madc.hi.cc pair one IMAD.WIDE.U32.X
__device__ __forceinline__ void mad3(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& h0, uint32_t& h1, const CFp& a, uint32_t s) {
    asm("mad.lo.cc.u32 %0, %5, %8, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %8, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %8, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %6, %8, %1;\n\t"
        "madc.hi.cc.u32 %2, %6, %8, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(h0), "+r"(h1)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(s));
}

// a, b in [0, 2p) -> a b / 2^384 mod p in [0, 2p).  Operand scanning; per limb of b: broadcast b_i, t += a_lane b_i, broadcast lane 0's
// low word, t += p_lane m, shift one limb down pulling the next lane's low word in; the carries that leave a lane's three limbs wait
// in (h0, h1) and are folded into the next lane once at the end.  One copy per kernel image.
static __device__ __noinline__ CFp cmul(CFp a, CFp b, CFp p, unsigned gl) {
    uint32_t t0 = 0, t1 = 0, t2 = 0, h0 = 0, h1 = 0, ypend = 0;
    uint32_t bb[12];
#pragma unroll
    for (int i = 0; i < 12; i++) bb[i] = __shfl_sync(FULL, b.v[i % 3], i / 3, 4);   // all broadcasts of b up front: off the critical path
#pragma unroll
    for (int i = 0; i < 12; i++) {
        mad3(t0, t1, t2, h0, h1, a, bb[i]);
        uint32_t m = __shfl_sync(FULL, t0, 0, 4);
        // the word shifted in from the next lane at the END of the previous row is added only now: a warp issues in order, so consuming
        // that shuffle right away would stall the whole row on its latency; one row later it has long arrived (the limb it belongs to is
        // still t2, and lane 0's t0 -- the only word m depends on -- is two rows away from it)
        asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;" : "+r"(t2), "+r"(h0), "+r"(h1) : "r"(ypend));
        m *= FpParams::M0;
        mad3(t0, t1, t2, h0, h1, p, m);
        ypend = __shfl_down_sync(FULL, t0, 1, 4);
        if (gl == 3) ypend = 0;
        t0 = t1; t1 = t2; t2 = h0; h0 = h1; h1 = 0;
    }
    asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, 0;" : "+r"(t2), "+r"(h0) : "r"(ypend));
    for (int pass = 0; pass < 3; pass++) {     // a second / third pass only if a carry ripples through a whole lane
        uint32_t cin = __shfl_up_sync(FULL, h0, 1, 4);
        if (gl == 0) cin = 0;
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;" : "+r"(t0), "+r"(t1), "+r"(t2), "=r"(h0) : "r"(cin));
        if (!__any_sync(FULL, h0 != 0)) break;
    }
    CFp r;
    r.v[0] = t0; r.v[1] = t1; r.v[2] = t2;
    return r;
}
__device__ __forceinline__ void cmul(const Ctx& c, CFp& r, const CFp& a, const CFp& b) { r = cmul(a, b, c.p, c.gl); }
__device__ __forceinline__ void csqr(const Ctx& c, CFp& r, const CFp& a) { r = cmul(a, a, c.p, c.gl); }

// TWO independent products, row by row in lock step: a lone product is a chain of 12 x (shuffle, multiply-adds, shuffle, multiply-adds,
// shuffle) in which the warp mostly waits; the second product's chain fills those gaps (the sync shuffles keep ptxas from interleaving
// two separate calls by itself).  The point formulas below issue their multiplications in independent pairs wherever they have them.
struct CFp2 { CFp a, b; };
static __device__ __noinline__ CFp2 cmul2(CFp a1, CFp b1, CFp a2, CFp b2, CFp p, unsigned gl) {
    uint32_t t0 = 0, t1 = 0, t2 = 0, h0 = 0, h1 = 0, ypend = 0;
    uint32_t u0 = 0, u1 = 0, u2 = 0, k0 = 0, k1 = 0, wpend = 0;
    uint32_t bb[12], cb[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        bb[i] = __shfl_sync(FULL, b1.v[i % 3], i / 3, 4);
        cb[i] = __shfl_sync(FULL, b2.v[i % 3], i / 3, 4);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) {
        mad3(t0, t1, t2, h0, h1, a1, bb[i]);
        uint32_t m = __shfl_sync(FULL, t0, 0, 4);
        mad3(u0, u1, u2, k0, k1, a2, cb[i]);
        uint32_t n = __shfl_sync(FULL, u0, 0, 4);
        asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;" : "+r"(t2), "+r"(h0), "+r"(h1) : "r"(ypend));   // (see cmul)
        asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;" : "+r"(u2), "+r"(k0), "+r"(k1) : "r"(wpend));
        m *= FpParams::M0;
        mad3(t0, t1, t2, h0, h1, p, m);
        ypend = __shfl_down_sync(FULL, t0, 1, 4);
        n *= FpParams::M0;
        mad3(u0, u1, u2, k0, k1, p, n);
        wpend = __shfl_down_sync(FULL, u0, 1, 4);
        if (gl == 3) { ypend = 0; wpend = 0; }
        t0 = t1; t1 = t2; t2 = h0; h0 = h1; h1 = 0;
        u0 = u1; u1 = u2; u2 = k0; k0 = k1; k1 = 0;
    }
    asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, 0;" : "+r"(t2), "+r"(h0) : "r"(ypend));
    asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, 0;" : "+r"(u2), "+r"(k0) : "r"(wpend));
    for (int pass = 0; pass < 3; pass++) {
        uint32_t cin = __shfl_up_sync(FULL, h0, 1, 4), din = __shfl_up_sync(FULL, k0, 1, 4);
        if (gl == 0) { cin = 0; din = 0; }
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;" : "+r"(t0), "+r"(t1), "+r"(t2), "=r"(h0) : "r"(cin));
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;" : "+r"(u0), "+r"(u1), "+r"(u2), "=r"(k0) : "r"(din));
        if (!__any_sync(FULL, (h0 | k0) != 0)) break;
    }
    CFp2 r;
    r.a.v[0] = t0; r.a.v[1] = t1; r.a.v[2] = t2;
    r.b.v[0] = u0; r.b.v[1] = u1; r.b.v[2] = u2;
    return r;
}
// r1 = a1 b1, r2 = a2 b2 (the results may alias the operands)
__device__ __forceinline__ void cmul2(const Ctx& c, CFp& r1, const CFp& a1, const CFp& b1, CFp& r2, const CFp& a2, const CFp& b2) {
    const CFp2 r = cmul2(a1, b1, a2, b2, c.p, c.gl);
    r1 = r.a;
    r2 = r.b;
}

// ---- points (the formulas of g1.cuh / g1_mul.cuh without their exceptional-case branches) ------------------------------------
__device__ __forceinline__ CJac from_jac(const Ctx& c, const G1Jac& p) {
    CJac r;
    r.x = from_fp(c, p.x); r.y = from_fp(c, p.y); r.z = from_fp(c, p.z);
    return r;
}
__device__ __forceinline__ G1Jac to_jac(const Ctx& c, const CJac& p) {
    G1Jac r;
    r.x = to_fp(c, p.x); r.y = to_fp(c, p.y); r.z = to_fp(c, p.z);
    return r;
}

// An addition or subtraction costs two ballot rounds here (~180 clocks, a fifth of a product), so the ladder does NOT use the
// add-heavy formulas of g1.cuh (dbl-2009-l: 14 additions, madd-2007-bl: 15) but forms of the same maps with the small constants moved
// into products and into the choice of the projective representative.  The points are the same; their Jacobian coordinates differ
// from the one-thread kernels' by a factor (lambda^2, lambda^3, lambda), which the final normalisation removes.
//
// Doubling, scaled by lambda = 1/2:  m = (3/2) X^2,  X3 = m^2 - 2 X Y^2,  Y3 = m (X Y^2 - X3) - Y^4,  Z3 = Y Z
// (from S = 4 X Y^2, M = 3 X^2, X3 = M^2 - 2 S, Y3 = M (S - X3) - 8 Y^4, Z3 = 2 Y Z): 8 products in 4 paired steps, 4 additions.
static __device__ __noinline__ void cjac_dbl(const Ctx& c, CJac& r, const CJac& p, const CFp& three_halves) {
    CFp a, b, m, xb, m2, bb, t, y, z;
    cmul2(c, a, p.x, p.x, b, p.y, p.y);
    cmul2(c, m, a, three_halves, xb, p.x, b);
    cmul2(c, m2, m, m, bb, b, b);
    csub(c, m2, m2, xb); csub(c, m2, m2, xb);   // X3
    csub(c, t, xb, m2);
    cmul2(c, y, m, t, z, p.y, p.z);
    csub(c, r.y, y, bb);
    r.x = m2;
    r.z = z;
}

// acc += (px, py) affine, acc neither the identity nor +-(px, py)  (madd-2004-hmv: H = px Z^2 - X, R = py Z^3 - Y,
// X3 = R^2 - H^3 - 2 X H^2, Y3 = R (X H^2 - X3) - Y H^3, Z3 = Z H): 11 products in 6 steps, 7 additions
static __device__ __noinline__ void cjac_madd(const Ctx& c, CJac& acc, const CFp& px, const CFp& py) {
    CFp zz, t, u2, s2, h, rr, hh, r2, hhh, v, x3, z3;
    cmul2(c, zz, acc.z, acc.z, t, py, acc.z);
    cmul2(c, u2, px, zz, s2, t, zz);
    csub(c, h, u2, acc.x);
    csub(c, rr, s2, acc.y);
    cmul2(c, hh, h, h, r2, rr, rr);
    cmul2(c, hhh, hh, h, v, acc.x, hh);
    cmul(c, z3, acc.z, h);
    csub(c, x3, r2, hhh); csub(c, x3, x3, v); csub(c, x3, x3, v);
    csub(c, v, v, x3);
    cmul2(c, v, v, rr, hhh, hhh, acc.y);
    csub(c, acc.y, v, hhh);
    acc.x = x3;
    acc.z = z3;
}

// r = a + (bx, by) affine with zr = Z3 / Z1 (madd-2004-hmv, jac_madd_zr): 11 products in 6 steps; r must not alias a
static __device__ __noinline__ void cjac_madd_zr(const Ctx& c, CJac& r, const CJac& a, const CFp& bx, const CFp& by, CFp& zr) {
    CFp t1, t2, t3, t4, x3;
    csqr(c, t1, a.z);
    cmul2(c, t2, t1, a.z, t1, t1, bx);
    csub(c, t1, t1, a.x);                      // H
    zr = t1;
    cmul2(c, t2, t2, by, r.z, a.z, t1);
    csub(c, t2, t2, a.y);                      // R
    cmul2(c, t3, t1, t1, x3, t2, t2);          // HH, R^2
    cmul2(c, t4, t3, t1, t3, t3, a.x);         // HHH, V
    cdbl(c, t1, t3);
    csub(c, x3, x3, t1);
    csub(c, x3, x3, t4);
    csub(c, t3, t3, x3);
    cmul2(c, t3, t3, t2, t4, t4, a.y);
    csub(c, r.y, t3, t4);
    r.x = x3;
}

// k P for the fixed scalar of an op list (jac_mul_ops): P a non-identity point of the prime-order subgroup, rows 0 and 64 excluded
static __device__ __noinline__ void cjac_mul_ops(const Ctx& c, CJac& out, const CJac& p, const uint16_t* ops) {
    CFp tx[8], ty[8], bx[8], zr[8];
    CFp zg;
    const CFp three_halves = const_fp(c, three_halves_limb);
    CJac d;
    cjac_dbl(c, d, p, three_halves);
    {
        CJac cur;
        CFp dz2, dz3;
        csqr(c, dz2, d.z);
        cmul2(c, dz3, dz2, d.z, cur.x, p.x, dz2);
        cmul(c, cur.y, p.y, dz3);
        cur.z = p.z;
        tx[0] = cur.x; ty[0] = cur.y;
        for (int i = 1; i < 8; i++) {
            CJac nxt;
            cjac_madd_zr(c, nxt, cur, d.x, d.y, zr[i]);
            cur = nxt;
            tx[i] = cur.x; ty[i] = cur.y;
        }
        zg = cur.z;
        CFp zs = zr[7];
        for (int i = 6; i >= 0; i--) {
            CFp z2, z3;
            csqr(c, z2, zs);
            cmul2(c, z3, z2, zs, tx[i], tx[i], z2);
            if (i) cmul2(c, ty[i], ty[i], z3, zs, zs, zr[i]);
            else cmul(c, ty[i], ty[i], z3);
        }
        const CFp beta = const_fp(c, beta_limb);
        for (int i = 0; i < 8; i += 2) cmul2(c, bx[i], tx[i], beta, bx[i + 1], tx[i + 1], beta);
    }
    CJac acc;
    const int n = ops[0];
    for (int k = 1; k <= n; k++) {
        const uint32_t op = ops[k];
        for (int s = op >> 8; s > 0; s--) cjac_dbl(c, acc, acc, three_halves);
        if (op & 0x20) {
            const int idx = op & 7;
            const CFp ex = (op & 0x10) ? bx[idx] : tx[idx];
            CFp ey = ty[idx];
            if (op & 8) cneg(c, ey, ey);
            if (k == 1) {                      // the accumulator is empty (the first op never doubles: tests/test_k5_coop_ops.py)
                acc.x = ex; acc.y = ey; acc.z = const_fp(c, one_limb);
            } else {
                cjac_madd(c, acc, ex, ey);
            }
        }
    }
    cmul(c, acc.z, acc.z, zg);
    cmul(c, acc.z, acc.z, d.z);
    out = acc;
}

}  // namespace coop
}  // namespace ekzg
#endif
