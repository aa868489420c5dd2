// The work units of the G1-NTT kernels (K5): the radix-2 butterfly of k_fk20_g1_ntts and the multiplication / combination units
// of the radix-4 latency-mode kernel k_fk20_g1_ntts_r4 (kzg_kernels.cu has the schedules and the write-up).  They live in a header
// so that tests/host_emu can run the very same code on the CPU, unit by unit in ticket order, on real curve points.
//   reference: the generic butterfly of polynomial/src/fft.rs:164-177 under Domain::ifft_g1_take_n / fft_g1
//   (polynomial/src/domain.rs:149-194).
// `tw` is the table of fixed-scalar op lists, row e = omega_128^e (twiddle_ops.inc): __constant__ on the device.
#pragma once
#include <cstddef>
#include "g1_mul.cuh"

namespace ekzg {

typedef const uint16_t (*TwiddleOps)[MULOPS_STRIDE];

#ifdef __CUDACC__
#define EKZG_NTT_UNIT static __device__ __noinline__
#define EKZG_NTT_INL __device__ __forceinline__
EKZG_NTT_INL G1Jac ld_pt(const G1Jac* p) {  // L2-coherent: the producer ran on another SM
    G1Jac r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(G1Jac) / 16); i++) {
        const uint4 q = __ldcg(s + i);
        limb_word(r, 4 * i) = q.x; limb_word(r, 4 * i + 1) = q.y; limb_word(r, 4 * i + 2) = q.z; limb_word(r, 4 * i + 3) = q.w;
    }
    return r;
}
EKZG_NTT_INL void st_pt(G1Jac* p, const G1Jac& v) {
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(G1Jac) / 16); i++)
        __stcg(d + i, make_uint4(limb_word(v, 4 * i), limb_word(v, 4 * i + 1), limb_word(v, 4 * i + 2), limb_word(v, 4 * i + 3)));
}
#else
#define EKZG_NTT_UNIT static inline
#define EKZG_NTT_INL static inline
EKZG_NTT_INL G1Jac ld_pt(const G1Jac* p) { return *p; }
EKZG_NTT_INL void st_pt(G1Jac* p, const G1Jac& v) { *p = v; }
#endif

EKZG_NTT_UNIT void g1_ntt_butterfly(G1Jac* __restrict__ pts, int B, int b, int t, int ph, TwiddleOps tw) {
    const int mode = ph >= 7, st = mode ? 13 - ph : ph;
    const int len = 1 << st;
    const int pos = t & (len - 1);
    const int i = ((t >> st) << (st + 1)) + pos, j = i + len;
    const int e = pos << (6 - st);  // twiddle exponent of omega_128
    G1Jac* pi = &pts[(size_t)i * B + b];
    G1Jac* pj = &pts[(size_t)j * B + b];
    if (mode == 0) {
        G1Jac u = ld_pt(pi), v = ld_pt(pj);
        if (e != 0 && !jac_is_inf(v)) jac_mul_ops(v, v, tw[(128 - e) & 127]);
        G1Jac s = u;
        jac_add(s, v);
        st_pt(pi, s);
        if (st != 6) {
            jac_neg(v, v);
            jac_add(u, v);
            st_pt(pj, u);
        }
    } else if (st == 6) {
        G1Jac u = ld_pt(pi);
        if (e != 0 && !jac_is_inf(u)) jac_mul_ops(u, u, tw[e]);
        st_pt(pj, u);
    } else {
        G1Jac u = ld_pt(pi), v = ld_pt(pj);
        G1Jac s = u;
        jac_add(s, v);
        jac_neg(v, v);
        jac_add(u, v);
        if (e != 0 && !jac_is_inf(u)) jac_mul_ops(u, u, tw[e]);
        st_pt(pi, s);
        st_pt(pj, u);
    }
}

constexpr int R4_SUPER = 7;
constexpr int R4_UNITS = 192;          // units per blob group and super-phase: 160 + 32, middle 128 + 64
constexpr int R4_TMP_POINTS = 160;     // products per blob and super-phase
EKZG_NTT_INL int r4_nmul(int sp) { return sp == 3 ? 128 : 160; }

// p <- omega_128^tw * p (tw in [0, 128))
EKZG_NTT_INL void r4_twiddle_mul(G1Jac& p, int e, TwiddleOps tw) {
    if (e == 0 || jac_is_inf(p)) return;
    if (e == 64) { jac_neg(p, p); return; }
    jac_mul_ops(p, p, tw[e]);
}
EKZG_NTT_INL void r4_sub(G1Jac& a, const G1Jac& b) {   // a -= b
    G1Jac n;
    jac_neg(n, b);
    jac_add(a, n);
}

// which point (or combination of points) a multiplication unit multiplies by which root of unity:
// p = pts[i0] - pts[i1] + pts[i2] - pts[i3] (absent terms: -1), then p *= omega^e
EKZG_NTT_INL void r4_mul_operands(int sp, int u, int& i0, int& i1, int& i2, int& i3, int& e) {
    i1 = i2 = i3 = -1;
    if (sp == 3) {           // middle: omega^-t x1 (even u) and omega^t x0 (odd u)
        const int t = u >> 1;
        if (u & 1) { i0 = t; e = t; } else { i0 = t + 64; e = (128 - t) & 127; }
    } else {
        const bool fwd = sp > 3;
        const int s = fwd ? 2 * (6 - sp) : 2 * sp, len = 1 << s;
        const int q = u / 5, which = u - 5 * q;
        const int pos = q & (len - 1), base = ((q >> s) << (s + 2)) + pos;
        const int ea = pos << (6 - s), eb = pos << (5 - s);
        e = which == 0 ? ea : which == 1 ? eb : which == 2 ? (fwd ? eb + 32 : ea + eb) : which == 3 ? (fwd ? ea + eb : eb + 32) : ea + eb + 32;
        if (!fwd) {
            i0 = base + (which == 0 ? 1 : (which == 1 || which == 3) ? 2 : 3) * len;
            e = (128 - e) & 127;
        } else {
            // which 0: (x0 - x1) + (x2 - x3); 1, 3: x0 - x2; 2, 4: x1 - x3
            const int a = (which == 2 || which == 4) ? 1 : 0;
            i0 = base + a * len;
            i1 = base + (which == 0 ? 1 : a + 2) * len;
            if (which == 0) { i2 = base + 2 * len; i3 = base + 3 * len; }
            e &= 127;
        }
    }
}
EKZG_NTT_INL G1Jac r4_mul_operand_point(const G1Jac* __restrict__ pts, int B, int b, int i0, int i1, int i2, int i3) {
    G1Jac p = ld_pt(&pts[(size_t)i0 * B + b]);
    if (i1 >= 0) r4_sub(p, ld_pt(&pts[(size_t)i1 * B + b]));
    if (i2 >= 0) {
        G1Jac d = ld_pt(&pts[(size_t)i2 * B + b]);
        r4_sub(d, ld_pt(&pts[(size_t)i3 * B + b]));
        jac_add(p, d);
    }
    return p;
}

EKZG_NTT_UNIT void r4_mul_unit(G1Jac* __restrict__ pts, G1Jac* __restrict__ tmp, int B, int b, int sp, int u, TwiddleOps tw) {
    int i0, i1, i2, i3, e;
    r4_mul_operands(sp, u, i0, i1, i2, i3, e);
    G1Jac p = r4_mul_operand_point(pts, B, b, i0, i1, i2, i3);
    r4_twiddle_mul(p, e, tw);
    st_pt(&tmp[(size_t)u * B + b], p);
}

#ifdef __CUDACC__
}  // namespace ekzg
#include "g1_coop.cuh"
namespace ekzg {
// The same unit for EIGHT blobs per warp, four lanes per blob (batches that leave most of the machine idle): the four lanes of a
// group form the operand point redundantly in one-thread arithmetic (a handful of additions), then run the fixed-scalar ladder --
// 98 % of the unit -- cooperatively, each on its three limbs of every coordinate (g1_coop.cuh: half the latency per dependent
// product).  Every lane of the warp must come here (the ladder shuffles and votes across the whole warp); groups without work
// (b >= B, identity operand) run the ladder on whatever they hold and drop the result.  Bit-identical to r4_mul_unit.
EKZG_NTT_UNIT void r4_mul_unit_coop(G1Jac* __restrict__ pts, G1Jac* __restrict__ tmp, int B, int b, int sp, int u, TwiddleOps tw) {
    int i0, i1, i2, i3, e;
    r4_mul_operands(sp, u, i0, i1, i2, i3, e);
    const bool valid = b < B;
    G1Jac p;
    if (valid) p = r4_mul_operand_point(pts, B, b, i0, i1, i2, i3);
    else jac_set_inf(p);
    if (e == 64) {
        jac_neg(p, p);
    } else if (e != 0) {     // e is the same for the whole warp
        const bool active = valid && !jac_is_inf(p);
        if (__any_sync(0xffffffffu, active)) {
            const coop::Ctx c = coop::make_ctx();
            coop::CJac r;
            coop::cjac_mul_ops(c, r, coop::from_jac(c, p), tw[e]);
            const G1Jac full = coop::to_jac(c, r);
            if (active) p = full;
        }
    }
    if (valid && (threadIdx.x & 3) == 0) st_pt(&tmp[(size_t)u * B + b], p);
}
#endif

#ifdef __CUDACC__
// The combination unit of the cooperative kernel (8 blobs per warp): the four lanes of a blob share the unit's additions.  Its 5 - 8
// additions are only two deep (a = x0 + p1, bm = x0 - p1, cc = p2 + p3, d = p4 - p5, then four independent sums), and in a batch this
// small the other lanes are idle anyway: every lane does ONE addition per level, its operands and sign picked by its role k = lane & 3
// -- one instruction stream for the whole warp --, the intermediate sums change hands through `xch` (this warp's 32 slots of shared
// memory).  All loads from pts happen before the first store to it.  Same sums as r4_combine_unit.
EKZG_NTT_UNIT void r4_combine_unit_par(G1Jac* __restrict__ pts, const G1Jac* __restrict__ tmp, int B, int b, int k, bool valid, int sp, int c,
                                       G1Jac* __restrict__ xch /* indexed by lane */) {
    const int lane = threadIdx.x & 31, base_lane = lane & 28;
    G1Jac pair[2];   // both operands of an addition in ONE array: see r4_combine_unit below
    if (sp == 3) {   // pts[c] += tmp[2c], pts[c + 64] += tmp[2c + 1]: two independent additions
        if (valid && k < 2) {
            G1Jac* o = &pts[(size_t)(c + 64 * k) * B + b];
            pair[0] = ld_pt(o);
            pair[1] = ld_pt(&tmp[(size_t)(2 * c + k) * B + b]);
            jac_add(pair[0], pair[1]);
            st_pt(o, pair[0]);
        }
        return;
    }
    const bool fwd = sp > 3;
    const int s = fwd ? 2 * (6 - sp) : 2 * sp, len = 1 << s;
    const int pos = c & (len - 1), base = ((c >> s) << (s + 2)) + pos;
    G1Jac* o[4];
    for (int q = 0; q < 4; q++) o[q] = &pts[(size_t)(base + q * len) * B + b];
    const G1Jac* t0 = &tmp[(size_t)(5 * c) * B + b];
    // level 1
    if (valid) {
        const G1Jac* pa;
        const G1Jac* pb;
        bool neg;
        if (!fwd) {
            pa = k < 2 ? o[0] : (k == 2 ? t0 + (size_t)B : t0 + (size_t)3 * B);
            pb = k < 2 ? t0 : (k == 2 ? t0 + (size_t)2 * B : t0 + (size_t)4 * B);
            neg = (k & 1) != 0;
        } else {
            pa = k == 0 ? o[0] : k == 1 ? o[2] : (k == 2 ? t0 + (size_t)B : t0 + (size_t)3 * B);
            pb = k == 0 ? o[1] : k == 1 ? o[3] : (k == 2 ? t0 + (size_t)2 * B : t0 + (size_t)4 * B);
            neg = k == 3;
        }
        pair[0] = ld_pt(pa);
        pair[1] = ld_pt(pb);
        jac_cneg(pair[1], pair[1], neg);
        jac_add(pair[0], pair[1]);
        xch[lane] = pair[0];
    }
    __syncwarp();
    // level 2
    if (valid) {
        if (!fwd) {          // k = 0: a + cc -> o0, 1: a - cc -> o2, 2: bm + d -> o1, 3: bm - d -> o3
            pair[0] = xch[base_lane + (k < 2 ? 0 : 1)];
            pair[1] = xch[base_lane + (k < 2 ? 2 : 3)];
            jac_cneg(pair[1], pair[1], (k & 1) != 0);
            jac_add(pair[0], pair[1]);
            st_pt(o[k == 0 ? 0 : k == 1 ? 2 : k == 2 ? 1 : 3], pair[0]);
        } else if (k == 0) { // (x0 + x1) + (x2 + x3) -> o0
            pair[0] = xch[base_lane];
            pair[1] = xch[base_lane + 1];
            jac_add(pair[0], pair[1]);
            st_pt(o[0], pair[0]);
        } else if (k == 1) {
            st_pt(o[1], ld_pt(t0));
        } else {             // the sums of level 1 are final: p2 + p3 -> o2, p4 - p5 -> o3
            st_pt(o[k], xch[lane]);
        }
    }
    __syncwarp();            // the next unit of this warp reuses xch
}
#endif

EKZG_NTT_UNIT void r4_combine_unit(G1Jac* __restrict__ pts, const G1Jac* __restrict__ tmp, int B, int b, int sp, int c) {
    if (sp == 3) {           // pts[c] = x0 + omega^-c x1, pts[c + 64] = x1 + omega^c x0
        // Both operands of the addition sit in ONE array on purpose.  As two separate locals (`G1Jac acc = ld_pt(o); const G1Jac p =
        // ld_pt(..); jac_add(acc, p);`, and before that as two named points plus two temporaries) nvcc 12.9 gave them the same
        // stack slot for sm_100a: jac_add then saw acc == p and returned 2 p -- bit-exact under host emulation, clean under
        // compute-sanitizer, wrong on the device (found with device printf and the prefix hook).
        G1Jac pair[2];
        for (int k = 0; k < 2; k++) {
            G1Jac* o = &pts[(size_t)(c + 64 * k) * B + b];
            pair[0] = ld_pt(o);
            pair[1] = ld_pt(&tmp[(size_t)(2 * c + k) * B + b]);
            jac_add(pair[0], pair[1]);
            st_pt(o, pair[0]);
        }
        return;
    }
    const bool fwd = sp > 3;
    const int s = fwd ? 2 * (6 - sp) : 2 * sp, len = 1 << s;
    const int pos = c & (len - 1), base = ((c >> s) << (s + 2)) + pos;
    G1Jac* o0 = &pts[(size_t)base * B + b];
    G1Jac* o1 = &pts[(size_t)(base + len) * B + b];
    G1Jac* o2 = &pts[(size_t)(base + 2 * len) * B + b];
    G1Jac* o3 = &pts[(size_t)(base + 3 * len) * B + b];
    const G1Jac* t0 = &tmp[(size_t)(5 * c) * B + b];
    if (!fwd) {
        G1Jac a = ld_pt(o0), bm = a;
        { const G1Jac p1 = ld_pt(t0); jac_add(a, p1); r4_sub(bm, p1); }
        G1Jac cc = ld_pt(t0 + (size_t)B);
        jac_add(cc, ld_pt(t0 + (size_t)2 * B));                 // p2 + p3
        G1Jac y = a;
        jac_add(y, cc);
        st_pt(o0, y);
        r4_sub(a, cc);
        st_pt(o2, a);
        G1Jac d = ld_pt(t0 + (size_t)3 * B);
        r4_sub(d, ld_pt(t0 + (size_t)4 * B));                   // p4 - p5
        y = bm;
        jac_add(y, d);
        st_pt(o1, y);
        r4_sub(bm, d);
        st_pt(o3, bm);
    } else {
        G1Jac y = ld_pt(o0);
        jac_add(y, ld_pt(o1));
        G1Jac z = ld_pt(o2);
        jac_add(z, ld_pt(o3));
        jac_add(y, z);
        st_pt(o0, y);
        st_pt(o1, ld_pt(t0));
        y = ld_pt(t0 + (size_t)B);
        jac_add(y, ld_pt(t0 + (size_t)2 * B));
        st_pt(o2, y);
        y = ld_pt(t0 + (size_t)3 * B);
        r4_sub(y, ld_pt(t0 + (size_t)4 * B));
        st_pt(o3, y);
    }
}


}  // namespace ekzg
