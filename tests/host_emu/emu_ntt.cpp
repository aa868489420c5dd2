// Host-emulation harness for the G1-NTT work units (g1_ntt_units.cuh): the radix-2 butterflies of k_fk20_g1_ntts and the
// multiplication / combination units of the radix-4 kernel, run one by one in ticket order on real curve points.
#include "g1_ntt_units.cuh"
#include <string.h>
#include <vector>
using namespace ekzg;

static void out48(uint8_t* out, const G1Jac& j) {
    G1Affine a;
    if (jac_is_inf(j)) g1a_set_inf(a);
    else { Fp zi; fp_inv(zi, j.z); jac_to_affine_with_inv(a, j, zi); }
    g1a_compress(out, a);
}

extern "C" {
// form 0: the first `phases` (0..14) radix-2 phases; form 1: the first phases / 2 radix-4 super-phases.
// in: 128 compressed points in storage order (what K4 leaves in pts[][b]); out: the 128 points afterwards, compressed.
int emu_g1_ntt(int form, int phases, const uint8_t* in128x48, uint8_t* out128x48) {
    std::vector<G1Jac> pts(128), tmp(R4_TMP_POINTS);
    for (int i = 0; i < 128; i++) {
        G1Affine a;
        if (g1a_decompress(a, in128x48 + 48 * i)) return 1;
        jac_from_affine(pts[i], a);
    }
    if (form == 0) {
        for (int ph = 0; ph < phases; ph++)
            for (int t = 0; t < 64; t++) g1_ntt_butterfly(pts.data(), 1, 0, t, ph, TWIDDLE_OPS_HOST);
    } else {
        for (int sp = 0; sp < phases / 2; sp++) {
            const int nmul = r4_nmul(sp);
            for (int u = 0; u < nmul; u++) r4_mul_unit(pts.data(), tmp.data(), 1, 0, sp, u, TWIDDLE_OPS_HOST);
            for (int c = 0; c < R4_UNITS - nmul; c++) r4_combine_unit(pts.data(), tmp.data(), 1, 0, sp, c);
        }
    }
    for (int i = 0; i < 128; i++) out48(out128x48 + 48 * i, pts[i]);
    return 0;
}
}


