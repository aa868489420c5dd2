// Host pairing check against the fixed G2 points of the verification keys (see host_pairing.cpp).
#pragma once
#include <stdint.h>
#include <string>

namespace ekzg {
namespace host {

enum class G2Sel : int { Gen = 0, Tau = 1, Tau64 = 2, NegGen = 3, NegTau = 4, NegTau64 = 5 };

struct PairingInput {
    uint64_t g1_x[6];  // affine coordinates, plain integers < p, little-endian 64-bit limbs
    uint64_t g1_y[6];
    bool g1_is_identity;
    G2Sel g2;
};

// prod_i e(P_i, Q_i) == 1
bool pairing_check(const PairingInput* in, int n);

// the same for G1 points as the device leaves them: Jacobian coordinates in Montgomery form (R = 2^384, 6 x 64-bit limbs);
// the host normalises (one inversion per point)
struct PairingInputJac {
    uint64_t x[6], y[6], z[6];
    bool g1_is_identity;
    G2Sel g2;
};
struct G2Keys;   // line tables of a caller-supplied setup's G2 points (nullptr = the embedded ceremony's)
bool pairing_check_jac(const PairingInputJac* in, int n, const G2Keys* keys = nullptr);
// g2_monomial of a caller-supplied trusted setup: `count` (must be 65) compressed points of 96 bytes.  Decompression, on-curve
// check, optional subgroup check of all of them; keeps [1]_2, [tau]_2, [tau^64]_2.  nullptr + *err on failure.
G2Keys* g2_keys_from_compressed(const uint8_t* g2, int count, bool subgroup_check, std::string* err);
void g2_keys_free(G2Keys* k);
int g2_decompress_plain(const uint8_t* in96, uint64_t* out24);   // test hook
// the sparse line product and the cyclotomic squaring against the general Fp12 routines (test hook)
bool pairing_selftest();

}  // namespace host
}  // namespace ekzg
