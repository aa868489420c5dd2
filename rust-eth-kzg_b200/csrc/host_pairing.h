// Host pairing check against the fixed G2 points of the verification keys (see host_pairing.cpp).
#pragma once
#include <stdint.h>

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

}  // namespace host
}  // namespace ekzg
