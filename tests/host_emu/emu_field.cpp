// Host-emulation harness: compiles the DEVICE field/curve headers with the PTX primitives
// replaced by their CPU emulation (mp.cuh) and exports them for ctypes (tests/test_host_emu.py).
// Test scaffolding only.
#include "field.cuh"
#include "fp_inv_gcd.cuh"
#include <string.h>
using namespace ekzg;
extern "C" {
void emu_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 48); memcpy(y.v, b, 48); fe_mul(z, x, y); memcpy(r, z.v, 48); }
void emu_fp_sqr(const uint32_t* a, uint32_t* r) { Fp x, z; memcpy(x.v, a, 48); fe_sqr(z, x); memcpy(r, z.v, 48); }
void emu_fr_sqr(const uint32_t* a, uint32_t* r) { Fr x, z; memcpy(x.v, a, 32); fe_sqr(z, x); memcpy(r, z.v, 32); }
void emu_fp_mul2(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
    Fp x, y, z, w, o; memcpy(x.v, a, 48); memcpy(y.v, b, 48); memcpy(z.v, c, 48); memcpy(w.v, d, 48); fp_mul2_add(o, x, y, z, w); memcpy(r, o.v, 48); }
void emu_fp_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 48); memcpy(y.v, b, 48); fe_add(z, x, y); memcpy(r, z.v, 48); }
void emu_fp_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 48); memcpy(y.v, b, 48); fe_sub(z, x, y); memcpy(r, z.v, 48); }
void emu_fp_neg(const uint32_t* a, uint32_t* r) { Fp x, z; memcpy(x.v, a, 48); fe_neg(z, x); memcpy(r, z.v, 48); }
void emu_fp_inv(const uint32_t* a, uint32_t* r) { Fp x, z; memcpy(x.v, a, 48); fp_inv(z, x); memcpy(r, z.v, 48); }
void emu_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fe_mul(z, x, y); memcpy(r, z.v, 32); }
void emu_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fe_add(z, x, y); memcpy(r, z.v, 32); }
void emu_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fe_sub(z, x, y); memcpy(r, z.v, 32); }
void emu_fr_inv(const uint32_t* a, uint32_t* r) { Fr x, z; memcpy(x.v, a, 32); fr_inv(z, x); memcpy(r, z.v, 32); }
int emu_fr_ge_mod(const uint32_t* a) { Fr x; memcpy(x.v, a, 32); return fe_plain_ge_mod(x); }
int emu_fp_gt_half(const uint32_t* a) { Fp x; memcpy(x.v, a, 48); return fe_plain_gt_half(x); }
// inversion by division steps (csrc/fp_inv_gcd.cuh), Montgomery form in and out
void emu_fp_inv_gcd(const uint32_t* a, uint32_t* r) {
    Fp x, y;
    memcpy(x.v, a, 48);
    fp_inv_gcd(y, x);
    memcpy(r, y.v, 48);
}
}
