// Host-emulation harness for the Fp interpreter of the hot kernels (csrc/fpvm.cuh); see emu_field.cpp.
#include "fpvm.cuh"
#include <string.h>
using namespace ekzg;
static const uint32_t PROG[] = FPVM_PROG_INIT;
extern "C" {
int emu_fpvm_words() { return fpvm::FPVM_PROG_WORDS; }
uint32_t emu_fpvm_word(int i) { return PROG[i]; }
// slots: 8 x 12 limbs, Montgomery form, in and out; runs instructions [pc, pc + n)
uint32_t emu_fpvm_run(int pc, int n, int reps, uint32_t* slots) {
    fpvm::HostMem m;
    for (int s = 0; s < 8; s++) memcpy(m.s[s].v, slots + 12 * s, 48);
    uint32_t z = fpvm::host_run(m, PROG, pc, n, reps);
    for (int s = 0; s < 8; s++) memcpy(slots + 12 * s, m.s[s].v, 48);
    return z;
}
}
