// Warp-cooperative Fp multiplier probe ("What's weak" #9 / "Next" #3): one BLS12-381 Fp element spread over FOUR
// lanes (3 x 32-bit limbs each) against the shipped one-element-per-thread multiplier (csrc/field.cuh), on the two figures that
// decide whether K5 (k_fk20_g1_ntts) should get a second multiplier:
//   * throughput (products/s) as a function of resident warps per SM sub-partition -- K5 holds 2 at 255 registers, a cooperative
//     kernel could hold 8-16;
//   * latency of ONE dependent product when a warp has its sub-partition to itself -- what K5's latency mode is bound by.
// The cooperative product is operand scanning with a lazily carried accumulator: per b-limb one broadcast of b_i, 3 wide
// multiply-adds a_lane * b_i, one broadcast of lane 0's low word for m = t0 * M0, 3 wide multiply-adds p_lane * m, and a shift by
// one limb that pulls the next lane's low word in (3 shuffles, 7 wide multiply-adds per lane and row; 84 per lane and product
// against 75 = 300 / 4).  Values stay in [0, 2p): with R = 2^384 > 4p the product of two such values is again below 2p, so no
// comparison across lanes is needed.  Bit-exactness is checked against fe_mul_inline on random inputs before anything is timed.
// Build: make -C rust-eth-kzg_b200 lib/coop_probe        Run: lib/coop_probe [iters]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "field.cuh"
#ifdef EKZG_PROBE_LIB
#include "g1_coop.cuh"   // the multiplier the library ships (shadow chain of lane 0: no shuffle on the critical path)
#define COOP_MUL coop_mul_lib
#elif defined(EKZG_PROBE_96)
#define COOP_MUL coop_mul96
#else
#define COOP_MUL coop_mul
#endif
using namespace ekzg;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct Fp3 { uint32_t x0, x1, x2; };
#ifdef EKZG_PROBE_LIB
__device__ __forceinline__ coop::CFp to_c(const Fp3& a) { coop::CFp r; r.v[0] = a.x0; r.v[1] = a.x1; r.v[2] = a.x2; return r; }
__device__ __forceinline__ Fp3 from_c(const coop::CFp& a) { Fp3 r; r.x0 = a.v[0]; r.x1 = a.v[1]; r.x2 = a.v[2]; return r; }
__device__ __forceinline__ Fp3 coop_mul_lib(const Fp3& a, const Fp3& b, const Fp3& p, unsigned gl) { return from_c(coop::cmul(to_c(a), to_c(b), to_c(p), gl)); }
#endif

__device__ __forceinline__ uint32_t p_limb(int i) { return FpParams::mod(i); }
__device__ __forceinline__ Fp3 p_of_lane(unsigned gl) {
    Fp3 p;
    p.x0 = gl == 0 ? p_limb(0) : gl == 1 ? p_limb(3) : gl == 2 ? p_limb(6) : p_limb(9);
    p.x1 = gl == 0 ? p_limb(1) : gl == 1 ? p_limb(4) : gl == 2 ? p_limb(7) : p_limb(10);
    p.x2 = gl == 0 ? p_limb(2) : gl == 1 ? p_limb(5) : gl == 2 ? p_limb(8) : p_limb(11);
    return p;
}

// (t0, t1, t2, h0, h1) += a(3 limbs) * s: two carry chains in the style of field.cuh (a0 s and a2 s at limbs 0 and 2, a1 s at limb 1),
// every mad.lo.cc / madc.hi.cc pair one IMAD.WIDE.U32.X; h0, h1 catch what leaves the lane's three limbs
__device__ __forceinline__ void mad3(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& h0, uint32_t& h1, const Fp3& a, uint32_t s) {
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
        : "r"(a.x0), "r"(a.x1), "r"(a.x2), "r"(s));
}

// a, b in [0, 2p), lane gl of a 4-lane group holds limbs 3 gl .. 3 gl + 2; returns a b / 2^384 mod p in [0, 2p).
// EKZG_PROBE_V1: the first schedule (the word shifted in from the next lane is consumed at once: the in-order warp then waits for two
// shuffles per row, 967 clocks per product); default: the schedule of csrc/g1_coop.cuh (b broadcast up front, the shifted-in word added
// one row later).
__device__ __forceinline__ Fp3 coop_mul(const Fp3& a, const Fp3& b, const Fp3& p, unsigned gl) {
    const unsigned full = 0xffffffffu;
    uint32_t t0 = 0, t1 = 0, t2 = 0, h0 = 0, h1 = 0;
#ifdef EKZG_PROBE_V1
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint32_t bsel = (i % 3 == 0) ? b.x0 : (i % 3 == 1) ? b.x1 : b.x2;
        const uint32_t bi = __shfl_sync(full, bsel, i / 3, 4);
        mad3(t0, t1, t2, h0, h1, a, bi);
        const uint32_t m = __shfl_sync(full, t0, 0, 4) * FpParams::M0;
        mad3(t0, t1, t2, h0, h1, p, m);
        uint32_t y = __shfl_down_sync(full, t0, 1, 4);   // the next lane's low word moves into this lane's top limb
        if (gl == 3) y = 0;
        t0 = t1;
        t1 = t2;
        asm("add.cc.u32 %0, %2, %3;\n\t addc.u32 %1, %4, 0;" : "=r"(t2), "=r"(h0) : "r"(h0), "r"(y), "r"(h1));
        h1 = 0;
    }
#else
    uint32_t ypend = 0, bb[12];
#pragma unroll
    for (int i = 0; i < 12; i++) bb[i] = __shfl_sync(full, (i % 3 == 0) ? b.x0 : (i % 3 == 1) ? b.x1 : b.x2, i / 3, 4);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        mad3(t0, t1, t2, h0, h1, a, bb[i]);
        uint32_t m = __shfl_sync(full, t0, 0, 4);
        asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;" : "+r"(t2), "+r"(h0), "+r"(h1) : "r"(ypend));
        m *= FpParams::M0;
        mad3(t0, t1, t2, h0, h1, p, m);
        ypend = __shfl_down_sync(full, t0, 1, 4);
        if (gl == 3) ypend = 0;
        t0 = t1; t1 = t2; t2 = h0; h0 = h1; h1 = 0;
    }
    asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, 0;" : "+r"(t2), "+r"(h0) : "r"(ypend));
#endif
    // fold the deferred carries into the next lane; a second and third pass only if a carry ripples through a whole lane
    for (int pass = 0; pass < 3; pass++) {
        uint32_t c = __shfl_up_sync(full, h0, 1, 4);
        if (gl == 0) c = 0;
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;" : "+r"(t0), "+r"(t1), "+r"(t2), "=r"(h0) : "r"(c));
        if (!__any_sync(full, h0 != 0)) break;
    }
    Fp3 r;
    r.x0 = t0; r.x1 = t1; r.x2 = t2;
    return r;
}

// ---- radix-2^96 variant: a lane's three limbs are ONE Montgomery digit, four rounds instead of twelve --------------------------------
// per round: broadcast the three limbs of b's digit j, T += a_lane * b_j (3x3 limbs), m = T_0 * (-p^-1 mod 2^96) from lane 0,
// T += p_lane * m, shift one DIGIT down (the next lane's low digit comes in).  Same work, a third of the dependent shuffle rounds.
__device__ __forceinline__ void mul3x3(uint32_t (&P)[6], const Fp3& a, uint32_t b0, uint32_t b1, uint32_t b2) {
    uint64_t c;
    c = (uint64_t)a.x0 * b0; P[0] = (uint32_t)c;
    c = (uint64_t)a.x1 * b0 + (c >> 32); P[1] = (uint32_t)c;
    c = (uint64_t)a.x2 * b0 + (c >> 32); P[2] = (uint32_t)c; P[3] = (uint32_t)(c >> 32);
    c = (uint64_t)a.x0 * b1 + P[1]; P[1] = (uint32_t)c;
    c = (uint64_t)a.x1 * b1 + P[2] + (c >> 32); P[2] = (uint32_t)c;
    c = (uint64_t)a.x2 * b1 + P[3] + (c >> 32); P[3] = (uint32_t)c; P[4] = (uint32_t)(c >> 32);
    c = (uint64_t)a.x0 * b2 + P[2]; P[2] = (uint32_t)c;
    c = (uint64_t)a.x1 * b2 + P[3] + (c >> 32); P[3] = (uint32_t)c;
    c = (uint64_t)a.x2 * b2 + P[4] + (c >> 32); P[4] = (uint32_t)c; P[5] = (uint32_t)(c >> 32);
}
__device__ __forceinline__ void add6(uint32_t (&T)[7], const uint32_t (&P)[6]) {
    asm("add.cc.u32 %0, %0, %7;\n\t addc.cc.u32 %1, %1, %8;\n\t addc.cc.u32 %2, %2, %9;\n\t addc.cc.u32 %3, %3, %10;\n\t"
        "addc.cc.u32 %4, %4, %11;\n\t addc.cc.u32 %5, %5, %12;\n\t addc.u32 %6, %6, 0;"
        : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6])
        : "r"(P[0]), "r"(P[1]), "r"(P[2]), "r"(P[3]), "r"(P[4]), "r"(P[5]));
}
__device__ __forceinline__ Fp3 coop_mul96(const Fp3& a, const Fp3& b, const Fp3& p, unsigned gl) {
    const unsigned full = 0xffffffffu;
    const uint32_t N0 = 0xfffcfffdu, N1 = 0x89f3fffcu, N2 = 0xd9d113e8u;   // -p^-1 mod 2^96
    uint32_t T[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t b0 = __shfl_sync(full, b.x0, j, 4), b1 = __shfl_sync(full, b.x1, j, 4), b2 = __shfl_sync(full, b.x2, j, 4);
        uint32_t P[6];
        mul3x3(P, a, b0, b1, b2);
        add6(T, P);
        // m = T[0..2] * N mod 2^96 (every lane on its own digit; lane 0's is the one that counts)
        uint32_t m0, m1, m2;
        {
            uint64_t c = (uint64_t)T[0] * N0; m0 = (uint32_t)c;
            c = (uint64_t)T[0] * N1 + (c >> 32) + (uint64_t)((uint32_t)(T[1] * N0)); m1 = (uint32_t)c;
            m2 = (uint32_t)(c >> 32) + __umulhi(T[1], N0) + T[0] * N2 + T[1] * N1 + T[2] * N0;
        }
        m0 = __shfl_sync(full, m0, 0, 4); m1 = __shfl_sync(full, m1, 0, 4); m2 = __shfl_sync(full, m2, 0, 4);
        mul3x3(P, p, m0, m1, m2);
        add6(T, P);
        uint32_t y0 = __shfl_down_sync(full, T[0], 1, 4), y1 = __shfl_down_sync(full, T[1], 1, 4), y2 = __shfl_down_sync(full, T[2], 1, 4);
        if (gl == 3) { y0 = 0; y1 = 0; y2 = 0; }
        asm("add.cc.u32 %0, %4, %7;\n\t addc.cc.u32 %1, %5, %8;\n\t addc.cc.u32 %2, %6, %9;\n\t addc.u32 %3, %10, 0;"
            : "=r"(T[0]), "=r"(T[1]), "=r"(T[2]), "=r"(T[3]) : "r"(T[3]), "r"(T[4]), "r"(T[5]), "r"(y0), "r"(y1), "r"(y2), "r"(T[6]));
        T[4] = 0; T[5] = 0; T[6] = 0;
    }
    for (int pass = 0; pass < 3; pass++) {
        uint32_t c = __shfl_up_sync(full, T[3], 1, 4);
        if (gl == 0) c = 0;
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, 0, 0;" : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "=r"(T[3]) : "r"(c));
        if (!__any_sync(full, T[3] != 0)) break;
    }
    Fp3 r;
    r.x0 = T[0]; r.x1 = T[1]; r.x2 = T[2];
    return r;
}

// ---- correctness: every group multiplies its own pair both ways ---------------------------------------------------------------
__global__ void k_check(const Fp* a, const Fp* b, uint32_t* bad, int n, int chain) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const unsigned gl = threadIdx.x & 3;
    if (g >= n) return;   // n is a multiple of 8: whole warps
    Fp x = a[g], y = b[g];
    Fp3 cx = {x.v[3 * gl], x.v[3 * gl + 1], x.v[3 * gl + 2]}, cy = {y.v[3 * gl], y.v[3 * gl + 1], y.v[3 * gl + 2]};
    const Fp3 p = p_of_lane(gl);
    for (int it = 0; it < chain; it++) {
        Fp r;
        fe_mul_inline<FpParams>(r, x, y);
        x = r;
        cx = COOP_MUL(cx, cy, p, gl);
    }
    // reassemble the cooperative result in every lane and bring it below p
    uint32_t w[12];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        w[3 * l] = __shfl_sync(0xffffffffu, cx.x0, l, 4);
        w[3 * l + 1] = __shfl_sync(0xffffffffu, cx.x1, l, 4);
        w[3 * l + 2] = __shfl_sync(0xffffffffu, cx.x2, l, 4);
    }
    fe_final_sub<FpParams>(w);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 12; j++) ok = ok && w[j] == x.v[j];
    if (!ok && gl == 0) atomicAdd(bad, 1u);
}

// ---- timing kernels: `iters` dependent products per thread (group) ------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_single(const Fp* a, const Fp* b, Fp* out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fp x = a[t & 1023], y = b[t & 1023];
    for (int it = 0; it < iters; it++) {
        Fp r;
        fe_mul_inline<FpParams>(r, x, y);
        x = r;
    }
    if (x.v[0] == 0x12345678u && x.v[5] == 77u) out[t & 1023] = x;   // keeps the loop alive, practically never writes
}
// the shipped form: the multiplier behind a call (one copy per kernel image)
__global__ void __launch_bounds__(128, 2) k_single_call(const Fp* a, const Fp* b, Fp* out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fp x = a[t & 1023], y = b[t & 1023];
    for (int it = 0; it < iters; it++) { Fp r; fe_mul(r, x, y); x = r; }   // fe_mul(Fp) = the call form on the device
    if (x.v[0] == 0x12345678u && x.v[5] == 77u) out[t & 1023] = x;
}
__global__ void __launch_bounds__(128) k_coop(const Fp* a, const Fp* b, Fp* out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned gl = threadIdx.x & 3;
    const Fp& xa = a[(t >> 2) & 1023];
    const Fp& ya = b[(t >> 2) & 1023];
    Fp3 x = {xa.v[3 * gl], xa.v[3 * gl + 1], xa.v[3 * gl + 2]}, y = {ya.v[3 * gl], ya.v[3 * gl + 1], ya.v[3 * gl + 2]};
    const Fp3 p = p_of_lane(gl);
    for (int it = 0; it < iters; it++) x = COOP_MUL(x, y, p, gl);
    if (x.x0 == 0x12345678u && x.x2 == 77u) out[(t >> 2) & 1023].v[gl] = x.x1;
}
// two independent products per group in flight (what a point formula offers: its multiplications come in independent pairs)
__global__ void __launch_bounds__(128) k_coop2(const Fp* a, const Fp* b, Fp* out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned gl = threadIdx.x & 3;
    const Fp& xa = a[(t >> 2) & 1023];
    const Fp& ya = b[(t >> 2) & 1023];
    Fp3 x = {xa.v[3 * gl], xa.v[3 * gl + 1], xa.v[3 * gl + 2]}, y = {ya.v[3 * gl], ya.v[3 * gl + 1], ya.v[3 * gl + 2]};
    Fp3 u = y, v = x;
    const Fp3 p = p_of_lane(gl);
    for (int it = 0; it < iters; it++) {
#ifdef EKZG_PROBE_LIB
        const coop::CFp2 r2 = coop::cmul2(to_c(x), to_c(y), to_c(u), to_c(v), to_c(p), gl);   // the dual multiplier
        x = from_c(r2.a);
        u = from_c(r2.b);
#else
        x = COOP_MUL(x, y, p, gl);
        u = COOP_MUL(u, v, p, gl);
#endif
    }
    if ((x.x0 ^ u.x0) == 0x12345678u && x.x2 == 77u) out[(t >> 2) & 1023].v[gl] = x.x1 + u.x1;
}

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    std::vector<Fp> ha(1024), hb(1024);
    for (int i = 0; i < 1024; i++)
        for (int j = 0; j < 12; j++) {
            ha[i].v[j] = rnd(); hb[i].v[j] = rnd();
            if (j == 11) { ha[i].v[j] %= 0x1a0111eau; hb[i].v[j] %= 0x1a0111eau; }   // below p
        }
    // adversarial limbs: all-ones / zero words, p - 1, so that the deferred carries ripple
    for (int j = 0; j < 12; j++) { ha[0].v[j] = j == 11 ? 0x1a0111e9u : 0xffffffffu; hb[0].v[j] = ha[0].v[j]; ha[1].v[j] = 0; hb[1].v[j] = j < 11 ? 0xffffffffu : 0x0a000000u; }
    for (int j = 0; j < 12; j++) { ha[2].v[j] = FpParams::mod(j); hb[2].v[j] = FpParams::mod(j); }
    ha[2].v[0] -= 1; hb[2].v[0] -= 1;
    for (int j = 0; j < 12; j++) { ha[3].v[j] = (j % 3 == 2) ? 0xffffffffu : 0u; hb[3].v[j] = (j % 3 == 0) ? 0xffffffffu : 0u; if (j == 11) { ha[3].v[j] = 0x1a000000u; } }
    Fp *da, *db, *dout;
    uint32_t* dbad;
    CK(cudaMalloc(&da, sizeof(Fp) * 1024)); CK(cudaMalloc(&db, sizeof(Fp) * 1024)); CK(cudaMalloc(&dout, sizeof(Fp) * 1024)); CK(cudaMalloc(&dbad, 4));
    CK(cudaMemcpy(da, ha.data(), sizeof(Fp) * 1024, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), sizeof(Fp) * 1024, cudaMemcpyHostToDevice));
    CK(cudaMemset(dbad, 0, 4));
    for (int chain : {1, 2, 17}) k_check<<<1024 * 4 / 128, 128>>>(da, db, dbad, 1024, chain);
    CK(cudaDeviceSynchronize());
    uint32_t bad = 0;
    CK(cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost));
    printf("{\"check\": {\"groups\": 1024, \"chains\": [1, 2, 17], \"mismatches\": %u}", bad);
    if (bad) { printf("}\n"); return 1; }

    auto rate = [&](double ms, double products) { return products / (ms * 1e-3) / 1e9; };   // G products / s
    printf(", \"sms\": %d, \"clock_mhz\": %d, \"iters\": %d, \"throughput_Gprod_s\": {", sms, clk_khz / 1000, iters);
    // single-thread multiplier: CTAs of 128 threads (one warp per sub-partition each); w CTAs per SM = w warps per sub-partition
    bool first = true;
    for (int w : {1, 2, 3, 4}) {
        const int blocks = sms * w;
        double ms = w <= 2 ? time_ms([&] { k_single<2><<<blocks, 128>>>(da, db, dout, iters); }) : time_ms([&] { k_single<4><<<blocks, 128>>>(da, db, dout, iters); });
        printf("%s\"single_inline_w%d\": %.2f", first ? "" : ", ", w, rate(ms, (double)blocks * 128 * iters));
        first = false;
    }
    for (int w : {1, 2}) {
        const int blocks = sms * w;
        double ms = time_ms([&] { k_single_call<<<blocks, 128>>>(da, db, dout, iters); });
        printf(", \"single_call_w%d\": %.2f", w, rate(ms, (double)blocks * 128 * iters));
    }
    for (int w : {1, 2, 4, 8, 12, 16}) {
        const int blocks = sms * w;
        double ms = time_ms([&] { k_coop<<<blocks, 128>>>(da, db, dout, iters); });
        printf(", \"coop4_w%d\": %.2f", w, rate(ms, (double)blocks * 32 * iters));
    }
    for (int w : {1, 2, 4, 8}) {
        const int blocks = sms * w;
        double ms = time_ms([&] { k_coop2<<<blocks, 128>>>(da, db, dout, iters); });
        printf(", \"coop4_pair_w%d\": %.2f", w, rate(ms, (double)blocks * 32 * 2 * iters));
    }
    // latency of one dependent product: one warp per sub-partition, nothing else on the SM
    double ms_s = time_ms([&] { k_single<2><<<sms, 128>>>(da, db, dout, iters); });
    double ms_sc = time_ms([&] { k_single_call<<<sms, 128>>>(da, db, dout, iters); });
    double ms_c = time_ms([&] { k_coop<<<sms, 128>>>(da, db, dout, iters); });
    double ms_c2 = time_ms([&] { k_coop2<<<sms, 128>>>(da, db, dout, iters); });
    const double clk = clk_khz * 1e3;
    printf("}, \"latency_clocks_per_dependent_product\": {\"single_inline\": %.0f, \"single_call\": %.0f, \"coop4\": %.0f, \"coop4_two_independent\": %.0f}",
           ms_s * 1e-3 * clk / iters, ms_sc * 1e-3 * clk / iters, ms_c * 1e-3 * clk / iters, ms_c2 * 1e-3 * clk / iters / 2);
    printf("}\n");
    return 0;
}
