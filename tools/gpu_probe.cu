// GPU probe: (1) integer-multiply issue-rate microbenchmarks (the roofline denominator for the
// field-arithmetic kernels: SURVEY.md §8d "IMAD peak ... must be measured first"), (2) Fp/Fr Montgomery
// multiplication throughput, (3) device-vs-host-emulation cross-check of the arithmetic headers: the
// same EKZG_HD functions run on the GPU (PTX carry chains) and on the CPU (emulated carry flag) and
// must agree bit for bit.  Build: make -C rust-eth-kzg_b200 probes.   Run on the GPU box: rust-eth-kzg_b200/lib/gpu_probe [out.json]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "g1_mul.cuh"
using namespace ekzg;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

// ---------------- (1) issue-rate microbenchmarks ----------------
template <int ILP>
__global__ void k_imad_lo(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_imad_hi(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_imad_wide(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    unsigned long long x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        // operands that change every iteration: with loop-invariant a, b ptxas folds the product out of the loop and the
        // kernel measures IADD3 (that is what the first version of this probe reported as "imad_wide": 17.7 T/s)
        for (int i = 0; i < ILP; i++) x[i] = (unsigned long long)(uint32_t)x[i] * (uint32_t)(x[i] >> 32) + x[(i + 1) % ILP];
    }
    unsigned long long s = 0;
    for (int i = 0; i < ILP; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// ILP independent lo/hi carry chains of length 8 (what ptxas fuses into IMAD.WIDE.U32.X)
template <int ILP>
__global__ void k_imad_chain(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[ILP][8];
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 8; j++) x[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                "madc.lo.cc.u32 %2, %8, %9, %2;\n\t madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                "madc.lo.cc.u32 %4, %8, %9, %4;\n\t madc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                "madc.lo.cc.u32 %6, %8, %9, %6;\n\t madc.hi.u32 %7, %8, %9, %7;"
                : "+r"(x[i][0]), "+r"(x[i][1]), "+r"(x[i][2]), "+r"(x[i][3]), "+r"(x[i][4]), "+r"(x[i][5]), "+r"(x[i][6]), "+r"(x[i][7])
                : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 8; j++) s ^= x[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class P, int ILP>
__global__ void k_fe_mul(Fe<P>* io, int iters) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    Fe<P> x[ILP], y = io[gid];
    for (int i = 0; i < ILP; i++) { x[i] = y; x[i].v[0] ^= i; x[i].v[P::N - 1] &= 0x0fffffffu; }
    y.v[P::N - 1] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) fe_mul(x[i], x[i], y);
    }
    Fe<P> s = x[0];
    for (int i = 1; i < ILP; i++) fe_add(s, s, x[i]);
    io[gid] = s;
}

template <class F>
static double time_ms(F launch, int reps = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
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

// ---------------- (3) cross-check kernels ----------------
struct Case { Fp a, b; Fr c, d; G1Affine p, q; G1Jac jp, jq; int e; uint32_t k[8]; };
struct Res { Fp mul, add, sub, neg; Fr rmul, radd, rsub; G1Xyzz madd_p, madd_n, xadd, xdbl; G1Jac jadd, jdbl, jmadd, jglv, jfromx; int booth[8]; uint8_t comp[48]; };

__host__ __device__ void run_case(const Case& c, Res& r) {
    fe_mul(r.mul, c.a, c.b); fe_add(r.add, c.a, c.b); fe_sub(r.sub, c.a, c.b); fe_neg(r.neg, c.a);
    fe_mul(r.rmul, c.c, c.d); fe_add(r.radd, c.c, c.d); fe_sub(r.rsub, c.c, c.d);
    G1Xyzz x; xyzz_from_affine(x, c.p); xyzz_madd(x, c.q, false); xyzz_madd(x, c.p, false); r.madd_p = x;   // 2p+q
    xyzz_from_affine(x, c.p); xyzz_madd(x, c.q, true); r.madd_n = x;
    G1Xyzz y; xyzz_from_affine(y, c.q); xyzz_madd(y, c.q, false);  // 2q via the doubling branch
    x = r.madd_p; xyzz_add(x, y); r.xadd = x;
    xyzz_dbl(r.xdbl, r.madd_p);
    G1Jac j = c.jp; jac_add(j, c.jq); r.jadd = j;
    jac_dbl(r.jdbl, c.jp);
    j = c.jp; jac_madd(j, c.q, true); r.jmadd = j;
    jac_from_xyzz(r.jfromx, r.xadd);
    for (int t = 0; t < 8; t++) r.booth[t] = booth_digit(c.k, t * 3 + 1, 8 + t);
}
__global__ void k_cases(const Case* cs, Res* rs, const int8_t* digits, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    run_case(cs[i], rs[i]);
    jac_mul_glv16(rs[i].jglv, cs[i].jp, digits + 66 * cs[i].e);
    G1Affine a; Fp zi; fp_inv(zi, rs[i].jglv.z); jac_to_affine_with_inv(a, rs[i].jglv, zi);
    g1a_compress(rs[i].comp, a);
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }
template <class P> static Fe<P> rnd_fe() {
    Fe<P> a; for (int i = 0; i < P::N; i++) a.v[i] = rnd();
    a.v[P::N - 1] &= (P::N == 12 ? 0x0fffffffu : 0x3fffffffu);  // < modulus
    return a;
}
static G1Jac rnd_point() {
    G1Jac g; for (int i = 0; i < 12; i++) { g.x.v[i] = FpParams::gen_x(i); g.y.v[i] = FpParams::gen_y(i); } fe_set_one(g.z);
    uint32_t k[8]; for (int i = 0; i < 8; i++) k[i] = rnd(); k[7] &= 0x3fffffffu;
    G1Jac r; jac_mul_u256(r, g, k); return r;
}
static G1Affine to_aff(const G1Jac& j) { G1Affine a; Fp zi; fp_inv(zi, j.z); jac_to_affine_with_inv(a, j, zi); return a; }

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d,\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    const int SMS = prop.multiProcessorCount;
    uint32_t* d_out; CK(cudaMalloc(&d_out, (size_t)SMS * 16 * 1024 * 4));
    const int iters = 4096;
    // (1)
    struct { const char* name; double gops; } rows[16]; int nrows = 0;
    auto bench = [&](const char* name, auto kern, int ilp, int ops_per_iter_per_ilp, int threads, int ctas_per_sm) {
        int grid = SMS * ctas_per_sm;
        double ms = time_ms([&] { kern<<<grid, threads>>>(d_out, 0x9e3779b9u, 0x7f4a7c15u, iters); });
        double ops = (double)grid * threads * iters * ilp * ops_per_iter_per_ilp;
        rows[nrows++] = {name, ops / ms / 1e6};
        printf(" \"%s\": %.1f,\n", name, ops / ms / 1e6);
    };
    bench("imad_lo_gops_ilp8_1024thr", k_imad_lo<8>, 8, 1, 256, 4);
    bench("imad_hi_gops_ilp8_1024thr", k_imad_hi<8>, 8, 1, 256, 4);
    bench("imad_wide_gops_ilp8_1024thr", k_imad_wide<8>, 8, 1, 256, 4);
    bench("imad_wide_gops_ilp8_256thr", k_imad_wide<8>, 8, 1, 256, 1);
    bench("imad_chain_wideops_gops_ilp4_1024thr", k_imad_chain<4>, 4, 4, 256, 4);
    bench("imad_chain_wideops_gops_ilp2_512thr", k_imad_chain<2>, 2, 4, 256, 2);
    bench("imad_chain_wideops_gops_ilp1_256thr", k_imad_chain<1>, 1, 4, 256, 1);
    // (2)
    {
        int threads = 128;
        for (int cps : {2, 4, 8}) {
            int grid = SMS * cps; size_t n = (size_t)grid * threads;
            std::vector<Fp> h(n); for (auto& x : h) x = rnd_fe<FpParams>();
            Fp* d; CK(cudaMalloc(&d, n * sizeof(Fp))); CK(cudaMemcpy(d, h.data(), n * sizeof(Fp), cudaMemcpyHostToDevice));
            double ms1 = time_ms([&] { k_fe_mul<FpParams, 1><<<grid, threads>>>(d, 2000); });
            double ms2 = time_ms([&] { k_fe_mul<FpParams, 2><<<grid, threads>>>(d, 2000); });
            printf(" \"fp_mul_gps_ilp1_%dthr_per_sm\": %.2f, \"fp_mul_gps_ilp2_%dthr_per_sm\": %.2f,\n", cps * threads, n * 2000.0 / ms1 / 1e6, cps * threads, n * 2.0 * 2000.0 / ms2 / 1e6);
            cudaFree(d);
        }
        int grid = SMS * 8; size_t n = (size_t)grid * threads;
        std::vector<Fr> h(n); for (auto& x : h) x = rnd_fe<FrParams>();
        Fr* d; CK(cudaMalloc(&d, n * sizeof(Fr))); CK(cudaMemcpy(d, h.data(), n * sizeof(Fr), cudaMemcpyHostToDevice));
        double ms = time_ms([&] { k_fe_mul<FrParams, 2><<<grid, threads>>>(d, 2000); });
        printf(" \"fr_mul_gps_ilp2_1024thr_per_sm\": %.2f,\n", n * 2.0 * 2000.0 / ms / 1e6);
        cudaFree(d);
    }
    // (3)
    const int NC = 24;
    std::vector<Case> cs(NC); std::vector<Res> host(NC), devr(NC);
    for (int i = 0; i < NC; i++) {
        Case& c = cs[i];
        c.a = rnd_fe<FpParams>(); c.b = rnd_fe<FpParams>(); c.c = rnd_fe<FrParams>(); c.d = rnd_fe<FrParams>();
        if (i == 0) { fe_set_zero(c.a); fe_set_zero(c.c); }
        if (i == 1) { for (int l = 0; l < 12; l++) c.a.v[l] = c.b.v[l] = FpParams::mod(l); c.a.v[0] -= 1; c.b.v[0] -= 1; }
        c.jp = rnd_point(); c.jq = rnd_point(); c.p = to_aff(c.jp); c.q = to_aff(c.jq);
        if (i == 2) { c.jq = c.jp; c.q = c.p; }
        jac_dbl(c.jq, c.jq);  // Z != 1
        c.e = (i * 11 + 1) & 127;
        for (int l = 0; l < 8; l++) c.k[l] = rnd();
        run_case(c, host[i]);
        jac_mul_glv16(host[i].jglv, c.jp, GLV_TWIDDLE_DIGITS_HOST[c.e]);
        G1Affine a = to_aff(host[i].jglv); g1a_compress(host[i].comp, a);
    }
    Case* dc; Res* dr; int8_t* dd;
    CK(cudaMalloc(&dc, NC * sizeof(Case))); CK(cudaMalloc(&dr, NC * sizeof(Res))); CK(cudaMalloc(&dd, 128 * 66));
    CK(cudaMemcpy(dc, cs.data(), NC * sizeof(Case), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dd, GLV_TWIDDLE_DIGITS_HOST, 128 * 66, cudaMemcpyHostToDevice));
    CK(cudaMemset(dr, 0, NC * sizeof(Res)));
    k_cases<<<1, 32>>>(dc, dr, dd, NC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(devr.data(), dr, NC * sizeof(Res), cudaMemcpyDeviceToHost));
    int bad = 0;
#define CMP(field) do { if (memcmp(&host[i].field, &devr[i].field, sizeof(host[i].field))) { bad++; printf(" \"mismatch_%d_" #field "\": 1,\n", i); } } while (0)
    for (int i = 0; i < NC; i++) {
        CMP(mul); CMP(add); CMP(sub); CMP(neg); CMP(rmul); CMP(radd); CMP(rsub); CMP(madd_p); CMP(madd_n); CMP(xadd); CMP(xdbl);
        CMP(jadd); CMP(jdbl); CMP(jmadd); CMP(jglv); CMP(jfromx); CMP(booth); CMP(comp);
    }
    printf(" \"crosscheck_cases\": %d, \"crosscheck_mismatches\": %d}\n", NC, bad);
    return bad ? 1 : 0;
}
