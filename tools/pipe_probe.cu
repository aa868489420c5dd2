// Pipe-concurrency probe for sm_100a: which multiply pipes exist next to the integer fmaheavy pipe and do they overlap?
// Measures, per SM and clock: (A) carry-chained IMAD.WIDE.U32.X, (B) plain IMAD.WIDE.U32 with distinct operands
// (the gpu_probe "imad_wide" kernel let ptxas fold the product away, so its number was an IADD3 rate),
// (C) DFMA, (D) IADD3, and the co-run of A with C and of C with D from different warps of the same CTA.
// Build: make -C rust-eth-kzg_b200 probes
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ void work_imadx(uint32_t (&x)[4][8], uint32_t a, uint32_t b) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        asm volatile(
            "mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %9, %2;\n\t madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
            "madc.lo.cc.u32 %4, %8, %9, %4;\n\t madc.hi.cc.u32 %5, %8, %9, %5;\n\t"
            "madc.lo.cc.u32 %6, %8, %9, %6;\n\t madc.hi.u32 %7, %8, %9, %7;"
            : "+r"(x[i][0]), "+r"(x[i][1]), "+r"(x[i][2]), "+r"(x[i][3]), "+r"(x[i][4]), "+r"(x[i][5]), "+r"(x[i][6]), "+r"(x[i][7])
            : "r"(a), "r"(b));
    }
}
// 16 wide multiply-adds per call
__device__ __forceinline__ void work_imadw(uint64_t (&y)[8]) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = (uint64_t)(uint32_t)y[i] * (uint32_t)(y[i] >> 32) + y[(i + 1) & 7];
}
// 16 DFMA per call
__device__ __forceinline__ void work_dfma(double (&d)[8], double a, double b) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = __fma_rz(d[i], a, b);
}
// 16 IADD3 per call
__device__ __forceinline__ void work_iadd(uint32_t (&z)[8], uint32_t a) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(a ^ i));
}

// mode bits per warp parity: what even warps do / what odd warps do.  0 = nothing, 1 = imadx, 2 = imadw, 3 = dfma, 4 = iadd,
// 5 = dfma + iadd interleaved in one warp, 6 = imadx + dfma interleaved in one warp
__global__ void __launch_bounds__(1024, 1) k_probe(uint32_t* out, int mode_even, int mode_odd, int iters, uint32_t a, uint32_t b, double da, double db) {
    int mode = ((threadIdx.x >> 7) & 1) ? mode_odd : mode_even;   // warps 0-3 / 4-7 / ...: every SM sub-partition gets both kinds
    uint32_t x[4][8]; uint64_t y[8]; double d[8]; uint32_t z[8];
    for (int i = 0; i < 8; i++) { y[i] = threadIdx.x * 0x9e3779b97f4a7c15ull + i; d[i] = 1.0 + threadIdx.x * 1e-9 + i; z[i] = threadIdx.x + i; for (int j = 0; j < 4; j++) x[j][i] = threadIdx.x + i + j; }
    if (mode == 1) for (int it = 0; it < iters; it++) work_imadx(x, a, b);          // 32 wide ops / iter
    else if (mode == 2) for (int it = 0; it < iters; it++) { work_imadw(y); work_imadw(y); }   // 32
    else if (mode == 3) for (int it = 0; it < iters; it++) { work_dfma(d, da, db); work_dfma(d, da, db); }  // 32
    else if (mode == 4) for (int it = 0; it < iters; it++) { work_iadd(z, a); work_iadd(z, a); }  // 32
    else if (mode == 5) for (int it = 0; it < iters; it++) { work_dfma(d, da, db); work_iadd(z, a); work_dfma(d, da, db); work_iadd(z, a); }  // 32 + 32
    else if (mode == 6) for (int it = 0; it < iters; it++) { work_imadx(x, a, b); work_dfma(d, da, db); work_dfma(d, da, db); }  // 32 + 32
    else if (mode == 7) for (int it = 0; it < iters; it++) { work_imadx(x, a, b); work_iadd(z, a); work_iadd(z, a); }  // 32 + 32
    else if (mode == 8) for (int it = 0; it < iters; it++) { work_imadx(x, a, b); work_iadd(z, a); }  // 32 + 16
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) { s ^= (uint32_t)y[i] ^ (uint32_t)(y[i] >> 32) ^ (uint32_t)__double_as_longlong(d[i]) ^ z[i]; for (int j = 0; j < 4; j++) s ^= x[j][i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double run(uint32_t* out, int me, int mo, int threads, int iters) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms = 148;
    k_probe<<<sms, threads>>>(out, me, mo, iters, 0x12345677u, 0x9abcdef1u, 1.0000001, 1e-9);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        k_probe<<<sms, threads>>>(out, me, mo, iters, 0x12345677u, 0x9abcdef1u, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    uint32_t* out; CK(cudaMalloc(&out, 148 * 1024 * 4));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const int iters = 20000;
    const char* names[] = {"none", "imad_wide_x", "imad_wide", "dfma", "iadd", "dfma+iadd(same warp)", "imad_wide_x+dfma(same warp)", "imad_wide_x+iadd(same warp)", "imad_wide_x+iadd/2(same warp)"};
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    fprintf(f, "{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"note\": \"ops = 32 per iteration per thread per active mode; lanes/clk/SM at the nominal clock\", \"runs\": [\n", pr.name, pr.multiProcessorCount, clk / 1000);
    struct Cfg { int me, mo, threads; } cfgs[] = {
        {1, 1, 512}, {1, 1, 256}, {2, 2, 512}, {3, 3, 512}, {3, 3, 256}, {4, 4, 512},
        {1, 0, 512}, {3, 0, 512}, {4, 0, 512}, {1, 3, 512}, {1, 3, 1024}, {4, 3, 512}, {1, 4, 512}, {5, 5, 512}, {5, 5, 256}, {6, 6, 512}, {7, 7, 512}, {8, 8, 512}, {8, 8, 256},
    };
    int n = sizeof(cfgs) / sizeof(cfgs[0]);
    for (int c = 0; c < n; c++) {
        double ms = run(out, cfgs[c].me, cfgs[c].mo, cfgs[c].threads, iters);
        // threads running each mode
        double thr_e = cfgs[c].threads / 2.0, thr_o = cfgs[c].threads / 2.0;
        auto ops = [&](int m) { return m == 0 ? 0.0 : (m == 8 ? 48.0 : (m >= 5 ? 64.0 : 32.0)); };
        double tot_e = thr_e * ops(cfgs[c].me) * iters * 148, tot_o = thr_o * ops(cfgs[c].mo) * iters * 148;
        double lanes = (tot_e + tot_o) / (ms * 1e-3) / 148 / (clk * 1e3);
        fprintf(f, " {\"even\": \"%s\", \"odd\": \"%s\", \"threads\": %d, \"ms\": %.3f, \"gops\": %.1f, \"ops_per_clk_per_sm\": %.1f}%s\n", names[cfgs[c].me], names[cfgs[c].mo],
                cfgs[c].threads, ms, (tot_e + tot_o) / (ms * 1e-3) / 1e9, lanes, c + 1 < n ? "," : "");
    }
    fprintf(f, "]}\n");
    if (f != stdout) fclose(f);
    return 0;
}
