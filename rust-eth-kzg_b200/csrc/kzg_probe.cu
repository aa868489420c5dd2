// Live measurement of the roofline denominator of the point-arithmetic kernels: the issue rate of carry-chained
// IMAD.WIDE.U32.X (ncu: sm__pipe_fmaheavy) on the device the context lives on.  bench.py runs it in the same process,
// on the same clocks, right before the timed region, instead of quoting a constant (tools/gpu_probe.cu is the
// standalone form of the same loop; profiles/r1_v4_gpu_probe.json).
#include <cuda_runtime.h>
#include <stdint.h>

namespace ekzg {

// four independent lo/hi carry chains of length 8 per thread: ptxas fuses each lo/hi pair into one IMAD.WIDE.U32.X
__global__ void __launch_bounds__(256, 4) k_probe_imad_chain(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) x[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; it++) {
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
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) s ^= x[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// wide multiply-adds per second on the current device (best of `reps` runs of ~10 ms), 0 on error
double probe_imad_wide_per_s(int reps) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    const int grid = sms * 4, threads = 256, iters = 32768;
    uint32_t* d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)grid * threads * 4) != cudaSuccess) return 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int r = 0; r < reps + 1; r++) {   // first run is the warm-up
        cudaEventRecord(e0, 0);
        k_probe_imad_chain<<<grid, threads>>>(d_out, 0x9e3779b9u, 0x7f4a7c15u, iters);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = 0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * threads * iters * 4 /*chains*/ * 4 /*wide ops per chain*/;
        if (r > 0 && ms > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

}  // namespace ekzg
