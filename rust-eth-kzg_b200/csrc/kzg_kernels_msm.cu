// K7  variable-base G1 multi-scalar multiplication by the bucket method
//   reference: g1_lincomb -> blst's Pippenger (crates/cryptography/bls12_381/src/lincomb.rs:7-30), as the verifier uses it for its
//   two random-linear-combination sums  sum rho_k pi_k  and  sum rho_k h_k^64 pi_k  (kzg_multi_open/src/fk20/verifier.rs:186-201).
//
//   sum_k s_k P_k  with  s_k = k1 + k2*lambda (GLV, both halves below 2^128)  =  sum over 13 windows of 10 bits of
//   2^(10 w) * sum_b b * B_(w,b),   B_(w,b) = sum of the +-P_k / +-phi(P_k) whose signed digit in window w is +-b   (b = 1 .. 512).
//   Both GLV halves share the buckets of a window (same weight), so there are 13 x 512 buckets holding 2 N points.
//     k_msm_digits      one thread per scalar: GLV split, 2 x 13 signed digits (int16, window-major)
//     k_msm_accumulate  one LANE per bucket: a warp owns 32 buckets of one window, scans that window's digits (32 at a time, handed
//                       round by shuffles), every lane collects the entries of its bucket in a shared-memory list, then the lanes
//                       add their lists in lock step (XYZZ mixed additions; x * beta for the phi half)
//     k_msm_reduce      one CTA per window: sum_b b * B_b.  Thread t folds its four buckets with running sums, the 128 partial
//                       pairs are combined by a block-wide suffix scan and two tree reductions through shared memory
//     k_msm_combine     2^(10 w) * R_w by doublings, one window per lane, then a tree sum
//   Cost at N = 16 384: 0.43 M mixed additions (against 32 768 ladders of ~2200 multiplications); the two dependent tails -- ~30
//   XYZZ additions in k_msm_reduce, 120 doublings in k_msm_combine -- are what bounds its latency.
#include "kzg_kernels.h"
#include "fr_ntt.cuh"
#include "g1_mul.cuh"

namespace ekzg {

constexpr int MSM_C = 10;                        // window bits
constexpr int MSM_W = 13;                        // windows of a 128-bit half (130 bits: the top one takes the carry)
constexpr int MSM_BUCKETS = 1 << (MSM_C - 1);    // |digit| in 1 .. 512
constexpr int MSM_LIST = 160;                    // list slots of a bucket before a flush (mean 2 N / 512 = 64 at N = 16 384)

size_t msm_bucket_scratch_bytes(int n, int sets) {
    return (size_t)sets * ((size_t)2 * MSM_W * n * sizeof(int16_t) + 256) + (size_t)sets * MSM_W * MSM_BUCKETS * sizeof(G1Xyzz) +
           (size_t)sets * MSM_W * sizeof(G1Jac) + 1024;
}

// digits[((set * 2 + half) * MSM_W + w) * n + k]
__global__ void __launch_bounds__(128)
k_msm_digits(const uint32_t* __restrict__ scalars0, const uint32_t* __restrict__ scalars1, int16_t* __restrict__ digits, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, set = blockIdx.y;
    if (k >= n) return;
    const uint32_t* sc = (set ? scalars1 : scalars0) + (size_t)k * 8;
    uint32_t s[8];
    for (int i = 0; i < 8; i++) s[i] = sc[i];
    uint32_t half[2][5];
    glv_split_halves(half[0], half[1], s);
    for (int h = 0; h < 2; h++) {
        int carry = 0;
        for (int w = 0; w < MSM_W; w++) {
            const int bit = w * MSM_C, word = bit >> 5, sh = bit & 31;
            uint32_t v = half[h][word] >> sh;
            if (sh > 32 - MSM_C && word + 1 < 5) v |= half[h][word + 1] << (32 - sh);
            int x = (int)(v & ((1u << MSM_C) - 1u)) + carry;
            carry = x > MSM_BUCKETS;
            if (carry) x -= 1 << MSM_C;
            digits[((size_t)(set * 2 + h) * MSM_W + w) * n + k] = (int16_t)x;
        }
    }
}

// buckets[(set * MSM_W + w) * MSM_BUCKETS + b - 1]
__global__ void __launch_bounds__(128)
k_msm_accumulate(const G1Affine* __restrict__ pts, const int16_t* __restrict__ digits, G1Xyzz* __restrict__ buckets, int n) {
    extern __shared__ uint32_t lists[];          // [warp][slot][lane]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.y, set = blockIdx.z;
    const int mine = (blockIdx.x * 4 + warp) * 32 + lane + 1;      // this lane's bucket: |digit| == mine
    uint32_t* my = lists + (size_t)warp * MSM_LIST * 32 + lane;
    int cnt = 0;
    G1Xyzz acc;
    xyzz_set_inf(acc);
    Fp beta;
#pragma unroll
    for (int j = 0; j < 12; j++) beta.v[j] = FpParams::beta(j);
    auto flush = [&]() {
        const int longest = __reduce_max_sync(0xffffffffu, cnt);
        for (int e = 0; e < longest; e++) {
            if (e < cnt) {
                const uint32_t ent = my[e * 32];
                G1Affine p = ld_vec(&pts[ent & 0x3fffffffu]);
                if (ent & 0x40000000u) fe_mul(p.x, p.x, beta);     // phi(P) = (beta x, y)
                xyzz_madd(acc, p, (ent & 0x80000000u) != 0);
            }
        }
        cnt = 0;
    };
    for (int h = 0; h < 2; h++) {
        const int16_t* dg = digits + ((size_t)(set * 2 + h) * MSM_W + w) * n;
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int d = k0 + lane < n ? (int)dg[k0 + lane] : 0;
            if (__any_sync(0xffffffffu, d != 0)) {
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const int dj = __shfl_sync(0xffffffffu, d, j);
                    const int a = dj < 0 ? -dj : dj;
                    if (a == mine) {
                        my[cnt * 32] = (uint32_t)(k0 + j) | (h ? 0x40000000u : 0u) | (dj < 0 ? 0x80000000u : 0u);
                        cnt++;
                    }
                }
            }
            if (__any_sync(0xffffffffu, cnt > MSM_LIST - 32)) flush();
        }
    }
    flush();
    st_vec(&buckets[((size_t)set * MSM_W + w) * MSM_BUCKETS + mine - 1], acc);
}

// window sums R[set * MSM_W + w] = sum_b b * B_b
__global__ void __launch_bounds__(128)
k_msm_reduce(const G1Xyzz* __restrict__ buckets, G1Jac* __restrict__ R) {
    __shared__ G1Xyzz sa[128], sb[128];
    const int t = threadIdx.x, w = blockIdx.x, set = blockIdx.y;
    const G1Xyzz* B = buckets + ((size_t)set * MSM_W + w) * MSM_BUCKETS + 4 * t;   // buckets 4t+1 .. 4t+4
    // (operands of the non-inlined additions live in ONE array: separate locals have been given one stack slot by nvcc 12.9, see
    // g1_ntt_units.cuh)
    G1Xyzz v[3];                         // run, T, the bucket
    xyzz_set_inf(v[0]);
    xyzz_set_inf(v[1]);
    for (int i = 3; i >= 0; i--) {      // T = sum (i + 1) * B[i], run = sum B[i]
        v[2] = ld_vec(&B[i]);
        xyzz_add(v[0], v[2]);
        xyzz_add(v[1], v[0]);
    }
    // sum_t T_t : tree over sb
    sb[t] = v[1];
    __syncthreads();
    for (int step = 64; step >= 1; step >>= 1) {
        if (t < step) { G1Xyzz x = sb[t]; xyzz_add(x, sb[t + step]); sb[t] = x; }
        __syncthreads();
    }
    if (t == 0) v[1] = sb[0];            // sum of the T_t
    __syncthreads();
    // sum_t 4 t * S_t = 4 * sum_{t >= 1} U_t with the suffix sums U_t = S_t + S_(t+1) + ..   (Hillis-Steele, double-buffered)
    sa[t] = v[0];
    __syncthreads();
    G1Xyzz* cur = sa;
    G1Xyzz* nxt = sb;
    for (int off = 1; off < 128; off <<= 1) {
        G1Xyzz x = cur[t];
        if (t + off < 128) xyzz_add(x, cur[t + off]);
        nxt[t] = x;
        __syncthreads();
        G1Xyzz* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (t == 0) xyzz_set_inf(cur[0]);    // the term t = 0 has weight 0
    __syncthreads();
    for (int step = 64; step >= 1; step >>= 1) {
        if (t < step) { G1Xyzz x = cur[t]; xyzz_add(x, cur[t + step]); cur[t] = x; }
        __syncthreads();
    }
    if (t == 0) {
        v[0] = cur[0];
        xyzz_dbl(v[2], v[0]);
        xyzz_dbl(v[0], v[2]);            // 4 * sum_{t>=1} U_t
        xyzz_add(v[0], v[1]);
        G1Jac r;
        jac_from_xyzz(r, v[0]);
        st_vec(&R[set * MSM_W + w], r);
    }
}

// out[set] = sum_w 2^(MSM_C w) R_w
__global__ void __launch_bounds__(32)
k_msm_combine(const G1Jac* __restrict__ R, G1Jac* __restrict__ out0, G1Jac* __restrict__ out1) {
    __shared__ G1Jac sm[16];
    const int w = threadIdx.x, set = blockIdx.x;
    G1Jac acc;
    jac_set_inf(acc);
    if (w < MSM_W) {
        acc = ld_vec(&R[set * MSM_W + w]);
        for (int i = 0; i < MSM_C * w; i++) jac_dbl(acc, acc);
    }
    for (int step = 8; step >= 1; step >>= 1) {
        if (w >= step && w < 2 * step) sm[w] = acc;
        __syncwarp();
        if (w < step) jac_add(acc, sm[w + step]);
        __syncwarp();
    }
    if (w == 0) st_vec(set ? out1 : out0, acc);
}

// sum_k scalars0[k] * pts[k] -> out0 and (scalars1 != nullptr) sum_k scalars1[k] * pts[k] -> out1; scratch: msm_bucket_scratch_bytes(n, sets)
cudaError_t launch_msm_bucket(const G1Affine* pts, const uint32_t* scalars0, const uint32_t* scalars1, int n, G1Jac* out0, G1Jac* out1,
                              void* scratch, cudaStream_t st) {
    const int sets = scalars1 ? 2 : 1;
    if (n > 0x3fffffff) return cudaErrorInvalidValue;
    uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
    int16_t* digits = reinterpret_cast<int16_t*>(p);
    p += ((size_t)sets * 2 * MSM_W * n * sizeof(int16_t) + 255) / 256 * 256;
    G1Xyzz* buckets = reinterpret_cast<G1Xyzz*>(p);
    p += (size_t)sets * MSM_W * MSM_BUCKETS * sizeof(G1Xyzz);
    G1Jac* R = reinterpret_cast<G1Jac*>(p);
    const size_t smem = (size_t)4 * MSM_LIST * 32 * sizeof(uint32_t);
    {   // per device, so not cached in a static (a DASContext may span several)
        cudaError_t e = cudaFuncSetAttribute(k_msm_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_msm_digits<<<dim3((n + 127) / 128, sets), 128, 0, st>>>(scalars0, scalars1, digits, n);
    EKZG_LAUNCH_CHECK();
    k_msm_accumulate<<<dim3(MSM_BUCKETS / 128, MSM_W, sets), 128, smem, st>>>(pts, digits, buckets, n);
    EKZG_LAUNCH_CHECK();
    k_msm_reduce<<<dim3(MSM_W, sets), 128, 0, st>>>(buckets, R);
    EKZG_LAUNCH_CHECK();
    k_msm_combine<<<sets, 32, 0, st>>>(R, out0, out1);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace ekzg
