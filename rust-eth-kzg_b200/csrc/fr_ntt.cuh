// Shared-memory radix-2 NTTs over Fr (device only).
//
// Replaces the reference's generic radix-2 transform (crates/cryptography/polynomial/src/fft.rs:46-177)
// and its explicit bit-reversal passes (fk20/cosets.rs:56-78): a decimation-in-time pass consumes
// bit-reversed input and a decimation-in-frequency pass produces bit-reversed output, so every
// `reverse_bit_order` of the reference call stack is absorbed into the choice of
// pass and never touches memory.
//
// Elements live in shared memory limb-major (limb l of element i at s[l*stride + i]) so that a warp
// touching 32 consecutive elements is bank-conflict free for every butterfly span >= 32.
#pragma once
#include "fr_consts.cuh"

namespace ekzg {

// 16-byte vector loads / stores of limb structs.  The LOCAL object is only ever touched through its own members
// (limb_word(obj, k): word k of a field element or point, resolved at compile time in the unrolled loops): no reinterpret_cast of a
// local.  Writing it through a uint4* and reading the limbs afterwards was a strict-aliasing violation, and nvcc does merge stack
// slots it believes unrelated (seen in the radix-4 G1-NTT combination unit, g1_ntt_units.cuh).
template <class P>
__host__ __device__ __forceinline__ uint32_t& limb_word(Fe<P>& a, int k) { return a.v[k]; }
template <class P>
__host__ __device__ __forceinline__ const uint32_t& limb_word(const Fe<P>& a, int k) { return a.v[k]; }

template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
        const uint4 q = s[i];
        limb_word(r, 4 * i) = q.x; limb_word(r, 4 * i + 1) = q.y; limb_word(r, 4 * i + 2) = q.z; limb_word(r, 4 * i + 3) = q.w;
    }
    return r;
}
template <class T>
__device__ __forceinline__ void st_vec(T* p, const T& v) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++)
        d[i] = make_uint4(limb_word(v, 4 * i), limb_word(v, 4 * i + 1), limb_word(v, 4 * i + 2), limb_word(v, 4 * i + 3));
}

__device__ __forceinline__ Fr smem_ld(const uint32_t* s, int stride, int i) {
    Fr r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = s[l * stride + i];
    return r;
}
__device__ __forceinline__ void smem_st(uint32_t* s, int stride, int i, const Fr& a) {
#pragma unroll
    for (int l = 0; l < 8; l++) s[l * stride + i] = a.v[l];
}

// `nbatch` independent transforms of size 2^LOGN stored back to back (transform q at element q<<LOGN).
// tw[i] = w^i (Montgomery), i < 2^(LOGN-1), w the primitive 2^LOGN-th root for this direction.
// DIT: input bit-reversed, output natural.  Ends with __syncthreads().
template <int LOGN>
__device__ void ntt_dit_shared(uint32_t* s, int stride, int nbatch, const Fr* __restrict__ tw, int tid, int nth) {
    constexpr int N = 1 << LOGN;
    for (int st = 0; st < LOGN; st++) {
        const int len = 1 << st;
        for (int t = tid; t < nbatch * (N / 2); t += nth) {
            int q = t >> (LOGN - 1), tt = t & (N / 2 - 1);
            int pos = tt & (len - 1);
            int i = (q << LOGN) + ((tt >> st) << (st + 1)) + pos, j = i + len;
            Fr a = smem_ld(s, stride, i), b = smem_ld(s, stride, j);
            if (pos) {
                Fr w = ld_vec(&tw[pos << (LOGN - 1 - st)]);
                fe_mul(b, b, w);
            }
            Fr u, v;
            fe_add(u, a, b);
            fe_sub(v, a, b);
            smem_st(s, stride, i, u);
            smem_st(s, stride, j, v);
        }
        __syncthreads();
    }
}

// DIF: input natural, output bit-reversed.  Ends with __syncthreads().
template <int LOGN>
__device__ void ntt_dif_shared(uint32_t* s, int stride, int nbatch, const Fr* __restrict__ tw, int tid, int nth) {
    constexpr int N = 1 << LOGN;
    for (int st = LOGN - 1; st >= 0; st--) {
        const int len = 1 << st;
        for (int t = tid; t < nbatch * (N / 2); t += nth) {
            int q = t >> (LOGN - 1), tt = t & (N / 2 - 1);
            int pos = tt & (len - 1);
            int i = (q << LOGN) + ((tt >> st) << (st + 1)) + pos, j = i + len;
            Fr a = smem_ld(s, stride, i), b = smem_ld(s, stride, j);
            Fr u, v;
            fe_add(u, a, b);
            fe_sub(v, a, b);
            if (pos) {
                Fr w = ld_vec(&tw[pos << (LOGN - 1 - st)]);
                fe_mul(v, v, w);
            }
            smem_st(s, stride, i, u);
            smem_st(s, stride, j, v);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// 32 big-endian bytes (16-byte aligned) -> plain little-endian limbs
__device__ __forceinline__ Fr fr_load_be(const uint8_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 hi = q[0], lo = q[1];
    Fr r;
    r.v[7] = bswap32(hi.x); r.v[6] = bswap32(hi.y); r.v[5] = bswap32(hi.z); r.v[4] = bswap32(hi.w);
    r.v[3] = bswap32(lo.x); r.v[2] = bswap32(lo.y); r.v[1] = bswap32(lo.z); r.v[0] = bswap32(lo.w);
    return r;
}
__device__ __forceinline__ void fr_store_be(uint8_t* p, const Fr& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(bswap32(a.v[7]), bswap32(a.v[6]), bswap32(a.v[5]), bswap32(a.v[4]));
    q[1] = make_uint4(bswap32(a.v[3]), bswap32(a.v[2]), bswap32(a.v[1]), bswap32(a.v[0]));
}

}  // namespace ekzg
