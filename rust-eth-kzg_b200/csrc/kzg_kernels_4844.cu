// sm_100a kernels of the EIP-4844 prover path (blob_to_kzg_commitment, compute_kzg_proof,
// compute_blob_kzg_proof) and point validation shared with the verifiers.
//   reference: crates/eip4844/src/prover.rs:17-88, crates/eip4844/src/verifier.rs:146-196,
//              crates/cryptography/kzg_single_open/src/prover.rs:33-65, crates/serialization/src/lib.rs:69-99.
// The two MSMs (4096 coefficients x monomial SRS) reuse the fixed-base kernel of the FK20 path: the SRS is
// fixed, so commitment and proof are table additions only (the reference calls blst's Pippenger per blob).
#include "kzg_kernels.h"
#include "fr_ntt.cuh"
#include "sha256.cuh"

namespace ekzg {

// 32 big-endian bytes -> Fr (Montgomery), value reduced mod r  (bls12_381/src/lib.rs:128-140 reduce_bytes_to_scalar_bias)
__device__ __forceinline__ Fr fr_from_be_reduce(const uint8_t* p) {
    Fr x;
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const uint8_t* q = p + 4 * (7 - l);
        x.v[l] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    // 2^256 < 3r: two conditional subtractions bring x below r (the Montgomery product needs an operand < r,
    // otherwise its running sum can exceed 8 limbs)
    fe_final_sub<FrParams>(x.v);
    fe_final_sub<FrParams>(x.v);
    fe_to_mont(x, x);
    return x;
}

// z = SHA-256("FSBLOBVERIFY_V1_" || u128_be(4096) || blob || commitment) mod r, one thread per blob
// (crates/eip4844/src/verifier.rs:155-196).  The chain is sequential per blob, so the kernel is latency-bound: the 64
// bytes of the next compression are fetched (four 16-byte loads) before the current one is computed.  The 32-byte
// header shifts the blob by half a block: compression k takes blob bytes [64k-32, 64k+32).
__device__ __forceinline__ void sha_words_from_u4(uint32_t* w, const uint4& v) {
    w[0] = __byte_perm(v.x, 0, 0x0123); w[1] = __byte_perm(v.y, 0, 0x0123); w[2] = __byte_perm(v.z, 0, 0x0123); w[3] = __byte_perm(v.w, 0, 0x0123);
}
__global__ void __launch_bounds__(32)
k_blob_challenge(const uint8_t* __restrict__ blobs, const uint8_t* __restrict__ commitments, Fr* __restrict__ z_out, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint4* src = reinterpret_cast<const uint4*>(blobs + (size_t)b * BYTES_PER_BLOB);   // 8192 x 16 B
    const uint4* com = reinterpret_cast<const uint4*>(commitments + (size_t)b * 48);
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    uint32_t w[16];
    // "FSBL" "OBVE" "RIFY" "_V1_" then u128_be(4096)
    w[0] = 0x4653424cu; w[1] = 0x4f425645u; w[2] = 0x52494659u; w[3] = 0x5f56315fu;
    w[4] = 0; w[5] = 0; w[6] = 0; w[7] = 0x00001000u;
    uint4 nxt[4];
    nxt[0] = src[0]; nxt[1] = src[1];
    sha_words_from_u4(w + 8, nxt[0]);
    sha_words_from_u4(w + 12, nxt[1]);
#pragma unroll
    for (int q = 0; q < 4; q++) nxt[q] = src[2 + q];
    sha256_compress_words(h, w);
#pragma unroll 1
    for (int k = 1; k < 2048; k++) {
#pragma unroll
        for (int q = 0; q < 4; q++) sha_words_from_u4(w + 4 * q, nxt[q]);
        if (k < 2047) {
#pragma unroll
            for (int q = 0; q < 4; q++) nxt[q] = src[4 * (k + 1) - 2 + q];
        } else {  // compression 2048: last 32 blob bytes + first 32 commitment bytes
            nxt[0] = src[8190]; nxt[1] = src[8191]; nxt[2] = com[0]; nxt[3] = com[1];
        }
        sha256_compress_words(h, w);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) sha_words_from_u4(w + 4 * q, nxt[q]);
    sha256_compress_words(h, w);
    // tail: last 16 commitment bytes, 0x80, zeros, bit length of 32 + 131072 + 48 bytes
    sha_words_from_u4(w, com[2]);
    w[4] = 0x80000000u;
#pragma unroll
    for (int q = 5; q < 15; q++) w[q] = 0;
    w[15] = (32u + (uint32_t)BYTES_PER_BLOB + 48u) * 8u;
    sha256_compress_words(h, w);
    uint8_t hb[32];
#pragma unroll
    for (int i = 0; i < 8; i++) { hb[4 * i] = (uint8_t)(h[i] >> 24); hb[4 * i + 1] = (uint8_t)(h[i] >> 16); hb[4 * i + 2] = (uint8_t)(h[i] >> 8); hb[4 * i + 3] = (uint8_t)h[i]; }
    st_vec(&z_out[b], fr_from_be_reduce(hb));
}

// user-supplied evaluation points (compute_kzg_proof): 32 BE bytes -> Fr, status |= 2 if not canonical
__global__ void k_scalars_from_be(const uint8_t* __restrict__ in, Fr* __restrict__ out, uint32_t* __restrict__ status, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[32];
    for (int c = 0; c < 32; c++) buf[c] = in[(size_t)i * 32 + c];
    Fr x;
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const uint8_t* q = buf + 4 * (7 - l);
        x.v[l] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    if (status && fe_plain_ge_mod(x)) atomicOr(&status[i], 2u);   // (status == nullptr: values known to be canonical)
    fe_to_mont(x, x);
    st_vec(&out[i], x);
}

// Quotient by (X - z) with Ruffini's rule (kzg_single_open/src/prover.rs:48-65):
// t_i = c_i + z*t_{i+1};  q_{i-1} = t_i (i = 4095..1);  y = t_0.  The reference runs the recurrence serially; here it
// is a weighted suffix scan, one CTA of 256 threads per blob: thread s owns coefficients 16s..16s+15, computes the
// local suffix sums l_j, the 256 segment heads are combined by a Hillis-Steele scan with multipliers z^(16*2^k), and
// t_i = l_i + z^(16(s+1)-i) * (value entering the segment from above).  q is written as plain integers in the MSM
// scalar layout [i][B]; q_4095 = 0 so the 4096-point MSM is the reference's 4095-point one.
constexpr int QT_THREADS = 256, QT_SEG = N_BLOB / QT_THREADS;
__global__ void __launch_bounds__(QT_THREADS)
k_quotient(const Fr* __restrict__ coeffs, const Fr* __restrict__ z_in, uint32_t* __restrict__ scalars, uint8_t* __restrict__ y_out, int B) {
    __shared__ Fr T[QT_THREADS];
    const int b = blockIdx.x, s = threadIdx.x, base = QT_SEG * s;
    const Fr* c = coeffs + (size_t)b * N_BLOB;
    const Fr z = ld_vec(&z_in[b]);
    Fr l[QT_SEG];
    Fr acc;
    fe_set_zero(acc);
#pragma unroll
    for (int j = QT_SEG - 1; j >= 0; j--) {
        Fr cj = ld_vec(&c[base + j]);
        fe_mul(acc, acc, z);
        fe_add(acc, acc, cj);
        l[j] = acc;
    }
    Fr mult = z;   // z^16
#pragma unroll
    for (int q = 0; q < 4; q++) fe_sqr(mult, mult);
    static_assert(QT_SEG == 16, "z^SEG by four squarings");
    T[s] = l[0];
    __syncthreads();
    for (int d = 1; d < QT_THREADS; d <<= 1) {
        Fr v = T[s];
        if (s + d < QT_THREADS) {
            Fr o = T[s + d];
            fe_mul(o, o, mult);
            fe_add(v, v, o);
        }
        __syncthreads();
        T[s] = v;
        __syncthreads();
        fe_sqr(mult, mult);
    }
    Fr tin;
    if (s + 1 < QT_THREADS) tin = T[s + 1]; else fe_set_zero(tin);
    Fr zpow = z;
#pragma unroll
    for (int j = QT_SEG - 1; j >= 0; j--) {
        Fr t;
        fe_mul(t, zpow, tin);
        fe_add(t, t, l[j]);
        if (j) fe_mul(zpow, zpow, z);
        const int i = base + j;
        Fr p;
        fe_from_mont(p, t);
        if (i >= 1) {
            uint4* d4 = reinterpret_cast<uint4*>(scalars + ((size_t)(i - 1) * B + b) * 8);
            d4[0] = make_uint4(p.v[0], p.v[1], p.v[2], p.v[3]);
            d4[1] = make_uint4(p.v[4], p.v[5], p.v[6], p.v[7]);
        } else if (y_out) {
            fr_store_be(y_out + (size_t)b * 32, p);
        }
    }
    if (s == QT_THREADS - 1) {
        uint4* d4 = reinterpret_cast<uint4*>(scalars + ((size_t)(N_BLOB - 1) * B + b) * 8);
        d4[0] = make_uint4(0, 0, 0, 0);
        d4[1] = make_uint4(0, 0, 0, 0);
    }
}

// DIRECT cell proofs for a blob or two (the latency path of compute_cells_and_kzg_proofs):  proof k of a blob is the commitment to
//   q_k(X) = f(X) div (X^64 - c_k),   c_k = omega_128^rev7(k) = h_k^64 for the coset h_k <omega_64> of cell k
// (kzg_multi_open/src/fk20/prover.rs:173-228 computes the same 128 commitments through the Toeplitz / G1-FFT route, whose 14-phase
// dependency chain costs ~10 ms however few blobs there are; 128 independent 4096-point MSMs over the fixed SRS tables are
// throughput work -- 10.5 M table additions, ~4 ms -- and win for one or two blobs).  The quotient by X^64 - c is 64 independent
// Horner recurrences of 63 steps: q[r + 64 m] = f[r + 64 (m+1)] + c q[r + 64 (m+1)].
// Thread (r, vb): residue r < 64, virtual blob vb = blob * 128 + k; scalars [4096][Bv] plain, the layout k_coeffs_to_scalars writes.
__global__ void __launch_bounds__(128)
k_coset_quotients(const Fr* __restrict__ coeffs, uint32_t* __restrict__ scalars, const Fr* __restrict__ tw128, int Bv) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)64 * Bv) return;
    const int r = (int)(gid / Bv), vb = (int)(gid % Bv), blob = vb >> 7, k = vb & 127;
    int e = 0;
    for (int bit = 0; bit < 7; bit++) e |= ((k >> bit) & 1) << (6 - bit);
    Fr c = ld_vec(&tw128[e & 63]);
    if (e >= 64) fe_neg(c, c);
    const Fr* f = coeffs + (size_t)blob * N_BLOB;
    Fr q;
    fe_set_zero(q);
    {   // block 63 of the quotient is zero
        uint4* d4 = reinterpret_cast<uint4*>(scalars + ((size_t)(r + 64 * 63) * Bv + vb) * 8);
        d4[0] = make_uint4(0, 0, 0, 0);
        d4[1] = make_uint4(0, 0, 0, 0);
    }
    for (int m = 62; m >= 0; m--) {
        const Fr fi = ld_vec(&f[r + 64 * (m + 1)]);
        fe_mul(q, q, c);
        fe_add(q, q, fi);
        Fr v;
        fe_from_mont(v, q);
        uint4* d4 = reinterpret_cast<uint4*>(scalars + ((size_t)(r + 64 * m) * Bv + vb) * 8);
        d4[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
        d4[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    }
}

// coefficients [B][4096] (Montgomery) -> plain MSM scalars [4096][B]
__global__ void k_coeffs_to_scalars(const Fr* __restrict__ coeffs, uint32_t* __restrict__ scalars, int B) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)N_BLOB * B) return;
    int i = (int)(gid / B), b = (int)(gid % B);
    Fr v = ld_vec(&coeffs[(size_t)b * N_BLOB + i]);
    fe_from_mont(v, v);
    uint4* d4 = reinterpret_cast<uint4*>(scalars + gid * 8);
    d4[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    d4[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

// pts[0][b] = sum_{i < count} pts[stride*i][b]   (the 64 partial sums the fixed-base kernel leaves at even positions)
__global__ void __launch_bounds__(64)
k_sum_positions(G1Jac* __restrict__ pts, int B, int count, int stride) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    G1Jac acc = ld_vec(&pts[b]);
    for (int i = 1; i < count; i++) {
        G1Jac q = ld_vec(&pts[(size_t)stride * i * B + b]);
        jac_add(acc, q);
    }
    st_vec(&pts[b], acc);
}

// the same as a tree, one CTA of 64 threads per blob (count <= 64): six levels of additions instead of 63 in a row -- for the
// single-blob calls (commitment, point proof), which are nothing but dependency chains
__global__ void __launch_bounds__(64)
k_sum_positions_tree(G1Jac* __restrict__ pts, int B, int count, int stride) {
    __shared__ G1Jac sm[64];
    const int b = blockIdx.x, t = threadIdx.x;
    G1Jac acc;
    if (t < count) acc = ld_vec(&pts[(size_t)stride * t * B + b]); else jac_set_inf(acc);
    for (int step = 32; step >= 1; step >>= 1) {
        if (t >= step && t < 2 * step) sm[t] = acc;
        __syncthreads();
        if (t < step) jac_add(acc, sm[t + step]);
        __syncthreads();
    }
    if (t == 0) st_vec(&pts[b], acc);
}

// r = BLS12-381 group order as plain limbs
__device__ __forceinline__ void fr_modulus(uint32_t* k) {
#pragma unroll
    for (int l = 0; l < 8; l++) k[l] = FrParams::mod(l);
}

// 48-byte compressed -> affine with curve and (optionally) prime-order subgroup check
// (serialization/src/lib.rs:69-81 -> blstrs from_compressed).  status: 0 ok, 1 malformed/not on curve, 2 not in G1.
__global__ void __launch_bounds__(64)
k_g1_validate(const uint8_t* __restrict__ in, G1Affine* __restrict__ out, uint32_t* __restrict__ status, int n, int check_subgroup) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t buf[48];
    for (int c = 0; c < 48; c++) buf[c] = in[(size_t)i * 48 + c];
    G1Affine a;
    uint32_t st = 0;
    if (g1a_decompress(a, buf)) {
        g1a_set_inf(a);
        st = 1;
    } else if (check_subgroup && !g1a_is_inf(a)) {
        if (!g1a_in_subgroup(a)) st = 2;
    }
    status[i] = st;
    st_vec(&out[i], a);
}

// the prime-order subgroup check on its own (points already decompressed by k_g1_validate with check_subgroup = 0): the verifiers
// run it on a side stream, beside the scalar multiplications that only need the coordinates.  status: 2 = not in G1.
// (two arrays in one launch: a small call is latency-bound, and a second launch behind the first would double it)
__global__ void __launch_bounds__(64)
k_g1_subgroup(const G1Affine* __restrict__ pts, uint32_t* __restrict__ status, int n, const G1Affine* __restrict__ pts2,
              uint32_t* __restrict__ status2, int n2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + n2) return;
    if (i >= n) { i -= n; pts = pts2; status = status2; }
    if (status[i] != 0) return;
    const G1Affine a = ld_vec(&pts[i]);
    if (!g1a_is_inf(a) && !g1a_in_subgroup(a)) status[i] = 2;
}

// ------------------------------------------------------------------------------------------------

cudaError_t launch_blob_challenge(const uint8_t* blobs, const uint8_t* commitments, Fr* z, int B, cudaStream_t st) {
    k_blob_challenge<<<(B + 31) / 32, 32, 0, st>>>(blobs, commitments, z, B);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_scalars_from_be(const uint8_t* in, Fr* out, uint32_t* status, int n, cudaStream_t st) {
    k_scalars_from_be<<<(n + 63) / 64, 64, 0, st>>>(in, out, status, n);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_quotient(const Fr* coeffs, const Fr* z, uint32_t* scalars, uint8_t* y_out, int B, cudaStream_t st) {
    k_quotient<<<B, QT_THREADS, 0, st>>>(coeffs, z, scalars, y_out, B);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_coset_quotients(const Fr* coeffs, uint32_t* scalars, const DevTables& T, int nblobs, cudaStream_t st) {
    const int Bv = 128 * nblobs;
    k_coset_quotients<<<(unsigned)(((size_t)64 * Bv + 127) / 128), 128, 0, st>>>(coeffs, scalars, T.tw128, Bv);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_coeffs_to_scalars(const Fr* coeffs, uint32_t* scalars, int B, cudaStream_t st) {
    size_t n = (size_t)N_BLOB * B;
    k_coeffs_to_scalars<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(coeffs, scalars, B);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_sum_positions(G1Jac* pts, int B, int count, int stride, cudaStream_t st) {
    if (B <= 256 && count <= 64) k_sum_positions_tree<<<B, 64, 0, st>>>(pts, B, count, stride);
    else k_sum_positions<<<(B + 63) / 64, 64, 0, st>>>(pts, B, count, stride);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_g1_subgroup(const G1Affine* pts, uint32_t* status, int n, const G1Affine* pts2, uint32_t* status2, int n2, cudaStream_t st) {
    k_g1_subgroup<<<(n + n2 + 63) / 64, 64, 0, st>>>(pts, status, n, pts2, status2, n2);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}
cudaError_t launch_g1_validate(const uint8_t* in, G1Affine* out, uint32_t* status, int n, bool check_subgroup, cudaStream_t st) {
    k_g1_validate<<<(n + 63) / 64, 64, 0, st>>>(in, out, status, n, check_subgroup ? 1 : 0);
    EKZG_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace ekzg
