// Device-side data layout shared by the kernels (kzg_kernels.cu) and the host runtime (kzg_runtime.cu).
#pragma once
#include <cuda_runtime.h>
#include "g1_mul.cuh"

namespace ekzg {

constexpr int N_BLOB = 4096;         // FIELD_ELEMENTS_PER_BLOB       (crates/serialization/src/constants.rs)
constexpr int N_EXT = 8192;          // FIELD_ELEMENTS_PER_EXT_BLOB
constexpr int CELL_ELEMS = 64;       // FIELD_ELEMENTS_PER_CELL
constexpr int N_CELLS = 128;         // CELLS_PER_EXT_BLOB
constexpr int BYTES_PER_BLOB = 131072;
constexpr int BYTES_PER_CELL = 2048;
constexpr int BYTES_PER_G1 = 48;
constexpr int FK20_POINTS = 64;      // points per fixed-base MSM  (fk20/prover.rs:95-104)
constexpr int FK20_MSMS = 128;       // MSMs per blob = circulant domain size (fk20/batch_toeplitz.rs:113)

// Fixed-base window table over `npoints` base points P_i: entry (i, t, m) = (m+1) * 2^(t*w) * P_i, affine,
// at index ((i*nw + t)*half + m).
struct MsmTable {
    const G1Affine* table;
    int w;      // window width in bits
    int nw;     // number of windows = 255/w + 1
    int half;   // entries per window = 2^(w-1)
    // The top window only sees the tb = 255 - w*(nw-1) leading bits of a scalar (< r < 2^255) plus the Booth carry: its digit
    // lies in [0, rtop) with rtop = 2^tb + 1 (9 for w = 14 and w = 12) and is never negative.  So the top digits of mg consecutive
    // points share ONE lookup: the top slice of the first point of each group of mg holds sum_i d_i * 2^(w(nw-1)) * P_i at index
    // (sum_i d_i * rtop^i) - 1, and a scalar costs nw - 1 + 1/mg additions instead of nw (w = 14: 18.25 instead of 19).
    int mg;     // points per merged top lookup: 4, 2 or 1 (the largest with rtop^mg - 1 <= half)
    int rtop;
    __host__ __device__ void set_window(int w_) {
        w = w_;
        nw = 255 / w_ + 1;
        half = 1 << (w_ - 1);
        rtop = (1 << (255 - w_ * (nw - 1))) + 1;
        mg = 1;
        for (int m = 4; m > 1; m >>= 1) {
            long v = 1;
            for (int i = 0; i < m; i++) v *= rtop;
            if (v - 1 <= half) { mg = m; break; }
        }
    }
};

// Read-only tables, built once per device at context creation.  All Fr/Fp values in Montgomery form.
struct DevTables {
    const Fr* tw4096;        // omega_4096^i,  i < 2048
    const Fr* tw4096_inv;    // omega_4096^-i, i < 2048
    const Fr* tw8192;        // omega_8192^i,  i < 4096   (coset shift of the odd half of the 8192-NTT, also the 8192 twiddles)
    const Fr* tw8192_inv;    // omega_8192^-i, i < 4096
    const Fr* tw128;         // omega_128^i,   i < 64
    const Fr* tw64_inv;      // omega_64^-i,   i < 32     (verifier: 64-point coset IFFT)
    // FK20 fixed-base tables: entry (j, k, t, m) = (m+1) * 2^(t*w) * F_k[j], affine.
    // F_k = NTT_128^{G1}(V_k || O^64)  (fk20/batch_toeplitz.rs:49-58); the reference's table holds only
    // the t = 0 slice and pays w doublings per window at MSM time (fixed_base_msm_window.rs:154-165);
    // here every window has its own slice, so an MSM is additions only.
    MsmTable fk20;                     // base point i = j*64 + k  ->  F_k[j]
    // Same construction over the 4096 monomial SRS points, for commitments and single-point proofs
    // (reference: g1_lincomb -> blst Pippenger, bls12_381/src/lincomb.rs:7-30; fixed bases make it additions only).
    MsmTable srs;
    const G1Affine* srs_g1;            // g1_monomial[4096]
    const G1Affine* srs_g1_lagrange;   // g1_lagrange[4096] in the JSON's (bit-reversed) order
};

}  // namespace ekzg
