// Fixed-base window table descriptor (plain C++: also compiled by the host-emulation tests).
#pragma once
#include "g1.cuh"

namespace ekzg {

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
    EKZG_HD void set_window(int w_) {
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

}  // namespace ekzg
