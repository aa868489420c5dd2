/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's EIP-7594 / EIP-4844 algorithms (rust-eth-kzg v0.9.1), used
 *   (1) as the parity checker for the CUDA path (tests/, __graft_entry__.smoke()), and
 *   (2) as the "port" CPU baseline timed by bench.py (cpu_baseline leg / --impl reference).
 * Nothing under rust-eth-kzg_b200/ may link, import or call this file.
 *
 * Each function cites the reference file:line whose behaviour it restates.  It follows the
 * reference's ALGORITHM (FK20 with width-8 Booth fixed-base tables and batched-affine bucket
 * sums, radix-2 NTTs, Pippenger lincomb) so that its timing is a fair stand-in for the
 * reference's CPU path; the code itself is written from scratch in C.
 *
 * Parity status: PINNED -- tests/test_oracle_vectors.py runs every consensus vector
 * (test_vectors/<fn>/kzg-mainnet/<case>/data.y*ml, glob includes the two .yml cases)
 * through the okzg_* entry points and requires byte-exact outputs / error classification.
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "pairing.h"
#include "sha256.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define N_BLOB 4096
#define N_EXT 8192
#define N_CELLS 128
#define CELL_ELEMS 64
#define BYTES_PER_CELL 2048
#define BYTES_PER_BLOB 131072
#define PRECOMP_W 8                      /* RECOMMENDED_PRECOMP_WIDTH, eip7594/src/constants.rs:65 */
#define PRECOMP_ENTRIES (1 << (PRECOMP_W - 1))
#define N_WINDOWS (255 / PRECOMP_W + 1)  /* fixed_base_msm_window.rs:113 */

enum { OKZG_OK = 0, OKZG_ERR_INPUT = 1, OKZG_ERR_INTERNAL = 2 };

/* ------------------------------------------------------------------ domains (domain.rs:41-107) */
typedef struct {
    int n, log_n;
    fr_t *roots;      /* omega^i, i < n */
    fr_t *roots_inv;  /* omega^-i */
    fr_t n_inv;
} domain_t;

static void domain_init(domain_t *d, int n) {
    d->n = n; d->log_n = 0; while ((1 << d->log_n) < n) d->log_n++;
    fr_t w; fr_from_plain(&w, FR_ROOT_2_32);
    for (int i = d->log_n; i < 32; i++) fr_mul(&w, &w, &w);   /* omega_n = root^(2^32/n) */
    fr_t wi; fr_inv(&wi, &w);
    d->roots = malloc(sizeof(fr_t) * n); d->roots_inv = malloc(sizeof(fr_t) * n);
    fr_set_one(&d->roots[0]); fr_set_one(&d->roots_inv[0]);
    for (int i = 1; i < n; i++) { fr_mul(&d->roots[i], &d->roots[i - 1], &w); fr_mul(&d->roots_inv[i], &d->roots_inv[i - 1], &wi); }
    fr_t nn; fr_from_u64(&nn, (uint64_t)n); fr_inv(&d->n_inv, &nn);
}

static inline uint32_t bitrev(uint32_t i, int bits) {
    uint32_t r = 0; for (int b = 0; b < bits; b++) { r = (r << 1) | (i & 1); i >>= 1; } return r;
}
/* cosets.rs:56-78 reverse_bit_order */
static void brp_fr(fr_t *v, int n) {
    int bits = 0; while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) { int j = bitrev(i, bits); if (i < j) { fr_t t = v[i]; v[i] = v[j]; v[j] = t; } }
}
static void brp_g1(g1_t *v, int n) {
    int bits = 0; while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) { int j = bitrev(i, bits); if (i < j) { g1_t t = v[i]; v[i] = v[j]; v[j] = t; } }
}

/* natural-order in, natural-order out radix-2 DIT (fft.rs:46-64 computes the same map) */
static void fft_fr_core(fr_t *v, int n, const fr_t *roots) {
    brp_fr(v, n);
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len)
            for (int i = 0; i < half; i++) {
                fr_t t; fr_mul(&t, &v[s + i + half], &roots[i * step]);
                fr_sub(&v[s + i + half], &v[s + i], &t); fr_add(&v[s + i], &v[s + i], &t);
            }
    }
}
/* domain.rs:117-125 fft_scalars (input zero-padded to n by the caller) */
static void fft_fr(const domain_t *d, fr_t *v) { fft_fr_core(v, d->n, d->roots); }
/* domain.rs:199-211 ifft_scalars */
static void ifft_fr(const domain_t *d, fr_t *v) {
    fft_fr_core(v, d->n, d->roots_inv);
    for (int i = 0; i < d->n; i++) fr_mul(&v[i], &v[i], &d->n_inv);
}
/* domain.rs:129-142 coset_fft_scalars: scale by g^i then fft */
static void coset_fft_fr(const domain_t *d, fr_t *v, const fr_t *g) {
    fr_t p; fr_set_one(&p);
    for (int i = 0; i < d->n; i++) { fr_mul(&v[i], &v[i], &p); fr_mul(&p, &p, g); }
    fft_fr(d, v);
}
/* domain.rs:214-223 coset_ifft_scalars: ifft then scale by g^-i */
static void coset_ifft_fr(const domain_t *d, fr_t *v, const fr_t *g_inv) {
    ifft_fr(d, v);
    fr_t p; fr_set_one(&p);
    for (int i = 0; i < d->n; i++) { fr_mul(&v[i], &v[i], &p); fr_mul(&p, &p, g_inv); }
}

/* windowed (w=4, signed) scalar multiplication: the stand-in for blst's G1 scalar mul */
static void g1_mul_fr_w4(g1_t *r, const g1_t *p, const fr_t *s) {
    uint64_t k[5]; fr_to_plain(k, s); k[4] = 0;
    g1_t tab[8]; tab[0] = *p; g1_dbl(&tab[1], p);
    for (int i = 2; i < 8; i++) g1_add(&tab[i], &tab[i - 1], p);
    g1_t acc; g1_set_inf(&acc);
    /* Booth digits over 4-bit windows, 64 windows cover 256 bits */
    for (int wdx = 63; wdx >= 0; wdx--) {
        for (int j = 0; j < 4; j++) g1_dbl(&acc, &acc);
        int lo = wdx * 4 - 1;
        uint32_t v;
        if (lo < 0) v = (uint32_t)(k[0] << 1) & 0x1F;
        else { v = (uint32_t)(k[lo / 64] >> (lo % 64)); if (lo % 64 > 59) v |= (uint32_t)(k[lo / 64 + 1] << (64 - lo % 64)); v &= 0x1F; }
        int d = (int)((v + 1) >> 1) - (int)((v >> 4) << 4);
        if (d > 0) g1_add(&acc, &acc, &tab[d - 1]);
        else if (d < 0) { g1_t n; g1_neg(&n, &tab[-d - 1]); g1_add(&acc, &acc, &n); }
    }
    *r = acc;
}

/* fft.rs:164-177 dit butterfly over G1 with the twiddle==1 and identity short-cuts */
static void fft_g1_core(g1_t *v, int n, const fr_t *roots) {
    brp_g1(v, n);
    fr_t one; fr_set_one(&one);
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len)
            for (int i = 0; i < half; i++) {
                g1_t t;
                const fr_t *w = &roots[i * step];
                if (g1_is_inf(&v[s + i + half])) g1_set_inf(&t);
                else if (fr_eq(w, &one)) t = v[s + i + half];
                else g1_mul_fr_w4(&t, &v[s + i + half], w);
                g1_t a = v[s + i];
                g1_add(&v[s + i], &a, &t); g1_sub(&v[s + i + half], &a, &t);
            }
    }
}
/* domain.rs:149-157 fft_g1 */
static void fft_g1(const domain_t *d, g1_t *v) { fft_g1_core(v, d->n, d->roots); }
/* domain.rs:172-194 ifft_g1_take_n: inverse transform, keep the first `take`, scale by 1/n */
static void ifft_g1_take_n(const domain_t *d, g1_t *v, int take) {
    fft_g1_core(v, d->n, d->roots_inv);
    for (int i = 0; i < take; i++) g1_mul_fr_w4(&v[i], &v[i], &d->n_inv);
}

/* ------------------------------------------------------------------ batch inversion (batch_inversion.rs:17-57) */
static void fr_batch_inverse(fr_t *v, int n) {
    fr_t *pre = malloc(sizeof(fr_t) * n);
    fr_t acc; fr_set_one(&acc);
    for (int i = 0; i < n; i++) { pre[i] = acc; fr_mul(&acc, &acc, &v[i]); }
    fr_inv(&acc, &acc);
    for (int i = n - 1; i >= 0; i--) { fr_t t; fr_mul(&t, &acc, &pre[i]); fr_mul(&acc, &acc, &v[i]); v[i] = t; }
    free(pre);
}
static void fp_batch_inverse(fp_t *v, fp_t *scratch, int n) {
    fp_t acc; fp_set_one(&acc);
    for (int i = 0; i < n; i++) { scratch[i] = acc; fp_mul(&acc, &acc, &v[i]); }
    fp_inv(&acc, &acc);
    for (int i = n - 1; i >= 0; i--) { fp_t t; fp_mul(&t, &acc, &scratch[i]); fp_mul(&acc, &acc, &v[i]); v[i] = t; }
}

/* ------------------------------------------------------------------ batched affine addition
 * batch_addition.rs:142-232 multi_batch_addition_binary_tree_stride: all lists are reduced
 * pairwise, one shared inversion per tree level.  Unlike the reference this version also handles
 * P+P and P+(-P) pairs (the reference assumes they never occur, batch_addition.rs:27-31). */
static void multi_batch_add(g1a_t **lists, int *counts, int nlists, g1_t *sums) {
    int total = 0; for (int l = 0; l < nlists; l++) total += counts[l];
    fp_t *den = malloc(sizeof(fp_t) * (total / 2 + 1)), *scr = malloc(sizeof(fp_t) * (total / 2 + 1));
    for (;;) {
        int np = 0;
        for (int l = 0; l < nlists; l++) {
            g1a_t *L = lists[l];
            /* drop identities */
            int c = 0; for (int i = 0; i < counts[l]; i++) if (!L[i].inf) L[c++] = L[i];
            counts[l] = c;
            for (int i = 0; i + 1 < c; i += 2) {
                fp_t d; fp_sub(&d, &L[i + 1].x, &L[i].x);
                if (fp_is_zero(&d)) {
                    if (fp_eq(&L[i].y, &L[i + 1].y) && !fp_is_zero(&L[i].y)) fp_dbl(&d, &L[i].y);   /* doubling: 2y */
                    else fp_set_one(&d);                                                         /* cancels: marker */
                }
                den[np++] = d;
            }
        }
        if (np == 0) break;
        fp_batch_inverse(den, scr, np);
        np = 0;
        for (int l = 0; l < nlists; l++) {
            g1a_t *L = lists[l]; int c = counts[l], o = 0;
            for (int i = 0; i + 1 < c; i += 2) {
                const g1a_t *p = &L[i], *q = &L[i + 1];
                fp_t lam, t; g1a_t r; r.inf = 0;
                fp_sub(&t, &q->x, &p->x);
                if (fp_is_zero(&t)) {
                    if (fp_eq(&p->y, &q->y) && !fp_is_zero(&p->y)) {
                        fp_sqr(&t, &p->x); fp_dbl(&lam, &t); fp_add(&lam, &lam, &t); fp_mul(&lam, &lam, &den[np]);
                    } else { np++; continue; }   /* P + (-P): contributes nothing */
                } else { fp_sub(&t, &q->y, &p->y); fp_mul(&lam, &t, &den[np]); }
                np++;
                fp_sqr(&r.x, &lam); fp_sub(&r.x, &r.x, &p->x); fp_sub(&r.x, &r.x, &q->x);
                fp_sub(&t, &p->x, &r.x); fp_mul(&r.y, &lam, &t); fp_sub(&r.y, &r.y, &p->y);
                L[o++] = r;
            }
            if (c & 1) L[o++] = L[c - 1];
            counts[l] = o;
        }
    }
    for (int l = 0; l < nlists; l++) {
        if (counts[l] == 0) g1_set_inf(&sums[l]); else g1_from_affine(&sums[l], &lists[l][0]);
    }
    free(den); free(scr);
}

/* booth_encoding.rs:4-46 get_booth_index, closed form d = ((v+1)>>1) - (v>>w)*2^w */
static int booth_digit(const uint64_t k[4], int window, int w) {
    int lo = window * w - 1;
    uint64_t v;
    if (lo < 0) v = k[0] << 1;
    else {
        int li = lo / 64, sh = lo % 64;
        v = li < 4 ? k[li] >> sh : 0;
        if (sh && li + 1 < 4) v |= k[li + 1] << (64 - sh);
    }
    v &= (1u << (w + 1)) - 1;
    return (int)((v + 1) >> 1) - (int)((v >> w) << w);
}

/* ------------------------------------------------------------------ Pippenger lincomb
 * lincomb.rs:7-30 g1_lincomb -> blstrs multi_exp (blst Pippenger); identity points and zero
 * scalars are filtered first (lincomb.rs:13-27). */
static void g1_lincomb(g1_t *out, const g1a_t *pts, const fr_t *sc, int n) {
    g1a_t *P = malloc(sizeof(g1a_t) * (n + 1)); uint64_t (*K)[4] = malloc(32 * (n + 1));
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (pts[i].inf || fr_is_zero(&sc[i])) continue;
        P[m] = pts[i]; fr_to_plain(K[m], &sc[i]); m++;
    }
    g1_t acc; g1_set_inf(&acc);
    if (m == 0) { *out = acc; free(P); free(K); return; }
    int c = m < 32 ? 3 : m < 512 ? 6 : m < 8192 ? 9 : 11;
    int nw = (255 + c) / c;   /* signed digits need one carry bit */
    int nb = 1 << (c - 1);
    g1_t *buckets = malloc(sizeof(g1_t) * nb);
    for (int w = nw - 1; w >= 0; w--) {
        for (int j = 0; j < c; j++) g1_dbl(&acc, &acc);
        for (int b = 0; b < nb; b++) g1_set_inf(&buckets[b]);
        for (int i = 0; i < m; i++) {
            int d = booth_digit(K[i], w, c);
            if (d > 0) g1_add_affine(&buckets[d - 1], &buckets[d - 1], &P[i]);
            else if (d < 0) { g1a_t nq; g1a_neg(&nq, &P[i]); g1_add_affine(&buckets[-d - 1], &buckets[-d - 1], &nq); }
        }
        g1_t run, sum; g1_set_inf(&run); g1_set_inf(&sum);
        for (int b = nb - 1; b >= 0; b--) { g1_add(&run, &run, &buckets[b]); g1_add(&sum, &sum, &run); }
        g1_add(&acc, &acc, &sum);
    }
    *out = acc; free(buckets); free(P); free(K);
}

/* ------------------------------------------------------------------ context */
typedef struct {
    g1a_t *srs_g1;            /* 4096 monomial points */
    g2a_t g2_gen, g2_tau, g2_tau64, g2_neg_gen;
    domain_t d4096, d8192, d128, d64;
    g1a_t *fk20_table;        /* [128 msm][64 point][128 multiple], fixed_base_msm_window.rs:69-82 */
    fr_t coset_gens[N_CELLS];       /* h_j = omega_8192^{rev7(j)} (cosets.rs:89-112) */
    fr_t coset_gens_inv[N_CELLS];
    fr_t coset_gens_pow_n[N_CELLS]; /* h_j^64 (verifier.rs:77-106) */
    fr_t rs_coset_gen, rs_coset_gen_inv; /* 7 (reed_solomon.rs:129) */
    int ready;
} ctx_t;

static ctx_t G;

static void fk20_setup(ctx_t *c);

/* trusted_setup/src/lib.rs:80-124: parse the ceremony points (no subgroup check, lib.rs:80-86) */
int okzg_init(const char *setup_bin) {
    if (G.ready) return OKZG_OK;
    FILE *f = fopen(setup_bin, "rb"); if (!f) return OKZG_ERR_INPUT;
    uint8_t hdr[16]; if (fread(hdr, 1, 16, f) != 16 || memcmp(hdr, "EKZGTS01", 8)) { fclose(f); return OKZG_ERR_INPUT; }
    uint8_t *g1m = malloc(48 * 4096), *g1l = malloc(48 * 4096), *g2m = malloc(96 * 65);
    if (fread(g1m, 48, 4096, f) != 4096 || fread(g1l, 48, 4096, f) != 4096 || fread(g2m, 96, 65, f) != 65) { fclose(f); return OKZG_ERR_INPUT; }
    fclose(f);
    ctx_t *c = &G;
    c->srs_g1 = malloc(sizeof(g1a_t) * 4096);
    int bad = 0;
#pragma omp parallel for reduction(| : bad)
    for (int i = 0; i < 4096; i++) bad |= !g1a_decompress(&c->srs_g1[i], g1m + 48 * i, 0);
    if (bad) return OKZG_ERR_INPUT;
    if (!g2a_decompress(&c->g2_gen, g2m) || !g2a_decompress(&c->g2_tau, g2m + 96) || !g2a_decompress(&c->g2_tau64, g2m + 96 * 64)) return OKZG_ERR_INPUT;
    g2a_neg(&c->g2_neg_gen, &c->g2_gen);
    free(g1m); free(g1l); free(g2m);
    domain_init(&c->d4096, 4096); domain_init(&c->d8192, 8192); domain_init(&c->d128, 128); domain_init(&c->d64, 64);
    for (int j = 0; j < N_CELLS; j++) {
        c->coset_gens[j] = c->d8192.roots[bitrev(j, 7)];
        c->coset_gens_inv[j] = c->d8192.roots_inv[bitrev(j, 7)];
        fr_pow_u64(&c->coset_gens_pow_n[j], &c->coset_gens[j], 64);
    }
    fr_from_u64(&c->rs_coset_gen, 7); fr_inv(&c->rs_coset_gen_inv, &c->rs_coset_gen);
    fk20_setup(c);
    c->ready = 1;
    return OKZG_OK;
}

/* fk20/prover.rs:88-108 + batch_toeplitz.rs:34-78 */
static void fk20_setup(ctx_t *c) {
    /* S = reverse(srs)[64..]  => S[i] = srs[4031 - i]; V_k = S[k], S[k+64], ... (63 points) || O */
    g1_t *F = malloc(sizeof(g1_t) * 64 * 128);
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < 64; k++) {
        g1_t *v = F + 128 * k;
        for (int m = 0; m < 128; m++) {
            int idx = k + 64 * m;
            if (m < 63 && idx < 4032) g1_from_affine(&v[m], &c->srs_g1[4031 - idx]); else g1_set_inf(&v[m]);
        }
        fft_g1(&c->d128, v);   /* batch_toeplitz.rs:49-58 */
    }
    /* transpose (batch_toeplitz.rs:61): bases of MSM j are F_k[j], k<64; tables of (m+1)*P, m<128 */
    c->fk20_table = malloc(sizeof(g1a_t) * 128 * 64 * PRECOMP_ENTRIES);
#pragma omp parallel for schedule(dynamic)
    for (int jk = 0; jk < 128 * 64; jk++) {
        int j = jk / 64, k = jk % 64;
        g1_t mult[PRECOMP_ENTRIES];
        mult[0] = F[128 * k + j];
        for (int m = 1; m < PRECOMP_ENTRIES; m++) g1_add(&mult[m], &mult[m - 1], &mult[0]);
        g1_batch_to_affine(c->fk20_table + (size_t)jk * PRECOMP_ENTRIES, mult, PRECOMP_ENTRIES);
    }
    free(F);
}

/* fixed_base_msm_window.rs:102-168 FixedBaseMSMPrecompWindow::msm */
static void fixed_base_msm(g1_t *out, const g1a_t *table /* [64][128] */, const fr_t *scalars) {
    uint64_t K[64][4];
    for (int i = 0; i < 64; i++) fr_to_plain(K[i], &scalars[i]);
    static __thread g1a_t store[N_WINDOWS][64];
    g1a_t *lists[N_WINDOWS]; int counts[N_WINDOWS];
    for (int w = 0; w < N_WINDOWS; w++) {
        lists[w] = store[w]; int c = 0;
        for (int i = 0; i < 64; i++) {
            int d = booth_digit(K[i], w, PRECOMP_W);
            if (d == 0) continue;
            const g1a_t *e = &table[i * PRECOMP_ENTRIES + (d > 0 ? d : -d) - 1];
            if (d > 0) store[w][c] = *e; else g1a_neg(&store[w][c], e);
            c++;
        }
        counts[w] = c;
    }
    g1_t sums[N_WINDOWS];
    multi_batch_add(lists, counts, N_WINDOWS, sums);
    g1_t r = sums[N_WINDOWS - 1];
    for (int w = N_WINDOWS - 2; w >= 0; w--) {
        for (int j = 0; j < PRECOMP_W; j++) g1_dbl(&r, &r);
        g1_add(&r, &r, &sums[w]);
    }
    *out = r;
}

/* serialization/src/lib.rs:36-63 */
static int blob_to_scalars(fr_t *out, const uint8_t *blob) {
    for (int i = 0; i < N_BLOB; i++) if (!fr_from_be(&out[i], blob + 32 * i)) return 0;
    return 1;
}
/* fk20/prover.rs:177-180 / eip4844/src/verifier.rs:146-150: c = INTT(BRP(e)) */
static void scalars_to_coeffs(const ctx_t *c, fr_t *v) { brp_fr(v, N_BLOB); ifft_fr(&c->d4096, v); }

/* fk20/prover.rs:158-165 compute_coset_evaluations + serialization/src/lib.rs:132-156 */
static void coeffs_to_cells(const ctx_t *c, const fr_t *coeffs, uint8_t *cells) {
    fr_t *e = malloc(sizeof(fr_t) * N_EXT);
    memcpy(e, coeffs, sizeof(fr_t) * N_BLOB);
    for (int i = N_BLOB; i < N_EXT; i++) fr_set_zero(&e[i]);
    fft_fr(&c->d8192, e); brp_fr(e, N_EXT);
    for (int i = 0; i < N_EXT; i++) fr_to_be(cells + 32 * i, &e[i]);
    free(e);
}

/* fk20/prover.rs:206-228 + h_poly.rs:18-57 + toeplitz.rs:132-144 + batch_toeplitz.rs:86-125 */
static void coeffs_to_proofs(const ctx_t *c, const fr_t *coeffs, uint8_t *proofs) {
    /* row_k[i] = p[k + 64 i], p = reverse(coeffs); circulant a_k = [row[0], 0^63, 0, row[63..1]] */
    fr_t (*A)[128] = malloc(sizeof(fr_t) * 64 * 128);
    for (int k = 0; k < 64; k++) {
        fr_t *a = A[k];
        for (int i = 0; i < 128; i++) fr_set_zero(&a[i]);
        a[0] = coeffs[4095 - k];
        for (int i = 1; i < 64; i++) a[128 - i] = coeffs[4095 - k - 64 * i];
        fft_fr(&c->d128, a);   /* batch_toeplitz.rs:103-106 */
    }
    g1_t R[128];
    for (int j = 0; j < 128; j++) {   /* batch_toeplitz.rs:113-117 */
        fr_t s[64];
        for (int k = 0; k < 64; k++) s[k] = A[k][j];
        fixed_base_msm(&R[j], c->fk20_table + (size_t)j * 64 * PRECOMP_ENTRIES, s);
    }
    free(A);
    ifft_g1_take_n(&c->d128, R, 64);                  /* batch_toeplitz.rs:123 */
    for (int i = 64; i < 128; i++) g1_set_inf(&R[i]);
    fft_g1(&c->d128, R);                              /* fk20/prover.rs:217 */
    brp_g1(R, 128);                                   /* fk20/prover.rs:222 */
    g1a_t aff[128]; g1_batch_to_affine(aff, R, 128);
    for (int i = 0; i < 128; i++) g1a_compress(proofs + 48 * i, &aff[i]);
}

/* eip7594/src/prover.rs:117-134 */
int okzg_compute_cells_and_kzg_proofs(const uint8_t *blob, uint8_t *cells, uint8_t *proofs) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    fr_t *v = malloc(sizeof(fr_t) * N_BLOB);
    if (!blob_to_scalars(v, blob)) { free(v); return OKZG_ERR_INPUT; }
    scalars_to_coeffs(&G, v);
    if (cells) coeffs_to_cells(&G, v, cells);
    if (proofs) coeffs_to_proofs(&G, v, proofs);
    free(v);
    return OKZG_OK;
}
/* eip7594/src/prover.rs:136-148 */
int okzg_compute_cells(const uint8_t *blob, uint8_t *cells) { return okzg_compute_cells_and_kzg_proofs(blob, cells, NULL); }

/* throughput helper for the CPU baseline: blobs are independent, one blob per thread */
int okzg_compute_cells_and_kzg_proofs_batch(int n, const uint8_t *blobs, uint8_t *cells, uint8_t *proofs, int nthreads) {
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic) reduction(| : err)
    for (int b = 0; b < n; b++)
        err |= okzg_compute_cells_and_kzg_proofs(blobs + (size_t)b * BYTES_PER_BLOB, cells + (size_t)b * N_CELLS * BYTES_PER_CELL, proofs + (size_t)b * N_CELLS * 48);
    return err;
}
int okzg_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* eip4844/src/prover.rs:17-31 (monomial SRS after an IFFT) */
int okzg_blob_to_kzg_commitment(const uint8_t *blob, uint8_t *out) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    fr_t *v = malloc(sizeof(fr_t) * N_BLOB);
    if (!blob_to_scalars(v, blob)) { free(v); return OKZG_ERR_INPUT; }
    scalars_to_coeffs(&G, v);
    g1_t cm; g1_lincomb(&cm, G.srs_g1, v, N_BLOB);
    g1_compress(out, &cm); free(v);
    return OKZG_OK;
}

/* kzg_single_open/src/prover.rs:33-65: Ruffini quotient from the top coefficient, proof = MSM */
static void compute_kzg_proof_poly(const fr_t *poly, const fr_t *z, uint8_t *proof_out, fr_t *y_out) {
    fr_t *q = malloc(sizeof(fr_t) * N_BLOB);
    fr_t k; fr_set_zero(&k);
    /* quotient pushed from the highest degree; q_rev[i] for coefficient 4095-i */
    for (int i = N_BLOB - 1; i >= 0; i--) {
        fr_t t; fr_add(&t, &poly[i], &k);
        q[i] = t;             /* q[i] will be the coefficient of X^(i-1); q[0] is the remainder */
        fr_mul(&k, z, &t);
    }
    *y_out = q[0];
    g1_t pr; g1_lincomb(&pr, G.srs_g1, q + 1, N_BLOB - 1);
    g1_compress(proof_out, &pr); free(q);
}

/* eip4844/src/prover.rs:37-56 */
int okzg_compute_kzg_proof(const uint8_t *blob, const uint8_t *z_bytes, uint8_t *proof, uint8_t *y_bytes) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    fr_t *v = malloc(sizeof(fr_t) * N_BLOB);
    if (!blob_to_scalars(v, blob)) { free(v); return OKZG_ERR_INPUT; }
    scalars_to_coeffs(&G, v);
    fr_t z, y;
    if (!fr_from_be(&z, z_bytes)) { free(v); return OKZG_ERR_INPUT; }
    compute_kzg_proof_poly(v, &z, proof, &y);
    fr_to_be(y_bytes, &y); free(v);
    return OKZG_OK;
}

/* eip4844/src/verifier.rs:155-196 */
static void blob_challenge(fr_t *z, const uint8_t *blob, const uint8_t *commitment) {
    sha256_ctx s; sha256_init(&s);
    sha256_update(&s, (const uint8_t *)"FSBLOBVERIFY_V1_", 16);
    uint8_t deg[16] = {0}; deg[14] = 0x10; /* u128_be(4096) */
    sha256_update(&s, deg, 16); sha256_update(&s, blob, BYTES_PER_BLOB); sha256_update(&s, commitment, 48);
    uint8_t h[32]; sha256_final(&s, h); fr_from_be_reduce(z, h);
}

/* eip4844/src/prover.rs:65-88 */
int okzg_compute_blob_kzg_proof(const uint8_t *blob, const uint8_t *commitment, uint8_t *proof) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    fr_t *v = malloc(sizeof(fr_t) * N_BLOB);
    if (!blob_to_scalars(v, blob)) { free(v); return OKZG_ERR_INPUT; }
    scalars_to_coeffs(&G, v);
    g1a_t cm; if (!g1a_decompress(&cm, commitment, 1)) { free(v); return OKZG_ERR_INPUT; }
    fr_t z, y; blob_challenge(&z, blob, commitment);
    compute_kzg_proof_poly(v, &z, proof, &y); free(v);
    return OKZG_OK;
}

/* kzg_single_open/src/verifier.rs:33-58 */
static int verify_kzg_proof_inner(const g1a_t *cm, const fr_t *z, const fr_t *y, const g1a_t *proof) {
    g1a_t gen = G.srs_g1[0];
    g1_t t, gy, c; g1_from_affine(&t, &gen); g1_mul_fr(&gy, &t, y);
    g1_from_affine(&c, cm); g1_sub(&c, &c, &gy);
    g1a_t ps[2]; g2a_t qs[2];
    g1_to_affine(&ps[0], &c); qs[0] = G.g2_neg_gen;
    ps[1] = *proof;
    g2a_t gz, ngz; g2a_mul_fr(&gz, &G.g2_gen, z); g2a_neg(&ngz, &gz); g2a_add(&qs[1], &G.g2_tau, &ngz);
    return pairing_check(ps, qs, 2);
}

/* eip4844/src/verifier.rs:19-48 */
int okzg_verify_kzg_proof(const uint8_t *commitment, const uint8_t *z_bytes, const uint8_t *y_bytes, const uint8_t *proof, int *ok) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    g1a_t cm, pr; fr_t z, y;
    if (!g1a_decompress(&cm, commitment, 1)) return OKZG_ERR_INPUT;
    if (!g1a_decompress(&pr, proof, 1)) return OKZG_ERR_INPUT;
    if (!fr_from_be(&z, z_bytes)) return OKZG_ERR_INPUT;
    if (!fr_from_be(&y, y_bytes)) return OKZG_ERR_INPUT;
    *ok = verify_kzg_proof_inner(&cm, &z, &y, &pr);
    return OKZG_OK;
}

/* poly_coeff.rs eval (Horner) */
static void poly_eval(fr_t *y, const fr_t *poly, int n, const fr_t *z) {
    fr_t acc; fr_set_zero(&acc);
    for (int i = n - 1; i >= 0; i--) { fr_mul(&acc, &acc, z); fr_add(&acc, &acc, &poly[i]); }
    *y = acc;
}

/* eip4844/src/verifier.rs:53-78 */
int okzg_verify_blob_kzg_proof(const uint8_t *blob, const uint8_t *commitment, const uint8_t *proof, int *ok) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    fr_t *v = malloc(sizeof(fr_t) * N_BLOB);
    if (!blob_to_scalars(v, blob)) { free(v); return OKZG_ERR_INPUT; }
    g1a_t cm, pr;
    if (!g1a_decompress(&cm, commitment, 1) || !g1a_decompress(&pr, proof, 1)) { free(v); return OKZG_ERR_INPUT; }
    fr_t z, y; blob_challenge(&z, blob, commitment);
    scalars_to_coeffs(&G, v); poly_eval(&y, v, N_BLOB, &z); free(v);
    *ok = verify_kzg_proof_inner(&cm, &z, &y, &pr);
    return OKZG_OK;
}

/* eip4844/src/verifier.rs:80-143 + :201-260 + kzg_single_open/src/verifier.rs:60-108 */
int okzg_verify_blob_kzg_proof_batch(int n, const uint8_t *blobs, const uint8_t *commitments, const uint8_t *proofs, int *ok) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    fr_t *zs = malloc(sizeof(fr_t) * (n + 1)), *ys = malloc(sizeof(fr_t) * (n + 1));
    g1a_t *cms = malloc(sizeof(g1a_t) * (n + 1)), *prs = malloc(sizeof(g1a_t) * (n + 1));
    fr_t *v = malloc(sizeof(fr_t) * N_BLOB);
    int rc = OKZG_OK;
    /* reference order: all blobs, then all commitments, then all proofs */
    for (int i = 0; i < n && rc == OKZG_OK; i++) if (!blob_to_scalars(v, blobs + (size_t)i * BYTES_PER_BLOB)) rc = OKZG_ERR_INPUT;
    for (int i = 0; i < n && rc == OKZG_OK; i++) if (!g1a_decompress(&cms[i], commitments + 48 * i, 1)) rc = OKZG_ERR_INPUT;
    for (int i = 0; i < n && rc == OKZG_OK; i++) if (!g1a_decompress(&prs[i], proofs + 48 * i, 1)) rc = OKZG_ERR_INPUT;
    if (rc != OKZG_OK) goto done;
    for (int i = 0; i < n; i++) {
        blob_to_scalars(v, blobs + (size_t)i * BYTES_PER_BLOB);
        blob_challenge(&zs[i], blobs + (size_t)i * BYTES_PER_BLOB, commitments + 48 * i);
        scalars_to_coeffs(&G, v); poly_eval(&ys[i], v, N_BLOB, &zs[i]);
    }
    {
        sha256_ctx s; sha256_init(&s);
        sha256_update(&s, (const uint8_t *)"RCKZGBATCH___V1_", 16);
        uint8_t u[8] = {0, 0, 0, 0, 0, 0, 0x10, 0}; sha256_update(&s, u, 8);
        for (int i = 0; i < 8; i++) u[i] = (uint8_t)((uint64_t)n >> (56 - 8 * i));
        sha256_update(&s, u, 8);
        for (int i = 0; i < n; i++) {
            uint8_t zb[32], yb[32]; fr_to_be(zb, &zs[i]); fr_to_be(yb, &ys[i]);
            sha256_update(&s, commitments + 48 * i, 48); sha256_update(&s, zb, 32); sha256_update(&s, yb, 32); sha256_update(&s, proofs + 48 * i, 48);
        }
        uint8_t h[32]; sha256_final(&s, h);
        fr_t r; fr_from_be_reduce(&r, h);
        /* lhs = sum r^i C_i - (sum r^i y_i) G + sum r^i z_i Q_i ; rhs = sum r^i Q_i */
        int m = 2 * n + 1;
        g1a_t *pts = malloc(sizeof(g1a_t) * m); fr_t *sc = malloc(sizeof(fr_t) * m), *rp = malloc(sizeof(fr_t) * (n + 1));
        fr_t p, ylin; fr_set_one(&p); fr_set_zero(&ylin);
        for (int i = 0; i < n; i++) {
            rp[i] = p; pts[i] = cms[i]; sc[i] = p; pts[n + 1 + i] = prs[i]; fr_mul(&sc[n + 1 + i], &p, &zs[i]);
            fr_t t; fr_mul(&t, &p, &ys[i]); fr_add(&ylin, &ylin, &t);
            fr_mul(&p, &p, &r);
        }
        pts[n] = G.srs_g1[0]; fr_neg(&sc[n], &ylin);
        g1_t lhs, rhs; g1_lincomb(&lhs, pts, sc, m); g1_lincomb(&rhs, prs, rp, n);
        g1a_t ps[2]; g2a_t qs[2];
        g1_to_affine(&ps[0], &lhs); qs[0] = G.g2_neg_gen;
        g1_to_affine(&ps[1], &rhs); qs[1] = G.g2_tau;
        *ok = pairing_check(ps, qs, 2);
        free(pts); free(sc); free(rp);
    }
done:
    free(zs); free(ys); free(cms); free(prs); free(v);
    return rc;
}

/* ------------------------------------------------------------------ recovery
 * eip7594/src/recovery.rs:22-146, fk20/cosets.rs:141-198, erasure_codes/src/reed_solomon.rs:220-384 */
static int recover_coeffs(const ctx_t *c, int n, const uint64_t *idx, const uint8_t *cells, fr_t *coeffs /*4096*/) {
    /* validation order: recovery.rs:90-146 (len equality is the caller's: flat arrays) */
    for (int i = 0; i < n; i++) if (idx[i] >= N_CELLS) return OKZG_ERR_INPUT;
    for (int i = 1; i < n; i++) if (!(idx[i - 1] < idx[i])) return OKZG_ERR_INPUT;
    if (n < N_CELLS / 2 || n > N_CELLS) return OKZG_ERR_INPUT;
    fr_t *e = calloc(N_EXT, sizeof(fr_t));
    int present[N_CELLS] = {0};
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < CELL_ELEMS; j++)
            if (!fr_from_be(&e[idx[i] * CELL_ELEMS + j], cells + (size_t)i * BYTES_PER_CELL + 32 * j)) { free(e); return OKZG_ERR_INPUT; }
        present[bitrev((uint32_t)idx[i], 7)] = 1;   /* cosets.rs:141-198: idx -> rev7(idx) */
    }
    brp_fr(e, N_EXT);
    /* vanishing polynomial over the missing 128-domain roots (poly_coeff.rs:109-115), stride 64 */
    fr_t z0[N_CELLS + 1]; int deg = 0; fr_set_one(&z0[0]);
    for (int m = 0; m < N_CELLS; m++) {
        if (present[m]) continue;
        /* z0 *= (X - w^m) */
        fr_t root = c->d128.roots[m];
        fr_set_zero(&z0[deg + 1]);
        for (int i = deg + 1; i >= 1; i--) { fr_t t; fr_mul(&t, &z0[i], &root); fr_sub(&z0[i], &z0[i - 1], &t); }
        { fr_t t; fr_mul(&t, &z0[0], &root); fr_neg(&z0[0], &t); }
        deg++;
    }
    /* NB: loop above computes new[i] = old[i-1] - root*old[i]; processed high->low so old values are intact */
    fr_t *zx = calloc(N_EXT, sizeof(fr_t)), *zeval = malloc(sizeof(fr_t) * N_EXT);
    for (int i = 0; i <= deg; i++) zx[i * (N_EXT / N_CELLS)] = z0[i];
    memcpy(zeval, zx, sizeof(fr_t) * N_EXT);
    fft_fr(&c->d8192, zeval);
    for (int i = 0; i < N_EXT; i++) fr_mul(&e[i], &e[i], &zeval[i]);   /* (E*Z) evals */
    ifft_fr(&c->d8192, e);                                            /* dz coeffs */
    coset_fft_fr(&c->d8192, e, &c->rs_coset_gen);
    coset_fft_fr(&c->d8192, zx, &c->rs_coset_gen);
    fr_batch_inverse(zx, N_EXT);
    for (int i = 0; i < N_EXT; i++) fr_mul(&e[i], &e[i], &zx[i]);
    coset_ifft_fr(&c->d8192, e, &c->rs_coset_gen_inv);
    int rc = OKZG_OK;
    for (int i = N_BLOB; i < N_EXT; i++) if (!fr_is_zero(&e[i])) rc = OKZG_ERR_INPUT;
    memcpy(coeffs, e, sizeof(fr_t) * N_BLOB);
    free(e); free(zx); free(zeval);
    return rc;
}

/* eip7594/src/prover.rs:156-171 */
int okzg_recover_cells_and_kzg_proofs(int n_idx, const uint64_t *idx, int n_cells, const uint8_t *cells_in, uint8_t *cells, uint8_t *proofs) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    if (n_idx != n_cells) return OKZG_ERR_INPUT;
    fr_t *coeffs = malloc(sizeof(fr_t) * N_BLOB);
    int rc = recover_coeffs(&G, n_idx, idx, cells_in, coeffs);
    if (rc == OKZG_OK) { coeffs_to_cells(&G, coeffs, cells); coeffs_to_proofs(&G, coeffs, proofs); }
    free(coeffs);
    return rc;
}

/* ------------------------------------------------------------------ batch cell verification
 * eip7594/src/verifier.rs:49-164 + fk20/verifier.rs:129-384 */
int okzg_verify_cell_kzg_proof_batch(int n_commitments, const uint8_t *commitments, int n_idx, const uint64_t *cell_idx,
                                     int n_cells, const uint8_t *cells, int n_proofs, const uint8_t *proofs, int *ok) {
    if (!G.ready) return OKZG_ERR_INTERNAL;
    /* dedup commitments by first occurrence (verifier.rs:49-65) */
    int n = n_commitments;
    int *row = malloc(sizeof(int) * (n + 1)); int *uniq = malloc(sizeof(int) * (n + 1)); int m = 0;
    for (int i = 0; i < n; i++) {
        int f = -1;
        for (int j = 0; j < m; j++) if (!memcmp(commitments + 48 * uniq[j], commitments + 48 * i, 48)) { f = j; break; }
        if (f < 0) { uniq[m] = i; f = m++; }
        row[i] = f;
    }
    int rc = OKZG_OK;
    g1a_t *cms = NULL, *prs = NULL; fr_t *ev = NULL;
    if (!(n == n_idx && n == n_cells && n == n_proofs)) { rc = OKZG_ERR_INPUT; goto done; }
    for (int i = 0; i < n; i++) if (cell_idx[i] >= N_CELLS) { rc = OKZG_ERR_INPUT; goto done; }
    if (n == 0) { *ok = 1; goto done; }
    cms = malloc(sizeof(g1a_t) * m); prs = malloc(sizeof(g1a_t) * n); ev = malloc(sizeof(fr_t) * n * CELL_ELEMS);
    for (int j = 0; j < m; j++) if (!g1a_decompress(&cms[j], commitments + 48 * uniq[j], 1)) { rc = OKZG_ERR_INPUT; goto done; }
    {
        int bad = 0;
#pragma omp parallel for reduction(| : bad)
        for (int i = 0; i < n; i++) bad |= !g1a_decompress(&prs[i], proofs + 48 * i, 1);
        if (bad) { rc = OKZG_ERR_INPUT; goto done; }
    }
    for (int i = 0; i < n * CELL_ELEMS; i++) if (!fr_from_be(&ev[i], cells + 32 * (size_t)i)) { rc = OKZG_ERR_INPUT; goto done; }
    {
        /* Fiat-Shamir (fk20/verifier.rs:269-328) */
        sha256_ctx s; sha256_init(&s);
        sha256_update(&s, (const uint8_t *)"RCKZGCBATCH__V1_", 16);
        uint64_t hdr[4] = {N_BLOB, CELL_ELEMS, (uint64_t)m, (uint64_t)n};
        for (int k = 0; k < 4; k++) { uint8_t u[8]; for (int i = 0; i < 8; i++) u[i] = (uint8_t)(hdr[k] >> (56 - 8 * i)); sha256_update(&s, u, 8); }
        for (int j = 0; j < m; j++) sha256_update(&s, commitments + 48 * uniq[j], 48);
        for (int k = 0; k < n; k++) {
            uint8_t u[16];
            for (int i = 0; i < 8; i++) { u[i] = (uint8_t)((uint64_t)row[k] >> (56 - 8 * i)); u[8 + i] = (uint8_t)(cell_idx[k] >> (56 - 8 * i)); }
            sha256_update(&s, u, 16); sha256_update(&s, cells + (size_t)k * BYTES_PER_CELL, BYTES_PER_CELL); sha256_update(&s, proofs + 48 * k, 48);
        }
        uint8_t h[32]; sha256_final(&s, h);
        fr_t r; fr_from_be_reduce(&r, h);
        fr_t *rp = malloc(sizeof(fr_t) * n), *wrp = malloc(sizeof(fr_t) * n), *wts = calloc(m, sizeof(fr_t));
        fr_t p; fr_set_one(&p);
        for (int k = 0; k < n; k++) {
            rp[k] = p; fr_mul(&wrp[k], &p, &G.coset_gens_pow_n[cell_idx[k]]);
            fr_add(&wts[row[k]], &wts[row[k]], &p);
            fr_mul(&p, &p, &r);
        }
        g1_t sum_proofs, wsum_proofs, sum_cm, icm;
        g1_lincomb(&sum_proofs, prs, rp, n); g1_lincomb(&wsum_proofs, prs, wrp, n); g1_lincomb(&sum_cm, cms, wts, m);
        /* compute_sum_interpolation_poly (fk20/verifier.rs:348-384) */
        fr_t I[CELL_ELEMS]; for (int i = 0; i < CELL_ELEMS; i++) fr_set_zero(&I[i]);
        for (int k = 0; k < n; k++) {
            fr_t t[CELL_ELEMS]; memcpy(t, ev + (size_t)k * CELL_ELEMS, sizeof(t));
            brp_fr(t, CELL_ELEMS);
            coset_ifft_fr(&G.d64, t, &G.coset_gens_inv[cell_idx[k]]);
            for (int i = 0; i < CELL_ELEMS; i++) { fr_t x; fr_mul(&x, &t[i], &rp[k]); fr_add(&I[i], &I[i], &x); }
        }
        g1_lincomb(&icm, G.srs_g1, I, CELL_ELEMS);
        g1_t rl; g1_sub(&rl, &sum_cm, &icm); g1_add(&rl, &rl, &wsum_proofs);
        g1a_t ps[2]; g2a_t qs[2];
        g1_to_affine(&ps[0], &sum_proofs); qs[0] = G.g2_tau64;
        g1_to_affine(&ps[1], &rl); qs[1] = G.g2_neg_gen;
        *ok = pairing_check(ps, qs, 2);
        free(rp); free(wrp); free(wts);
    }
done:
    free(row); free(uniq); free(cms); free(prs); free(ev);
    return rc;
}

/* ------------------------------------------------------------------ primitive hooks for tests */
/* n dependent Fp Montgomery multiplications on one core; returns nanoseconds per multiplication (bench.py prints it
 * beside the CPU baseline so that this port's distance from blst's hand-written assembly is a number, not a guess) */
double okzg_time_fp_mul(int n, uint8_t sink[48]) {
    fp_t x, y;
    memcpy(x.l, FP_R2, 48);
    memcpy(y.l, FP_R2, 48);
    y.l[0] ^= 0x9e3779b97f4a7c15ull; y.l[5] &= 0x0fffffffffffffffull;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n; i++) fp_mul(&x, &x, &y);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fp_to_be(sink, &x);
    return ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / n;
}
void okzg_test_fp_mul(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) {
    fp_t x, y, z; fp_from_be(&x, a); fp_from_be(&y, b); fp_mul(&z, &x, &y); fp_to_be(out, &z);
}
void okzg_test_fr_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    fr_t x, y, z; fr_from_be(&x, a); fr_from_be(&y, b); fr_mul(&z, &x, &y); fr_to_be(out, &z);
}
int okzg_test_g1_mul(const uint8_t p[48], const uint8_t k[32], uint8_t out[48]) {
    g1a_t a; if (!g1a_decompress(&a, p, 0)) return 1;
    fr_t s; fr_from_be_reduce(&s, k);
    g1_t j, r1, r2; g1_from_affine(&j, &a); g1_mul_fr(&r1, &j, &s); g1_mul_fr_w4(&r2, &j, &s);
    uint8_t o2[48]; g1_compress(out, &r1); g1_compress(o2, &r2);
    return memcmp(out, o2, 48) ? 2 : 0;
}
int okzg_test_g1_lincomb(int n, const uint8_t *pts, const uint8_t *ks, uint8_t out[48]) {
    g1a_t *P = malloc(sizeof(g1a_t) * n); fr_t *S = malloc(sizeof(fr_t) * n);
    for (int i = 0; i < n; i++) { if (!g1a_decompress(&P[i], pts + 48 * i, 0)) return 1; fr_from_be_reduce(&S[i], ks + 32 * i); }
    g1_t r; g1_lincomb(&r, P, S, n); g1_compress(out, &r); free(P); free(S); return 0;
}
int okzg_test_g1_decompress(const uint8_t p[48], int check_subgroup) { g1a_t a; return g1a_decompress(&a, p, check_subgroup); }
void okzg_test_fft_fr(int n, int inverse, uint8_t *io /* n*32 BE */) {
    domain_t d; domain_init(&d, n);
    fr_t *v = malloc(sizeof(fr_t) * n);
    for (int i = 0; i < n; i++) fr_from_be(&v[i], io + 32 * i);
    if (inverse) ifft_fr(&d, v); else fft_fr(&d, v);
    for (int i = 0; i < n; i++) fr_to_be(io + 32 * i, &v[i]);
    free(v); free(d.roots); free(d.roots_inv);
}
void okzg_test_sha256(const uint8_t *p, size_t n, uint8_t out[32]) { sha256_ctx s; sha256_init(&s); sha256_update(&s, p, n); sha256_final(&s, out); }
/* FK20 MSM stage only (for kernel-level parity): scalars s_j[k] as 128*64 BE32, outputs 128 compressed R_j */
void okzg_test_fk20_msm(const uint8_t *scalars_be, uint8_t *out) {
    for (int j = 0; j < 128; j++) {
        fr_t s[64]; for (int k = 0; k < 64; k++) fr_from_be_reduce(&s[k], scalars_be + 32 * (j * 64 + k));
        g1_t r; fixed_base_msm(&r, G.fk20_table + (size_t)j * 64 * PRECOMP_ENTRIES, s); g1_compress(out + 48 * j, &r);
    }
}
/* FK20 intermediates of one blob for stage-by-stage kernel parity (tests/test_gpu_stages.py):
 * scalars: A_k[j] as BE32 at [(j*64+k)*32]; msm: 128 compressed R_j; h: 64 compressed h-commitments
 * (fk20/batch_toeplitz.rs:94-124). */
int okzg_test_fk20_stages(const uint8_t *blob, uint8_t *scalars, uint8_t *msm, uint8_t *h) {
    fr_t *c = malloc(sizeof(fr_t) * N_BLOB);
    if (!blob_to_scalars(c, blob)) { free(c); return OKZG_ERR_INPUT; }
    scalars_to_coeffs(&G, c);
    fr_t (*A)[128] = malloc(sizeof(fr_t) * 64 * 128);
    for (int k = 0; k < 64; k++) {
        fr_t *a = A[k];
        for (int i = 0; i < 128; i++) fr_set_zero(&a[i]);
        a[0] = c[4095 - k];
        for (int i = 1; i < 64; i++) a[128 - i] = c[4095 - k - 64 * i];
        fft_fr(&G.d128, a);
    }
    g1_t R[128];
    for (int j = 0; j < 128; j++) {
        fr_t s[64];
        for (int k = 0; k < 64; k++) { s[k] = A[k][j]; fr_to_be(scalars + 32 * (j * 64 + k), &s[k]); }
        fixed_base_msm(&R[j], G.fk20_table + (size_t)j * 64 * PRECOMP_ENTRIES, s);
        g1_compress(msm + 48 * j, &R[j]);
    }
    ifft_g1_take_n(&G.d128, R, 64);
    for (int i = 0; i < 64; i++) g1_compress(h + 48 * i, &R[i]);
    free(A); free(c);
    return OKZG_OK;
}
