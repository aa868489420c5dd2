/* Native load generator for the reference's own per-item symbols (include/c_eth_kzg.h):
 *   T POSIX threads share ONE DASContext and each calls eth_kzg_compute_cells_and_kzg_proofs (or
 *   eth_kzg_recover_cells_and_proofs) in a loop, every thread with its own 128 + 128 separately allocated output buffers --
 *   exactly what a binding hands over (bindings/c/src/pointer_utils.rs:53-62 writes through such pointer arrays;
 *   bindings/node/src/lib.rs:92-130 calls from a thread pool).  No Python, no GIL: the number is the library's.
 *
 *   abi_load [--threads T] [--calls C] [--mode compute|recover] [--precomp 0|1] [--check 0|1]
 * prints one JSON line: blobs/s over all threads, per-call latency, and (with --check) whether every thread's last result
 * equals what the batch entry point gives for the same blob.
 * Build: make -C rust-eth-kzg_b200 abi_load   (gcc, links libc_eth_kzg_b200.so) */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "c_eth_kzg.h"

enum { BLOB = 131072, CELL = 2048, PROOF = 48, NCELLS = 128 };

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* deterministic canonical blob: 4096 big-endian field elements below 2^254 */
static void make_blob(uint8_t* out, uint64_t seed) {
    uint64_t x = seed * 0x9e3779b97f4a7c15ull + 0x632be59bd9b4e019ull;
    for (int i = 0; i < BLOB; i += 8) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        memcpy(out + i, &x, 8);
    }
    for (int i = 0; i < BLOB; i += 32) out[i] &= 0x3f;
}

typedef struct {
    const DASContext* ctx;
    int id, calls, mode;
    const uint8_t* blob;            /* compute: the blob; recover: unused */
    const uint8_t* const* in_cells; /* recover: 64 cell pointers */
    const uint64_t* in_idx;
    uint8_t* cells[NCELLS];
    uint8_t* proofs[NCELLS];
    pthread_barrier_t* start;
    double t_first, t_last, lat_sum, lat_max;
    int failed;
} Worker;

static void* run(void* arg) {
    Worker* w = (Worker*)arg;
    pthread_barrier_wait(w->start);
    w->t_first = now_s();
    for (int c = 0; c < w->calls; c++) {
        const double t0 = now_s();
        CResult r = w->mode == 0 ? eth_kzg_compute_cells_and_kzg_proofs(w->ctx, w->blob, w->cells, w->proofs)
                                 : eth_kzg_recover_cells_and_proofs(w->ctx, 64, w->in_cells, 64, w->in_idx, w->cells, w->proofs);
        const double dt = now_s() - t0;
        w->lat_sum += dt;
        if (dt > w->lat_max) w->lat_max = dt;
        if (r.status != Ok) {
            if (!w->failed) fprintf(stderr, "thread %d: %s\n", w->id, r.error_msg ? r.error_msg : "error");
            eth_kzg_free_error_message(r.error_msg);
            w->failed = 1;
        }
    }
    w->t_last = now_s();
    return NULL;
}

int main(int argc, char** argv) {
    int T = 64, calls = 4, mode = 0, precomp = 1, check = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--threads")) T = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--calls")) calls = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--mode")) mode = !strcmp(argv[i + 1], "recover");
        else if (!strcmp(argv[i], "--precomp")) precomp = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--check")) check = atoi(argv[i + 1]);
    }
    if (T < 1 || calls < 1) return 2;
    const double t_init0 = now_s();
    DASContext* ctx = eth_kzg_das_context_new(precomp != 0);
    if (!ctx) { fprintf(stderr, "no context (no CUDA device? there is no CPU fallback)\n"); return 3; }
    const double t_init = now_s() - t_init0;

    /* distinct blobs (at most 256 different ones: enough to catch a mixed-up result, cheap to prepare) */
    const int nb = T < 256 ? T : 256;
    uint8_t* blobs = (uint8_t*)malloc((size_t)nb * BLOB);
    for (int i = 0; i < nb; i++) make_blob(blobs + (size_t)i * BLOB, 1000 + i);
    uint8_t* ref_cells = (uint8_t*)malloc((size_t)nb * NCELLS * CELL);
    uint8_t* ref_proofs = (uint8_t*)malloc((size_t)nb * NCELLS * PROOF);
    {
        CResult r = eth_kzg_b200_compute_cells_and_kzg_proofs_batch(ctx, nb, blobs, ref_cells, ref_proofs, NULL);
        if (r.status != Ok) { fprintf(stderr, "reference batch failed: %s\n", r.error_msg); return 4; }
    }
    uint64_t idx[64];
    for (int i = 0; i < 64; i++) idx[i] = 2 * i + 1;   /* every other cell is missing */

    Worker* ws = (Worker*)calloc(T, sizeof(Worker));
    const uint8_t*** in_ptrs = (const uint8_t***)calloc(T, sizeof(void*));
    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, T + 1);
    for (int t = 0; t < T; t++) {
        Worker* w = &ws[t];
        w->ctx = ctx; w->id = t; w->calls = calls; w->mode = mode; w->start = &start;
        w->blob = blobs + (size_t)(t % nb) * BLOB;
        in_ptrs[t] = (const uint8_t**)malloc(64 * sizeof(void*));
        for (int i = 0; i < 64; i++) in_ptrs[t][i] = ref_cells + ((size_t)(t % nb) * NCELLS + idx[i]) * CELL;
        w->in_cells = in_ptrs[t];
        w->in_idx = idx;
        for (int i = 0; i < NCELLS; i++) {   /* scattered destinations, one allocation each */
            w->cells[i] = (uint8_t*)malloc(CELL);
            w->proofs[i] = (uint8_t*)malloc(PROOF);
        }
    }
    pthread_t* th = (pthread_t*)malloc(T * sizeof(pthread_t));
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 512 * 1024);
    /* warm-up pass (workspaces, pinned staging), then the timed pass */
    double rate = 0, lat_avg = 0, lat_max = 0, wall = 0;
    int failed = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int t = 0; t < T; t++) { ws[t].lat_sum = ws[t].lat_max = 0; ws[t].calls = pass ? calls : 1; }
        for (int t = 0; t < T; t++)
            if (pthread_create(&th[t], &attr, run, &ws[t]) != 0) { fprintf(stderr, "pthread_create failed at %d\n", t); return 5; }
        pthread_barrier_wait(&start);
        const double t0 = now_s();
        for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
        wall = now_s() - t0;
        if (pass) {
            rate = (double)T * calls / wall;
            for (int t = 0; t < T; t++) {
                lat_avg += ws[t].lat_sum / calls / T;
                if (ws[t].lat_max > lat_max) lat_max = ws[t].lat_max;
                failed |= ws[t].failed;
            }
        }
    }
    int mismatches = 0;
    if (check) {
        for (int t = 0; t < T; t++) {
            const uint8_t* rc = ref_cells + (size_t)(t % nb) * NCELLS * CELL;
            const uint8_t* rp = ref_proofs + (size_t)(t % nb) * NCELLS * PROOF;
            int bad = 0;
            for (int i = 0; i < NCELLS; i++) bad |= memcmp(ws[t].cells[i], rc + (size_t)i * CELL, CELL) != 0 || memcmp(ws[t].proofs[i], rp + (size_t)i * PROOF, PROOF) != 0;
            mismatches += bad;
        }
    }
    printf("{\"tool\": \"abi_load\", \"symbol\": \"%s\", \"threads\": %d, \"calls_per_thread\": %d, \"blobs_per_s\": %.1f, \"wall_s\": %.4f, "
           "\"latency_ms_avg\": %.3f, \"latency_ms_max\": %.3f, \"context_init_s\": %.2f, \"fk20_window_bits\": %d, \"devices\": %d, "
           "\"checked_threads\": %d, \"mismatches\": %d, \"failed\": %s}\n",
           mode ? "eth_kzg_recover_cells_and_proofs" : "eth_kzg_compute_cells_and_kzg_proofs", T, calls, rate, wall, 1e3 * lat_avg, 1e3 * lat_max, t_init,
           eth_kzg_b200_context_window(ctx), eth_kzg_b200_context_device_count(ctx), check ? T : 0, mismatches, failed ? "true" : "false");
    eth_kzg_das_context_free(ctx);
    return (failed || mismatches) ? 1 : 0;
}
