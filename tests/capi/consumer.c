/* A plain C99 consumer of include/c_eth_kzg.h, as a binding's native side would be (bindings/golang/prover.go:4-9 is cgo over
   this header).  Without a GPU it can only check what needs no device: constants, null-safe frees, and that context creation
   fails cleanly (NULL) instead of falling back to anything.  With a GPU (argv[1] == "gpu") it runs one blob through the ABI and
   checks that cells 0..63 reproduce the blob. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "c_eth_kzg.h"

int main(int argc, char **argv) {
    if (eth_kzg_constant_bytes_per_cell() != 2048 || eth_kzg_constant_bytes_per_proof() != 48 || eth_kzg_constant_cells_per_ext_blob() != 128) return 10;
    eth_kzg_das_context_free(NULL);
    eth_kzg_free_error_message(NULL);
    DASContext *ctx = eth_kzg_das_context_new(false);
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
        if (ctx != NULL) { eth_kzg_das_context_free(ctx); printf("context created (a device is present)\n"); return 0; }
        printf("no device: context creation refused\n");
        return 0;
    }
    if (!ctx) return 11;
    static uint8_t blob[131072];
    for (int i = 0; i < 4096; i++) { blob[32 * i + 31] = (uint8_t)i; blob[32 * i + 30] = (uint8_t)(i >> 8); }
    uint8_t *cells[128], *proofs[128];
    for (int i = 0; i < 128; i++) { cells[i] = malloc(2048); proofs[i] = malloc(48); }
    CResult r = eth_kzg_compute_cells_and_kzg_proofs(ctx, blob, cells, proofs);
    if (r.status != Ok) { fprintf(stderr, "error: %s\n", r.error_msg ? r.error_msg : "?"); eth_kzg_free_error_message(r.error_msg); return 12; }
    for (int i = 0; i < 64; i++) if (memcmp(cells[i], blob + 2048 * i, 2048) != 0) return 13;
    blob[0] = 0xff;   /* first field element >= r: must be an Err with a message, not a crash */
    r = eth_kzg_compute_cells_and_kzg_proofs(ctx, blob, cells, proofs);
    if (r.status != Err || r.error_msg == NULL) return 14;
    eth_kzg_free_error_message(r.error_msg);
    eth_kzg_das_context_free(ctx);
    printf("gpu path ok\n");
    return 0;
}
