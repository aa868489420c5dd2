#!/usr/bin/env python3
"""Large-scale self-consistency soak on the GPU: the same synthetic blobs through two contexts with DIFFERENT table layouts
(FK20 window 14 + merged top window / SRS window 13  vs  FK20 window 10 / SRS window 9) must give identical cells, proofs,
commitments and blob proofs.  The two paths share the field arithmetic but feed it different operands, so a rare arithmetic
slip (a lost carry is a 2^-32 event per multiplication row on random data) shows up as a mismatch: N = 16384 blobs are
~4 * 10^10 Fp multiplications per context.  The small-window context is itself pinned to the oracle and the consensus vectors
by tests/test_gpu_fk20.py.   python tools/soak_parity.py [--blobs N] > profiles/..."""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blobs", type=int, default=16384)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--mixed", type=int, default=1, help="cycle the batch size through 1024 / 256 / 96 / 33 / 512 / 7 (the latency-mode "
                    "G1-NTT kernel on context A below 257 blobs, the wide kernel forced on context B) instead of --batch every round")
    ap.add_argument("--small", type=int, default=0, help="cycle the batch size through 3 / 5 / 8 / 13 / 16 / 24 / 40 / 64 / 80 instead: context A runs "
                    "its defaults (one-point-per-thread MSM, cooperative latency-mode G1 NTTs), context B the 16-slice MSM and the radix-2 kernel")
    args = ap.parse_args()
    pkg = __graft_entry__.load_package()
    import importlib
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    os.environ["EKZG_FK20_WINDOW"], os.environ["EKZG_SRS_WINDOW"] = "14", "13"
    a = pkg.DASContext(use_precomp=True)
    os.environ["EKZG_FK20_WINDOW"], os.environ["EKZG_SRS_WINDOW"] = "10", "9"
    b = pkg.DASContext(use_precomp=True)
    assert (a.window, b.window) == (14, 10)
    import numpy as np
    # one batch of real synthetic blobs ...
    base_arr = np.frombuffer(b"".join(syn.blob(70000 + i) for i in range(args.batch)), dtype=np.uint8).reshape(-1, 32)
    mism = {"cells": 0, "proofs": 0, "commitments": 0, "blob_proofs": 0}
    t0 = time.time()
    done = 0
    rnd = 0
    while done < args.blobs:
        # ... re-keyed per round by XOR-ing a round-dependent 31-byte pattern into every field element (top byte untouched: canonical)
        key = np.frombuffer(hashlib.sha256(b"soak" + rnd.to_bytes(4, "little")).digest()[1:], dtype=np.uint8)
        arr = base_arr.copy()
        arr[:, 1:] ^= key
        n = (1024, 256, 96, 33, 512, 7)[rnd % 6] if args.mixed else args.batch
        if args.small:
            n = (3, 5, 8, 13, 16, 24, 40, 64, 80)[rnd % 9]
        n = min(n, args.batch)
        flat = arr[:n * 4096].tobytes()
        os.environ.pop("EKZG_K5_R4_MAX", None)
        os.environ.pop("EKZG_K4_NO_TINY", None)
        ca, pa, _ = a.compute_cells_and_kzg_proofs_batch(flat, n)
        os.environ["EKZG_K5_R4_MAX"] = "0"      # context B: always the radix-2 G1-NTT kernel (and never the 64-slice MSM)
        os.environ["EKZG_K4_NO_TINY"] = "1"
        cb, pb, _ = b.compute_cells_and_kzg_proofs_batch(flat, n)
        os.environ.pop("EKZG_K5_R4_MAX", None)
        os.environ.pop("EKZG_K4_NO_TINY", None)
        ka, _ = a.blob_to_kzg_commitment_batch(flat, n)
        kb, _ = b.blob_to_kzg_commitment_batch(flat, n)
        qa, _ = a.compute_blob_kzg_proof_batch(flat, ka, n)
        qb, _ = b.compute_blob_kzg_proof_batch(flat, kb, n)
        for i in range(n):
            mism["cells"] += ca[i * 262144:(i + 1) * 262144] != cb[i * 262144:(i + 1) * 262144]
            mism["proofs"] += pa[i * 6144:(i + 1) * 6144] != pb[i * 6144:(i + 1) * 6144]
            mism["commitments"] += ka[i * 48:(i + 1) * 48] != kb[i * 48:(i + 1) * 48]
            mism["blob_proofs"] += qa[i * 48:(i + 1) * 48] != qb[i * 48:(i + 1) * 48]
        done += n
        rnd += 1
    print(json.dumps({"check": "two table layouts, identical outputs", "blobs": done, "rounds": rnd, "batch_sizes": "small (3..80)" if args.small else ("mixed" if args.mixed else args.batch),
                      "windows": [[14, 13], [10, 9]], "mismatches": mism,
                      "fp_multiplications_per_context": "~%.1e" % (done * 2.6e6), "seconds": round(time.time() - t0, 1)}))
    a.close()
    b.close()
    sys.exit(1 if any(mism.values()) else 0)


if __name__ == "__main__":
    main()
