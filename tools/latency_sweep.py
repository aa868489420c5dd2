#!/usr/bin/env python3
"""Small-batch latency of compute_cells_and_kzg_proofs through the host-buffer batch entry point (copies included), by route:
direct (every proof its own SRS MSM), FK20 with the radix-2 G1-NTT kernel, with the radix-4 latency-mode kernel, and with the
latter's cooperative form (four lanes per field element).  One JSON line per batch size; median of 9 calls after 3 warm-ups."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

pkg = __graft_entry__.load_package()
import importlib  # noqa: E402
syn = importlib.import_module("eth_kzg_b200.synthetic")
ctx = pkg.DASContext(use_precomp=True)
blobs_all = syn.blobs(128)
ROUTES = {
    "direct": {"EKZG_DIRECT_MAX": "8"},
    "fk20_radix2": {"EKZG_DIRECT_MAX": "0", "EKZG_K5_R4_MAX": "0"},
    "fk20_radix4": {"EKZG_DIRECT_MAX": "0", "EKZG_K5_R4_MAX": "256", "EKZG_K5_COOP_MAX": "0"},
    "fk20_radix4_coop": {"EKZG_DIRECT_MAX": "0", "EKZG_K5_R4_MAX": "256", "EKZG_K5_COOP_MAX": "256"},
    "default": {},
}
for n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2,3,4,6,8,16,32,64,128".split(","))]:
    flat = blobs_all[:n * 131072]
    row = {"blobs": n}
    ref = None
    for name, env in ROUTES.items():
        if name == "direct" and n > 8:
            continue
        for k in ("EKZG_DIRECT_MAX", "EKZG_K5_R4_MAX", "EKZG_K5_COOP_MAX"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ts = []
        for it in range(12):
            t0 = time.perf_counter()
            out = ctx.compute_cells_and_kzg_proofs_batch(flat, n)
            ts.append((time.perf_counter() - t0) * 1e3)
        if ref is None:
            ref = out[1]
        row[name + "_ms"] = round(statistics.median(ts[3:]), 3)
        row[name + "_same"] = out[1] == ref
    print(json.dumps(row), flush=True)
ctx.close()
