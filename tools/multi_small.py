#!/usr/bin/env python3
"""Small batches through ONE context over several devices (EKZG_DEVICES=all) against one device: the shards are groups of 8 blobs,
so every device runs the cooperative latency-mode G1-NTT kernel.  One JSON line per batch size; median of 9 calls after 3 warm-ups,
preallocated host buffers through eth_kzg_b200_compute_cells_and_kzg_proofs_batch."""
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__  # noqa: E402

pkg = __graft_entry__.load_package()
import importlib  # noqa: E402
syn = importlib.import_module("eth_kzg_b200.synthetic")
lib = pkg.load_library()
ndev = torch.cuda.device_count()
blobs_all = syn.blobs(128)
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "8,16,32,64,128".split(","))]
rows = {n: {"blobs": n, "devices": ndev} for n in sizes}
proofs = {}
# one context at a time: two contexts on one device would share its HBM and the second would get narrower tables
for name, devs in (("one_device_ms", None), ("all_devices_ms", "all")):
    os.environ.pop("EKZG_DEVICES", None)
    if devs:
        os.environ["EKZG_DEVICES"] = devs
    ctx = pkg.DASContext(use_precomp=True)
    os.environ.pop("EKZG_DEVICES", None)
    for n in sizes:
        src = bytearray(blobs_all[:n * 131072])
        bc, bp, bs = bytearray(n * 262144), bytearray(n * 6144), bytearray(n)
        cb, cc, cp, cs = [(C.c_char * len(x)).from_buffer(x) for x in (src, bc, bp, bs)]
        ts = []
        for it in range(12):
            t0 = time.perf_counter()
            res = lib.eth_kzg_b200_compute_cells_and_kzg_proofs_batch(C.c_void_p(ctx.handle), C.c_uint64(n), cb, cc, cp, cs)
            ts.append((time.perf_counter() - t0) * 1e3)
            assert res.status == 0
        rows[n][name] = round(statistics.median(ts[3:]), 3)
        rows[n]["same"] = proofs.setdefault(n, bytes(bp)) == bytes(bp)
    ctx.close()
for n in sizes:
    print(json.dumps(rows[n]), flush=True)
