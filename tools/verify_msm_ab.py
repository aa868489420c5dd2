#!/usr/bin/env python3
"""verify_cell_kzg_proof_batch: the two random-linear-combination sums as one ladder per point against the bucket-method MSM (K7),
EKZG_VERIFY_MSM=ladder|bucket, at several batch sizes; EKZG_TRACE gives the device time after the transcript hash"""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402
pkg = __graft_entry__.load_package()
import importlib  # noqa: E402
import ctypes as C  # noqa: E402
syn = importlib.import_module("eth_kzg_b200.synthetic")
ctx = pkg.DASContext(use_precomp=True)
lib = pkg.load_library()
H = C.c_void_p(ctx.handle)
nb = 256
blobs = syn.blobs(nb)
cells_flat, proofs_flat, st = ctx.compute_cells_and_kzg_proofs_batch(blobs, nb)
comms, st2 = ctx.blob_to_kzg_commitment_batch(blobs, nb)
N = nb * 128
b_comm, b_cells, b_proofs = C.create_string_buffer(comms, nb * 48), C.create_string_buffer(cells_flat, N * 2048), C.create_string_buffer(proofs_flat, N * 48)
a0, a1, a2 = C.addressof(b_comm), C.addressof(b_cells), C.addressof(b_proofs)
pc = (C.c_void_p * N)(*[a0 + 48 * (k // 128) for k in range(N)])
pl = (C.c_void_p * N)(*[a1 + 2048 * k for k in range(N)])
pp = (C.c_void_p * N)(*[a2 + 48 * k for k in range(N)])
idx = (C.c_uint64 * N)(*[k % 128 for k in range(N)])
flag = C.c_bool(False)


def call(count):
    r = lib.eth_kzg_verify_cell_kzg_proof_batch(H, C.c_uint64(count), pc, C.c_uint64(count), idx, C.c_uint64(count), pl, C.c_uint64(count), pp, C.byref(flag))
    assert r.status == 0 and flag.value is True


for count in (128, 1024, 4096, 8192, 16384, 32768):
    row = {"cells": count}
    for form in ("ladder", "bucket"):
        os.environ["EKZG_VERIFY_MSM"] = form
        call(count)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            call(count)
            ts.append(time.perf_counter() - t0)
        row[form + "_ms"] = round(1e3 * sorted(ts)[2], 3)
    print(json.dumps(row), flush=True)
ctx.close()
