#!/usr/bin/env python3
"""verify_cell_kzg_proof_batch on ONE blob's 128 cells (the reference's own bench shape, crates/eip7594/benches/benchmark-mt.rs:77-101)
with the host-phase trace on (EKZG_TRACE=1): where the milliseconds go"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["EKZG_TRACE"] = "1"
import __graft_entry__  # noqa: E402
pkg = __graft_entry__.load_package()
import importlib  # noqa: E402
syn = importlib.import_module("eth_kzg_b200.synthetic")
ctx = pkg.DASContext(use_precomp=True)
blob = syn.blob(5)
cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
cm = ctx.blob_to_kzg_commitment(blob)
for n in (128, 8):
    for rep in range(4):
        t0 = time.perf_counter()
        ok = ctx.verify_cell_kzg_proof_batch([cm] * n, list(range(n)), cells[:n], proofs[:n])
        print("== %d cells: %s in %.3f ms (python wall clock incl. ctypes marshalling)" % (n, ok, 1e3 * (time.perf_counter() - t0)), file=sys.stderr, flush=True)
ctx.close()
