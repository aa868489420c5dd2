#!/usr/bin/env python3
"""Small end-to-end run of every entry point, meant to be executed under compute-sanitizer on the GPU box:
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
(the kernels are 10-100x slower under the tool, so the batch sizes are tiny; EKZG_FK20_WINDOW / EKZG_SRS_WINDOW pick the table
widths, default here: 10 / 9 with use_precomp so that the merged-top-window code paths are NOT taken, then 12 / 12 so that they are)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402


def run(pkg, syn, fk20_w, srs_w, nb):
    os.environ["EKZG_FK20_WINDOW"], os.environ["EKZG_SRS_WINDOW"] = fk20_w, srs_w
    ctx = pkg.DASContext(use_precomp=True)
    blobs = [syn.blob(40 + i) for i in range(nb)]
    flat = b"".join(blobs)
    cells, proofs, st = ctx.compute_cells_and_kzg_proofs_batch(flat, nb)
    c1, p1 = ctx.compute_cells_and_kzg_proofs(blobs[0])
    assert b"".join(c1) == cells[:262144] and b"".join(p1) == proofs[:6144]
    com, _ = ctx.blob_to_kzg_commitment_batch(flat, nb)
    prf, _ = ctx.compute_blob_kzg_proof_batch(flat, com, nb)
    assert ctx.verify_blob_kzg_proof(blobs[0], com[:48], prf[:48]) is True
    idx = list(range(0, 128, 2))
    rc, rp = ctx.recover_cells_and_kzg_proofs(idx, [c1[j] for j in idx])
    assert rc == c1 and rp == p1
    z = (777).to_bytes(32, "big")
    pz, y = ctx.compute_kzg_proof(blobs[0], z)
    assert ctx.verify_kzg_proof(com[:48], z, y, pz) is True
    assert ctx.verify_blob_kzg_proof_batch(blobs[:2], [com[:48], com[48:96]], [prf[:48], prf[48:96]]) is True
    os.environ["EKZG_K5_R4_MAX"] = "0"          # the wide (radix-2) G1-NTT kernel on the same blob
    c2, p2 = ctx.compute_cells_and_kzg_proofs(blobs[0])
    del os.environ["EKZG_K5_R4_MAX"]
    assert p2 == p1
    # latency mode of the G1 NTTs with identity points in some lanes: cooperative form (four lanes per field element) against the
    # lane-per-blob form, 9 blobs = one full group of eight + a ragged one
    edge = [syn.blob(70 + i) for i in range(5)] + list(syn.edge_blobs())[:4]
    os.environ["EKZG_DIRECT_MAX"] = "0"
    got = ctx.compute_cells_and_kzg_proofs_batch(b"".join(edge), 9)
    os.environ["EKZG_K5_COOP_MAX"] = "0"
    want = ctx.compute_cells_and_kzg_proofs_batch(b"".join(edge), 9)
    del os.environ["EKZG_K5_COOP_MAX"], os.environ["EKZG_DIRECT_MAX"]
    assert got[1] == want[1] and got[0] == want[0]
    sel = [0, 5, 64, 127]
    assert ctx.verify_cell_kzg_proof_batch([com[:48]] * 4, sel, [c1[j] for j in sel], [p1[j] for j in sel]) is True
    assert ctx.verify_cell_kzg_proof_batch([com[:48]] * 4, sel, [c1[j] for j in sel], [p1[j] for j in (5, 0, 64, 127)]) is False
    ctx.close()
    print("ok", fk20_w, srs_w, nb, flush=True)


def main():
    pkg = __graft_entry__.load_package()
    import importlib
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    run(pkg, syn, "10", "9", 3)
    run(pkg, syn, "10", "8", 2)      # SRS tables without a merged top window: the 64-slice MSM of the single-blob calls
    run(pkg, syn, "12", "12", 40)
    if len(sys.argv) > 1:   # a full-size batch as well (two-piece scheduler path): slow under the tools
        run(pkg, syn, "10", "9", int(sys.argv[1]))


if __name__ == "__main__":
    main()
