"""GPU parity tests of the verifiers through the C ABI: verify_cell_kzg_proof_batch (BASELINE.json config #5) and the
EIP-4844 verify_kzg_proof / verify_blob_kzg_proof / verify_blob_kzg_proof_batch -- every consensus vector of the reference
(crates/eip7594/tests/verify_cell_kzg_proof_batch.rs, crates/eip4844/tests/*.rs): true / false / Err classification."""
import importlib

import pytest

from tests import vectors

pytestmark = pytest.mark.gpu


def _cases(fn):
    return [pytest.param(n, i, o, id=n) for n, i, o in vectors.load(fn)]


def _run(pkg, f, *a):
    try:
        return f(*a)
    except pkg.KzgError:
        return None


@pytest.mark.parametrize("name,inp,expected", _cases("verify_cell_kzg_proof_batch"))
def test_verify_cell_kzg_proof_batch_vectors(vec_ctx, pkg, name, inp, expected):
    assert _run(pkg, vec_ctx.verify_cell_kzg_proof_batch, inp["commitments"], inp["cell_indices"], inp["cells"], inp["proofs"]) == expected


@pytest.mark.parametrize("name,inp,expected", _cases("verify_kzg_proof"))
def test_verify_kzg_proof_vectors(vec_ctx, pkg, name, inp, expected):
    assert _run(pkg, vec_ctx.verify_kzg_proof, inp["commitment"], inp["z"], inp["y"], inp["proof"]) == expected


@pytest.mark.parametrize("name,inp,expected", _cases("verify_blob_kzg_proof"))
def test_verify_blob_kzg_proof_vectors(vec_ctx, pkg, name, inp, expected):
    assert _run(pkg, vec_ctx.verify_blob_kzg_proof, inp["blob"], inp["commitment"], inp["proof"]) == expected


@pytest.mark.parametrize("name,inp,expected", _cases("verify_blob_kzg_proof_batch"))
def test_verify_blob_kzg_proof_batch_vectors(vec_ctx, pkg, name, inp, expected):
    assert _run(pkg, vec_ctx.verify_blob_kzg_proof_batch, inp["blobs"], inp["commitments"], inp["proofs"]) == expected


def test_verify_cells_of_many_blobs(das_ctx, pkg):
    """config #5 shape at reduced size (16 blobs x 128 cells = 2048 openings, 16 distinct commitments, commitments passed
    duplicated per cell as the API expects): accepts honest data, rejects one corrupted cell / proof / commitment."""
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    nb = 16
    blobs = [syn.blob(500 + i) for i in range(nb)]
    flat = b"".join(blobs)
    cms, st = das_ctx.blob_to_kzg_commitment_batch(flat, nb)
    cells, proofs, st2 = das_ctx.compute_cells_and_kzg_proofs_batch(flat, nb)
    assert st == [0] * nb and st2 == [0] * nb
    C, I, CL, PR = [], [], [], []
    for b in range(nb):
        for k in range(128):
            C.append(cms[48 * b:48 * b + 48]); I.append(k)
            CL.append(cells[b * 262144 + k * 2048: b * 262144 + (k + 1) * 2048])
            PR.append(proofs[b * 6144 + k * 48: b * 6144 + (k + 1) * 48])
    assert das_ctx.verify_cell_kzg_proof_batch(C, I, CL, PR) is True
    # a shuffled subset still verifies (openings are independent)
    sel = list(range(0, len(C), 7))[::-1]
    assert das_ctx.verify_cell_kzg_proof_batch([C[i] for i in sel], [I[i] for i in sel], [CL[i] for i in sel], [PR[i] for i in sel]) is True
    bad = list(CL)
    bad[1000] = bad[1000][:31] + bytes([bad[1000][31] ^ 1]) + bad[1000][32:]
    assert das_ctx.verify_cell_kzg_proof_batch(C, I, bad, PR) is False
    badp = list(PR)
    badp[77] = PR[78]
    assert das_ctx.verify_cell_kzg_proof_batch(C, I, CL, badp) is False
    badc = list(C)
    badc[5] = cms[48:96]
    assert das_ctx.verify_cell_kzg_proof_batch(badc, I, CL, PR) is False
    # the oracle agrees on a small slice
    from oracle import cref
    assert cref.verify_cell_kzg_proof_batch(C[:40], I[:40], CL[:40], PR[:40]) is True
    assert cref.verify_cell_kzg_proof_batch(C[980:1020], I[980:1020], CL[980:1020], PR[980:1020]) is True
    assert cref.verify_cell_kzg_proof_batch(C[980:1020], I[980:1020], bad[980:1020], PR[980:1020]) is False   # the slice holding the corrupted cell
    assert das_ctx.verify_cell_kzg_proof_batch(C[980:1020], I[980:1020], bad[980:1020], PR[980:1020]) is False


def _openings(ctx, pkg, nb, first):
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    flat = b"".join(syn.blob(first + i) for i in range(nb))
    cms, st = ctx.blob_to_kzg_commitment_batch(flat, nb)
    cells, proofs, st2 = ctx.compute_cells_and_kzg_proofs_batch(flat, nb)
    assert not any(st) and not any(st2)
    C = [cms[48 * (k // 128):48 * (k // 128) + 48] for k in range(nb * 128)]
    I = [k % 128 for k in range(nb * 128)]
    CL = [cells[k * 2048:(k + 1) * 2048] for k in range(nb * 128)]
    PR = [proofs[k * 48:(k + 1) * 48] for k in range(nb * 128)]
    return C, I, CL, PR


def _invalid_proofs():
    """the proofs of the reference's `invalid_proof` vectors (not on the curve / not in G1 / x >= p / bad flags)"""
    return [i["proofs"][0] for n, i, o in vectors.load("verify_cell_kzg_proof_batch") if "_invalid_proof_" in n and len(i["proofs"][0]) == 48]


@pytest.mark.parametrize("nb", [128, 264], ids=["config5_128x128", "above_32768_cells_column_sums"])
def test_verify_config5_full_size(vec_ctx, pkg, nb):
    """BASELINE config #5 at its real size (128 blobs x 128 cells = 16 384 openings in ONE call) and one call above 32 768 cells,
    where the verifier switches to per-column sums on its own (no EKZG_VERIFY_COLUMN_SUMS): honest data -> true, one corrupted
    cell -> false, one corrupted proof -> false, every malformed proof of the reference's vectors -> Err.  The oracle checks the
    128-cell slice that holds the corrupted cell (it needs ~0.1 s per 128 cells, the whole batch would take minutes)."""
    C, I, CL, PR = _openings(vec_ctx, pkg, nb, 7000)
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, CL, PR) is True
    k = len(C) - 300
    bad = list(CL)
    bad[k] = bad[k][:2047] + bytes([bad[k][2047] ^ 1])
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, bad, PR) is False
    badp = list(PR)
    badp[5], badp[6] = PR[6], PR[5]
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, CL, badp) is False
    inv = _invalid_proofs()
    assert len(inv) >= 2
    for q, pr in enumerate(inv):
        badp = list(PR)
        badp[(q * 4099 + 17) % len(PR)] = pr
        with pytest.raises(pkg.KzgError):
            vec_ctx.verify_cell_kzg_proof_batch(C, I, CL, badp)
    from oracle import cref
    lo = k - k % 128
    sl = slice(lo, lo + 128)
    assert cref.verify_cell_kzg_proof_batch(C[sl], I[sl], CL[sl], PR[sl]) is True
    assert cref.verify_cell_kzg_proof_batch(C[sl], I[sl], bad[sl], PR[sl]) is False


@pytest.mark.parametrize("chunk,nb", [(None, 9), ("4", 9), (None, 20), ("8", 20)], ids=["one_chunk", "chunks_of_4", "device_challenges", "device_challenges_chunked"])
def test_verify_blob_proof_batch_synthetic(das_ctx, pkg, chunk, nb, monkeypatch):
    """(chunks: the blobs of a batch pass through the workspace in chunks, EKZG_CHUNK -- 9 blobs = 4 + 4 + 1; up to 16 blobs the
    Fiat-Shamir challenges are hashed on the host, above that by the device kernel: 20 blobs take that path)"""
    if chunk:
        monkeypatch.setenv("EKZG_CHUNK", chunk)
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    blobs = [syn.blob(900 + i) for i in range(nb)]
    flat = b"".join(blobs)
    cms, _ = das_ctx.blob_to_kzg_commitment_batch(flat, nb)
    prs, _ = das_ctx.compute_blob_kzg_proof_batch(flat, cms, nb)
    cl = [cms[48 * i:48 * i + 48] for i in range(nb)]
    pl = [prs[48 * i:48 * i + 48] for i in range(nb)]
    assert das_ctx.verify_blob_kzg_proof_batch(blobs, cl, pl) is True
    assert all(das_ctx.verify_blob_kzg_proof(blobs[i], cl[i], pl[i]) for i in range(3))
    pl[4], pl[5] = pl[5], pl[4]
    assert das_ctx.verify_blob_kzg_proof_batch(blobs, cl, pl) is False


def test_verify_cell_batch_column_sum_path():
    """EKZG_VERIFY_COLUMN_SUMS=1 selects the throughput form of the two random-linear-combination sums (one scalar-multiplication
    pass + per-column sums + 128 fixed-twiddle multiplications) that batches above 32 768 cells use.  The switch is read once
    per process, so the vectors run in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import __graft_entry__ as g\n"
        "from tests import vectors\n"
        "pkg = g.load_package(); ctx = pkg.DASContext(use_precomp=False)\n"
        "bad = 0\n"
        "for name, inp, exp in vectors.load('verify_cell_kzg_proof_batch'):\n"
        "    try: got = ctx.verify_cell_kzg_proof_batch(inp['commitments'], inp['cell_indices'], inp['cells'], inp['proofs'])\n"
        "    except pkg.KzgError: got = None\n"
        "    bad += got != exp\n"
        "ctx.close(); print('mismatches', bad); sys.exit(1 if bad else 0)\n" % root)
    env = dict(os.environ, EKZG_VERIFY_COLUMN_SUMS="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_verify_cell_batch_bucket_msm(vec_ctx, pkg, monkeypatch):
    """EKZG_VERIFY_MSM=bucket sends the two random-linear-combination sums through the bucket-method MSM (K7, kzg_kernels_msm.cu)
    instead of one ladder per point: every consensus vector, then 40 blobs x 128 cells (5120 openings: every bucket of every
    window is hit) honest / one corrupted cell / swapped proofs, with the ladder path as the cross-check."""
    monkeypatch.setenv("EKZG_VERIFY_MSM", "bucket")
    for name, inp, expected in vectors.load("verify_cell_kzg_proof_batch"):
        got = _run(pkg, vec_ctx.verify_cell_kzg_proof_batch, inp["commitments"], inp["cell_indices"], inp["cells"], inp["proofs"])
        assert got == expected, name
    C, I, CL, PR = _openings(vec_ctx, pkg, 40, 7700)
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, CL, PR) is True
    sel = list(range(0, len(C), 3))[::-1]       # a ragged, reordered subset
    assert vec_ctx.verify_cell_kzg_proof_batch([C[i] for i in sel], [I[i] for i in sel], [CL[i] for i in sel], [PR[i] for i in sel]) is True
    bad = list(CL)
    bad[3333] = bad[3333][:100] + bytes([bad[3333][100] ^ 4]) + bad[3333][101:]
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, bad, PR) is False
    badp = list(PR)
    badp[10], badp[11] = PR[11], PR[10]
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, CL, badp) is False
    monkeypatch.setenv("EKZG_VERIFY_MSM", "ladder")
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, CL, PR) is True
    assert vec_ctx.verify_cell_kzg_proof_batch(C, I, bad, PR) is False
