"""GPU parity tests of the EIP-4844 prover side through the C ABI (BASELINE.json config #2):
blob_to_kzg_commitment, compute_blob_kzg_proof, compute_kzg_proof -- consensus vectors of the reference
(crates/eip4844/tests/*.rs read the same YAML) and the oracle on synthetic / edge blobs.  Bit-exact."""
import importlib

import pytest

from tests import vectors

pytestmark = pytest.mark.gpu


def _cases(fn):
    return [pytest.param(n, i, o, id=n) for n, i, o in vectors.load(fn)]


@pytest.mark.parametrize("name,inp,expected", _cases("blob_to_kzg_commitment"))
def test_blob_to_kzg_commitment_vectors(vec_ctx, pkg, name, inp, expected):
    try:
        got = vec_ctx.blob_to_kzg_commitment(inp["blob"])
    except pkg.KzgError:
        got = None
    assert got == expected


@pytest.mark.parametrize("name,inp,expected", _cases("compute_blob_kzg_proof"))
def test_compute_blob_kzg_proof_vectors(vec_ctx, pkg, name, inp, expected):
    try:
        got = vec_ctx.compute_blob_kzg_proof(inp["blob"], inp["commitment"])
    except pkg.KzgError:
        got = None
    assert got == expected


@pytest.mark.parametrize("name,inp,expected", _cases("compute_kzg_proof"))
def test_compute_kzg_proof_vectors(vec_ctx, pkg, name, inp, expected):
    try:
        got = list(vec_ctx.compute_kzg_proof(inp["blob"], inp["z"]))
    except pkg.KzgError:
        got = None
    assert got == (list(expected) if expected is not None else None)


def test_commit_and_proof_batch_matches_oracle(das_ctx, pkg):
    """config #2 shape at oracle-friendly size: ragged batch of synthetic + edge blobs"""
    from oracle import cref
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    blobs = syn.edge_blobs() + [syn.blob(100 + i) for i in range(29)]
    n = len(blobs)
    flat = b"".join(blobs)
    cms, st = das_ctx.blob_to_kzg_commitment_batch(flat, n)
    assert st == [0] * n
    for i in range(n):
        assert cms[48 * i:48 * i + 48] == cref.blob_to_kzg_commitment(blobs[i]), "commitment %d" % i
    prs, st = das_ctx.compute_blob_kzg_proof_batch(flat, cms, n)
    assert st == [0] * n
    for i in range(n):
        assert prs[48 * i:48 * i + 48] == cref.compute_blob_kzg_proof(blobs[i], cms[48 * i:48 * i + 48]), "proof %d" % i
    # the pair (commitment, proof) verifies under the oracle's pairing check
    for i in (0, 2, 5, n - 1):
        assert cref.verify_blob_kzg_proof(blobs[i], cms[48 * i:48 * i + 48], prs[48 * i:48 * i + 48])


def test_batch_invalid_items_are_isolated(das_ctx, pkg):
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    good = syn.blob(5)
    bad = bytearray(good)
    bad[0:32] = b"\xff" * 32
    cms, st = das_ctx.blob_to_kzg_commitment_batch(good + bytes(bad) + good, 3)
    assert st == [0, 1, 0] and cms[:48] == cms[96:144]
    # a commitment that is on the curve but not in the prime-order subgroup / not on the curve at all
    not_on_curve = bytes([0x80]) + bytes(46) + bytes([0x01])
    prs, st = das_ctx.compute_blob_kzg_proof_batch(good + good, cms[:48] + not_on_curve, 2)
    assert st[0] == 0 and st[1] == 2
