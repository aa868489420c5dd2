"""Pins the C oracle (oracle/src/kzg.c) against ALL consensus vectors of the reference
(test_vectors/*, 311 cases) -- the same checks as crates/eip7594/tests/*.rs and
crates/eip4844/tests/*.rs: byte-exact outputs, `output: null` <=> error, verification booleans."""
import pytest

from oracle import cref
from tests import vectors


def _run(fn, inp):
    if fn == "blob_to_kzg_commitment":
        return cref.blob_to_kzg_commitment(inp["blob"])
    if fn == "compute_blob_kzg_proof":
        return cref.compute_blob_kzg_proof(inp["blob"], inp["commitment"])
    if fn == "compute_cells_and_kzg_proofs":
        c, p = cref.compute_cells_and_kzg_proofs(inp["blob"])
        return [c, p]
    if fn == "compute_kzg_proof":
        p, y = cref.compute_kzg_proof(inp["blob"], inp["z"])
        return [p, y]
    if fn == "recover_cells_and_kzg_proofs":
        c, p = cref.recover_cells_and_kzg_proofs(inp["cell_indices"], inp["cells"])
        return [c, p]
    if fn == "verify_blob_kzg_proof":
        return cref.verify_blob_kzg_proof(inp["blob"], inp["commitment"], inp["proof"])
    if fn == "verify_blob_kzg_proof_batch":
        return cref.verify_blob_kzg_proof_batch(inp["blobs"], inp["commitments"], inp["proofs"])
    if fn == "verify_cell_kzg_proof_batch":
        return cref.verify_cell_kzg_proof_batch(inp["commitments"], inp["cell_indices"], inp["cells"], inp["proofs"])
    if fn == "verify_kzg_proof":
        return cref.verify_kzg_proof(inp["commitment"], inp["z"], inp["y"], inp["proof"])
    raise AssertionError(fn)


def _cases():
    for fn in vectors.FUNCTIONS:
        for name, inp, out in vectors.load(fn):
            yield pytest.param(fn, inp, out, id=name)


@pytest.mark.parametrize("fn,inp,expected", list(_cases()))
def test_oracle_matches_consensus_vector(fn, inp, expected):
    try:
        got = _run(fn, inp)
    except cref.OracleError:
        got = None
    if isinstance(expected, list):
        expected = [list(x) if isinstance(x, list) else x for x in expected]
    assert got == expected
