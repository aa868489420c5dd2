"""GPU parity tests of recover_cells_and_kzg_proofs through the C ABI (BASELINE.json config #4): the reference's
consensus vectors (crates/eip7594/tests/recover_cells_and_kzg_proofs.rs, incl. the two `data.yml` shuffled-order cases),
the oracle, and encode -> erase -> recover round trips over the erasure patterns of the vector classes. Bit-exact."""
import importlib
import random

import pytest

from tests import vectors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("route", ["direct", "fk20"])
@pytest.mark.parametrize("name,inp,expected", [pytest.param(n, i, o, id=n) for n, i, o in vectors.load("recover_cells_and_kzg_proofs")])
def test_recover_vectors(vec_ctx, pkg, name, inp, expected, route, monkeypatch):
    """(both routes the proofs of a single recovered blob can take: direct and, with EKZG_DIRECT_MAX=0, FK20)"""
    if route == "fk20":
        monkeypatch.setenv("EKZG_DIRECT_MAX", "0")
    try:
        cells, proofs = vec_ctx.recover_cells_and_kzg_proofs(inp["cell_indices"], inp["cells"])
        got = [cells, proofs]
    except pkg.KzgError:
        got = None
    if expected is not None:
        expected = [list(expected[0]), list(expected[1])]
    assert got == expected


def _patterns():
    rng = random.Random(11)
    return {
        "first_half": list(range(64)),                 # the reference bench's "worst case" (benchmark-mt.rs:57-60)
        "second_half": list(range(64, 128)),
        "every_other": list(range(0, 128, 2)),
        "random_64": sorted(rng.sample(range(128), 64)),
        "random_100": sorted(rng.sample(range(128), 100)),
        "all": list(range(128)),
    }


def test_encode_erase_recover_roundtrip(das_ctx, pkg):
    """batch of blobs x erasure patterns: recovery reproduces exactly what compute_cells_and_kzg_proofs produced"""
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    blobs = [syn.blob(300 + i) for i in range(5)] + syn.edge_blobs()
    n = len(blobs)
    cells, proofs, st = das_ctx.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n)
    assert st == [0] * n
    pats = list(_patterns().items())
    idx_lists, cell_lists, expect = [], [], []
    for b in range(n):
        for name, keep in pats:
            idx_lists.append(keep)
            cell_lists.append([cells[b * 262144 + k * 2048: b * 262144 + (k + 1) * 2048] for k in keep])
            expect.append((b, name))
    rc, rp, st = das_ctx.recover_cells_and_kzg_proofs_batch(idx_lists, cell_lists)
    assert st == [0] * len(idx_lists)
    for i, (b, name) in enumerate(expect):
        assert rc[i * 262144:(i + 1) * 262144] == cells[b * 262144:(b + 1) * 262144], "cells blob %d pattern %s" % (b, name)
        assert rp[i * 6144:(i + 1) * 6144] == proofs[b * 6144:(b + 1) * 6144], "proofs blob %d pattern %s" % (b, name)


def test_recover_matches_oracle(das_ctx, pkg):
    from oracle import cref
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    blob = syn.blob(777)
    ocells, _ = cref.compute_cells_and_kzg_proofs(blob)
    keep = _patterns()["random_64"]
    oc, op = cref.recover_cells_and_kzg_proofs(keep, [ocells[k] for k in keep])
    gc, gp = das_ctx.recover_cells_and_kzg_proofs(keep, [ocells[k] for k in keep])
    assert gc == oc and gp == op


def test_recover_batch_invalid_items(das_ctx, pkg):
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    blob = syn.blob(9)
    cells, proofs, _ = das_ctx.compute_cells_and_kzg_proofs_batch(blob, 1)
    cl = [cells[k * 2048:(k + 1) * 2048] for k in range(128)]
    good = list(range(0, 128, 2))
    bad_cell = list(cl)
    bad_cell[4] = b"\xff" * 32 + bad_cell[4][32:]
    # inconsistent (not from one low-degree polynomial): swap two cells
    swapped = list(cl)
    swapped[0], swapped[2] = swapped[2], swapped[0]
    items = [
        (good, [cl[k] for k in good], 0),
        (list(range(63)), [cl[k] for k in range(63)], 3),                    # not enough cells
        ([1, 0] + list(range(2, 64)), [cl[k] for k in [1, 0] + list(range(2, 64))], 3),  # not ascending
        (list(range(63)) + [128], [cl[k] for k in range(63)] + [cl[0]], 3),  # index out of range
        (good, [bad_cell[k] for k in good], 1),                              # non-canonical scalar
        (list(range(70)), [swapped[k] for k in range(70)], 4),               # degree check fails
        (good, [cl[k] for k in good], 0),
    ]
    rc, rp, st = das_ctx.recover_cells_and_kzg_proofs_batch([i[0] for i in items], [i[1] for i in items])
    assert st == [i[2] for i in items]
    for i in (0, 6):
        assert rc[i * 262144:(i + 1) * 262144] == cells and rp[i * 6144:(i + 1) * 6144] == proofs
