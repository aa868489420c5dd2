"""Differential fuzzing of the verifiers and provers through the C ABI against the CPU oracle: valid inputs with one random byte
(or one random bit) changed somewhere -- in a commitment, a proof, an evaluation point, a value, a cell, a blob -- must be classified
exactly as the oracle classifies them: true / false / Err (`None` here).  Seeded, so a failure is reproducible; the consensus vectors
cover the named edge cases, this covers the unnamed ones (flags in the top bits of a compressed point, x just above p, a value just
above r, an off-curve x, a point of the curve outside G1 ...)."""
import importlib
import random

import pytest

pytestmark = pytest.mark.gpu


def _call(pkg, f, *a):
    try:
        return f(*a)
    except pkg.KzgError:
        return None


def _ocall(f, *a):
    try:
        return f(*a)
    except Exception:
        return None


def _mutate(rng, b):
    b = bytearray(b)
    i = rng.randrange(len(b))
    if rng.random() < 0.5:
        b[i] ^= 1 << rng.randrange(8)
    else:
        b[i] = rng.randrange(256)
    return bytes(b)


def _mutate_head(rng, b, head):
    """mutations concentrated in the first `head` bytes (the flag bits and the most significant limbs of a point / scalar)"""
    b = bytearray(b)
    i = rng.randrange(head)
    b[i] ^= 1 << rng.randrange(8)
    return bytes(b)


@pytest.fixture(scope="module")
def material(das_ctx, pkg):
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    blob = syn.blob(4242)
    cm = das_ctx.blob_to_kzg_commitment(blob)
    z = (987654321).to_bytes(32, "big")
    pz, y = das_ctx.compute_kzg_proof(blob, z)
    pb = das_ctx.compute_blob_kzg_proof(blob, cm)
    cells, proofs = das_ctx.compute_cells_and_kzg_proofs(blob)
    return {"blob": blob, "cm": cm, "z": z, "y": y, "pz": pz, "pb": pb, "cells": cells, "proofs": proofs}


def test_fuzz_verify_kzg_proof(das_ctx, pkg, material):
    from oracle import cref
    rng = random.Random(20260)
    m = material
    assert das_ctx.verify_kzg_proof(m["cm"], m["z"], m["y"], m["pz"]) is True
    seen = {True: 0, False: 0, None: 0}
    for it in range(120):
        args = [m["cm"], m["z"], m["y"], m["pz"]]
        k = rng.randrange(4)
        args[k] = _mutate_head(rng, args[k], 4) if it % 3 == 0 else _mutate(rng, args[k])
        got = _call(pkg, das_ctx.verify_kzg_proof, *args)
        want = _ocall(cref.verify_kzg_proof, *args)
        assert got == want, "case %d (argument %d mutated): GPU %r, oracle %r" % (it, k, got, want)
        seen[got] += 1
    assert seen[None] > 0 and seen[False] > 0      # both rejection classes were exercised


def test_fuzz_verify_blob_kzg_proof_and_provers(das_ctx, pkg, material):
    from oracle import cref
    rng = random.Random(777)
    m = material
    for it in range(40):
        blob, cm, pb = m["blob"], m["cm"], m["pb"]
        k = rng.randrange(3)
        if k == 0:
            # a field element of the blob: most significant bytes (canonicity) or anywhere
            e = rng.randrange(4096)
            blob = bytearray(blob)
            blob[32 * e + (0 if it % 2 else rng.randrange(32))] ^= 1 << rng.randrange(8)
            blob = bytes(blob)
        elif k == 1:
            cm = _mutate_head(rng, cm, 3) if it % 2 else _mutate(rng, cm)
        else:
            pb = _mutate_head(rng, pb, 3) if it % 2 else _mutate(rng, pb)
        got = _call(pkg, das_ctx.verify_blob_kzg_proof, blob, cm, pb)
        want = _ocall(cref.verify_blob_kzg_proof, blob, cm, pb)
        assert got == want, "verify_blob_kzg_proof case %d (%d): GPU %r, oracle %r" % (it, k, got, want)
        if it < 12:      # the provers on the same mutated inputs
            assert _call(pkg, das_ctx.blob_to_kzg_commitment, blob) == _ocall(cref.blob_to_kzg_commitment, blob)
            assert _call(pkg, das_ctx.compute_blob_kzg_proof, blob, cm) == _ocall(cref.compute_blob_kzg_proof, blob, cm)


def test_fuzz_verify_cell_kzg_proof_batch(das_ctx, pkg, material):
    from oracle import cref
    rng = random.Random(31337)
    m = material
    for it in range(30):
        n = rng.choice([1, 3, 17])
        sel = [rng.randrange(128) for _ in range(n)]
        C, I = [m["cm"]] * n, list(sel)
        CL, PR = [m["cells"][i] for i in sel], [m["proofs"][i] for i in sel]
        k, j = rng.randrange(4), rng.randrange(n)
        if k == 0:
            C = list(C); C[j] = _mutate_head(rng, C[j], 3) if it % 2 else _mutate(rng, C[j])
        elif k == 1:
            PR[j] = _mutate_head(rng, PR[j], 3) if it % 2 else _mutate(rng, PR[j])
        elif k == 2:
            e = rng.randrange(64)
            c = bytearray(CL[j]); c[32 * e + (0 if it % 2 else rng.randrange(32))] ^= 1 << rng.randrange(8); CL[j] = bytes(c)
        else:
            I[j] = rng.choice([I[j] ^ 1, 128, 129, 2 ** 40])
        got = _call(pkg, das_ctx.verify_cell_kzg_proof_batch, C, I, CL, PR)
        want = _ocall(cref.verify_cell_kzg_proof_batch, C, I, CL, PR)
        assert got == want, "case %d (kind %d, item %d of %d): GPU %r, oracle %r" % (it, k, j, n, got, want)
