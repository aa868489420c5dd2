"""GPU tests of eth_kzg_b200_das_context_new_from_json -- DASContext::new(&TrustedSetup::from_json[_unchecked](json), ..) of the
reference (crates/trusted_setup/src/lib.rs:112-127, crates/eip7594/src/trusted_setup.rs:6-23).

* The mainnet file through the loader must give the embedded context's results (consensus vectors).
* A DIFFERENT setup whose outputs are known exactly without knowing any secret: negate every odd-index point of the ceremony,
  i.e. the setup of secret -tau.  A commitment to p under -tau is the commitment to p(-X) under tau, and p(-X) in the blob's
  bit-reversed evaluation order is the blob with adjacent elements swapped (brp(i) + 2048 flips the lowest bit of i).  Cell k's coset
  is closed under negation (-1 = w_64^32), so its proof under -tau is the mainnet proof of the swapped blob as well; the cells do not
  depend on the setup at all.  Expected values come from the mainnet oracle; verification exercises [(-tau)]_2 and [tau^64]_2.
* Invalid points where the reference panics: Err, no context."""
import json
import random

import pytest

from oracle import cref
from tests import setup_util as su
from tests import vectors

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _swap_pairs(blob):
    out = bytearray(len(blob))
    for i in range(0, 4096, 2):
        out[32 * i:32 * i + 32] = blob[32 * (i + 1):32 * (i + 2)]
        out[32 * (i + 1):32 * (i + 2)] = blob[32 * i:32 * i + 32]
    return bytes(out)


@pytest.fixture(scope="module")
def mainnet_json():
    g1m, g1l, g2m = su.mainnet_points()
    return su.setup_json(g1m, g2m, g1_lagrange=g1l)


@pytest.fixture(scope="module")
def neg_ctx(pkg, das_ctx):
    g1m, _, g2m = su.mainnet_points()
    ctx = pkg.DASContext.from_json(su.setup_json(su.negate_odd(g1m), su.negate_odd(g2m)), use_precomp=False)
    yield ctx
    ctx.close()


def test_mainnet_json_equals_embedded(pkg, das_ctx, mainnet_json):
    ctx = pkg.DASContext.from_json(mainnet_json, use_precomp=False)
    try:
        assert ctx.window == das_ctx.window
        for name, inp, expected in vectors.load("compute_cells_and_kzg_proofs")[:4]:
            if expected is None:
                with pytest.raises(pkg.KzgError):
                    ctx.compute_cells_and_kzg_proofs(inp["blob"])
            else:
                cells, proofs = ctx.compute_cells_and_kzg_proofs(inp["blob"])
                assert [cells, proofs] == [list(expected[0]), list(expected[1])], name
        for suite, call in (("verify_cell_kzg_proof_batch", lambda c, i: c.verify_cell_kzg_proof_batch(i["commitments"], i["cell_indices"], i["cells"], i["proofs"])),
                            ("verify_kzg_proof", lambda c, i: c.verify_kzg_proof(i["commitment"], i["z"], i["y"], i["proof"])),
                            ("verify_blob_kzg_proof", lambda c, i: c.verify_blob_kzg_proof(i["blob"], i["commitment"], i["proof"]))):
            for name, inp, expected in vectors.load(suite)[:12]:
                try:
                    got = call(ctx, inp)
                except pkg.KzgError:
                    got = None
                assert got == expected, (suite, name)
    finally:
        ctx.close()


def test_negated_setup_against_mainnet_oracle(pkg, das_ctx, neg_ctx):
    rng = random.Random(77)
    blob = b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(4096))
    swapped = _swap_pairs(blob)
    want_commitment = cref.blob_to_kzg_commitment(swapped)
    want_cells, _ = cref.compute_cells_and_kzg_proofs(blob)
    _, want_proofs = cref.compute_cells_and_kzg_proofs(swapped)

    commitment = neg_ctx.blob_to_kzg_commitment(blob)
    assert commitment == want_commitment
    assert commitment != das_ctx.blob_to_kzg_commitment(blob)
    cells, proofs = neg_ctx.compute_cells_and_kzg_proofs(blob)
    assert cells == want_cells
    assert proofs == want_proofs
    # the batch entry point takes the FK20 route (one blob alone takes the direct one): both must agree under the custom setup
    bc, bp, _ = neg_ctx.compute_cells_and_kzg_proofs_batch(blob * 3, 3)
    for b in range(3):
        assert bytes(bc[b * 128 * 2048:(b + 1) * 128 * 2048]) == b"".join(want_cells)
        assert bytes(bp[b * 128 * 48:(b + 1) * 128 * 48]) == b"".join(want_proofs)

    # cell proofs: [tau^64]_2 (unchanged by the negation) + the custom G1 side
    idx = list(range(128))
    assert neg_ctx.verify_cell_kzg_proof_batch([commitment] * 128, idx, cells, proofs) is True
    assert das_ctx.verify_cell_kzg_proof_batch([commitment] * 128, idx, cells, proofs) is False
    bad = list(cells)
    bad[5] = bad[5][:-1] + bytes([bad[5][-1] ^ 1])
    assert neg_ctx.verify_cell_kzg_proof_batch([commitment] * 128, idx, bad, proofs) is False

    # recovery from half of the cells: the proofs come from the custom tables again
    keep = list(range(1, 128, 2))
    rc, rp = neg_ctx.recover_cells_and_kzg_proofs(keep, [cells[i] for i in keep])
    assert rc == cells and rp == proofs

    # single-point proofs: [-tau]_2 is the one verification key the negation changes
    z = rng.randrange(R).to_bytes(32, "big")
    proof, y = neg_ctx.compute_kzg_proof(blob, z)
    mz = ((R - int.from_bytes(z, "big")) % R).to_bytes(32, "big")
    # q(-X) = (p(-X) - y) / (-X - z) = -(quotient of p(-X) at -z): the mainnet proof of the swapped blob at -z, negated
    want_proof, want_y = cref.compute_kzg_proof(swapped, mz)
    want_proof = want_proof if want_proof[0] & 0x40 else bytes([want_proof[0] ^ 0x20]) + want_proof[1:]
    assert y == want_y and proof == want_proof
    assert neg_ctx.verify_kzg_proof(commitment, z, y, proof) is True
    assert das_ctx.verify_kzg_proof(commitment, z, y, proof) is False
    bproof = neg_ctx.compute_blob_kzg_proof(blob, commitment)
    assert neg_ctx.verify_blob_kzg_proof(blob, commitment, bproof) is True
    assert das_ctx.verify_blob_kzg_proof(blob, commitment, bproof) is False
    assert neg_ctx.verify_blob_kzg_proof_batch([blob, blob], [commitment] * 2, [bproof] * 2) is True


def test_rotated_setup_against_mainnet_oracle(pkg, das_ctx):
    """a setup whose points are all NEW ones (not sign flips): secret s * tau with s = w_4096^64, a 64th root of unity.  p(s X) is the blob
    with its domain rotated by 64 places; s^64 = 1 keeps every cell's coset and X^64 - c_k in place, so commitment and cell proofs under
    s * tau are the mainnet ones of the rotated blob, and a point proof at z is s^-1 times the mainnet proof of the rotated blob at z / s.
    G1 points scaled by the oracle's curve arithmetic, G2 points by the affine arithmetic of tests/setup_util.py."""
    g1m, _, g2m = su.mainnet_points()
    w4096 = pow(7, (R - 1) // 4096, R)
    s_ = pow(w4096, 64, R)
    assert pow(s_, 64, R) == 1 and pow(s_, 32, R) != 1
    g1, g2 = su.rotated_setup(g1m, g2m, s_, cref.g1_mul)
    assert g2[64] == g2m[64] and g2[1] != g2m[1] and g1[1] != g1m[1]
    ctx = pkg.DASContext.from_json(su.setup_json(g1, g2), use_precomp=False)
    try:
        rng = random.Random(79)
        blob = b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(4096))
        rot = su.rotate_blob(blob, 64)
        commitment = ctx.blob_to_kzg_commitment(blob)
        assert commitment == cref.blob_to_kzg_commitment(rot)
        cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
        want_cells, _ = cref.compute_cells_and_kzg_proofs(blob)
        _, want_proofs = cref.compute_cells_and_kzg_proofs(rot)
        assert cells == want_cells and proofs == want_proofs
        bc, bp, _ = ctx.compute_cells_and_kzg_proofs_batch(blob * 2, 2)        # FK20 route
        assert bytes(bp[:128 * 48]) == b"".join(want_proofs) and bytes(bp[128 * 48:]) == b"".join(want_proofs)
        idx = list(range(128))
        assert ctx.verify_cell_kzg_proof_batch([commitment] * 128, idx, cells, proofs) is True
        assert das_ctx.verify_cell_kzg_proof_batch([commitment] * 128, idx, cells, proofs) is False
        z = rng.randrange(R)
        proof, y = ctx.compute_kzg_proof(blob, z.to_bytes(32, "big"))
        s_inv = pow(s_, R - 2, R)
        mp, my = cref.compute_kzg_proof(rot, (z * s_inv % R).to_bytes(32, "big"))
        assert y == my and proof == cref.g1_mul(mp, s_inv)
        assert ctx.verify_kzg_proof(commitment, z.to_bytes(32, "big"), y, proof) is True       # e(.., [s tau]_2): a G2 point the loader decompressed
        assert das_ctx.verify_kzg_proof(commitment, z.to_bytes(32, "big"), y, proof) is False
        bproof = ctx.compute_blob_kzg_proof(blob, commitment)
        assert ctx.verify_blob_kzg_proof(blob, commitment, bproof) is True
    finally:
        ctx.close()


def test_custom_setup_on_wide_windows(pkg, das_ctx, neg_ctx, monkeypatch):
    """use_precomp = true with a caller's setup: merged top window + the wider SRS tables built from the custom points"""
    monkeypatch.setenv("EKZG_FK20_WINDOW", "10")
    monkeypatch.setenv("EKZG_SRS_WINDOW", "9")
    g1m, _, g2m = su.mainnet_points()
    ctx = pkg.DASContext.from_json(su.setup_json(su.negate_odd(g1m), su.negate_odd(g2m)), use_precomp=True, subgroup_check=False)
    try:
        assert ctx.window == 10 and ctx.srs_window == 9
        rng = random.Random(78)
        blobs = [b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(4096)) for _ in range(2)]
        bc, bp, _ = ctx.compute_cells_and_kzg_proofs_batch(b"".join(blobs) + blobs[0], 3)
        for b, blob in enumerate(blobs):
            cells, proofs = neg_ctx.compute_cells_and_kzg_proofs(blob)
            assert bytes(bc[b * 128 * 2048:(b + 1) * 128 * 2048]) == b"".join(cells)
            assert bytes(bp[b * 128 * 48:(b + 1) * 128 * 48]) == b"".join(proofs)
            assert ctx.blob_to_kzg_commitment(blob) == neg_ctx.blob_to_kzg_commitment(blob)
    finally:
        ctx.close()


def _g1_on_curve_off_subgroup(rng):
    while True:
        x = rng.randrange(su.P)
        y = su.fp_sqrt((x * x * x + 4) % su.P)
        if y is not None:
            enc = bytearray(x.to_bytes(48, "big"))
            enc[0] |= 0x80 | (0x20 if y > (su.P - 1) // 2 else 0)
            return bytes(enc)      # a random point of E(Fp): in the r-torsion with probability ~2^-126


def test_invalid_setups_are_refused(pkg, das_ctx):
    g1m, _, g2m = su.mainnet_points()
    rng = random.Random(9)

    def attempt(g1, g2, **kw):
        ctx = pkg.DASContext.from_json(su.setup_json(g1, g2), use_precomp=False, **kw)
        ctx.close()

    rogue = list(g1m)
    rogue[1234] = _g1_on_curve_off_subgroup(rng)
    with pytest.raises(pkg.KzgError, match="G1 point 1234"):
        attempt(rogue, g2m)
    attempt(rogue, g2m, subgroup_check=False)            # from_json_unchecked accepts a point that only satisfies the curve equation
    off = list(g1m)
    x = 5
    while su.fp_sqrt((x ** 3 + 4) % su.P) is not None:
        x += 1
    off[7] = bytes([0x80]) + x.to_bytes(47, "big")
    for check in (True, False):
        with pytest.raises(pkg.KzgError, match="G1 point 7"):
            attempt(off, g2m, subgroup_check=check)
    ident = list(g1m)
    ident[100] = bytes([0xc0]) + bytes(47)
    with pytest.raises(pkg.KzgError, match=r"g1_monomial\[100\] is the point at infinity"):
        attempt(ident, g2m)
    with pytest.raises(pkg.KzgError, match="4096 points"):
        attempt(g1m[:4095], g2m)
    with pytest.raises(pkg.KzgError, match="65 points"):
        attempt(g1m, g2m[:2])
    badg2 = list(g2m)
    badg2[1] = bytes([badg2[1][0] & 0x7f]) + badg2[1][1:]
    with pytest.raises(pkg.KzgError, match=r"g2_monomial\[1\]"):
        attempt(g1m, badg2)
    with pytest.raises(pkg.KzgError, match="JSON"):
        pkg.DASContext.from_json("{", use_precomp=False)
    # the session context is unaffected by the failed attempts
    name, inp, expected = [c for c in vectors.load("blob_to_kzg_commitment") if c[2] is not None][0]
    assert das_ctx.blob_to_kzg_commitment(inp["blob"]) == expected
