"""GPU parity tests of the hot path (compute_cells_and_kzg_proofs) through the C ABI, against
(1) the reference's consensus vectors (tests/golden, from test_vectors/compute_cells_and_kzg_proofs),
(2) the CPU oracle on seeded synthetic + edge blobs, stage by stage and end to end,
(3) size-independent properties at the full batch size of BASELINE.json config #3.
Bit-exact everywhere: this is integer arithmetic."""
import hashlib

import pytest

from tests import vectors

pytestmark = pytest.mark.gpu

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _synth(pkg):
    import importlib
    return importlib.import_module("eth_kzg_b200.synthetic")


def test_fk20_stages_match_oracle(das_ctx, pkg):
    """scalars (Toeplitz NTT), MSM outputs and h commitments of one blob vs the oracle's intermediates.
    The device folds the 1/128 of the inverse G1 NTT into the scalars, so scalars and MSM outputs are
    compared after scaling the oracle's by 128^-1."""
    from oracle import cref
    blob = _synth(pkg).blob(7)
    sc, msm, h = das_ctx.debug_fk20_stages(blob)
    osc, omsm, oh = cref.fk20_stages(blob)
    inv128 = pow(128, -1, R)
    bad = [(j, k) for j in range(128) for k in range(64) if sc[j][k] != osc[j][k] * inv128 % R]
    assert not bad, "scalar mismatches (first 5): %r" % bad[:5]
    badm = [j for j in range(128) if msm[j] != cref.g1_mul(omsm[j], inv128)]
    assert not badm, "MSM mismatches at j = %r" % badm[:10]
    badh = [i for i in range(64) if h[i] != oh[i]]
    assert not badh, "h mismatches at i = %r" % badh[:10]


@pytest.mark.parametrize("route", ["direct", "fk20"])
@pytest.mark.parametrize("name,inp,expected", [pytest.param(n, i, o, id=n) for n, i, o in vectors.load("compute_cells_and_kzg_proofs")])
def test_consensus_vectors(vec_ctx, pkg, name, inp, expected, route, monkeypatch):
    """crates/eip7594/tests/compute_cells_and_kzg_proofs.rs: byte-exact cells+proofs, `output: null` <=> Err -- on both routes a
    single-blob call can take: the direct one (128 SRS MSMs, the default for one blob) and FK20 (EKZG_DIRECT_MAX=0)"""
    if route == "fk20":
        monkeypatch.setenv("EKZG_DIRECT_MAX", "0")
    try:
        cells, proofs = vec_ctx.compute_cells_and_kzg_proofs(inp["blob"])
        got = [cells, proofs]
    except pkg.KzgError:
        got = None
    if expected is not None:
        expected = [list(expected[0]), list(expected[1])]
    assert got == expected


@pytest.mark.parametrize("name,inp,expected", [pytest.param(n, i, o, id=n) for n, i, o in vectors.load("compute_cells_and_kzg_proofs")])
def test_consensus_vectors_compute_cells(vec_ctx, pkg, name, inp, expected):
    try:
        got = vec_ctx.compute_cells(inp["blob"])
    except pkg.KzgError:
        got = None
    assert got == (list(expected[0]) if expected is not None else None)


def test_batch_matches_oracle(das_ctx, pkg):
    """ragged batch (not a multiple of the warp or chunk size) of synthetic + edge blobs vs the oracle"""
    from oracle import cref
    syn = _synth(pkg)
    blobs = syn.edge_blobs() + [syn.blob(i) for i in range(33)]
    n = len(blobs)
    cells, proofs, status = das_ctx.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n)
    assert status == [0] * n
    ocells, oproofs = cref.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n)
    for i in range(n):
        assert cells[i * 262144:(i + 1) * 262144] == ocells[i * 262144:(i + 1) * 262144], "cells of blob %d" % i
        assert proofs[i * 6144:(i + 1) * 6144] == oproofs[i * 6144:(i + 1) * 6144], "proofs of blob %d" % i
    # the constant blob: every proof is the identity
    assert proofs[2 * 6144:3 * 6144] == (b"\xc0" + bytes(47)) * 128


def test_batch_invalid_blob_is_isolated(das_ctx, pkg):
    """one non-canonical element poisons only its own blob; the call reports Err like the reference"""
    syn = _synth(pkg)
    good = syn.blob(1)
    bad = bytearray(syn.blob(2))
    bad[32 * 100:32 * 101] = R.to_bytes(32, "big")
    cells, proofs, status = das_ctx.compute_cells_and_kzg_proofs_batch(good + bytes(bad) + good, 3)
    assert status == [0, 1, 0]
    assert cells[:262144] == cells[2 * 262144:] and proofs[:6144] == proofs[2 * 6144:]
    c1, p1 = das_ctx.compute_cells_and_kzg_proofs(good)
    assert b"".join(c1) == cells[:262144] and b"".join(p1) == proofs[:6144]
    with pytest.raises(pkg.KzgError):
        das_ctx.compute_cells_and_kzg_proofs(bytes(bad))


def test_full_size_batch_properties(das_ctx, pkg):
    """BASELINE config #3 size (1024 blobs): properties that need no oracle run --
    cells 0..63 reproduce the blob; the batch is a permutation-equivariant map (blob order does not
    matter); results equal the single-blob ABI on sampled blobs; checksum is stable across chunkings."""
    syn = _synth(pkg)
    n = 1024
    base = syn.edge_blobs() + [syn.blob(i) for i in range(60)]   # zero, all r-1, constant (identity proofs), dummy_blob first
    blobs = [base[(i * 37) % 64] for i in range(n)]
    cells, proofs, status = das_ctx.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n)
    assert status == [0] * n
    first = {}
    for i in range(n):
        c = cells[i * 262144:(i + 1) * 262144]
        p = proofs[i * 6144:(i + 1) * 6144]
        assert c[:131072] == blobs[i]
        key = (i * 37) % 64
        if key in first:
            assert first[key] == (hashlib.sha256(c).digest(), p), "blob %d differs from its duplicate" % i
        else:
            first[key] = (hashlib.sha256(c).digest(), p)
    assert first[2][1] == (b"\xc0" + bytes(47)) * 128
    for key in (0, 1, 2, 3, 17, 63):
        c1, p1 = das_ctx.compute_cells_and_kzg_proofs(base[key])
        assert first[key] == (hashlib.sha256(b"".join(c1)).digest(), b"".join(p1))


@pytest.mark.parametrize("fk20_w,srs_w", [("14", "13"), ("12", "12"), ("10", "9")])
def test_precomputed_window_variants(das_ctx, pkg, fk20_w, srs_w, monkeypatch, precomp_holder):
    """use_precomp contexts: per-window tables at several widths incl. the production one (w = 14, four top digits per lookup;
    SRS w = 13) and the pair-merged one (w = 12).  The session context (use_precomp=False: w = 8, no merged top window) is
    itself pinned by the consensus vectors and the oracle above, so it serves as the reference here -- plus the vectors
    again, directly."""
    precomp_holder.release()   # the shared production context holds ~144 GiB: this test needs the memory for its own layout
    monkeypatch.setenv("EKZG_FK20_WINDOW", fk20_w)
    monkeypatch.setenv("EKZG_SRS_WINDOW", srs_w)
    ctx = pkg.DASContext(use_precomp=True)
    try:
        assert ctx.window == int(fk20_w)
        syn = _synth(pkg)
        # edge scalars exercise the top window: all r-1, zero, constant, and the reference's dummy blob
        # the production widths get a large batch: ~10^9 field multiplications per context, computed through two different
        # table layouts -- a rare arithmetic slip (a lost carry is a 2^-32 event) would show as a mismatch
        n = 512 if fk20_w == "14" else 40
        blobs = [syn.blob(300 + i) for i in range(n - 4)] + list(syn.edge_blobs())[:4]
        flat = b"".join(blobs[:n])
        want = das_ctx.compute_cells_and_kzg_proofs_batch(flat, len(blobs[:n]))
        got = ctx.compute_cells_and_kzg_proofs_batch(flat, len(blobs[:n]))
        assert got[0] == want[0], "cells differ"
        assert got[1] == want[1], "proofs differ from the w = 8 context"
        for name, inp, expected in [c for c in vectors.load("compute_cells_and_kzg_proofs") if c[2] is not None][:3]:
            cells, proofs = ctx.compute_cells_and_kzg_proofs(inp["blob"])
            assert [cells, proofs] == [list(expected[0]), list(expected[1])], name
        for b in blobs[:6] + blobs[-4:]:
            c = ctx.blob_to_kzg_commitment(b)
            assert c == das_ctx.blob_to_kzg_commitment(b)
            assert ctx.compute_blob_kzg_proof(b, c) == das_ctx.compute_blob_kzg_proof(b, c)
        for name, inp, expected in [c for c in vectors.load("blob_to_kzg_commitment") if c[2] is not None][:3]:
            assert ctx.blob_to_kzg_commitment(inp["blob"]) == expected, name
    finally:
        ctx.close()


@pytest.mark.parametrize("coop", ["0", "256"], ids=["lane_per_blob", "four_lanes_per_element"])
@pytest.mark.parametrize("n", [1, 2, 7, 9, 31, 33, 128, 160])
def test_k5_radix4_equals_radix2(vec_ctx, pkg, n, coop, monkeypatch):
    """the latency-mode G1-NTT kernel (radix-4 super-phases, small batches) against the radix-2 kernel the vectors pin, on the
    same batch: ragged group sizes, identity points everywhere (zero blob), and the oracle on the first two blobs -- with its
    multiplication units one lane per blob and in the cooperative form (four lanes per field element, eight blobs per warp:
    csrc/g1_coop.cuh; the default up to 64 blobs)"""
    monkeypatch.setenv("EKZG_K5_COOP_MAX", coop)
    syn = _synth(pkg)
    blobs = [syn.blob(6100 + i) for i in range(n)]
    if n >= 7:
        blobs[3:7] = list(syn.edge_blobs())[:4]
    flat = b"".join(blobs)
    monkeypatch.setenv("EKZG_DIRECT_MAX", "0")      # (one or two blobs would otherwise not reach the G1 transforms at all)
    monkeypatch.setenv("EKZG_K5_R4_MAX", "0")
    want = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    monkeypatch.setenv("EKZG_K5_R4_MAX", "256")
    got = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    assert got[2] == want[2] == [0] * n
    assert got[0] == want[0]
    assert got[1] == want[1], "radix-4 proofs differ from the radix-2 kernel's"
    from oracle import cref
    for i in range(min(n, 2)):
        oc, op = cref.compute_cells_and_kzg_proofs(blobs[i])
        assert got[1][i * 6144:(i + 1) * 6144] == b"".join(op)


@pytest.mark.parametrize("n", [2, 5, 16])
def test_k4_one_point_per_thread_form(vec_ctx, pkg, n, monkeypatch):
    """small batches run the fixed-base MSM with ONE point per thread (64 slices); on tables with a merged top window the digits of a
    merge group are then gathered across neighbouring threads -- against the 16-slice form, the register form and the oracle"""
    syn = _synth(pkg)
    blobs = [syn.blob(8800 + i) for i in range(n)]
    if n >= 5:
        blobs[1:5] = list(syn.edge_blobs())[:4]
    flat = b"".join(blobs)
    monkeypatch.setenv("EKZG_DIRECT_MAX", "0")
    got = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    monkeypatch.setenv("EKZG_K4_NO_TINY", "1")
    want = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    monkeypatch.setenv("EKZG_K4", "r")
    want_r = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    assert got[2] == want[2] == [0] * n
    assert got[1] == want[1] == want_r[1]
    from oracle import cref
    oc, op = cref.compute_cells_and_kzg_proofs(blobs[0])
    assert got[1][:6144] == b"".join(op) and got[0][:262144] == b"".join(oc)


@pytest.mark.parametrize("form", ["r", "a"], ids=["register_xyzz", "batched_affine"])
def test_k4_alternative_forms_equal_default(vec_ctx, pkg, form, monkeypatch):
    """the A/B forms of the fixed-base MSM kernel that stay in the tree (EKZG_K4: register XYZZ of round 1; batched affine with a
    division-step inverter warp) against the default shared-memory-operand form, incl. the edge blobs
    (identity accumulators, equal-x cases of the affine additions)"""
    syn = _synth(pkg)
    n = 40
    blobs = [syn.blob(8200 + i) for i in range(n - 4)] + list(syn.edge_blobs())[:4]
    flat = b"".join(blobs)
    want = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    monkeypatch.setenv("EKZG_K4", form)
    monkeypatch.setenv("EKZG_K4A_MIN", "0")      # the affine kernel is normally reserved for wide launches
    got = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    assert got[2] == want[2] == [0] * n
    assert got[1] == want[1], "proofs of the '%s' form differ" % form
    c = vec_ctx.blob_to_kzg_commitment(blobs[0])
    monkeypatch.delenv("EKZG_K4")
    assert c == vec_ctx.blob_to_kzg_commitment(blobs[0])


def test_direct_proofs_equal_fk20(vec_ctx, pkg, monkeypatch):
    """the latency path for one or two blobs -- every proof as its own 4096-point MSM of f div (X^64 - c_k) over the SRS tables
    (EKZG_DIRECT_MAX, default 1) -- against the FK20 route on the same blobs: synthetic, all-zero, all r-1, constant (128 identity
    proofs) and the reference bench's blob; through the batch entry point, the single-blob symbol, recovery, and the oracle"""
    from oracle import cref
    syn = _synth(pkg)
    cases = [syn.blob(9100), syn.blob(9101)] + list(syn.edge_blobs())[:4]
    for i in range(0, len(cases), 2):
        pair = cases[i:i + 2]
        flat = b"".join(pair)
        monkeypatch.setenv("EKZG_DIRECT_MAX", "0")
        want = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, 2)
        want1 = vec_ctx.compute_cells_and_kzg_proofs(pair[0])
        monkeypatch.setenv("EKZG_DIRECT_MAX", "2")
        got = vec_ctx.compute_cells_and_kzg_proofs_batch(flat, 2)
        got1 = vec_ctx.compute_cells_and_kzg_proofs(pair[0])
        assert got == want, "direct proofs of pair %d differ from the FK20 route" % i
        assert got1 == want1
        keep = list(range(1, 128, 2))
        rc, rp = vec_ctx.recover_cells_and_kzg_proofs(keep, [got1[0][j] for j in keep])
        assert (rc, rp) == (got1[0], got1[1])
    oc, op = cref.compute_cells_and_kzg_proofs(cases[0])
    monkeypatch.setenv("EKZG_DIRECT_MAX", "2")
    c, p = vec_ctx.compute_cells_and_kzg_proofs(cases[0])
    assert (c, p) == (list(oc), list(op))
