"""CPU test of the library's HOST pairing check (csrc/host_pairing.cpp) -- needs no GPU.
Checks bilinearity/non-degeneracy and the SRS relations e([tau^k]_1, [1]_2) == e([1]_1, [tau^k]_2) for k = 1, 64 on
the mainnet trusted setup, with G1 points produced by the oracle's curve arithmetic."""
import ctypes
import random

import pytest

from oracle import pyref


@pytest.fixture(scope="module")
def lib(pkg):
    import os
    if not os.path.exists(pkg.library_path()):
        pkg.build_library()
    return ctypes.CDLL(pkg.library_path())


def _xy(pt):
    if pt is None:
        return bytes(96)
    return pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little")


def _check(lib, pairs):
    buf = b"".join(_xy(p) for p, _ in pairs)
    sel = (ctypes.c_int * len(pairs))(*[s for _, s in pairs])
    return lib.eth_kzg_b200_debug_pairing_check(len(pairs), buf, sel)


def _srs_point(i):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ts = open(os.path.join(root, "rust-eth-kzg_b200", "data", "trusted_setup_4096.bin"), "rb").read()
    return pyref.g1_decompress(ts[16 + 48 * i:16 + 48 * i + 48], check_subgroup=False)


def test_pairing_bilinearity(lib):
    rng = random.Random(3)
    G = pyref.G1_GEN
    a, b = rng.randrange(1, pyref.R), rng.randrange(1, pyref.R)
    aG, bG, abG = pyref.g1_mul(G, a), pyref.g1_mul(G, b), pyref.g1_mul(G, a * b % pyref.R)
    GEN, TAU, TAU64, NEG = 0, 1, 2, 3
    assert _check(lib, [(aG, GEN), (pyref.g1_neg(aG), GEN)]) == 1
    assert _check(lib, [(aG, GEN), (aG, GEN + NEG)]) == 1
    assert _check(lib, [(aG, GEN), (pyref.g1_neg(bG), GEN)]) == 0
    assert _check(lib, [(aG, TAU), (bG, TAU64), (pyref.g1_neg(aG), TAU), (bG, TAU64 + NEG)]) == 1
    assert _check(lib, [(aG, TAU)]) == 0                       # non-degenerate
    assert _check(lib, [(None, TAU), (None, GEN)]) == 1        # identity pairs are skipped
    assert _check(lib, [(abG, GEN), (None, TAU), (pyref.g1_neg(abG), GEN)]) == 1


def test_pairing_srs_relations(lib):
    GEN, TAU, TAU64, NEG = 0, 1, 2, 3
    s0, s1, s64 = _srs_point(0), _srs_point(1), _srs_point(64)
    assert s0 == pyref.G1_GEN
    assert _check(lib, [(s1, GEN + NEG), (s0, TAU)]) == 1       # e([tau]_1, -[1]_2) e([1]_1, [tau]_2) = 1
    assert _check(lib, [(s64, GEN + NEG), (s0, TAU64)]) == 1
    assert _check(lib, [(s64, GEN + NEG), (s0, TAU)]) == 0
    assert _check(lib, [(s1, TAU64), (s64, TAU + NEG), (None, GEN)]) == 1   # tau * tau^64 on both sides


def test_pairing_shortcuts_agree_with_general_routines(lib):
    """the sparse line product (36 instead of 54 Fp multiplications) and the Granger-Scott cyclotomic squaring (18 instead of 36)
    against the general Fp12 product / squaring, on values from a real Miller loop"""
    assert lib.eth_kzg_b200_debug_pairing_selftest() == 1
