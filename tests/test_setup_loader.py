"""CPU tests of the caller-supplied-setup loader's host side (needs no GPU): the JSON parser (csrc/trusted_setup_json.cpp) and the
G2 decompression / curve / subgroup checks (csrc/host_pairing.cpp) behind eth_kzg_b200_das_context_new_from_json.
Reference behaviour: TrustedSetup::from_json[_unchecked] (crates/trusted_setup/src/lib.rs:40-127) -- it panics where this returns Err.
G2 decompression is checked against an independent big-integer restatement of the ZCash encoding written here."""
import ctypes
import json
import random

import pytest

from tests import setup_util as su


@pytest.fixture(scope="module")
def lib(pkg):
    import os
    if not os.path.exists(pkg.library_path()):
        pkg.build_library()
    return pkg.load_library()


def _parse(lib, pkg, text):
    data = text.encode() if isinstance(text, str) else text
    n1, n2 = ctypes.c_uint64(), ctypes.c_uint64()
    g1 = ctypes.create_string_buffer(48)
    g2 = ctypes.create_string_buffer(96)
    res = lib.eth_kzg_b200_debug_parse_trusted_setup_json(data, ctypes.c_uint64(len(data)), ctypes.byref(n1), ctypes.byref(n2), g1, g2)
    if res.status != 0:
        msg = ctypes.cast(res.error_msg, ctypes.c_char_p).value.decode()
        lib.eth_kzg_free_error_message(res.error_msg)
        raise pkg.KzgError(msg)
    return n1.value, n2.value, g1.raw, g2.raw


def test_parse_mainnet_json(lib, pkg):
    g1m, g1l, g2m = su.mainnet_points()
    text = su.setup_json(g1m, g2m, g1_lagrange=g1l)
    assert _parse(lib, pkg, text) == (4096, 65, g1m[0], g2m[64])
    # key order, whitespace and unknown keys of any shape do not matter (serde ignores unknown fields)
    doc = {"comment": {"nested": [1, 2.5e3, None, True, {"a": "b\\\"c"}]}, "g2_monomial": ["0x" + x.hex() for x in g2m],
           "g1_lagrange": [], "g1_monomial": ["0x" + x.hex().upper() for x in g1m[:7]]}
    assert _parse(lib, pkg, json.dumps(doc, indent=3)) == (7, 65, g1m[0], g2m[64])


@pytest.mark.parametrize("mutate,needle", [
    (lambda d: d.pop("g1_monomial"), "missing field g1_monomial"),
    (lambda d: d.pop("g2_monomial"), "missing field g2_monomial"),
    (lambda d: d["g1_monomial"].__setitem__(3, d["g1_monomial"][3][2:]), "does not start with 0x"),
    (lambda d: d["g1_monomial"].__setitem__(3, d["g1_monomial"][3] + "00"), "expected 48"),
    (lambda d: d["g2_monomial"].__setitem__(1, d["g2_monomial"][1][:-2]), "expected 96"),
    (lambda d: d["g1_monomial"].__setitem__(0, "0x" + "zz" * 48), "not hexadecimal"),
    (lambda d: d["g1_monomial"].__setitem__(0, 17), "not a string"),
    (lambda d: d.__setitem__("g2_monomial", "0x00"), "not an array"),
])
def test_parse_rejects(lib, pkg, mutate, needle):
    g1m, _, g2m = su.mainnet_points()
    doc = {"g1_monomial": ["0x" + x.hex() for x in g1m[:16]], "g2_monomial": ["0x" + x.hex() for x in g2m]}
    mutate(doc)
    with pytest.raises(pkg.KzgError, match=needle):
        _parse(lib, pkg, json.dumps(doc))


@pytest.mark.parametrize("text", ["", "[]", "{", '{"g1_monomial": ["0x00"', '{"g1_monomial": [], "g2_monomial": []} x', '{"a" 1}'])
def test_parse_malformed_json(lib, pkg, text):
    with pytest.raises(pkg.KzgError):
        _parse(lib, pkg, text)


def _decompress(lib, b):
    out = (ctypes.c_uint64 * 24)()
    rc = lib.eth_kzg_b200_debug_g2_decompress(bytes(b), out)
    vals = [sum(out[6 * k + i] << (64 * i) for i in range(6)) for k in range(4)]
    return rc, vals


def test_g2_decompress_matches_bigint_restatement(lib):
    _, _, g2m = su.mainnet_points()
    for i in (0, 1, 2, 33, 64):
        rc, vals = _decompress(lib, g2m[i])
        (x0, x1), (y0, y1) = su.g2_decompress(g2m[i])
        assert rc == 0 and vals == [x0, x1, y0, y1], i
        # both signs: flipping the sign flag negates y and stays in the subgroup
        flipped = bytes([g2m[i][0] ^ 0x20]) + g2m[i][1:]
        rc, vals = _decompress(lib, flipped)
        assert rc == 0 and vals == [x0, x1, (-y0) % su.P, (-y1) % su.P]
        assert su.g2_compress(((x0, x1), (y0, y1))) == g2m[i]


def test_g2_decompress_rejects(lib):
    _, _, g2m = su.mainnet_points()
    good = g2m[1]
    assert _decompress(lib, bytes([good[0] & 0x7f]) + good[1:])[0] == 1            # uncompressed flag
    assert _decompress(lib, bytes([0xc0]) + bytes(95))[0] == 4                      # infinity
    assert _decompress(lib, bytes([0xc0]) + bytes(94) + b"\x01")[0] == 1            # infinity with payload
    assert _decompress(lib, bytes([0x9f]) + b"\xff" * 95)[0] == 1                   # x.c1 >= p
    rng = random.Random(11)
    off_curve = on_curve_off_subgroup = None
    while off_curve is None or on_curve_off_subgroup is None:
        x = (rng.randrange(su.P), rng.randrange(su.P))
        enc = bytes([0x80 | (x[1] >> 376)]) + (x[1] & ((1 << 376) - 1)).to_bytes(47, "big") + x[0].to_bytes(48, "big")
        if su.f2_sqrt(su.g2_rhs(x)) is None:
            off_curve = enc
        else:
            on_curve_off_subgroup = enc      # a random point of E'(Fp2) lies in the r-torsion with probability ~2^-256
    assert _decompress(lib, off_curve)[0] == 2
    assert _decompress(lib, on_curve_off_subgroup)[0] == 3


def test_g2_keys_of_a_setup(lib, pkg):
    _, _, g2m = su.mainnet_points()

    def keys(points, check=True):
        res = lib.eth_kzg_b200_debug_g2_keys(b"".join(points), len(points), ctypes.c_bool(check))
        if res.status != 0:
            msg = ctypes.cast(res.error_msg, ctypes.c_char_p).value.decode()
            lib.eth_kzg_free_error_message(res.error_msg)
            raise pkg.KzgError(msg)

    keys(g2m)
    keys(su.negate_odd(g2m))
    with pytest.raises(pkg.KzgError, match="65 points"):
        keys(g2m[:64])
    rng = random.Random(5)
    while True:
        x = (rng.randrange(su.P), rng.randrange(su.P))
        if su.f2_sqrt(su.g2_rhs(x)) is not None:
            break
    rogue = bytes([0x80 | (x[1] >> 376)]) + (x[1] & ((1 << 376) - 1)).to_bytes(47, "big") + x[0].to_bytes(48, "big")
    bad = list(g2m)
    bad[40] = rogue
    with pytest.raises(pkg.KzgError, match=r"g2_monomial\[40\] is outside the prime-order subgroup"):
        keys(bad)
    keys(bad, check=False)                   # from_json_unchecked: curve equation only
    bad[40] = bytes([0xc0]) + bytes(95)
    keys(bad)                                # the identity is in the subgroup; only the three points in use must not be it
    bad[64] = bytes([0xc0]) + bytes(95)
    with pytest.raises(pkg.KzgError, match=r"g2_monomial\[64\] is the point at infinity"):
        keys(bad)
