"""Helpers for the trusted-setup loader tests: the mainnet ceremony points from the packed file the library embeds, the
consensus-specs JSON layout built from them, a big-integer restatement of the ZCash G2 encoding, and the derived setup with
secret -tau (negate every odd-index point: [(-tau)^i] = (-1)^i [tau^i]) whose outputs are predictable from mainnet ones."""
import json
import os

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_cache = None


def mainnet_points():
    """(g1_monomial[4096], g1_lagrange[4096], g2_monomial[65]) as compressed byte strings (tools/convert_trusted_setup.py layout)"""
    global _cache
    if _cache is None:
        ts = open(os.path.join(ROOT, "rust-eth-kzg_b200", "data", "trusted_setup_4096.bin"), "rb").read()
        assert ts[:8] == b"EKZGTS01"
        n1, n2 = int.from_bytes(ts[8:12], "little"), int.from_bytes(ts[12:16], "little")
        o = 16
        g1m = [ts[o + 48 * i:o + 48 * i + 48] for i in range(n1)]
        o += 48 * n1
        g1l = [ts[o + 48 * i:o + 48 * i + 48] for i in range(n1)]
        o += 48 * n1
        g2m = [ts[o + 96 * i:o + 96 * i + 96] for i in range(n2)]
        _cache = (g1m, g1l, g2m)
    return _cache


def setup_json(g1_monomial, g2_monomial, g1_lagrange=None):
    doc = {"g1_monomial": ["0x" + x.hex() for x in g1_monomial]}
    if g1_lagrange is not None:
        doc["g1_lagrange"] = ["0x" + x.hex() for x in g1_lagrange]
    doc["g2_monomial"] = ["0x" + x.hex() for x in g2_monomial]
    return json.dumps(doc)


def negate_odd(points):
    """compressed points with the sign flag (0x20) of every odd-index one flipped: the setup of secret -tau"""
    return [bytes([p[0] ^ 0x20]) + p[1:] if i & 1 else p for i, p in enumerate(points)]


# ---- Fp2 = Fp[u]/(u^2 + 1), elements as (c0, c1) -------------------------------------------------
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp_sqrt(a):
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


def f2_sqrt(a):
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt(-a0 % P)
        return None if s is None else (0, s)
    s = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if s is None:
        return None
    inv2 = pow(2, P - 2, P)
    for d in ((a0 + s) * inv2 % P, (a0 - s) * inv2 % P):
        x0 = fp_sqrt(d)
        if x0 is not None and x0 != 0:
            x = (x0, a1 * pow(2 * x0, P - 2, P) % P)
            if f2_mul(x, x) == (a0, a1):
                return x
    return None


def g2_rhs(x):
    return f2_add(f2_mul(f2_mul(x, x), x), (4, 4))


def _larger(y):
    """y is lexicographically larger than -y (c1 is the more significant coordinate)"""
    return y[1] > (P - 1) // 2 if y[1] else y[0] > (P - 1) // 2


def g2_decompress(b):
    assert b[0] & 0x80 and not b[0] & 0x40
    x1 = int.from_bytes(bytes([b[0] & 0x1f]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    y = f2_sqrt(g2_rhs((x0, x1)))
    assert y is not None
    if _larger(y) != bool(b[0] & 0x20):
        y = (-y[0] % P, -y[1] % P)
    return (x0, x1), y


def g2_compress(pt):
    (x0, x1), y = pt
    out = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    out[0] |= 0x80 | (0x20 if _larger(y) else 0)
    return bytes(out)


# ---- G2 affine arithmetic over Fp2 (test-side derivation of a rotated setup; slow and simple) ----------------------------------
def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, -a[1] * n % P)


def g2_add(p1, p2):
    """affine addition on y^2 = x^3 + 4(1 + u); None is the identity"""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)


def g2_mul(pt, k):
    acc = None
    for bit in bin(k)[2:]:
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def brp12(k):
    return int(format(k, "012b")[::-1], 2)


def rotated_setup(g1_monomial, g2_monomial, s, g1_mul):
    """the setup of secret s * tau: point i scaled by s^i (g1_mul: compressed G1 point x integer -> compressed point, the oracle's)"""
    g1 = [g1_mul(pt, pow(s, i, R_ORDER)) if i else pt for i, pt in enumerate(g1_monomial)]
    g2 = [g2_compress(g2_mul(g2_decompress(pt), pow(s, i, R_ORDER))) if pow(s, i, R_ORDER) != 1 else pt for i, pt in enumerate(g2_monomial)]
    return g1, g2


def rotate_blob(blob, shift):
    """evaluation form of p(w^shift X) for the blob of p: the domain in natural order moves by `shift` places"""
    return b"".join(blob[32 * brp12((brp12(k) + shift) % 4096):32 * brp12((brp12(k) + shift) % 4096) + 32] for k in range(4096))
