"""Pure-Python bigint restatement of the BLS12-381 / KZG primitives (TEST INFRASTRUCTURE ONLY).

This module is part of the oracle: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it.  It is the slow, obviously-correct layer used to
 (a) derive every constant the C oracle and the CUDA code embed (tools/gen_constants.py), and
 (b) cross-check the C oracle's primitives on small cases.

The arithmetic itself lives in third-party blst (>=0.3.16) / blstrs 0.7.1 in the reference
(crates/cryptography/bls12_381/Cargo.toml:17-23); it is restated here from the BLS12-381
standard.  Algorithm structure follows:
  - NTT:            crates/cryptography/polynomial/src/fft.rs:46-177, domain.rs:84-125
  - cells:          crates/cryptography/kzg_multi_open/src/fk20/prover.rs:158-181
  - naive proofs:   crates/cryptography/kzg_multi_open/src/fk20/naive.rs:31-93
  - serialization:  crates/serialization/src/lib.rs:36-156
Parity status: pinned (tests/test_oracle_vectors.py checks against the consensus vectors).
"""

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
B_G1 = 4
# generator of G1
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
FR_GENERATOR = 7
BLS_X = 0xD201000000010000  # |x|, x is negative

FIELD_ELEMENTS_PER_BLOB = 4096
FIELD_ELEMENTS_PER_CELL = 64
CELLS_PER_EXT_BLOB = 128
FIELD_ELEMENTS_PER_EXT_BLOB = 8192


def root_of_unity(n):
    """omega_n = 7^((r-1)/n)  (domain.rs:84-102)"""
    assert (R - 1) % n == 0
    return pow(FR_GENERATOR, (R - 1) // n, R)


def bit_reverse(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def brp(v):
    n = len(v)
    bits = n.bit_length() - 1
    return [v[bit_reverse(i, bits)] for i in range(n)]


def ntt(v, w=None, mod=R):
    """natural-order DFT: out[j] = sum_i v[i] w^(ij)"""
    n = len(v)
    if w is None:
        w = root_of_unity(n)
    if n == 1:
        return list(v)
    e = ntt(v[0::2], w * w % mod, mod)
    o = ntt(v[1::2], w * w % mod, mod)
    out = [0] * n
    t = 1
    for i in range(n // 2):
        x = t * o[i] % mod
        out[i] = (e[i] + x) % mod
        out[i + n // 2] = (e[i] - x) % mod
        t = t * w % mod
    return out


def intt(v):
    n = len(v)
    w = pow(root_of_unity(n), R - 2, R)
    ninv = pow(n, R - 2, R)
    return [x * ninv % R for x in ntt(v, w)]


# ---------------------------------------------------------------- G1 (Jacobian, ints)
INF = None


def g1_is_on_curve(pt):
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - B_G1) % P == 0


def g1_add(a, b):
    if a is INF:
        return b
    if b is INF:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_neg(a):
    if a is INF:
        return INF
    return (a[0], (-a[1]) % P)


def _jac_dbl(X, Y, Z):
    if Y == 0 or Z == 0:
        return (1, 1, 0)
    A = X * X % P
    Bq = Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) ** 2 - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jac_add_affine(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    rr = (S2 - Y1) % P
    if H == 0:
        if rr == 0:
            return _jac_dbl(X1, Y1, Z1)
        return (1, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (rr * rr - HHH - 2 * V) % P
    Y3 = (rr * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def g1_mul(pt, k):
    k %= R
    if pt is INF or k == 0:
        return INF
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(*acc)
        if bit == "1":
            acc = _jac_add_affine(*acc, pt[0], pt[1])
    X, Y, Z = acc
    if Z == 0:
        return INF
    zi = pow(Z, P - 2, P)
    return (X * zi * zi % P, Y * zi * zi * zi % P)


def g1_lincomb(points, scalars):
    acc = INF
    for pt, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


def g1_compress(pt):
    """ZCash format: bit7 compressed, bit6 infinity, bit5 y > (p-1)/2 (serialization/src/lib.rs:85)"""
    if pt is INF:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_decompress(b, check_subgroup=True):
    """returns point or raises ValueError (serialization/src/lib.rs:69-81 -> blst)"""
    if len(b) != 48:
        raise ValueError("len")
    c, inf, sign = b[0] >> 7, (b[0] >> 6) & 1, (b[0] >> 5) & 1
    if not c:
        raise ValueError("uncompressed flag")
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if inf:
        if x != 0 or sign:
            raise ValueError("bad infinity")
        return INF
    if x >= P:
        raise ValueError("x >= p")
    y2 = (x * x * x + B_G1) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise ValueError("not on curve")
    if (y > (P - 1) // 2) != bool(sign):
        y = P - y
    pt = (x, y)
    if check_subgroup and g1_mul_raw(pt, R) is not INF:
        raise ValueError("not in subgroup")
    return pt


def g1_mul_raw(pt, k):
    """scalar mul without reducing k mod r (for subgroup checks)"""
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(*acc)
        if bit == "1":
            acc = _jac_add_affine(*acc, pt[0], pt[1])
    X, Y, Z = acc
    if Z == 0:
        return INF
    zi = pow(Z, P - 2, P)
    return (X * zi * zi % P, Y * zi * zi * zi % P)


# ---------------------------------------------------------------- protocol pieces
def blob_to_scalars(blob):
    """serialization/src/lib.rs:36-63"""
    if len(blob) != 32 * FIELD_ELEMENTS_PER_BLOB:
        raise ValueError("blob length")
    out = []
    for i in range(FIELD_ELEMENTS_PER_BLOB):
        v = int.from_bytes(blob[32 * i : 32 * i + 32], "big")
        if v >= R:
            raise ValueError("non-canonical scalar")
        out.append(v)
    return out


def blob_to_coeffs(blob):
    """fk20/prover.rs:177-180: c = INTT_4096(BRP(e))"""
    return intt(brp(blob_to_scalars(blob)))


def compute_cells(blob):
    """fk20/prover.rs:158-165: E = BRP(NTT_8192(c || 0)); 128 chunks of 64"""
    c = blob_to_coeffs(blob)
    E = brp(ntt(c + [0] * FIELD_ELEMENTS_PER_BLOB))
    cells = []
    for k in range(CELLS_PER_EXT_BLOB):
        cells.append(b"".join(x.to_bytes(32, "big") for x in E[64 * k : 64 * k + 64]))
    return cells


def naive_h_commitments(coeffs, srs):
    """fk20/naive.rs:31-59: h_i = sum_j c[j+64 i] [tau^j], i = 1..64"""
    hs = []
    for i in range(1, 65):
        sub = coeffs[64 * i :]
        hs.append(g1_lincomb(srs[: len(sub)], sub))
    return hs


def g1_ntt(points, w):
    n = len(points)
    if n == 1:
        return list(points)
    e = g1_ntt(points[0::2], w * w % R)
    o = g1_ntt(points[1::2], w * w % R)
    out = [INF] * n
    t = 1
    for i in range(n // 2):
        x = g1_mul(o[i], t)
        out[i] = g1_add(e[i], x)
        out[i + n // 2] = g1_add(e[i], g1_neg(x))
        t = t * w % R
    return out


def naive_proofs(coeffs, srs):
    """fk20/naive.rs:61-93: proofs = BRP(NTT_128^{G1}(h_1..h_64, O^64))"""
    hs = naive_h_commitments(coeffs, srs)
    pr = brp(g1_ntt(hs + [INF] * 64, root_of_unity(128)))
    return [g1_compress(p) for p in pr]


def load_trusted_setup_json(path):
    import json

    with open(path) as f:
        d = json.load(f)
    return d
