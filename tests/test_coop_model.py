"""A limb-exact Python model of the warp-cooperative arithmetic of csrc/g1_coop.cuh (four lanes per field element, three 32-bit limbs
each), checked against big-integer arithmetic: the CPU-side statement of what the CUDA code does, word for word --
  * the Montgomery multiplier: operand scanning, deferred carries in two spare words per lane, the word shifted in from the next lane
    consumed ONE ROW LATER, carries folded into the next lane at the end; inputs and outputs in [0, 2p);
  * carry-lookahead between lanes from "generate" / "propagate" votes: carries into the lanes = (X + G) ^ X ^ G, X = G | P;
  * semi-reduced addition / subtraction (one conditional -+ 2p);
  * the addition-light point formulas (doubling scaled by 1/2 with 3/2 folded into a product, madd-2004-hmv) against affine
    arithmetic on the curve.
The GPU tests compare the kernel itself with the radix-2 kernel and the oracle (tests/test_gpu_fk20.py)."""
import random

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 1 << 384
M0 = 0xfffcfffd
W = 0xffffffff
assert (P * M0 + 1) % (1 << 32) == 0


def lanes(x):
    """value -> 4 lanes x 3 limbs"""
    return [[(x >> (32 * (3 * l + k))) & W for k in range(3)] for l in range(4)]


def value(v):
    return sum(v[l][k] << (32 * (3 * l + k)) for l in range(4) for k in range(3))


P_L, P2_L = lanes(P), lanes(2 * P)


def mad3(t, a, s):
    """(t0, t1, t2, h0, h1) += a(3 limbs) * s, as the two carry chains of the kernel (exact: 160-bit accumulator)"""
    acc = sum(t[i] << (32 * i) for i in range(5)) + (a[0] + (a[1] << 32) + (a[2] << 64)) * s
    assert acc < 1 << 160
    for i in range(5):
        t[i] = (acc >> (32 * i)) & W


def cmul(a, b):
    T = [[0] * 5 for _ in range(4)]
    ypend = [0] * 4
    for i in range(12):
        bi = b[i // 3][i % 3]                                   # broadcast from lane i / 3
        for l in range(4):
            mad3(T[l], a[l], bi)
        m = (T[0][0] * M0) & W                                  # lane 0's low word, broadcast
        for l in range(4):                                      # the word shifted in at the end of the PREVIOUS row lands in t2 now
            acc = T[l][2] + (T[l][3] << 32) + (T[l][4] << 64) + ypend[l]
            T[l][2], T[l][3], T[l][4] = acc & W, (acc >> 32) & W, (acc >> 64) & W
        for l in range(4):
            mad3(T[l], P_L[l], m)
        assert T[0][0] == 0
        ypend = [T[l + 1][0] if l < 3 else 0 for l in range(4)]  # shuffle down
        for l in range(4):
            T[l] = [T[l][1], T[l][2], T[l][3], T[l][4], 0]
            assert T[l][3] < 8                                   # the deferred carries stay tiny
    for l in range(4):
        acc = T[l][2] + (T[l][3] << 32) + ypend[l]
        T[l][2], T[l][3] = acc & W, acc >> 32
    passes = 0
    while True:                                                  # fold: at most three passes
        cin = [0] + [T[l][3] for l in range(3)]
        for l in range(4):
            acc = T[l][0] + (T[l][1] << 32) + (T[l][2] << 64) + cin[l]
            T[l][0], T[l][1], T[l][2], T[l][3] = acc & W, (acc >> 32) & W, (acc >> 64) & W, acc >> 96
        passes += 1
        if not any(T[l][3] for l in range(3)):
            break
        assert passes < 3
    assert T[3][3] == 0
    return [T[l][:3] for l in range(4)]


def resolve(gen, prop):
    """carries INTO the four lanes and out of the group from the generate / propagate votes"""
    G = sum(1 << l for l in range(4) if gen[l])
    Pm = sum(1 << l for l in range(4) if prop[l])
    X = G | Pm
    C = (X + G) ^ X ^ G
    return [(C >> l) & 1 for l in range(4)], (C >> 4) & 1


def add_raw(a, b):
    s = [a[l][0] + (a[l][1] << 32) + (a[l][2] << 64) + b[l][0] + (b[l][1] << 32) + (b[l][2] << 64) for l in range(4)]
    gen = [x >> 96 for x in s]
    s = [x & ((1 << 96) - 1) for x in s]
    cin, out = resolve(gen, [x == (1 << 96) - 1 for x in s])
    s = [(s[l] + cin[l]) & ((1 << 96) - 1) for l in range(4)]
    return [[(x >> (32 * k)) & W for k in range(3)] for x in s], out


def sub_raw(a, b):
    d = [a[l][0] + (a[l][1] << 32) + (a[l][2] << 64) - (b[l][0] + (b[l][1] << 32) + (b[l][2] << 64)) for l in range(4)]
    gen = [x < 0 for x in d]
    d = [x % (1 << 96) for x in d]
    bin_, out = resolve(gen, [x == 0 for x in d])
    d = [(d[l] - bin_[l]) % (1 << 96) for l in range(4)]
    return [[(x >> (32 * k)) & W for k in range(3)] for x in d], out


def cadd(a, b):
    s, out = add_raw(a, b)
    assert out == 0
    d, below = sub_raw(s, P2_L)
    return s if below else d


def csub(a, b):
    d, below = sub_raw(a, b)
    r, _ = add_raw(d, P2_L if below else lanes(0))
    return r


EDGE = [0, 1, P - 1, P, P + 1, 2 * P - 1, (1 << 96) - 1, (1 << 192) - 1, ((1 << 96) - 1) << 96, (1 << 381) - 1, 2 * P - (1 << 96), (1 << 288)]


def _samples(rng, n):
    xs = [e for e in EDGE if e < 2 * P]
    return xs + [rng.randrange(2 * P) for _ in range(n)]


def test_resolve_is_a_carry_chain():
    for G in range(16):
        for Pm in range(16):
            gen, prop = [(G >> l) & 1 for l in range(4)], [(Pm >> l) & 1 for l in range(4)]
            c, want = 0, []
            for l in range(4):
                want.append(c)
                c = 1 if gen[l] else (c if prop[l] else 0)
            assert resolve(gen, prop) == (want, c)


def test_multiplier_semi_reduced():
    rng = random.Random(1)
    xs = _samples(rng, 40)
    rinv = pow(R, -1, P)
    for x in xs:
        for y in xs[::3]:
            r = value(cmul(lanes(x), lanes(y)))
            assert r < 2 * P and r % P == x * y * rinv % P, (x, y)


def test_add_sub_semi_reduced():
    rng = random.Random(2)
    xs = _samples(rng, 40)
    for x in xs:
        for y in xs:
            s, d = value(cadd(lanes(x), lanes(y))), value(csub(lanes(x), lanes(y)))
            assert s < 2 * P and s % P == (x + y) % P
            assert d < 2 * P and d % P == (x - y) % P


# ---- the point formulas, on field elements mod p (the representation is tested above) ----------------------------------------
def _aff_add(p1, p2):
    (x1, y1), (x2, y2) = p1, p2
    lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P if p1 == p2 else (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def _to_affine(X, Y, Z):
    zi = pow(Z, -1, P)
    return X * zi * zi % P, Y * zi * zi * zi % P


def test_addition_light_formulas():
    rng = random.Random(3)
    G = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
         0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
    assert (G[1] ** 2 - G[0] ** 3 - 4) % P == 0
    Q = _aff_add(G, G)
    three_halves = 3 * pow(2, -1, P) % P
    for _ in range(20):
        # a random Jacobian representative of Q
        z = rng.randrange(1, P)
        X, Y, Z = Q[0] * z * z % P, Q[1] * z ** 3 % P, z
        # doubling scaled by 1/2: m = (3/2) X^2, X3 = m^2 - 2 X Y^2, Y3 = m (X Y^2 - X3) - Y^4, Z3 = Y Z
        a, b = X * X % P, Y * Y % P
        m, xb = a * three_halves % P, X * b % P
        X3 = (m * m - 2 * xb) % P
        Y3 = (m * (xb - X3) - b * b) % P
        Z3 = Y * Z % P
        assert _to_affine(X3, Y3, Z3) == _aff_add(Q, Q)
        # madd-2004-hmv with the affine point G
        zz = Z * Z % P
        h, r = (G[0] * zz - X) % P, (G[1] * Z * zz - Y) % P
        hh = h * h % P
        hhh, v = hh * h % P, X * hh % P
        X3 = (r * r - hhh - 2 * v) % P
        Y3 = (r * (v - X3) - Y * hhh) % P
        assert _to_affine(X3, Y3, Z * h % P) == _aff_add(Q, G)
