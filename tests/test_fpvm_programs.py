"""The straight-line Fp programs of the hot kernels (tools/fpvm_asm.py -> csrc/fpvm_programs.inc) against the affine
group law, on Python integers -- no GPU, no compiled code.  The interpreter itself (csrc/fpvm.cuh) is checked on the
CPU by tests/test_host_emu.py::test_fpvm_* and on the device by the -m gpu parity tests."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fpvm_asm as A  # noqa: E402
from oracle import pyref  # noqa: E402

P = pyref.P
G = pyref.G1_GEN
rng = random.Random(7)


def rand_pt():
    return pyref.g1_mul(G, rng.randrange(1, pyref.R))


def jac(pt, z=None):
    z = z or rng.randrange(1, P)
    return [pt[0] * z * z % P, pt[1] * pow(z, 3, P) % P, z]


def from_jac(X, Y, Z):
    if Z % P == 0:
        return pyref.INF
    zi = pow(Z, -1, P)
    return (X * zi * zi % P, Y * pow(zi, 3, P) % P)


def xyzz(pt):
    z = rng.randrange(1, P)
    return [pt[0] * z * z % P, pt[1] * pow(z, 3, P) % P, z * z % P, pow(z, 3, P)]


def from_xyzz(X, Y, ZZ, ZZZ):
    if ZZ % P == 0:
        return pyref.INF
    return (X * pow(ZZ, -1, P) % P, Y * pow(ZZZ, -1, P) % P)


def test_inc_file_is_current():
    words = [w for p in A.PROGRAMS.values() for w in p]
    txt = open(os.path.join(ROOT, "rust-eth-kzg_b200", "csrc", "fpvm_programs.inc")).read()
    assert "FPVM_PROG_WORDS = %d;" % len(words) in txt
    assert ", ".join("0x%08xu" % w for w in words) in txt, "run tools/fpvm_asm.py"
    for prog in A.PROGRAMS.values():
        for ins in prog:  # eight slots per thread
            fields = [(ins >> s) & 15 for s in (20, 16, 12, 8, 4, 0)]
            assert all(f < 8 or f == A.NONE for f in fields)


def test_k4_xyzz_madd():
    for _ in range(20):
        a, e = rand_pt(), rand_pt()
        s = xyzz(a) + [e[0], e[1], 0, 0]
        m = A.emulate(A.K4["XYZZ_MADD_A"], s)
        assert m == 0
        A.emulate(A.K4["XYZZ_MADD_B"], s)
        assert from_xyzz(*s[:4]) == pyref.g1_add(a, e)
        assert (s[2] * s[2] * s[2] - s[3] * s[3]) % P == 0  # ZZ^3 == ZZZ^2 stays true
    a = rand_pt()
    s = xyzz(a) + [a[0], a[1], 0, 0]
    assert A.emulate(A.K4["XYZZ_MADD_A"], s) == 3          # same point: P == 0 and R == 0
    s = xyzz(a) + [a[0], P - a[1], 0, 0]
    assert A.emulate(A.K4["XYZZ_MADD_A"], s) == 1          # opposite point: only P == 0


def test_k4_xyzz_add_and_to_jac():
    for _ in range(20):
        a, b = rand_pt(), rand_pt()
        s = xyzz(a) + xyzz(b)
        assert A.emulate(A.K4["XYZZ_ADD_A"], s) & 0b1010 == 0
        A.emulate(A.K4["XYZZ_ADD_B"], s)
        assert from_xyzz(*s[:4]) == pyref.g1_add(a, b)
        A.emulate(A.K4["XYZZ_TO_JAC"], s)
        assert from_jac(*s[:3]) == pyref.g1_add(a, b)
    a = rand_pt()
    s = xyzz(a) + xyzz(a)
    assert A.emulate(A.K4["XYZZ_ADD_A"], s) & 0b1010 == 0b1010
    s = xyzz(a) + xyzz(pyref.g1_neg(a))
    assert A.emulate(A.K4["XYZZ_ADD_A"], s) & 0b1010 == 0b0010


def test_k5_dbl_madd_add():
    for _ in range(20):
        a, e = rand_pt(), rand_pt()
        s = jac(a) + [0] * 5
        A.emulate(A.K5["JAC_DBL"], s)
        assert from_jac(*s[:3]) == pyref.g1_add(a, a)
        s = jac(a) + [e[0], e[1], 0, 0, 0]
        assert A.emulate(A.K5["JAC_MADD_A"], s) & 0b1010 == 0
        A.emulate(A.K5["JAC_MADD_B"], s)
        assert from_jac(*s[:3]) == pyref.g1_add(a, e)
        s = jac(a) + jac(e) + [0, 0]
        assert A.emulate(A.K5["JAC_ADD_A"], s) & 0x88 == 0
        A.emulate(A.K5["JAC_ADD_B"], s)
        assert from_jac(*s[:3]) == pyref.g1_add(a, e)
    a = rand_pt()
    s = jac(a) + [a[0], a[1], 0, 0, 0]
    assert A.emulate(A.K5["JAC_MADD_A"], s) & 0b1010 == 0b1010
    A.emulate(A.K5["JAC_DBL"], s)                          # part A leaves the accumulator intact: the kernel doubles it
    assert from_jac(*s[:3]) == pyref.g1_add(a, a)
    s = jac(a) + [a[0], P - a[1], 0, 0, 0]
    assert A.emulate(A.K5["JAC_MADD_A"], s) & 0b1010 == 0b0010
    # equal points in the general addition: (U1, S1, Z1*Z2) is P1 again, the kernel doubles that
    s = jac(a) + jac(a) + [0, 0]
    assert A.emulate(A.K5["JAC_ADD_A"], s) & 0x88 == 0x88
    A.emulate(A.K5["MUL_Z_QZ"], s)
    assert from_jac(*s[:3]) == a
    A.emulate(A.K5["JAC_DBL"], s)
    assert from_jac(*s[:3]) == pyref.g1_add(a, a)
    s = jac(a) + jac(pyref.g1_neg(a)) + [0, 0]
    assert A.emulate(A.K5["JAC_ADD_A"], s) & 0x88 == 0x08


def load_twiddle_ops():
    rows = []
    for ln in open(os.path.join(ROOT, "rust-eth-kzg_b200", "csrc", "twiddle_ops.inc")):
        ln = ln.strip()
        if ln.startswith("{") and len(ln) > 2:
            v = [int(x) for x in ln.strip("{},").split(",")]
            rows.append(v[1:1 + v[0]])
    return rows


BETA = 0x1a0111ea397fe699ec02408663d4de85aa0d857d89759ad4897d29650fb85f9b409427eb4f49fffd8bfd00000000aaac


def ladder(pt, ops):
    """k5_mul_ops_vm of csrc/kzg_kernels.cu, statement by statement, on the emulator"""
    G_ = {}
    s = jac(pt) + [0] * 5
    G_[21], G_[22], G_[23] = s[0], s[1], s[2]
    A.emulate(A.K5["JAC_DBL"], s)
    s[3], s[4] = G_[21], G_[22]
    A.emulate(A.K5["TBL_ISO"], s)
    G_[25] = s[2]
    s[3], s[4] = s[0], s[1]
    s[0], s[1] = s[5], s[6]
    G_[0], G_[1] = s[0], s[1]
    s[2] = G_[23]
    for i in range(1, 8):
        A.emulate(A.K5["TBL_MADDZR_A"], s)
        G_[3 * i + 2] = s[5]
        A.emulate(A.K5["TBL_MADDZR_B"], s)
        G_[3 * i], G_[3 * i + 1] = s[0], s[1]
    s[5], s[6] = s[2], G_[25]
    A.emulate(A.K5["MUL_T0_T1"], s)
    G_[24] = s[5]
    s[0] = G_[23]
    s[6] = BETA
    s[3] = G_[21]
    A.emulate(A.K5["TBL_BETA"], s)
    G_[23] = s[7]
    for i in range(6, -1, -1):
        s[3], s[4] = G_[3 * i], G_[3 * i + 1]
        if i:
            s[5] = G_[3 * i + 2]
        A.emulate(A.K5["TBL_RESCALE"], s)
        G_[3 * i], G_[3 * i + 1], G_[3 * i + 2] = s[3], s[4], s[7]
    inf = True
    for op in ops:
        d = op >> 8
        if d and not inf:
            for _ in range(d):
                A.emulate(A.K5["JAC_DBL"], s)
        if op & 0x20:
            idx = op & 7
            ex = G_[3 * idx + (2 if op & 0x10 else 0)]
            ey = G_[3 * idx + 1]
            if op & 8:
                ey = (P - ey) % P
            if inf:
                s[0], s[1], s[2] = ex, ey, 1
                inf = False
            else:
                s[3], s[4] = ex, ey
                m = A.emulate(A.K5["JAC_MADD_A"], s)
                assert m & 2 == 0
                A.emulate(A.K5["JAC_MADD_B"], s)
    if inf:
        return pyref.INF
    s[5] = G_[24]
    A.emulate(A.K5["MUL_Z_T0"], s)
    return from_jac(*s[:3])


@pytest.mark.parametrize("e", [1, 2, 31, 32, 64, 65, 127])
def test_k5_ladder_against_scalar_mul(e):
    assert pow(BETA, 3, P) == 1 and BETA != 1
    ops = load_twiddle_ops()
    w = pow(pyref.root_of_unity(128), e, pyref.R)
    pt = rand_pt()
    assert ladder(pt, ops[e]) == pyref.g1_mul(pt, w)
