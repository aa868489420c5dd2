"""CPU checks of the DEVICE arithmetic (rust-eth-kzg_b200/csrc/{mp,field,g1,g1_mul}.cuh) compiled with the
PTX carry-chain primitives replaced by their host emulation (tests/host_emu/*.cpp).  The expected
values come from Python big integers and the oracle's affine curve arithmetic (oracle/pyref.py).
This is the only way to exercise the limb-level algorithms in the GPU-less build container; the
same functions are re-checked on the real device by tests/test_gpu_primitives.py."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rust-eth-kzg_b200", "csrc")
P, R = pyref.P, pyref.R


def _build(name, defines=(), suffix=""):
    src = os.path.join(ROOT, "tests", "host_emu", name + ".cpp")
    out = os.path.join(ROOT, "tests", "host_emu", name + suffix + ".so")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-w", "-x", "c++", "-std=c++17", "-shared", "-fPIC", "-I" + CSRC] + ["-D" + d for d in defines] + ["-o", out, src])
    return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def emu_field():
    return _build("emu_field")


@pytest.fixture(scope="module")
def emu_g1():
    return _build("emu_g1")


def arr(x, n):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


@pytest.mark.parametrize("mod,n,pre", [(P, 12, "fp"), (R, 8, "fr")])
def test_field_ops(emu_field, mod, n, pre):
    rng = random.Random(1)
    Rm = 1 << (32 * n)
    Ri = pow(Rm, -1, mod)
    edge = [0, 1, 2, mod - 1, mod - 2, Rm % mod, (mod - 1) // 2, (mod + 1) // 2]
    vals = edge + [rng.randrange(mod) for _ in range(200)]
    out = (ctypes.c_uint32 * n)()
    for a in vals + [rng.randrange(mod) | (1 << 31) | (1 << 63) | (1 << 95) for _ in range(50)] + [mod - 1 - (1 << k) for k in range(0, 32 * n - 4, 7)]:
        a %= mod
        getattr(emu_field, "emu_%s_sqr" % pre)(arr(a, n), out)
        assert val(out) == a * a * Ri % mod, hex(a)
    if pre == "fp":  # fused a*b + c*d with one reduction
        for _ in range(300):
            a, b, c, d = (rng.choice(vals) if rng.random() < 0.3 else rng.randrange(mod) for _ in range(4))
            emu_field.emu_fp_mul2(arr(a, n), arr(b, n), arr(c, n), arr(d, n), out)
            assert val(out) == (a * b + c * d) * Ri % mod
        emu_field.emu_fp_mul2(arr(mod - 1, n), arr(mod - 1, n), arr(mod - 1, n), arr(mod - 1, n), out)
        assert val(out) == 2 * (mod - 1) ** 2 * Ri % mod
    for a in vals:
        for b in rng.sample(vals, 6) + edge:
            getattr(emu_field, "emu_%s_mul" % pre)(arr(a, n), arr(b, n), out)
            assert val(out) == a * b * Ri % mod
            getattr(emu_field, "emu_%s_add" % pre)(arr(a, n), arr(b, n), out)
            assert val(out) == (a + b) % mod
            getattr(emu_field, "emu_%s_sub" % pre)(arr(a, n), arr(b, n), out)
            assert val(out) == (a - b) % mod
    for a in vals[1:30]:
        getattr(emu_field, "emu_%s_inv" % pre)(arr(a, n), out)
        assert val(out) == pow(a * Ri % mod, -1, mod) * Rm % mod
    for a in [0, 1, R - 1, R, R + 1, 2**256 - 1]:
        assert emu_field.emu_fr_ge_mod(arr(a, 8)) == (1 if a >= R else 0)
    for a in [0, (P - 1) // 2, (P - 1) // 2 + 1, P - 1]:
        assert emu_field.emu_fp_gt_half(arr(a, 12)) == (1 if a > (P - 1) // 2 else 0)


def _pts(rng, n):
    return [pyref.g1_mul(pyref.G1_GEN, rng.randrange(1, R)) for _ in range(n)]


def test_g1_compress_roundtrip(emu_g1):
    rng = random.Random(2)
    out = ctypes.create_string_buffer(48)
    for pt in _pts(rng, 8) + [None]:
        b = pyref.g1_compress(pt)
        assert emu_g1.emu_g1_decompress_roundtrip(b, out) == 0
        assert out.raw == b
    # malformed encodings: no compression flag, x >= p, not on curve, dirty infinity
    bad = [bytes(48), b"\x9f" + b"\xff" * 47, b"\xc0" + b"\x00" * 46 + b"\x01", b"\xe0" + b"\x00" * 47]
    x = 1
    while True:  # find an x with no curve point
        if pow((x**3 + 4) % P, (P - 1) // 2, P) != 1:
            break
        x += 1
    bad.append(bytes([0x80]) + x.to_bytes(48, "big")[1:])
    for b in bad:
        assert emu_g1.emu_g1_decompress_roundtrip(b, out) != 0


def test_g1_binops(emu_g1):
    rng = random.Random(3)
    pts = _pts(rng, 4)
    out = ctypes.create_string_buffer(48)
    cases = [(a, b) for a in pts[:3] for b in pts[:3]] + [(pts[0], None), (None, pts[1]), (None, None), (pts[0], pyref.g1_neg(pts[0]))]
    for a, b in cases:
        pa, pb = pyref.g1_compress(a), pyref.g1_compress(b)
        s = pyref.g1_compress(pyref.g1_add(a, b))
        d = pyref.g1_compress(pyref.g1_add(a, pyref.g1_neg(b)))
        dbl = pyref.g1_compress(pyref.g1_add(a, a))
        for op, exp in [(0, s), (1, d), (2, s), (3, s), (4, s), (5, d), (6, dbl), (7, dbl)]:
            assert emu_g1.emu_g1_binop(op, pa, pb, out) == 0
            assert out.raw == exp, (op, a is None, b is None)


def test_g1_chains(emu_g1):
    rng = random.Random(4)
    out = ctypes.create_string_buffer(48)
    for trial in range(3):
        n = 10
        pts = _pts(rng, n)
        if trial == 1:  # force P+P, P-P and identity inside the chain
            pts[3] = pts[2]; pts[5] = None; pts[1] = pts[0]
        negs = [rng.randrange(2) for _ in range(n)]
        if trial == 1:
            negs[0], negs[1] = 0, 1   # P - P at the start -> identity accumulator
            negs[2], negs[3] = 1, 1   # -P -P -> doubling
        exp = None
        for p, s in zip(pts, negs):
            exp = pyref.g1_add(exp, pyref.g1_neg(p) if s else p)
        buf = b"".join(pyref.g1_compress(p) for p in pts)
        for mode in range(4):
            assert emu_g1.emu_g1_chain(mode, n, buf, bytes(negs), out) == 0
            assert out.raw == pyref.g1_compress(exp), (trial, mode)
    # two equal half-chains (xyzz_add / jac_add doubling branch) and opposite half-chains (identity)
    pts = _pts(rng, 3)
    buf = b"".join(pyref.g1_compress(p) for p in pts + pts)
    tot = None
    for p in pts:
        tot = pyref.g1_add(tot, p)
    for mode in (2, 3):
        assert emu_g1.emu_g1_chain(mode, 6, buf, bytes(6), out) == 0
        assert out.raw == pyref.g1_compress(pyref.g1_add(tot, tot))
        assert emu_g1.emu_g1_chain(mode, 6, buf, bytes([0, 0, 0, 1, 1, 1]), out) == 0
        assert out.raw == pyref.g1_compress(None)


def test_g1_scalar_mul(emu_g1):
    rng = random.Random(5)
    out = ctypes.create_string_buffer(48)
    pt = _pts(rng, 1)[0]
    pb = pyref.g1_compress(pt)
    for k in [0, 1, 2, R - 1, rng.randrange(R), rng.randrange(2**256)]:
        assert emu_g1.emu_g1_mul_u256(pb, arr(k, 8), out) == 0
        assert out.raw == pyref.g1_compress(pyref.g1_mul(pt, k % R))
    # GLV split of a variable scalar (the verifiers' scalar multiplications): edge values around multiples of lambda
    lam = 0xac45a4010001a40200000000ffffffff
    for k in [0, 1, 8, 9, lam - 1, lam, lam + 1, 2 * lam - 1, 2 * lam, lam * lam, lam * lam + lam, R - 1, R - 2, (1 << 128) - 1, 1 << 128,
              0x8888888888888888888888888888888888888888888888888888888888888888 % R] + [rng.randrange(R) for _ in range(6)]:
        assert emu_g1.emu_g1_mul_fr_glv(pb, arr(k, 8), out) == 0
        assert out.raw == pyref.g1_compress(pyref.g1_mul(pt, 2 * k % R)), hex(k)
    w128 = pyref.root_of_unity(128)
    for e in [0, 1, 32, 63, 64, 127]:
        assert emu_g1.emu_g1_mul_twiddle(pb, e, out) == 0
        assert out.raw == pyref.g1_compress(pyref.g1_mul(pt, 2 * pow(w128, e, R) % R)), e
    # the op-list ladder (common-Z odd-multiples table + width-5 NAF) over every twiddle
    for e in range(128):
        assert emu_g1.emu_g1_mul_twiddle_ops(pb, e, out) == 0
        assert out.raw == pyref.g1_compress(pyref.g1_mul(pt, 2 * pow(w128, e, R) % R)), e


def _booth_ref(s, t, w):
    """literal transcription target: SURVEY.md Appendix A.4 closed form"""
    v = ((s << 1) >> (t * w)) & ((1 << (w + 1)) - 1)
    return ((v + 1) >> 1) - ((v >> w) << w)


def test_booth_digits(emu_g1):
    rng = random.Random(6)
    for s in [0, 1, R - 1, 2**255 - 1, 2**256 - 1] + [rng.randrange(R) for _ in range(40)]:
        for w in (4, 8, 9, 12, 13, 16):
            nw = 255 // w + 1
            ds = [emu_g1.emu_booth_digit(arr(s, 8), t, w) for t in range(nw + 1)]
            assert ds == [_booth_ref(s, t, w) for t in range(nw + 1)]
            if s < 2**255:
                assert sum(d << (t * w) for t, d in enumerate(ds[:nw])) == s
                assert all(abs(d) <= 1 << (w - 1) for d in ds)


def test_subgroup_check_matches_naive(emu_g1):
    """g1a_in_subgroup (x^2 P == phi(P) + P) against the definition r*P == O, on points of G1, random curve points
    (almost surely outside G1), pure cofactor-subgroup points r*Q, the order-3 points (0, +-2) and sums of both kinds."""
    rng = random.Random(11)

    def curve_point():
        while True:
            x = rng.randrange(P)
            y2 = (x * x * x + 4) % P
            y = pow(y2, (P + 1) // 4, P)
            if y * y % P == y2:
                return (x, y if rng.random() < 0.5 else P - y)

    pts = [pyref.G1_GEN, pyref.g1_mul(pyref.G1_GEN, rng.randrange(1, R)), (0, 2), (0, P - 2)]
    for _ in range(6):
        q = curve_point()
        pts.append(q)
        cof = pyref.g1_mul_raw(q, R)                      # in the cofactor subgroup
        if cof is not pyref.INF:
            ca = cof
            pts.append(ca)
            pts.append(pyref.g1_add(ca, pyref.g1_mul(pyref.G1_GEN, rng.randrange(1, R))))
    n_in = n_out = 0
    for pt in pts:
        want = 1 if pyref.g1_mul_raw(pt, R) is pyref.INF else 0
        got = emu_g1.emu_g1_in_subgroup(pyref.g1_compress(pt))
        assert got == want, pt
        n_in += want
        n_out += 1 - want
    assert n_in >= 2 and n_out >= 10


def test_msm_table_window_rule(emu_g1):
    """merged top window (csrc/kzg_device.cuh MsmTable::set_window): rtop = 2^(255 - w(nw-1)) + 1 values of the top digit,
    mg = largest of 4, 2, 1 points whose combined top digits still index one slice (rtop^mg - 1 <= 2^(w-1))"""
    out = (ctypes.c_int * 4)()
    for w, want in {8: (32, 128, 129, 1), 10: (26, 512, 33, 1), 12: (22, 2048, 9, 2), 13: (20, 4096, 257, 1), 14: (19, 8192, 9, 4),
                    15: (18, 16384, 2, 4), 16: (16, 32768, 32769, 1)}.items():
        emu_g1.emu_set_window(w, out)
        assert tuple(out) == want, (w, tuple(out))
        nw, half, rtop, mg = out
        assert w * (nw - 1) <= 255 < w * nw + 1 and rtop ** mg - 1 <= half


@pytest.mark.parametrize("mod,n,pre", [(P, 12, "fp"), (R, 8, "fr")])
def test_field_mul_adversarial_limbs(emu_field, mod, n, pre):
    """the shipped multipliers on limb patterns that random operands never produce (all-ones / zero / top-bit limbs): every
    carry of the interleaved Montgomery rows ripples as far as it can.  A carry lost into a full limb is a 2^-32 event per
    row on random data, invisible to random tests and to a few hundred parity blobs."""
    rng = random.Random(77)
    Ri = pow(1 << (32 * n), -1, mod)
    B = 1 << 32
    pats = [0xFFFFFFFF, 0, 0xFFFFFFFE, 1, 0x80000000, 0x7FFFFFFF]
    out = (ctypes.c_uint32 * n)()

    def operand():
        return sum((rng.choice(pats) if rng.random() < 0.75 else rng.randrange(B)) << (32 * i) for i in range(n)) % mod

    for _ in range(1500):
        a, b, c, d = operand(), operand(), operand(), operand()
        getattr(emu_field, "emu_%s_mul" % pre)(arr(a, n), arr(b, n), out)
        assert val(out) == a * b * Ri % mod, (hex(a), hex(b))
        getattr(emu_field, "emu_%s_sqr" % pre)(arr(a, n), out)
        assert val(out) == a * a * Ri % mod, hex(a)
        if pre == "fp":
            emu_field.emu_fp_mul2(arr(a, n), arr(b, n), arr(c, n), arr(d, n), out)
            assert val(out) == (a * b + c * d) * Ri % mod, (hex(a), hex(b), hex(c), hex(d))
        getattr(emu_field, "emu_%s_add" % pre)(arr(a, n), arr(b, n), out)
        assert val(out) == (a + b) % mod
        getattr(emu_field, "emu_%s_sub" % pre)(arr(a, n), arr(b, n), out)
        assert val(out) == (a - b) % mod


# ---------------------------------------------------------------------------------------------------------------
# the shared-memory-operand Fp interpreter of K4 / K5 (csrc/fpvm.cuh) on the CPU: the same `step` the device runs,
# every program of tools/fpvm_asm.py, against the Python emulation of the programs (tests/test_fpvm_programs.py
# checks those against the group law)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_fpvm():
    return _build("emu_fpvm")


def test_fpvm_interpreter_matches_program_emulation(emu_fpvm):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fpvm_asm as A
    words = [w for p in A.PROGRAMS.values() for w in p]
    assert emu_fpvm.emu_fpvm_words() == len(words)
    emu_fpvm.emu_fpvm_word.restype = ctypes.c_uint32
    assert [emu_fpvm.emu_fpvm_word(i) for i in range(len(words))] == words
    emu_fpvm.emu_fpvm_run.restype = ctypes.c_uint32
    rng = random.Random(11)
    Rm = 1 << 384
    pc = 0
    for name, prog in A.PROGRAMS.items():
        for trial in range(6):
            # plain values; trial 0 plants equal operands / zeros so the zero masks and the borrow paths are hit
            vals = [rng.randrange(P) for _ in range(8)]
            if trial == 0:
                vals[3] = vals[0]; vals[4] = vals[1]; vals[5] = 0; vals[6] = P - 1
            exp = list(vals)
            mask = A.emulate(prog, exp)
            buf = (ctypes.c_uint32 * 96)()
            for s in range(8):
                m = vals[s] * Rm % P
                for l in range(12):
                    buf[12 * s + l] = (m >> (32 * l)) & 0xFFFFFFFF
            got_mask = emu_fpvm.emu_fpvm_run(pc, len(prog), 1, buf)
            Ri = pow(Rm, -1, P)
            got = [sum(int(buf[12 * s + l]) << (32 * l) for l in range(12)) for s in range(8)]
            assert all(g < P for g in got), name
            assert [g * Ri % P for g in got] == exp, name
            assert got_mask == mask, name
        pc += len(prog)


def test_fp_inversion_by_division_steps(emu_field):
    """csrc/fp_inv_gcd.cuh (the shared inversion of the batched-affine MSM): Montgomery form in and out, against pow(-1)"""
    rng = random.Random(7)
    Rm = 1 << 384
    out = (ctypes.c_uint32 * 12)()
    vals = [1, 2, 3, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, Rm % P, pow(Rm, -1, P), 1 << 380, (1 << 381) - 1 - P % 7]
    vals += [1 << k for k in range(0, 381, 13)] + [P - (1 << k) for k in range(0, 380, 17)]
    vals += [rng.randrange(1, P) for _ in range(2000)]
    for x in vals:
        x %= P
        emu_field.emu_fp_inv_gcd(arr(x * Rm % P, 12), out)
        assert val(out) == pow(x, -1, P) * Rm % P, hex(x)
    emu_field.emu_fp_inv_gcd(arr(0, 12), out)
    assert val(out) == 0


@pytest.fixture(scope="module")
def emu_ntt():
    return _build("emu_ntt")


def test_g1_ntt_radix4_units_equal_radix2_units_on_points(emu_ntt):
    """the device code of both G1-NTT kernels (g1_ntt_units.cuh), unit by unit on the CPU: every prefix of radix-4 super-phases
    leaves the 128 points exactly where the radix-2 phases leave them -- random points, identities, repeated and opposite points
    (the doubling / cancellation branches of the additions)"""
    rng = random.Random(11)
    gens = [pyref.g1_mul(pyref.G1_GEN, rng.randrange(1, R)) for _ in range(12)]
    for trial in range(2):
        if trial == 0:
            pts = [pyref.g1_mul(gens[i % 12], rng.randrange(1, 1 << 20)) for i in range(128)]
        else:       # few distinct values: plenty of P + P, P - P and identities on the way
            pts = [None if i % 5 == 0 else (gens[i % 2] if i % 3 else pyref.g1_neg(gens[i % 2])) for i in range(128)]
        buf = b"".join(pyref.g1_compress(p) for p in pts)
        for phases in (2, 6, 8, 10, 14):
            a, b = ctypes.create_string_buffer(128 * 48), ctypes.create_string_buffer(128 * 48)
            assert emu_ntt.emu_g1_ntt(0, phases, buf, a) == 0
            assert emu_ntt.emu_g1_ntt(1, phases, buf, b) == 0
            bad = [i for i in range(128) if a.raw[48 * i:48 * i + 48] != b.raw[48 * i:48 * i + 48]]
            assert not bad, "trial %d, %d phases: positions %r differ" % (trial, phases, bad[:16])
