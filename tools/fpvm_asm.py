#!/usr/bin/env python3
"""Assembler + reference emulator for the Fp "virtual machine" of the hot kernels (rust-eth-kzg_b200/csrc/fpvm.cuh).

The point formulas of K4 (XYZZ mixed addition) and K5 (Jacobian doubling / mixed addition / addition, odd-multiples
table) are straight-line programs over EIGHT 48-byte slots per thread that live in shared memory; one interpreter
(fpvm_run) holds the only copy of the Montgomery multiplier / squarer in the kernel image and executes them.
Here the programs are written down symbolically, encoded into csrc/fpvm_programs.inc, and can be executed on
Python integers (emulate) so that tests/test_fpvm_programs.py checks every program against the affine group law
without a GPU.

Instruction word:  op[31:28] flag[27:24] dst[23:20] f1[19:16] f2[15:12] f3[11:8] f4[7:4] f5[3:0]   (slot 15 = none)
  MUL  : dst = (f1 [- f2]) * f3 [- f4]; flag bit 0: dst = 2 * dst
  SQR  : dst = (f1 [+ f2])^2 [- f3] [- f4] [- f5]
  MUL2 : dst = f1 * (f2 - f3) - f4 * f5          (one Montgomery reduction for both products)
  LIN  : flag 0: f1 + f2, 1: f1 - f2, 2: 2*f1, 3: 3*f1, 4: 4*f1, 5: 8*f1, 6: -f1, 7: f1

    python tools/fpvm_asm.py          # rewrites rust-eth-kzg_b200/csrc/fpvm_programs.inc
"""
import os

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
NONE = 15
OP_MUL, OP_SQR, OP_MUL2, OP_LIN = 0, 1, 2, 3
LIN_ADD, LIN_SUB, LIN_DBL, LIN_TRI, LIN_QUAD, LIN_OCT, LIN_NEG, LIN_COPY = range(8)


def enc(op, flag, dst, f1, f2=NONE, f3=NONE, f4=NONE, f5=NONE):
    for v in (dst, f1, f2, f3, f4, f5):
        assert 0 <= v <= 15
    return (op << 28) | (flag << 24) | (dst << 20) | (f1 << 16) | (f2 << 12) | (f3 << 8) | (f4 << 4) | f5


def MUL(dst, a, b, sub=NONE, pre=NONE, dbl=False):
    """dst = (a [- pre]) * b [- sub], doubled if dbl"""
    return enc(OP_MUL, 1 if dbl else 0, dst, a, pre, b, sub)


def SQR(dst, a, *subs, plus=NONE):
    """dst = (a [+ plus])^2 - subs..."""
    s = list(subs) + [NONE] * (3 - len(subs))
    return enc(OP_SQR, 0, dst, a, plus, s[0], s[1], s[2])


def MUL2(dst, a, b, b2, c, d):
    """dst = a * (b - b2) - c * d"""
    return enc(OP_MUL2, 0, dst, a, b, b2, c, d)


def LIN(kind, dst, a, b=NONE):
    return enc(OP_LIN, kind, dst, a, b)


# ---------------------------------------------------------------------------------------------------------------
# K4: XYZZ accumulator += affine table entry (madd-2008-s, 8M + 2S, the two products of Y3 share one reduction)
#   slots: X Y ZZ ZZZ | EX EY (the entry, y already negated for a negative digit) | T PPP
# ---------------------------------------------------------------------------------------------------------------
X, Y, ZZ, ZZZ, EX, EY, T, PPP = range(8)
K4 = {}
K4["XYZZ_MADD_A"] = [          # bit 0 of the result mask: P == 0, bit 1: R == 0  (doubling / cancellation: rare path)
    MUL(EX, EX, ZZ, sub=X),    # P  = x2*ZZ1 - X1
    MUL(EY, EY, ZZZ, sub=Y),   # R  = y2*ZZZ1 - Y1
]
K4["XYZZ_MADD_B"] = [
    SQR(T, EX),                # PP
    MUL(PPP, EX, T),           # PPP
    MUL(EX, X, T),             # Q = X1*PP
    MUL(ZZ, ZZ, T),
    MUL(ZZZ, ZZZ, PPP),
    SQR(X, EY, PPP, EX, EX),   # X3 = R^2 - PPP - 2Q
    MUL2(Y, EY, EX, X, Y, PPP),  # Y3 = R*(Q - X3) - Y1*PPP
]
# acc (slots 0..3) += q (slots 4..7), both XYZZ, in place (add-2008-s, 12M + 2S); q is destroyed
X2, Y2, ZZ2, ZZZ2 = 4, 5, 6, 7
K4["XYZZ_ADD_A"] = [
    MUL(X, X, ZZ2),                # U1
    MUL(X2, X2, ZZ, sub=X),        # P = U2 - U1      (bit 1)
    MUL(Y, Y, ZZZ2),               # S1
    MUL(Y2, Y2, ZZZ, sub=Y),       # R = S2 - S1      (bit 3)
]
K4["XYZZ_ADD_B"] = [
    MUL(ZZ, ZZ, ZZ2),
    MUL(ZZZ, ZZZ, ZZZ2),
    SQR(ZZ2, X2),                  # PP
    MUL(ZZZ2, X2, ZZ2),            # PPP
    MUL(X2, X, ZZ2),               # Q = U1*PP
    MUL(ZZ, ZZ, ZZ2),
    MUL(ZZZ, ZZZ, ZZZ2),
    SQR(X, Y2, ZZZ2, X2, X2),      # X3
    MUL2(Y, Y2, X2, X, Y, ZZZ2),   # Y3 = R*(Q - X3) - S1*PPP
]
# XYZZ -> Jacobian (Z = ZZ*ZZZ = z^5): result in slots X, Y, ZZ
K4["XYZZ_TO_JAC"] = [
    SQR(EX, ZZZ),
    MUL(EX, EX, ZZ),               # a = ZZ*ZZZ^2
    SQR(EY, ZZ),
    MUL(EY, EX, EY),               # c = ZZ^3*ZZZ^2
    MUL(X, X, EX),
    MUL(Y, Y, EY),
    MUL(ZZ, ZZ, ZZZ),
]

# ---------------------------------------------------------------------------------------------------------------
# K5: Jacobian arithmetic.  slots: X Y Z | EX EY | T0 T1 T2
# ---------------------------------------------------------------------------------------------------------------
JX, JY, JZ, JEX, JEY, T0, T1, T2 = range(8)
K5 = {}
# acc = 2*acc (dbl-2009-l, a = 0: 2M + 5S); temps A=EX B=EY C=T0 D=T1
K5["JAC_DBL"] = [
    SQR(JEX, JX),                        # A
    SQR(JEY, JY),                        # B
    SQR(T0, JEY),                        # C
    SQR(T1, JX, JEX, T0, plus=JEY),      # (X+B)^2 - A - C
    MUL(JZ, JY, JZ, dbl=True),           # Z3 = 2*Y*Z
    LIN(LIN_DBL, T1, T1),                # D
    LIN(LIN_TRI, JEX, JEX),              # E = 3A
    SQR(JX, JEX, T1, T1),                # X3 = E^2 - 2D
    LIN(LIN_OCT, T0, T0),                # 8C
    MUL(JY, T1, JEX, sub=T0, pre=JX),    # Y3 = (D - X3)*E - 8C
]
# acc += (EX, EY) affine (madd-2007-bl: 7M + 4S)
K5["JAC_MADD_A"] = [
    SQR(T0, JZ),                         # Z1Z1
    MUL(T1, JEX, T0, sub=JX),            # H = U2 - X1          (bit 1)
    MUL(JEX, JEY, JZ),
    MUL(JEX, JEX, T0, sub=JY, dbl=True),  # r = 2*(S2 - Y1)      (bit 3)
]
K5["JAC_MADD_B"] = [
    SQR(JEY, T1),                        # HH
    SQR(JZ, JZ, T0, JEY, plus=T1),       # Z3 = (Z1+H)^2 - Z1Z1 - HH
    LIN(LIN_QUAD, JEY, JEY),             # I
    MUL(T0, T1, JEY),                    # J
    MUL(T1, JX, JEY),                    # V
    SQR(JX, JEX, T0, T1, T1),            # X3 = r^2 - J - 2V
    MUL(JEY, JY, T0, dbl=True),          # 2*Y1*J
    MUL(JY, T1, JEX, sub=JEY, pre=JX),   # Y3 = r*(V - X3) - 2*Y1*J
]
# acc (slots 0..2) += q (slots 3..5 = X2 Y2 Z2), both Jacobian (add-2007-bl: 11M + 5S); temps 6, 7; q is destroyed
QX, QY, QZ, U0, U1 = 3, 4, 5, 6, 7
K5["JAC_ADD_A"] = [
    SQR(U0, JZ),                         # Z1Z1
    SQR(U1, QZ),                         # Z2Z2
    MUL(JX, JX, U1),                     # U1
    MUL(QX, QX, U0, sub=JX),             # H = U2 - U1          (bit 3)
    MUL(JY, JY, QZ),
    MUL(JY, JY, U1),                     # S1
    MUL(QY, QY, JZ),
    MUL(QY, QY, U0, sub=JY, dbl=True),   # r = 2*(S2 - S1)      (bit 7)
]
K5["JAC_ADD_B"] = [
    SQR(JZ, JZ, U0, U1, plus=QZ),        # (Z1+Z2)^2 - Z1Z1 - Z2Z2
    MUL(JZ, JZ, QX),                     # Z3
    LIN(LIN_DBL, U0, QX),
    SQR(U0, U0),                         # I = (2H)^2
    MUL(U1, QX, U0),                     # J
    MUL(QZ, JX, U0),                     # V
    SQR(JX, QY, U1, QZ, QZ),             # X3 = r^2 - J - 2V
    MUL(U0, JY, U1, dbl=True),           # 2*S1*J
    MUL(JY, QZ, QY, sub=U0, pre=JX),     # Y3 = r*(V - X3) - 2*S1*J
]
# Odd-multiples table of the G1 butterfly ladder (jac_mul_ops in g1_mul.cuh), step by step:
#  TBL_ISO: after JAC_DBL the slots hold d = 2P; EX, EY <- P.x, P.y (re-staged by the kernel).  Map P onto the curve
#           isomorphic by d.z:  T0 = P.x*dz^2, T1 = P.y*dz^3.
K5["TBL_ISO"] = [
    SQR(T2, JZ),
    MUL(T0, JEX, T2),
    MUL(T2, T2, JZ),
    MUL(T1, JEY, T2),
]
#  TBL_MADDZR: cur (slots 0..2, Jacobian) += (EX, EY) = 2P as an affine point of the isomorphic curve
#           (madd-2004-hmv: 8M + 3S, Z3 = Z1*H).  After part A slot T0 holds H = Z3/Z1 (kept by the kernel).
K5["TBL_MADDZR_A"] = [
    SQR(T0, JZ),
    MUL(T1, T0, JZ),
    MUL(T0, T0, JEX, sub=JX),            # H
    MUL(T1, T1, JEY, sub=JY),            # R
    MUL(JZ, JZ, T0),
]
K5["TBL_MADDZR_B"] = [
    SQR(T2, T0),                         # HH
    MUL(T0, T2, T0),                     # HHH
    MUL(T2, T2, JX),                     # V
    SQR(JX, T1, T2, T2, T0),             # X3 = R^2 - 2V - HHH
    MUL(T0, T0, JY),
    MUL(JY, T2, T1, sub=T0, pre=JX),     # Y3 = (V - X3)*R - HHH*Y1
]
#  TBL_RESCALE: slots ZS=0 (running z-ratio), 1, 2 scratch, EX EY = entry, T0 = zr_i, T1 = beta:
#           entry *= (zs^2, zs^3); T2 = beta * x; zs *= zr_i
K5["TBL_RESCALE"] = [
    SQR(1, 0),
    MUL(2, 1, 0),
    MUL(JEX, JEX, 1),
    MUL(JEY, JEY, 2),
    MUL(T2, JEX, T1),
    MUL(0, 0, T0),
]
K5["TBL_BETA"] = [MUL(T2, JEX, T1)]
K5["MUL_Z_T0"] = [MUL(JZ, JZ, T0)]       # Z *= T0
K5["MUL_T0_T1"] = [MUL(T0, T0, T1)]      # T0 *= T1
K5["MUL_Z_QZ"] = [MUL(JZ, JZ, QZ)]       # Z *= Z2: after JAC_ADD_A, (U1, S1, Z1*Z2) is P1 again (the P1 == P2 case doubles it)

PROGRAMS = {"K4_" + k: v for k, v in K4.items()}
PROGRAMS.update({"K5_" + k: v for k, v in K5.items()})


def emulate(prog, slots):
    """run a program on Python integers mod P (plain field values: Montgomery form is transparent to the formulas);
    returns the mask of instructions whose result is zero, like fpvm_run"""
    mask = 0
    for i, ins in enumerate(prog):
        op, fl, d = ins >> 28, (ins >> 24) & 15, (ins >> 20) & 15
        f = [(ins >> s) & 15 for s in (16, 12, 8, 4, 0)]
        g = lambda k: slots[f[k]]
        if op == OP_MUL:
            a = g(0) - (g(1) if f[1] != NONE else 0)
            r = a * g(2) - (g(3) if f[3] != NONE else 0)
            if fl & 1:
                r *= 2
        elif op == OP_SQR:
            a = g(0) + (g(1) if f[1] != NONE else 0)
            r = a * a - sum(g(k) for k in (2, 3, 4) if f[k] != NONE)
        elif op == OP_MUL2:
            r = g(0) * (g(1) - g(2)) - g(3) * g(4)
        else:
            a = g(0)
            r = {LIN_ADD: lambda: a + g(1), LIN_SUB: lambda: a - g(1), LIN_DBL: lambda: 2 * a, LIN_TRI: lambda: 3 * a,
                 LIN_QUAD: lambda: 4 * a, LIN_OCT: lambda: 8 * a, LIN_NEG: lambda: -a, LIN_COPY: lambda: a}[fl]()
        slots[d] = r % P
        if slots[d] == 0:
            mask |= 1 << i
    return mask


def write_inc(path):
    words, lines = [], []
    for name, prog in PROGRAMS.items():
        lines.append("constexpr int PROG_%s = %d, PROG_%s_LEN = %d;" % (name, len(words), name, len(prog)))
        words += prog
    with open(path, "w") as f:
        f.write("// generated by tools/fpvm_asm.py -- do not edit\n")
        f.write("\n".join(lines) + "\n")
        f.write("constexpr int FPVM_PROG_WORDS = %d;\n" % len(words))
        f.write("#define FPVM_PROG_INIT { " + ", ".join("0x%08xu" % w for w in words) + " }\n")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "rust-eth-kzg_b200", "csrc", "fpvm_programs.inc")
    write_inc(out)
    print("wrote", os.path.normpath(out), sum(len(p) for p in PROGRAMS.values()), "instructions")
