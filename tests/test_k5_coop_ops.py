"""The cooperative (4 lanes per field element) fixed-scalar ladder of K5's latency mode (csrc/g1_coop.cuh) is straight-line code: it has
no branch for the exceptional cases of a point addition (accumulator = +-table entry, sum = identity).  That is sound because the
scalars are FIXED -- the 128 op lists of csrc/twiddle_ops.inc -- and the input point has prime order r: an exceptional case would be
a relation between integers mod r, independent of the point.  This test replays every op list on integers and proves none occurs
(rows 0 and 64 never reach the ladder: omega^0 = 1 is skipped and omega^64 = -1 is a negation)."""
import os
import re

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xd201000000010000
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rows():
    text = open(os.path.join(ROOT, "rust-eth-kzg_b200", "csrc", "twiddle_ops.inc")).read()
    rows = [[int(v) for v in m.split(",")] for m in re.findall(r"\{([0-9,]+)\}", text)]
    assert len(rows) == 128 and all(len(r) == 64 for r in rows)
    return rows


def test_no_exceptional_addition_in_any_twiddle_ladder():
    lam = (BLS_X * BLS_X - 1) % R
    w128 = pow(7, (R - 1) // 128, R)
    assert pow(w128, 64, R) == R - 1
    rows = _rows()
    # the table generator may use any primitive 128th root: identify it from row 1
    def replay(row):
        acc = 0
        for op in row[1:1 + row[0]]:
            acc <<= op >> 8
            if op & 0x20:
                t = (2 * (op & 7) + 1) * (lam if op & 0x10 else 1)
                acc += -t if op & 8 else t
        return acc % R
    w = replay(rows[1])
    assert pow(w, 128, R) == 1 and pow(w, 64, R) == R - 1
    for e, row in enumerate(rows):
        assert replay(row) == pow(w, e, R), e
        if e in (0, 64):
            continue
        n, ops = row[0], row[1:1 + row[0]]
        assert n >= 2 and ops[0] >> 8 == 0 and ops[0] & 0x20, "the ladder starts by loading a table entry into the empty accumulator"
        acc = 0
        for i, op in enumerate(ops):
            acc = (acc << (op >> 8)) % R
            if i > 0:
                assert acc != 0, (e, i, "accumulator is the identity before a doubling/addition")
            if op & 0x20:
                t = (2 * (op & 7) + 1) * (lam if op & 0x10 else 1) % R
                if op & 8:
                    t = R - t
                if i > 0:
                    assert acc != t, (e, i, "addition of equal points")
                    assert (acc + t) % R != 0, (e, i, "addition of opposite points")
                acc = (acc + t) % R
        assert acc == pow(w, e, R) and acc != 0
    # the odd multiples 3P .. 15P are built by adding 2P to (2i-1)P: distinct, non-opposite multiples below r
    for i in range(1, 8):
        assert (2 * i - 1) % R not in (2, R - 2)
