#!/usr/bin/env python3
"""Static issue-cost estimate of a device function from its SASS control codes -- no GPU needed.
Every sm_100a instruction carries, in bits 105..108 of its 128-bit encoding, the number of cycles the warp must wait before issuing
the next one (the compiler's answer to fixed-latency dependences).  Their sum over a straight-line function is the time a LONE warp
needs to issue it, not counting scoreboard waits (shuffles, loads): a cheap way to compare two formulations of a latency-bound
routine before spending GPU time (used on the cooperative multiplier).
  python tools/sass_stalls.py <library.so | file.cubin> <substring of the function label> [...]"""
import os
import re
import subprocess
import sys
import tempfile


def disassemble(path):
    if path.endswith(".cubin"):
        return subprocess.run(["nvdisasm", "-c", "-hex", path], capture_output=True, text=True, check=True).stdout
    out = ""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, capture_output=True, check=True)
        for f in sorted(os.listdir(d)):
            if f.endswith(".cubin"):
                out += subprocess.run(["nvdisasm", "-c", "-hex", os.path.join(d, f)], capture_output=True, text=True).stdout
    return out


def main():
    text = disassemble(sys.argv[1]).split("\n")
    ins_re = re.compile(r"/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/")
    hi_re = re.compile(r"^\s*/\* (0x[0-9a-f]+) \*/")
    for needle in sys.argv[2:]:
        starts = [i for i, l in enumerate(text) if needle in l and l.rstrip().endswith(":") and not l.startswith(".L")]
        for start in starts:
            ops, total, n, i = {}, 0, 0, start + 1
            while i < len(text):
                m = ins_re.search(text[i])
                if m:
                    m2 = hi_re.search(text[i + 1]) if i + 1 < len(text) else None
                    stall = ((int(m2.group(1), 16) if m2 else 0) >> 41) & 0xF
                    words = m.group(2).split()
                    op = words[1] if words[0].startswith("@") else words[0]
                    ops.setdefault(op, [0, 0])
                    ops[op][0] += 1
                    ops[op][1] += stall
                    total += stall
                    n += 1
                    i += 2
                    if op.startswith("RET") or op.startswith("EXIT"):
                        break
                    continue
                i += 1
            print("%s\n  %d instructions, static stall sum %d cycles" % (text[start].strip()[:160], n, total))
            for k, v in sorted(ops.items(), key=lambda kv: -kv[1][1])[:12]:
                print("    %-18s %4d instr %5d cycles" % (k, v[0], v[1]))


if __name__ == "__main__":
    main()
