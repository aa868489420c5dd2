#!/usr/bin/env python3
"""Convert the public Ethereum KZG ceremony output (JSON, compressed hex points) to the packed
binary the library embeds/loads.

Source:  <reference>/crates/trusted_setup/data/trusted_setup_4096.json  (keys g1_monomial[4096],
g1_lagrange[4096], g2_monomial[65]; crates/trusted_setup/src/lib.rs:9-14,100-124).
This is ceremony DATA, not reference source code.

Layout (little-endian header, then raw compressed points exactly as in the JSON):
  8 B  magic "EKZGTS01"
  4 B  n_g1 (4096)      4 B  n_g2 (65)
  48*n_g1  g1_monomial    48*n_g1  g1_lagrange    96*n_g2  g2_monomial
"""
import json, struct, sys

def main(src, dst):
    d = json.load(open(src))
    g1m = [bytes.fromhex(x[2:]) for x in d["g1_monomial"]]
    g1l = [bytes.fromhex(x[2:]) for x in d["g1_lagrange"]]
    g2m = [bytes.fromhex(x[2:]) for x in d["g2_monomial"]]
    assert len(g1m) == len(g1l) == 4096 and len(g2m) == 65
    assert all(len(x) == 48 for x in g1m + g1l) and all(len(x) == 96 for x in g2m)
    with open(dst, "wb") as f:
        f.write(b"EKZGTS01" + struct.pack("<II", len(g1m), len(g2m)))
        for x in g1m + g1l + g2m:
            f.write(x)

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/crates/trusted_setup/data/trusted_setup_4096.json",
         sys.argv[2] if len(sys.argv) > 2 else "rust-eth-kzg_b200/data/trusted_setup_4096.bin")
