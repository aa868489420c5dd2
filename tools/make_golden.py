#!/usr/bin/env python3
"""Pack the consensus-spec KZG test vectors into small committed fixtures.

Reads   <reference>/test_vectors/<fn>/kzg-mainnet/<case>/data.y*ml   (both .yaml and .yml -- two
recover cases use .yml) and writes tests/golden/<fn>.msgpack.gz.  Byte strings are
de-duplicated through a table (most cases reuse the same 7 blobs), so all 311 cases fit in a few MB.
The vectors are consensus-spec test DATA (ethereum/consensus-spec-tests), not reference source.

Format (msgpack, gzip):  {"table": [bytes...], "cases": [{"name": str, "input": X, "output": X}]}
where X is the YAML value with every 0x-hex string replaced by {"b": index-into-table}.
"""
import glob, gzip, os, sys
import msgpack, yaml

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test_vectors"
DST = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
Loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)

def pack(fn):
    table, index = [], {}
    def enc(v):
        if isinstance(v, str) and v.startswith("0x"):
            b = bytes.fromhex(v[2:])
            if b not in index:
                index[b] = len(table); table.append(b)
            return {"b": index[b]}
        if isinstance(v, list):
            return [enc(x) for x in v]
        if isinstance(v, dict):
            return {k: enc(x) for k, x in v.items()}
        return v
    cases = []
    for f in sorted(glob.glob(os.path.join(SRC, fn, "kzg-mainnet", "*", "data.y*ml"))):
        d = yaml.load(open(f), Loader=Loader)
        cases.append({"name": os.path.basename(os.path.dirname(f)), "input": enc(d["input"]), "output": enc(d["output"])})
    raw = msgpack.packb({"table": table, "cases": cases}, use_bin_type=True)
    out = os.path.join(DST, fn + ".msgpack.gz")
    with gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as g:
        g.write(raw)
    return len(cases), os.path.getsize(out)

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for fn in sorted(os.listdir(SRC)):
        if os.path.isdir(os.path.join(SRC, fn)):
            print(fn, *pack(fn))
