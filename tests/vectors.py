"""Loader for the committed consensus-vector fixtures (tests/golden/*.msgpack.gz, written by
tools/make_golden.py from <reference>/test_vectors).  Mirrors the reference harness
crates/eip7594/tests/common.rs:12-51: every case has an `input` dict and an `output` (None => the
API must return an error)."""
import functools
import gzip
import os

import msgpack

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FUNCTIONS = [
    "blob_to_kzg_commitment", "compute_blob_kzg_proof", "compute_cells_and_kzg_proofs", "compute_kzg_proof",
    "recover_cells_and_kzg_proofs", "verify_blob_kzg_proof", "verify_blob_kzg_proof_batch",
    "verify_cell_kzg_proof_batch", "verify_kzg_proof",
]


@functools.lru_cache(maxsize=None)
def load(fn):
    with gzip.open(os.path.join(GOLDEN, fn + ".msgpack.gz"), "rb") as g:
        d = msgpack.unpackb(g.read(), raw=False, strict_map_key=False)
    table = d["table"]

    def dec(v):
        if isinstance(v, dict):
            if set(v.keys()) == {"b"}:
                return table[v["b"]]
            return {k: dec(x) for k, x in v.items()}
        if isinstance(v, list):
            return [dec(x) for x in v]
        return v

    return [(c["name"], dec(c["input"]), dec(c["output"])) for c in d["cases"]]


def case_ids(fn):
    return [c[0] for c in load(fn)]
