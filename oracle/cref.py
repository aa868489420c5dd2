"""ctypes binding for oracle/libkzg_oracle.so (the C restatement, oracle/src/kzg.c).
TEST INFRASTRUCTURE ONLY.  Function names mirror the reference API
(crates/eip7594/src/{prover,verifier}.rs, crates/eip4844/src/{prover,verifier}.rs); every call
returns the outputs or raises OracleError where the reference returns Err."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkzg_oracle.so")
_SETUP = os.path.join(os.path.dirname(_HERE), "rust-eth-kzg_b200", "data", "trusted_setup_4096.bin")

BYTES_PER_BLOB, BYTES_PER_CELL, CELLS = 131072, 2048, 128


class OracleError(Exception):
    pass


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        rc = _lib.okzg_init(_SETUP.encode())
        if rc != 0:
            raise RuntimeError("oracle init failed rc=%d" % rc)
    return _lib


def _len(x, n, what):
    """type-level length checks of the reference API (BlobRef = &[u8; 131072] etc.)"""
    if len(x) != n:
        raise OracleError("%s must be %d bytes, got %d" % (what, n, len(x)))


def _chk(rc):
    if rc != 0:
        raise OracleError("oracle rc=%d" % rc)


def num_threads():
    return lib().okzg_num_threads()


def compute_cells_and_kzg_proofs(blob):
    _len(blob, BYTES_PER_BLOB, "blob")
    cells = C.create_string_buffer(CELLS * BYTES_PER_CELL)
    proofs = C.create_string_buffer(CELLS * 48)
    _chk(lib().okzg_compute_cells_and_kzg_proofs(bytes(blob), cells, proofs))
    return ([cells.raw[i * 2048:(i + 1) * 2048] for i in range(CELLS)], [proofs.raw[i * 48:(i + 1) * 48] for i in range(CELLS)])


def time_fp_mul(n=2000000):
    """nanoseconds per Fp Montgomery multiplication of this port on one core (dependent chain of n)"""
    f = lib().okzg_time_fp_mul
    f.restype = C.c_double
    sink = C.create_string_buffer(48)
    f(20000, sink)
    return float(f(int(n), sink))


def compute_cells(blob):
    _len(blob, BYTES_PER_BLOB, "blob")
    cells = C.create_string_buffer(CELLS * BYTES_PER_CELL)
    _chk(lib().okzg_compute_cells(bytes(blob), cells))
    return [cells.raw[i * 2048:(i + 1) * 2048] for i in range(CELLS)]


def compute_cells_and_kzg_proofs_batch(blobs_flat, n, nthreads=0):
    """blobs_flat: bytes of n*131072; returns (cells_flat, proofs_flat) bytes"""
    cells = C.create_string_buffer(n * CELLS * BYTES_PER_CELL)
    proofs = C.create_string_buffer(n * CELLS * 48)
    _chk(lib().okzg_compute_cells_and_kzg_proofs_batch(n, bytes(blobs_flat), cells, proofs, nthreads))
    return cells.raw, proofs.raw


def blob_to_kzg_commitment(blob):
    _len(blob, BYTES_PER_BLOB, "blob")
    out = C.create_string_buffer(48)
    _chk(lib().okzg_blob_to_kzg_commitment(bytes(blob), out))
    return out.raw


def compute_kzg_proof(blob, z):
    _len(blob, BYTES_PER_BLOB, "blob"); _len(z, 32, "z")
    proof, y = C.create_string_buffer(48), C.create_string_buffer(32)
    _chk(lib().okzg_compute_kzg_proof(bytes(blob), bytes(z), proof, y))
    return proof.raw, y.raw


def compute_blob_kzg_proof(blob, commitment):
    _len(blob, BYTES_PER_BLOB, "blob"); _len(commitment, 48, "commitment")
    proof = C.create_string_buffer(48)
    _chk(lib().okzg_compute_blob_kzg_proof(bytes(blob), bytes(commitment), proof))
    return proof.raw


def verify_kzg_proof(commitment, z, y, proof):
    _len(commitment, 48, "commitment"); _len(z, 32, "z"); _len(y, 32, "y"); _len(proof, 48, "proof")
    ok = C.c_int(0)
    _chk(lib().okzg_verify_kzg_proof(bytes(commitment), bytes(z), bytes(y), bytes(proof), C.byref(ok)))
    return bool(ok.value)


def verify_blob_kzg_proof(blob, commitment, proof):
    _len(blob, BYTES_PER_BLOB, "blob"); _len(commitment, 48, "commitment"); _len(proof, 48, "proof")
    ok = C.c_int(0)
    _chk(lib().okzg_verify_blob_kzg_proof(bytes(blob), bytes(commitment), bytes(proof), C.byref(ok)))
    return bool(ok.value)


def verify_blob_kzg_proof_batch(blobs, commitments, proofs):
    if not (len(blobs) == len(commitments) == len(proofs)):
        raise OracleError("length mismatch")
    for b in blobs: _len(b, BYTES_PER_BLOB, "blob")
    for c in commitments: _len(c, 48, "commitment")
    for p in proofs: _len(p, 48, "proof")
    ok = C.c_int(0)
    _chk(lib().okzg_verify_blob_kzg_proof_batch(len(blobs), b"".join(blobs), b"".join(commitments), b"".join(proofs), C.byref(ok)))
    return bool(ok.value)


def recover_cells_and_kzg_proofs(cell_indices, cells):
    for c in cells: _len(c, BYTES_PER_CELL, "cell")
    idx = (C.c_uint64 * len(cell_indices))(*cell_indices)
    oc = C.create_string_buffer(CELLS * BYTES_PER_CELL)
    op = C.create_string_buffer(CELLS * 48)
    _chk(lib().okzg_recover_cells_and_kzg_proofs(len(cell_indices), idx, len(cells), b"".join(cells), oc, op))
    return ([oc.raw[i * 2048:(i + 1) * 2048] for i in range(CELLS)], [op.raw[i * 48:(i + 1) * 48] for i in range(CELLS)])


def verify_cell_kzg_proof_batch(commitments, cell_indices, cells, proofs):
    for c in cells: _len(c, BYTES_PER_CELL, "cell")
    for c in commitments: _len(c, 48, "commitment")
    for p in proofs: _len(p, 48, "proof")
    idx = (C.c_uint64 * len(cell_indices))(*cell_indices)
    ok = C.c_int(0)
    _chk(lib().okzg_verify_cell_kzg_proof_batch(len(commitments), b"".join(commitments), len(cell_indices), idx,
                                                 len(cells), b"".join(cells), len(proofs), b"".join(proofs), C.byref(ok)))
    return bool(ok.value)


def fk20_stages(blob):
    """(scalars[128][64] ints, msm[128] compressed R_j, h[64] compressed) of one blob -- stage parity hook"""
    sc = C.create_string_buffer(128 * 64 * 32)
    msm = C.create_string_buffer(128 * 48)
    h = C.create_string_buffer(64 * 48)
    _chk(lib().okzg_test_fk20_stages(bytes(blob), sc, msm, h))
    scalars = [[int.from_bytes(sc.raw[32 * (j * 64 + k):32 * (j * 64 + k) + 32], "big") for k in range(64)] for j in range(128)]
    return scalars, [msm.raw[48 * j:48 * j + 48] for j in range(128)], [h.raw[48 * i:48 * i + 48] for i in range(64)]


def g1_mul(p48, k):
    out = C.create_string_buffer(48)
    rc = lib().okzg_test_g1_mul(bytes(p48), int(k).to_bytes(32, "big"), out)
    if rc != 0:
        raise OracleError("g1_mul rc=%d" % rc)
    return out.raw
