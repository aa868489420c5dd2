"""The JNI shim (rust-eth-kzg_b200/shims/jni/eth_kzg_jni.cpp = the reference's bindings/java/rust_code/src/lib.rs on top of the C ABI)
driven through a mock JNIEnv (tests/jni/jni_mock.cpp): no JVM exists in this image.

CPU part: the 12 Java_ethereum_cryptography_LibEthKZG_* symbols of LibEthKZG.java:216-239 are exported, the function table has the
specification's 234 slots, and the argument checks that never reach the device throw IllegalArgumentException with the reference's
message (lib.rs:509-541).  GPU part: the consensus vectors through the shim, result objects as the Java classes expect them."""
import ctypes as C
import os
import subprocess

import pytest

from tests import vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVES = ["DASContextNew", "DASContextDestroy", "computeCellsAndKZGProofs", "computeCells", "blobToKZGCommitment", "verifyCellKZGProofBatch",
           "recoverCellsAndKZGProofs", "computeKzgProof", "computeBlobKzgProof", "verifyKzgProof", "verifyBlobKzgProof", "verifyBlobKzgProofBatch"]


class Jni:
    def __init__(self, pkg):
        libdir = os.path.join(ROOT, "rust-eth-kzg_b200", "lib")
        if not os.path.exists(os.path.join(libdir, "libjava_eth_kzg.so")):
            pkg.build_library()
        mock_so = os.path.join(ROOT, "tests", "jni", "libjni_mock.so")
        src = os.path.join(ROOT, "tests", "jni", "jni_mock.cpp")
        if not os.path.exists(mock_so) or os.path.getmtime(mock_so) < os.path.getmtime(src):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", mock_so, src])
        self.shim = C.CDLL(os.path.join(libdir, "libjava_eth_kzg.so"))
        self.mock = m = C.CDLL(mock_so)
        for f in ("mock_env_new", "mock_new_bytes", "mock_new_longs", "mock_new_array", "mock_elem"):
            getattr(m, f).restype = C.c_void_p
        m.mock_env_free.argtypes = [C.c_void_p]
        m.mock_new_bytes.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        m.mock_new_longs.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long]
        m.mock_new_array.argtypes = [C.c_void_p, C.c_long]
        m.mock_array_set.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        m.mock_len.argtypes = [C.c_void_p]
        m.mock_len.restype = C.c_long
        m.mock_bytes_get.argtypes = [C.c_void_p, C.c_char_p]
        m.mock_elem.argtypes = [C.c_void_p, C.c_long]
        m.mock_name.argtypes = [C.c_void_p]
        m.mock_name.restype = C.c_char_p
        m.mock_kind.argtypes = [C.c_void_p]
        m.mock_take_exception.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        m.mock_max_local_refs.argtypes = [C.c_void_p]
        m.mock_max_local_refs.restype = C.c_long
        self.env = m.mock_env_new()
        for n in NATIVES:
            fn = getattr(self.shim, "Java_ethereum_cryptography_LibEthKZG_" + n)
            fn.restype = C.c_int64 if n == "DASContextNew" else None if n == "DASContextDestroy" else C.c_uint8 if n.startswith("verify") else C.c_void_p

    # ---- Java values ----
    def barr(self, b):
        return self.mock.mock_new_bytes(self.env, bytes(b), len(b))

    def larr(self, xs):
        a = (C.c_int64 * max(len(xs), 1))(*[x if x < 1 << 63 else x - (1 << 64) for x in xs])
        return self.mock.mock_new_longs(self.env, a, len(xs))

    def barr2(self, items):
        arr = self.mock.mock_new_array(self.env, len(items))
        for i, b in enumerate(items):
            self.mock.mock_array_set(arr, i, self.barr(b))
        return arr

    def bytes_of(self, o):
        buf = C.create_string_buffer(self.mock.mock_len(o))
        self.mock.mock_bytes_get(o, buf)
        return buf.raw

    def list_of(self, o):
        return [self.bytes_of(self.mock.mock_elem(o, i)) for i in range(self.mock.mock_len(o))]

    def exception(self):
        buf = C.create_string_buffer(4096)
        return buf.value.decode() if self.mock.mock_take_exception(self.env, buf, 4096) else None

    def call(self, name, *args):
        fn = getattr(self.shim, "Java_ethereum_cryptography_LibEthKZG_" + name)
        cargs = [C.c_void_p(self.env), C.c_void_p(None)]
        for a in args:
            cargs.append(a if isinstance(a, (C.c_int64, C.c_uint8)) else C.c_void_p(a))
        return fn(*cargs)


@pytest.fixture(scope="module")
def jni(pkg):
    j = Jni(pkg)
    yield j
    j.mock.mock_env_free(j.env)


def test_symbols_and_table_layout(jni):
    for n in NATIVES:
        assert hasattr(jni.shim, "Java_ethereum_cryptography_LibEthKZG_" + n)
    assert jni.mock.mock_table_slots() == 234      # JNI specification, Java SE 9+: GetModule is slot 233


@pytest.mark.parametrize("native,args,needle", [
    ("computeCellsAndKZGProofs", lambda j: [j.barr(b"\0" * 100)], "function computeCellsAndKZGProofs has thrown an exception, with reason: blob is not the correct size. expected: 131072\ngot: 100"),
    ("computeCells", lambda j: [j.barr(b"")], "function computeCells has thrown an exception, with reason: blob is not the correct size. expected: 131072\ngot: 0"),
    ("blobToKZGCommitment", lambda j: [j.barr(b"\0" * 131073)], "blob is not the correct size. expected: 131072\ngot: 131073"),
    ("computeKzgProof", lambda j: [j.barr(b"\0" * 131072), j.barr(b"\0" * 31)], "function computeKzgProof has thrown an exception, with reason: z is not the correct size. expected: 32\ngot: 31"),
    ("computeBlobKzgProof", lambda j: [j.barr(b"\0" * 131072), j.barr(b"\0" * 49)], "commitment is not the correct size. expected: 48\ngot: 49"),
    ("verifyKzgProof", lambda j: [j.barr(b"\0" * 48), j.barr(b"\0" * 32), j.barr(b"\0" * 33), j.barr(b"\0" * 48)], "y is not the correct size. expected: 32\ngot: 33"),
    ("verifyBlobKzgProof", lambda j: [j.barr(b"\0" * 131072), j.barr(b"\0" * 48), j.barr(b"\0" * 47)], "proof is not the correct size. expected: 48\ngot: 47"),
    ("verifyBlobKzgProofBatch", lambda j: [j.barr2([b"\0" * 131072, b"\0" * 5]), j.barr2([]), j.barr2([])], "blob is not the correct size. expected: 131072\ngot: 5"),
    ("verifyCellKZGProofBatch", lambda j: [j.barr2([b"\0" * 48]), j.larr([0]), j.barr2([b"\0" * 2047]), j.barr2([b"\0" * 48])], "cell is not the correct size. expected: 2048\ngot: 2047"),
    ("recoverCellsAndKZGProofs", lambda j: [j.larr([0, 1]), j.barr2([b"\0" * 2048, b"\0" * 2049])], "cell is not the correct size. expected: 2048\ngot: 2049"),
])
def test_size_errors_throw_like_the_reference(jni, native, args, needle):
    ret = jni.call(native, C.c_int64(0), *args(jni))     # the context is never touched: the size check comes first
    assert not ret
    exc = jni.exception()
    assert exc is not None and exc.startswith("java/lang/IllegalArgumentException: function " + native + " has thrown an exception, with reason: ")
    assert needle in exc


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def jctx(jni):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    h = jni.call("DASContextNew", C.c_uint8(0))
    assert h
    yield C.c_int64(h)
    jni.call("DASContextDestroy", C.c_int64(h))


@pytest.mark.gpu
def test_compute_cells_and_proofs_vectors(jni, jctx):
    for name, inp, expected in vectors.load("compute_cells_and_kzg_proofs"):
        obj = jni.call("computeCellsAndKZGProofs", jctx, jni.barr(inp["blob"]))
        exc = jni.exception()
        if expected is None:
            assert not obj and exc and exc.startswith("java/lang/IllegalArgumentException: function computeCellsAndKZGProofs has thrown"), name
            continue
        assert exc is None and jni.mock.mock_name(obj) == b"ethereum/cryptography/CellsAndProofs([[B[[B)V"
        assert jni.list_of(jni.mock.mock_elem(obj, 0)) == list(expected[0]), name
        assert jni.list_of(jni.mock.mock_elem(obj, 1)) == list(expected[1]), name
        cells = jni.call("computeCells", jctx, jni.barr(inp["blob"]))
        assert jni.mock.mock_name(cells) == b"ethereum/cryptography/Cells([[B)V"
        assert jni.list_of(jni.mock.mock_elem(cells, 0)) == list(expected[0])


@pytest.mark.gpu
def test_commitment_and_proof_vectors(jni, jctx):
    for name, inp, expected in vectors.load("blob_to_kzg_commitment"):
        out = jni.call("blobToKZGCommitment", jctx, jni.barr(inp["blob"]))
        exc = jni.exception()
        assert (None if not out else jni.bytes_of(out)) == expected and (exc is None) == (expected is not None), name
    for name, inp, expected in vectors.load("compute_kzg_proof")[:20]:
        out = jni.call("computeKzgProof", jctx, jni.barr(inp["blob"]), jni.barr(inp["z"]))
        exc = jni.exception()
        got = None if not out else tuple(jni.list_of(out))
        assert got == (None if expected is None else tuple(expected)) and (exc is None) == (expected is not None), name
    for name, inp, expected in vectors.load("compute_blob_kzg_proof"):
        out = jni.call("computeBlobKzgProof", jctx, jni.barr(inp["blob"]), jni.barr(inp["commitment"]))
        exc = jni.exception()
        assert (None if not out else jni.bytes_of(out)) == expected and (exc is None) == (expected is not None), name


@pytest.mark.gpu
def test_verifier_vectors(jni, jctx):
    def run(native, args, expected, name):
        ret = jni.call(native, jctx, *args)
        exc = jni.exception()
        if expected is None:
            assert exc is not None and not ret, (native, name)
        else:
            assert exc is None and bool(ret) == expected, (native, name)

    for name, inp, expected in vectors.load("verify_cell_kzg_proof_batch"):
        if any(len(c) != 2048 for c in inp["cells"]) or any(len(c) != 48 for c in inp["commitments"] + inp["proofs"]):
            continue      # wrong-length items are refused by the size check (covered on the CPU above)
        run("verifyCellKZGProofBatch", [jni.barr2(inp["commitments"]), jni.larr(inp["cell_indices"]), jni.barr2(inp["cells"]), jni.barr2(inp["proofs"])], expected, name)
    for name, inp, expected in vectors.load("verify_kzg_proof")[:40]:
        if [len(inp[k]) for k in ("commitment", "z", "y", "proof")] != [48, 32, 32, 48]:
            continue
        run("verifyKzgProof", [jni.barr(inp[k]) for k in ("commitment", "z", "y", "proof")], expected, name)
    for name, inp, expected in vectors.load("verify_blob_kzg_proof")[:12]:
        if [len(inp[k]) for k in ("blob", "commitment", "proof")] != [131072, 48, 48]:
            continue
        run("verifyBlobKzgProof", [jni.barr(inp[k]) for k in ("blob", "commitment", "proof")], expected, name)
    for name, inp, expected in vectors.load("verify_blob_kzg_proof_batch")[:10]:
        if any(len(b) != 131072 for b in inp["blobs"]) or any(len(c) != 48 for c in inp["commitments"] + inp["proofs"]):
            continue
        run("verifyBlobKzgProofBatch", [jni.barr2(inp["blobs"]), jni.barr2(inp["commitments"]), jni.barr2(inp["proofs"])], expected, name)


@pytest.mark.gpu
def test_recover_vectors(jni, jctx):
    for name, inp, expected in vectors.load("recover_cells_and_kzg_proofs"):
        if any(len(c) != 2048 for c in inp["cells"]):
            continue
        obj = jni.call("recoverCellsAndKZGProofs", jctx, jni.larr(inp["cell_indices"]), jni.barr2(inp["cells"]))
        exc = jni.exception()
        if expected is None:
            assert not obj and exc is not None, name
        else:
            assert exc is None
            assert [jni.list_of(jni.mock.mock_elem(obj, 0)), jni.list_of(jni.mock.mock_elem(obj, 1))] == [list(expected[0]), list(expected[1])], name
