"""The Node-API addon (rust-eth-kzg_b200/shims/node/eth_kzg_node.cpp = the reference's bindings/node/src/lib.rs on top of the C ABI)
driven through a mock Node-API (tests/napi/napi_mock.cpp): no node exists in this image.

CPU part: module registration exports what bindings/node/index.d.ts declares (constants, CellsAndProofs, DasContextJs with create and
the 20 methods).  GPU part: consensus vectors through the synchronous methods, argument errors with the reference's messages
(lib.rs:440-453), number | bigint cell indices, and the async* forms: promises resolved / rejected after the work ran on pool threads,
several in flight at once (they meet in the library's coalescing queues)."""
import ctypes as C
import os
import subprocess

import pytest

from tests import vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METHODS = ["blobToKzgCommitment", "computeCellsAndKzgProofs", "computeCells", "recoverCellsAndKzgProofs", "verifyCellKzgProofBatch", "computeKzgProof",
           "computeBlobKzgProof", "verifyKzgProof", "verifyBlobKzgProof", "verifyBlobKzgProofBatch"]
T_UNDEF, T_NULL, T_BOOL, T_NUMBER, T_BIGINT, T_STRING, T_OBJECT, T_ARRAY, T_U8, T_AB, T_CLASS, T_ERROR, T_PROMISE = range(13)


class Node:
    def __init__(self, pkg):
        libdir = os.path.join(ROOT, "rust-eth-kzg_b200", "lib")
        if not os.path.exists(os.path.join(libdir, "node_eth_kzg.node")):
            pkg.build_library()
        mock_so = os.path.join(ROOT, "tests", "napi", "libnapi_mock.so")
        src = os.path.join(ROOT, "tests", "napi", "napi_mock.cpp")
        if not os.path.exists(mock_so) or os.path.getmtime(mock_so) < os.path.getmtime(src):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", mock_so, src, "-lpthread"])
        self.m = m = C.CDLL(mock_so, mode=C.RTLD_GLOBAL)          # the addon's napi_* references resolve here
        self.addon = C.CDLL(os.path.join(libdir, "node_eth_kzg.node"))
        for f in ("mock_napi_env_new", "mock_napi_load", "mock_napi_u8", "mock_napi_number", "mock_napi_bigint", "mock_napi_bool", "mock_napi_string",
                  "mock_napi_array", "mock_napi_object", "mock_napi_get_prop", "mock_napi_elem", "mock_napi_new", "mock_napi_call", "mock_napi_promise_result"):
            getattr(m, f).restype = C.c_void_p
        for f in ("mock_napi_text", "mock_napi_class_name", "mock_napi_method_name"):
            getattr(m, f).restype = C.c_char_p
        m.mock_napi_number_value.restype = C.c_double
        m.mock_napi_len.restype = C.c_long
        m.mock_napi_env_free.argtypes = [C.c_void_p]
        m.mock_napi_load.argtypes = [C.c_void_p, C.c_void_p]
        m.mock_napi_u8.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        m.mock_napi_number.argtypes = [C.c_void_p, C.c_double]
        m.mock_napi_bigint.argtypes = [C.c_void_p, C.c_uint64]
        m.mock_napi_bool.argtypes = [C.c_void_p, C.c_int]
        m.mock_napi_string.argtypes = [C.c_void_p, C.c_char_p]
        m.mock_napi_array.argtypes = [C.c_void_p, C.c_long]
        m.mock_napi_array_set.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        m.mock_napi_object.argtypes = [C.c_void_p]
        m.mock_napi_set_prop.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        m.mock_napi_get_prop.argtypes = [C.c_void_p, C.c_char_p]
        for f in ("mock_napi_type", "mock_napi_len", "mock_napi_number_value", "mock_napi_bool_value", "mock_napi_text", "mock_napi_class_name",
                  "mock_napi_method_count", "mock_napi_promise_state", "mock_napi_promise_result", "mock_napi_drain"):
            getattr(m, f).argtypes = [C.c_void_p]
        m.mock_napi_u8_get.argtypes = [C.c_void_p, C.c_char_p]
        m.mock_napi_elem.argtypes = [C.c_void_p, C.c_long]
        m.mock_napi_method_name.argtypes = [C.c_void_p, C.c_int]
        m.mock_napi_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        m.mock_napi_call.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        m.mock_napi_take_exception.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        self.env = m.mock_napi_env_new()
        self.exports = m.mock_napi_load(self.env, C.cast(self.addon.napi_register_module_v1, C.c_void_p))
        assert self.exports

    def close(self):
        self.m.mock_napi_env_free(self.env)

    # ---- JS values ----
    def u8(self, b):
        return self.m.mock_napi_u8(self.env, bytes(b), len(b))

    def arr(self, items):
        a = self.m.mock_napi_array(self.env, len(items))
        for i, v in enumerate(items):
            self.m.mock_napi_array_set(a, i, v)
        return a

    def u8s(self, items):
        return self.arr([self.u8(b) for b in items])

    def nums(self, xs, bigint=False):
        return self.arr([self.m.mock_napi_bigint(self.env, x) if bigint or x >= 1 << 32 else self.m.mock_napi_number(self.env, float(x)) for x in xs])

    def prop(self, obj, name):
        return self.m.mock_napi_get_prop(obj, name.encode())

    def py(self, v):
        """a JS value as Python: Uint8Array -> bytes, Array -> list, boolean -> bool, CellsAndProofs -> (cells, proofs)"""
        t = self.m.mock_napi_type(v)
        if t == T_U8:
            buf = C.create_string_buffer(self.m.mock_napi_len(v))
            self.m.mock_napi_u8_get(v, buf)
            return buf.raw
        if t == T_ARRAY:
            return [self.py(self.m.mock_napi_elem(v, i)) for i in range(self.m.mock_napi_len(v))]
        if t == T_BOOL:
            return bool(self.m.mock_napi_bool_value(v))
        if t == T_OBJECT and self.m.mock_napi_class_name(v) == b"CellsAndProofs":
            return self.py(self.prop(v, "cells")), self.py(self.prop(v, "proofs"))
        if t == T_ERROR:
            return RuntimeError(self.m.mock_napi_text(v).decode())
        raise AssertionError("unexpected JS value type %d" % t)

    def exception(self):
        buf = C.create_string_buffer(4096)
        return buf.value.decode() if self.m.mock_napi_take_exception(self.env, buf, 4096) else None

    def call(self, target, name, *args):
        argv = (C.c_void_p * max(len(args), 1))(*args)
        return self.m.mock_napi_call(self.env, target, name.encode(), len(args), argv)

    def call_py(self, target, name, *args):
        """(value, None) or (None, exception message)"""
        r = self.call(target, name, *args)
        exc = self.exception()
        return (None, exc) if exc is not None else (self.py(r), None)

    def new(self, cls, *args):
        argv = (C.c_void_p * max(len(args), 1))(*args)
        return self.m.mock_napi_new(self.env, cls, len(args), argv)


@pytest.fixture(scope="module")
def node(pkg):
    n = Node(pkg)
    yield n
    n.close()


def test_exports_match_index_d_ts(node):
    m = node.m
    for name, want in (("BYTES_PER_COMMITMENT", 48), ("BYTES_PER_PROOF", 48), ("BYTES_PER_FIELD_ELEMENT", 32), ("BYTES_PER_BLOB", 131072),
                       ("MAX_NUM_COLUMNS", 128), ("BYTES_PER_CELL", 2048)):
        assert m.mock_napi_number_value(node.prop(node.exports, name)) == want
    assert m.mock_napi_type(node.prop(node.exports, "CellsAndProofs")) == T_CLASS
    cls = node.prop(node.exports, "DasContextJs")
    assert m.mock_napi_type(cls) == T_CLASS
    names = sorted(m.mock_napi_method_name(cls, i).decode() for i in range(m.mock_napi_method_count(cls)))
    assert names == sorted(["create"] + METHODS + ["async" + x[0].upper() + x[1:] for x in METHODS])


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def jsctx(node):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cls = node.prop(node.exports, "DasContextJs")
    opts = node.m.mock_napi_object(node.env)
    node.m.mock_napi_set_prop(opts, b"usePrecomp", node.m.mock_napi_bool(node.env, 0))
    ctx = node.call(cls, "create", opts)        # DasContextJs.create({usePrecomp: false})
    assert ctx and node.exception() is None
    assert node.m.mock_napi_class_name(ctx) == b"DasContextJs"
    return ctx


@pytest.mark.gpu
def test_sync_methods_on_vectors(node, jsctx):
    for name, inp, expected in vectors.load("compute_cells_and_kzg_proofs"):
        got, exc = node.call_py(jsctx, "computeCellsAndKzgProofs", node.u8(inp["blob"]))
        if expected is None:
            want = "failed to compute compute_cells_and_kzg_proofs: " if len(inp["blob"]) == 131072 else "blob must have size 131072, found size %d" % len(inp["blob"])
            assert got is None and exc.startswith(want), name
        else:
            assert exc is None and [list(got[0]), list(got[1])] == [list(expected[0]), list(expected[1])], name
            cells, exc = node.call_py(jsctx, "computeCells", node.u8(inp["blob"]))
            assert cells == list(expected[0])
    for name, inp, expected in vectors.load("blob_to_kzg_commitment"):
        got, exc = node.call_py(jsctx, "blobToKzgCommitment", node.u8(inp["blob"]))
        assert got == expected and (exc is None) == (expected is not None), name
    for name, inp, expected in vectors.load("compute_kzg_proof")[:16]:
        got, exc = node.call_py(jsctx, "computeKzgProof", node.u8(inp["blob"]), node.u8(inp["z"]))
        assert (None if got is None else tuple(got)) == (None if expected is None else tuple(expected)), name
    for name, inp, expected in vectors.load("compute_blob_kzg_proof"):
        if len(inp["commitment"]) != 48:
            continue
        got, exc = node.call_py(jsctx, "computeBlobKzgProof", node.u8(inp["blob"]), node.u8(inp["commitment"]))
        assert got == expected, name
    for name, inp, expected in vectors.load("recover_cells_and_kzg_proofs"):
        if any(len(c) != 2048 for c in inp["cells"]):
            continue
        got, exc = node.call_py(jsctx, "recoverCellsAndKzgProofs", node.nums(inp["cell_indices"]), node.u8s(inp["cells"]))
        assert (None if got is None else [list(got[0]), list(got[1])]) == (None if expected is None else [list(expected[0]), list(expected[1])]), name


@pytest.mark.gpu
def test_verifiers_on_vectors(node, jsctx):
    for i, (name, inp, expected) in enumerate(vectors.load("verify_cell_kzg_proof_batch")):
        if any(len(c) != 2048 for c in inp["cells"]) or any(len(c) != 48 for c in inp["commitments"] + inp["proofs"]):
            continue
        got, exc = node.call_py(jsctx, "verifyCellKzgProofBatch", node.u8s(inp["commitments"]), node.nums(inp["cell_indices"], bigint=bool(i & 1)),
                                node.u8s(inp["cells"]), node.u8s(inp["proofs"]))
        assert got == expected and (exc is None) == (expected is not None), name
    for name, inp, expected in vectors.load("verify_kzg_proof")[:40]:
        if [len(inp[k]) for k in ("commitment", "z", "y", "proof")] != [48, 32, 32, 48]:
            continue
        got, exc = node.call_py(jsctx, "verifyKzgProof", *[node.u8(inp[k]) for k in ("commitment", "z", "y", "proof")])
        assert got == expected, name
    for name, inp, expected in vectors.load("verify_blob_kzg_proof")[:12]:
        if [len(inp[k]) for k in ("blob", "commitment", "proof")] != [131072, 48, 48]:
            continue
        got, exc = node.call_py(jsctx, "verifyBlobKzgProof", *[node.u8(inp[k]) for k in ("blob", "commitment", "proof")])
        assert got == expected, name
    for name, inp, expected in vectors.load("verify_blob_kzg_proof_batch")[:10]:
        if any(len(b) != 131072 for b in inp["blobs"]) or any(len(c) != 48 for c in inp["commitments"] + inp["proofs"]):
            continue
        got, exc = node.call_py(jsctx, "verifyBlobKzgProofBatch", node.u8s(inp["blobs"]), node.u8s(inp["commitments"]), node.u8s(inp["proofs"]))
        assert got == expected, name


@pytest.mark.gpu
def test_argument_errors(node, jsctx):
    got, exc = node.call_py(jsctx, "blobToKzgCommitment", node.u8(b"\0" * 100))
    assert exc == "blob must have size 131072, found size 100\n err:could not convert slice to array"
    got, exc = node.call_py(jsctx, "computeKzgProof", node.u8(b"\0" * 131072), node.u8(b"\0" * 33))
    assert exc == "z must have size 32, found size 33\n err:could not convert slice to array"
    got, exc = node.call_py(jsctx, "verifyCellKzgProofBatch", node.u8s([b"\0" * 48]), node.nums([0]), node.u8s([b"\0" * 2047]), node.u8s([b"\0" * 48]))
    assert exc == "cell must have size 2048, found size 2047\n err:could not convert slice to array"
    got, exc = node.call_py(jsctx, "blobToKzgCommitment", node.m.mock_napi_number(node.env, 3.0))
    assert exc == "blob must be a Uint8Array"
    got, exc = node.call_py(jsctx, "asyncComputeCells", node.u8(b"\0" * 5))       # argument errors of the async forms are synchronous
    assert exc is not None and exc.startswith("blob must have size 131072, found size 5")
    got, exc = node.call_py(jsctx, "recoverCellsAndKzgProofs", node.arr([node.m.mock_napi_string(node.env, b"x")]), node.u8s([]))
    assert exc == "cell indices must be an array of number | bigint"
    cls = node.prop(node.exports, "DasContextJs")
    assert node.call(cls, "create", node.m.mock_napi_number(node.env, 1.0)) is None
    assert node.exception() == "options must be an object {usePrecomp: boolean}"


@pytest.mark.gpu
def test_async_methods_share_batches(node, jsctx):
    valid = [(n, i, o) for n, i, o in vectors.load("compute_cells_and_kzg_proofs") if o is not None]
    invalid = [(n, i, o) for n, i, o in vectors.load("compute_cells_and_kzg_proofs") if o is None and len(i["blob"]) == 131072]
    promises = []
    for k in range(12):
        name, inp, expected = valid[k % len(valid)]
        p = node.call(jsctx, "asyncComputeCellsAndKzgProofs", node.u8(inp["blob"]))
        assert p and node.exception() is None and node.m.mock_napi_type(p) == T_PROMISE and node.m.mock_napi_promise_state(p) == 0
        promises.append((p, expected))
    bad = node.call(jsctx, "asyncComputeCellsAndKzgProofs", node.u8(invalid[0][1]["blob"]))
    name, inp, expected = vectors.load("verify_kzg_proof")[0]
    vp = node.call(jsctx, "asyncVerifyKzgProof", *[node.u8(inp[k]) for k in ("commitment", "z", "y", "proof")])
    assert node.m.mock_napi_drain(node.env) == 14            # the event loop turns: pool threads joined, completions on this thread
    for p, want in promises:
        assert node.m.mock_napi_promise_state(p) == 1
        cells, proofs = node.py(node.m.mock_napi_promise_result(p))
        assert [cells, proofs] == [list(want[0]), list(want[1])]
    assert node.m.mock_napi_promise_state(bad) == 2          # an invalid blob rejects its own promise only
    err = node.py(node.m.mock_napi_promise_result(bad))
    assert isinstance(err, RuntimeError) and str(err).startswith("failed to compute compute_cells_and_kzg_proofs: ")
    assert node.m.mock_napi_promise_state(vp) == 1 and node.py(node.m.mock_napi_promise_result(vp)) == expected


@pytest.mark.gpu
def test_default_constructor_uses_precomp(node, jsctx, monkeypatch):
    """new DasContextJs() = DASContextOptions::default() = {usePrecomp: true} (lib.rs:49-60); window widths pinned small here so that the
    test does not compete with the session's production-layout context for HBM"""
    monkeypatch.setenv("EKZG_FK20_WINDOW", "9")
    monkeypatch.setenv("EKZG_SRS_WINDOW", "9")
    cls = node.prop(node.exports, "DasContextJs")
    ctx = node.new(cls)
    assert ctx and node.exception() is None
    name, inp, expected = [c for c in vectors.load("blob_to_kzg_commitment") if c[2] is not None][0]
    got, exc = node.call_py(ctx, "blobToKzgCommitment", node.u8(inp["blob"]))
    assert got == expected and exc is None
