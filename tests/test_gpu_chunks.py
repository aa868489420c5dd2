"""The host batch scheduler across chunk boundaries (csrc/kzg_runtime.cu): a batch cut into several chunks that
alternate between two workspaces, with pageable caller memory (ctypes buffers) and with pinned caller memory (torch
pinned tensors), must give the bytes of the one-chunk run.  EKZG_CHUNK is read on every call."""
import ctypes as C
import os

import pytest

pytestmark = pytest.mark.gpu


def _synth(pkg):
    import importlib
    return importlib.import_module("eth_kzg_b200.synthetic")


@pytest.fixture
def small_chunks():
    old = os.environ.get("EKZG_CHUNK")
    os.environ["EKZG_CHUNK"] = "24"
    yield 24
    if old is None:
        del os.environ["EKZG_CHUNK"]
    else:
        os.environ["EKZG_CHUNK"] = old


def test_empty_batches(das_ctx):
    assert das_ctx.compute_cells_and_kzg_proofs_batch(b"", 0) == (b"", b"", [])
    assert das_ctx.blob_to_kzg_commitment_batch(b"", 0) == (b"", [])
    assert das_ctx.verify_cell_kzg_proof_batch([], [], [], []) is True


def test_chunked_equals_unchunked(das_ctx, pkg, small_chunks):
    syn = _synth(pkg)
    n = 3 * small_chunks + 5                       # 4 chunks, the last one ragged
    flat = b"".join(syn.blob(500 + i) for i in range(n))
    got = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    comm, st = das_ctx.blob_to_kzg_commitment_batch(flat, n)
    proofs, st2 = das_ctx.compute_blob_kzg_proof_batch(flat, comm, n)
    os.environ["EKZG_CHUNK"] = "1024"
    want = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    assert got == want
    comm1, _ = das_ctx.blob_to_kzg_commitment_batch(flat, n)
    proofs1, _ = das_ctx.compute_blob_kzg_proof_batch(flat, comm1, n)
    assert (comm, proofs) == (comm1, proofs1) and not any(st) and not any(st2)
    # spot-check against the single-blob ABI
    for i in (0, small_chunks - 1, small_chunks, n - 1):
        b = flat[i * 131072:(i + 1) * 131072]
        cells, prf = das_ctx.compute_cells_and_kzg_proofs(b)
        assert b"".join(cells) == got[0][i * 262144:(i + 1) * 262144] and b"".join(prf) == got[1][i * 6144:(i + 1) * 6144]
        assert das_ctx.blob_to_kzg_commitment(b) == comm[48 * i:48 * i + 48]


def test_pinned_and_pageable_callers_agree(das_ctx, pkg, small_chunks):
    import torch
    syn = _synth(pkg)
    n = 2 * small_chunks + 3
    flat = b"".join(syn.blob(700 + i) for i in range(n))
    want = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)     # ctypes string buffers: pageable
    lib = pkg.load_library()
    h_in = torch.frombuffer(bytearray(flat), dtype=torch.uint8).pin_memory()
    h_cells = torch.empty(n * 262144, dtype=torch.uint8).pin_memory()
    h_proofs = torch.empty(n * 6144, dtype=torch.uint8).pin_memory()
    h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
    res = lib.eth_kzg_b200_compute_cells_and_kzg_proofs_batch(C.c_void_p(das_ctx.handle), C.c_uint64(n), C.c_void_p(h_in.data_ptr()),
                                                              C.c_void_p(h_cells.data_ptr()), C.c_void_p(h_proofs.data_ptr()), C.c_void_p(h_st.data_ptr()))
    assert res.status == 0
    assert bytes(h_cells.numpy()) == want[0] and bytes(h_proofs.numpy()) == want[1] and list(h_st.numpy()) == [0] * n


def test_recover_batch_chunked(das_ctx, pkg, small_chunks):
    syn = _synth(pkg)
    n = small_chunks + 7
    flat = b"".join(syn.blob(900 + i) for i in range(n))
    os.environ["EKZG_CHUNK"] = "1024"
    cells, proofs, st = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    os.environ["EKZG_CHUNK"] = str(small_chunks)
    idx = [list(range(i % 2, 128, 2)) if i % 3 else list(range(64, 128)) for i in range(n)]
    cl = [[cells[(b * 128 + j) * 2048:(b * 128 + j + 1) * 2048] for j in idx[b]] for b in range(n)]
    oc, op, st = das_ctx.recover_cells_and_kzg_proofs_batch(idx, cl)
    assert not any(st) and oc == cells and op == proofs
