"""Re-entrancy of the C ABI on ONE shared DASContext (SURVEY.md §8b "Threading": the reference's context is immutable and
its bindings call it from thread pools -- bindings/node/src/lib.rs:92-130).  ctypes releases the GIL during the calls,
so these host threads really are concurrently inside the library."""
import threading

import pytest

pytestmark = pytest.mark.gpu


def _synth(pkg):
    import importlib
    return importlib.import_module("eth_kzg_b200.synthetic")


def test_concurrent_callers_share_one_context(das_ctx, pkg):
    syn = _synth(pkg)
    blobs = [syn.blob(100 + i) for i in range(6)]
    # serial ground truth
    want = []
    for b in blobs:
        cells, proofs = das_ctx.compute_cells_and_kzg_proofs(b)
        c = das_ctx.blob_to_kzg_commitment(b)
        p = das_ctx.compute_blob_kzg_proof(b, c)
        want.append((cells, proofs, c, p))
    errors = []

    def worker(tid):
        try:
            for rep in range(3):
                for k in range(len(blobs)):
                    i = (k + tid) % len(blobs)
                    b = blobs[i]
                    cells, proofs = das_ctx.compute_cells_and_kzg_proofs(b)
                    assert (cells, proofs) == want[i][:2]
                    assert das_ctx.blob_to_kzg_commitment(b) == want[i][2]
                    assert das_ctx.compute_blob_kzg_proof(b, want[i][2]) == want[i][3]
                    assert das_ctx.verify_blob_kzg_proof(b, want[i][2], want[i][3]) is True
                    idx = list(range(tid % 2, 128, 2))
                    rc, rp = das_ctx.recover_cells_and_kzg_proofs(idx, [cells[j] for j in idx])
                    assert (rc, rp) == want[i][:2]
                    assert das_ctx.verify_cell_kzg_proof_batch([want[i][2]] * 4, [0, 5, 64, 127], [cells[j] for j in (0, 5, 64, 127)],
                                                               [proofs[j] for j in (0, 5, 64, 127)]) is True
        except Exception as ex:  # noqa: BLE001 - surfaced below
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
        assert not t.is_alive(), "a caller thread hung"
    assert not errors, errors


def test_concurrent_batches(das_ctx, pkg):
    """two host threads each pushing a multi-chunk-sized batch through the scheduler at once"""
    syn = _synth(pkg)
    n = 96
    flat = [b"".join(syn.blob(1000 * t + i) for i in range(n)) for t in range(2)]
    want = [das_ctx.compute_cells_and_kzg_proofs_batch(f, n) for f in flat]
    got = [None, None]

    def worker(t):
        got[t] = das_ctx.compute_cells_and_kzg_proofs_batch(flat[t], n)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
        assert not t.is_alive()
    assert got == want


def test_coalesced_single_blob_callers(das_ctx, pkg):
    """48 host threads inside eth_kzg_compute_cells_and_kzg_proofs / eth_kzg_compute_cells at once: the library coalesces
    them into shared batches (kzg_runtime.h compute_cells_and_kzg_proofs_one).  Every caller must get ITS blob's result, and
    a non-canonical blob among them must fail alone."""
    syn = _synth(pkg)
    nthreads = 48
    blobs = [syn.blob(5000 + i) for i in range(nthreads)]
    bad = 17
    blobs[bad] = b"\xff" * 32 + blobs[bad][32:]          # first field element >= r
    flat = b"".join(b for i, b in enumerate(blobs) if i != bad)
    cells_flat, proofs_flat, _ = das_ctx.compute_cells_and_kzg_proofs_batch(flat, nthreads - 1)
    want = {}
    for pos, i in enumerate(j for j in range(nthreads) if j != bad):
        want[i] = (cells_flat[pos * 262144:(pos + 1) * 262144], proofs_flat[pos * 6144:(pos + 1) * 6144])
    results, errors = [None] * nthreads, []
    start = threading.Barrier(nthreads)

    def worker(i):
        try:
            start.wait()
            for rep in range(2):
                try:
                    cells, proofs = das_ctx.compute_cells_and_kzg_proofs(blobs[i])
                    only_cells = das_ctx.compute_cells(blobs[i])
                    results[i] = (b"".join(cells), b"".join(proofs), b"".join(only_cells))
                except pkg.KzgError:
                    results[i] = "err"
        except Exception as ex:  # noqa: BLE001
            errors.append((i, repr(ex)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
        assert not t.is_alive(), "a caller thread hung"
    assert not errors, errors
    for i in range(nthreads):
        if i == bad:
            assert results[i] == "err"
        else:
            assert results[i] == (want[i][0], want[i][1], want[i][0]), i


def test_coalesced_recover_callers(das_ctx, pkg):
    """32 host threads inside eth_kzg_recover_cells_and_proofs at once, each with its own blob and erasure pattern; one of them
    hands in a non-canonical cell and must fail alone, one has shuffled indices and must get the index error without ever
    joining a batch."""
    syn = _synth(pkg)
    nthreads = 32
    flat = b"".join(syn.blob(7000 + i) for i in range(nthreads))
    cells_flat, proofs_flat, _ = das_ctx.compute_cells_and_kzg_proofs_batch(flat, nthreads)
    want = [(cells_flat[i * 262144:(i + 1) * 262144], proofs_flat[i * 6144:(i + 1) * 6144]) for i in range(nthreads)]
    patterns = [list(range(64)), list(range(64, 128)), list(range(0, 128, 2)), list(range(1, 128, 2)), list(range(20, 100)), list(range(128))]
    bad_cell, bad_idx = 9, 21
    results, errors = [None] * nthreads, []
    start = threading.Barrier(nthreads)

    def worker(i):
        try:
            idx = list(patterns[i % len(patterns)])
            cells = [want[i][0][j * 2048:(j + 1) * 2048] for j in idx]
            if i == bad_cell:
                cells[3] = b"\xff" * 32 + cells[3][32:]
            if i == bad_idx:
                idx[0], idx[1] = idx[1], idx[0]
                cells[0], cells[1] = cells[1], cells[0]
            start.wait()
            for rep in range(2):
                try:
                    rc, rp = das_ctx.recover_cells_and_kzg_proofs(idx, cells)
                    results[i] = (b"".join(rc), b"".join(rp))
                except pkg.KzgError as ex:
                    results[i] = "err:" + str(ex)
        except Exception as ex:  # noqa: BLE001
            errors.append((i, repr(ex)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
        assert not t.is_alive(), "a caller thread hung"
    assert not errors, errors
    for i in range(nthreads):
        if i == bad_cell:
            assert isinstance(results[i], str) and "ScalarNotCanonical" in results[i], results[i]
        elif i == bad_idx:
            assert isinstance(results[i], str) and "NotUniquelyOrdered" in results[i], results[i]
        else:
            assert results[i] == want[i], i
