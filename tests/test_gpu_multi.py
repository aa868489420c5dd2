"""One DASContext over several devices (csrc/kzg_multi.cu, EKZG_DEVICES): a batch is cut into contiguous shards of whole
blob groups (32 blobs, or 8 where a share is small), one per member context, each on a host thread of its own, results written straight into the caller's
buffers.  On a one-GPU box the member contexts share device 0 (EKZG_DEVICES=0,0 -- same code path, two sets of tables and
queues); with two or more GPUs visible the same tests also run on devices 0,1.  Everything is compared with the
single-device session context, which the consensus vectors pin."""
import os
import threading

import pytest

pytestmark = pytest.mark.gpu


def _synth(pkg):
    import importlib
    return importlib.import_module("eth_kzg_b200.synthetic")


def _device_lists():
    import torch
    lists = ["0,0", "0,0,0"]
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        lists.append("0,1")
    if torch.cuda.is_available() and torch.cuda.device_count() >= 4:
        lists.append("all")
    return lists


@pytest.fixture(scope="module", params=_device_lists())
def multi_ctx(request, pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    old = os.environ.get("EKZG_DEVICES")
    os.environ["EKZG_DEVICES"] = request.param
    try:
        ctx = pkg.DASContext(use_precomp=False)
    finally:
        if old is None:
            del os.environ["EKZG_DEVICES"]
        else:
            os.environ["EKZG_DEVICES"] = old
    yield ctx
    ctx.close()


def test_device_list(multi_ctx):
    import torch
    devs = multi_ctx.devices
    assert len(devs) >= 2 and all(0 <= d < torch.cuda.device_count() for d in devs)



def test_sharded_batch_equals_single_device(das_ctx, multi_ctx, pkg):
    import torch
    syn = _synth(pkg)
    before = torch.cuda.current_device()
    for n in (12, 33, 100, 64 * len(multi_ctx.devices) + 5, 170):
        flat = b"".join(syn.blob(3000 + i) for i in range(n))
        want = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
        got = multi_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
        assert got == want, "n=%d" % n
        cells_only = multi_ctx.compute_cells_and_kzg_proofs_batch(flat, n, want_proofs=False)
        assert cells_only[0] == want[0] and cells_only[1] is None
        comm, st = multi_ctx.blob_to_kzg_commitment_batch(flat, n)
        comm1, st1 = das_ctx.blob_to_kzg_commitment_batch(flat, n)
        assert (comm, st) == (comm1, st1)
        prf, st = multi_ctx.compute_blob_kzg_proof_batch(flat, comm, n)
        prf1, st1 = das_ctx.compute_blob_kzg_proof_batch(flat, comm1, n)
        assert (prf, st) == (prf1, st1)
    assert torch.cuda.current_device() == before, "the library must leave the caller's current device alone"


def test_invalid_blob_in_a_later_shard(das_ctx, multi_ctx, pkg):
    syn = _synth(pkg)
    n = 70
    blobs = [syn.blob(3500 + i) for i in range(n)]
    bad = 66                                                   # lands in the last shard
    blobs[bad] = b"\xff" * 32 + blobs[bad][32:]
    flat = b"".join(blobs)
    cells, proofs, st = multi_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    assert st == [1 if i == bad else 0 for i in range(n)]
    wc, wp, wst = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    assert st == wst
    for i in range(n):
        if i != bad:
            assert cells[i * 262144:(i + 1) * 262144] == wc[i * 262144:(i + 1) * 262144]
            assert proofs[i * 6144:(i + 1) * 6144] == wp[i * 6144:(i + 1) * 6144]


def test_sharded_recovery(das_ctx, multi_ctx, pkg):
    syn = _synth(pkg)
    n = 40
    flat = b"".join(syn.blob(3700 + i) for i in range(n))
    cells, proofs, _ = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    # ragged counts so that the shards start at different offsets of the concatenated inputs
    idx = [list(range(i % 2, 128, 2)) if i % 3 == 0 else (list(range(64, 128)) if i % 3 == 1 else list(range(0, 100))) for i in range(n)]
    cl = [[cells[(b * 128 + j) * 2048:(b * 128 + j + 1) * 2048] for j in idx[b]] for b in range(n)]
    oc, op, st = multi_ctx.recover_cells_and_kzg_proofs_batch(idx, cl)
    assert not any(st) and oc == cells and op == proofs
    idx[37] = idx[37][:10]                                      # too few cells: that item alone fails
    cl[37] = cl[37][:10]
    oc, op, st = multi_ctx.recover_cells_and_kzg_proofs_batch(idx, cl)
    assert st == [3 if i == 37 else 0 for i in range(n)]
    assert oc[:37 * 262144] == cells[:37 * 262144] and op[38 * 6144:] == proofs[38 * 6144:]


def test_single_item_calls_round_robin(das_ctx, multi_ctx, pkg):
    syn = _synth(pkg)
    blobs = [syn.blob(3900 + i) for i in range(6)]
    want = [das_ctx.compute_cells_and_kzg_proofs(b) for b in blobs]
    got = [None] * len(blobs)
    errs = []

    def work(i):
        try:
            got[i] = multi_ctx.compute_cells_and_kzg_proofs(blobs[i])
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(blobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs and got == want
    comm = multi_ctx.blob_to_kzg_commitment(blobs[0])
    assert comm == das_ctx.blob_to_kzg_commitment(blobs[0])
    cells, proofs = want[1]
    assert multi_ctx.verify_cell_kzg_proof_batch([das_ctx.blob_to_kzg_commitment(blobs[1])] * 4, [0, 5, 64, 127],
                                                 [cells[0], cells[5], cells[64], cells[127]], [proofs[0], proofs[5], proofs[64], proofs[127]]) is True


def test_device_entry_point_finds_the_owning_member(das_ctx, multi_ctx, pkg):
    import torch
    syn = _synth(pkg)
    n = 8
    flat = b"".join(syn.blob(4100 + i) for i in range(n))
    want = das_ctx.compute_cells_and_kzg_proofs_batch(flat, n)
    for dev in sorted(set(multi_ctx.devices)):
        with torch.cuda.device(dev):
            d_in = torch.frombuffer(bytearray(flat), dtype=torch.uint8).cuda()
            d_cells = torch.empty(n * 262144, dtype=torch.uint8, device="cuda")
            d_proofs = torch.empty(n * 6144, dtype=torch.uint8, device="cuda")
            d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
            multi_ctx.compute_cells_and_kzg_proofs_device(n, d_in.data_ptr(), d_cells.data_ptr(), d_proofs.data_ptr(), d_st.data_ptr(),
                                                          torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert bytes(d_cells.cpu().numpy()) == want[0] and bytes(d_proofs.cpu().numpy()) == want[1] and int(d_st.sum()) == 0


def _sharded_worker(rank, world, port, n, backend, q):
    try:
        _sharded_worker_body(rank, world, port, n, backend, q)
    except BaseException as ex:   # the parent must hear about it at once, not after its queue timeout
        q.put(("error", "rank %d: %r" % (rank, ex)))
        raise


def _sharded_worker_body(rank, world, port, n, backend, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import importlib
    import torch
    import torch.distributed as dist
    import __graft_entry__
    pkg = __graft_entry__.load_package()
    sh = importlib.import_module("eth_kzg_b200.sharding")
    syn = importlib.import_module("eth_kzg_b200.synthetic")
    torch.cuda.set_device(rank % torch.cuda.device_count())
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    ctx = pkg.DASContext(use_precomp=False)
    blobs = b"".join(syn.blob(4000 + i) for i in range(n))
    res = sh.compute_cells_and_kzg_proofs_sharded(lambda shard, cnt: ctx.compute_cells_and_kzg_proofs_batch(shard, cnt), blobs, n)
    # the tensor path bench.py times: device entry point on the shard, rows gathered to rank 0
    lo, cnt = sh.shard_bounds(n, world, rank)
    counts = [sh.shard_bounds(n, world, r)[1] for r in range(world)]
    d_in = torch.frombuffer(bytearray(blobs[lo * 131072:(lo + cnt) * 131072]), dtype=torch.uint8).cuda()
    d_cells = torch.empty((cnt, 128 * 2048), dtype=torch.uint8, device="cuda")
    d_proofs = torch.empty((cnt, 128 * 48), dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(cnt, dtype=torch.int32, device="cuda")
    ctx.compute_cells_and_kzg_proofs_device(cnt, d_in.data_ptr(), d_cells.data_ptr(), d_proofs.data_ptr(), d_status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    on_group = (lambda t: t) if backend == "nccl" else (lambda t: t.cpu())
    g_cells, g_proofs = sh.gather_rows(on_group(d_cells), counts), sh.gather_rows(on_group(d_proofs), counts)
    if rank == 0:
        want = ctx.compute_cells_and_kzg_proofs_batch(blobs, n)
        ok_bytes = res[0] == want[0] and res[1] == want[1] and list(res[2]) == list(want[2])
        ok_rows = bytes(g_cells.cpu().numpy().tobytes()) == want[0] and bytes(g_proofs.cpu().numpy().tobytes()) == want[1]
        q.put((ok_bytes, ok_rows))
    else:
        assert res is None and g_cells is None
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [33, 70])
def test_sharded_processes_gather_real_compute(n, precomp_holder):
    """sharding.py with the REAL per-shard computation (two processes, one DASContext each): NCCL over NVLink when the box has
    two GPUs, otherwise two processes on device 0 gathering through gloo.  Rank 0 compares the gathered batch with what one
    context computes for the whole batch."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    precomp_holder.release()   # the children build contexts of their own: give the 144 GiB of the shared production context back first
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_sharded_worker, args=(r, 2, port, n, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _queue
    import time
    res, deadline = None, time.monotonic() + 600
    while res is None and time.monotonic() < deadline:
        try:
            res = q.get(timeout=2)
        except _queue.Empty:
            if all(p.exitcode is not None for p in procs):   # both children are gone and said nothing
                break
    for p in procs:
        p.join(timeout=120 if res is not None and res[0] != "error" else 5)
        if p.is_alive():
            p.terminate()
    assert res is not None, "the sharding processes ended without a result (exit codes %r)" % [p.exitcode for p in procs]
    assert res[0] != "error", res[1]
    ok_bytes, ok_rows = res
    for p in procs:
        assert p.exitcode == 0
    assert ok_bytes, "byte-level gather differs from the single-context batch"
    assert ok_rows, "tensor-level gather differs from the single-context batch"
