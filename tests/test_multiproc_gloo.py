"""Multi-process (world_size 2, gloo, CPU) checks of the shard bookkeeping behind the N > 1 path
(rust-eth-kzg_b200/sharding.py): the partition covers every blob once, the gather restores blob order for ragged
batch sizes, and times reduce as max over ranks.  The per-shard computation is a stand-in (a digest of each blob): the
real one is DASContext.compute_cells_and_kzg_proofs_batch on the rank's GPU (tests/test_gpu_fk20.py)."""
import hashlib
import os
import socket

import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _standin(shard, cnt):
    cells, proofs, status = bytearray(), bytearray(), []
    for i in range(cnt):
        b = shard[i * 131072:(i + 1) * 131072]
        d = hashlib.sha256(b).digest()
        cells += (d * (128 * 2048 // 32))
        proofs += (d[:16] * (128 * 48 // 16))
        status.append(b[0] & 1)
    return bytes(cells), bytes(proofs), status


def _worker(rank, world, port, n, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import importlib
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.load_package()
    sh = importlib.import_module("eth_kzg_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blobs = b"".join(bytes([i % 251]) * 131072 for i in range(n))
    res = sh.compute_cells_and_kzg_proofs_sharded(_standin, blobs, n)
    slow = sh.max_over_ranks(10.0 + rank)
    # the tensor-level gather of the timed path (bench.py "strong"): ragged shards, rows in rank order on rank 0 only
    import torch
    lo, cnt = sh.shard_bounds(n, world, rank)
    counts = [sh.shard_bounds(n, world, r)[1] for r in range(world)]
    rows = torch.arange(lo, lo + cnt, dtype=torch.uint8).view(cnt, 1).repeat(1, 48)
    got = sh.gather_rows(rows, counts)
    if rank == 0:
        assert got.shape == (n, 48) and bool((got[:, 0] == torch.arange(n, dtype=torch.uint8)).all())
    else:
        assert got is None
    if rank == 0:
        q.put((res == _standin(blobs, n), slow))
    else:
        assert res is None
    dist.destroy_process_group()


def test_shard_bounds_partition():
    import importlib
    import __graft_entry__
    __graft_entry__.load_package()
    sh = importlib.import_module("eth_kzg_b200.sharding")
    for n in (0, 1, 2, 7, 128, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [sh.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (lo, c), (lo2, _) in zip(spans, spans[1:]):
                assert lo + c == lo2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_gather_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    for n in (5, 2, 1):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
        for p in procs:
            p.start()
        same, slow = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert same, "gathered batch differs from the single-process result (n=%d)" % n
        assert slow == 11.0
