"""Multi-GPU sharding of a batch of blobs: blobs are independent, so a batch of n blobs is cut into
`world` contiguous shards, one per process / GPU (tables are replicated per device), every rank runs its shard through
its own DASContext, and the results are gathered to rank 0.  There is no data-path collective; torch.distributed is
only the plumbing for the gather (NCCL over NVLink when the group is an NCCL group, gloo otherwise).

Replaces the role of the reference's rayon fan-out across CPU cores (crates/maybe_rayon/src/multi_threaded.rs:9-39) at
the level above one device."""
import torch
import torch.distributed as dist

BYTES_PER_BLOB = 131072
CELLS_BYTES = 128 * 2048
PROOFS_BYTES = 128 * 48


def shard_bounds(n, world, rank):
    """contiguous partition: GPU g gets blobs [g*n/world, (g+1)*n/world) (sizes differ by at most one)"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi - lo


def _group_device(group):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def gather_bytes(local, sizes, group=None, dst=0):
    """gather per-rank byte strings of known sizes to rank dst; returns the concatenation there, None elsewhere"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _group_device(group)
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if len(local):
        buf[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    # all_gather works on every backend (NCCL has no native gather-to-one in older torch builds)
    out = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    if rank != dst:
        return None
    return b"".join(bytes(out[r][:sizes[r]].cpu().numpy()) for r in range(world))


def compute_cells_and_kzg_proofs_sharded(compute_batch, blobs_flat, n, group=None, dst=0):
    """Every rank passes the SAME n blobs (or at least its own shard's bytes at the right offsets); rank r computes
    shard r with compute_batch(shard_bytes, count) -> (cells, proofs, status) -- in production
    DASContext.compute_cells_and_kzg_proofs_batch -- and rank dst gets (cells, proofs, status) of the whole batch in
    blob order."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, cnt = shard_bounds(n, world, rank)
    if cnt:
        cells, proofs, status = compute_batch(blobs_flat[lo * BYTES_PER_BLOB:(lo + cnt) * BYTES_PER_BLOB], cnt)
    else:
        cells, proofs, status = b"", b"", []
    counts = [shard_bounds(n, world, r)[1] for r in range(world)]
    all_cells = gather_bytes(cells, [c * CELLS_BYTES for c in counts], group, dst)
    all_proofs = gather_bytes(proofs, [c * PROOFS_BYTES for c in counts], group, dst)
    all_status = gather_bytes(bytes(status), counts, group, dst)
    if rank != dst:
        return None
    return all_cells, all_proofs, list(all_status)


def gather_rows(local, counts, group=None, dst=0):
    """Tensor-level gather for the timed path: `local` is this rank's [counts[rank], row_bytes] uint8 tensor on the group's
    device (CUDA for NCCL -- the bytes move GPU to GPU over NVLink -- or CPU for gloo).  Rank dst gets one
    [sum(counts), row_bytes] tensor in rank order, the others None.  Uneven shards are padded to the largest."""
    if not dist.is_initialized():
        return local
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return local
    mx = max(counts)
    row = local.shape[1]
    if local.shape[0] != mx:
        pad = torch.zeros((mx, row), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        local = pad
    local = local.contiguous()
    if rank == dst:
        out = torch.empty((world, mx, row), dtype=local.dtype, device=local.device)
        dist.gather(local, list(out.unbind(0)), dst=dst, group=group)
        if all(c == mx for c in counts):
            return out.view(world * mx, row)
        return torch.cat([out[r, :counts[r]] for r in range(world)], 0)
    dist.gather(local, None, dst=dst, group=group)
    return None


def max_over_ranks(value, group=None):
    """the timing rule of bench.py: a multi-GPU time is the MAX over ranks"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=_group_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
