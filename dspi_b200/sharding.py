"""Channel / instance sharding across the GPUs of one box.

The path has no cross-channel dependency, so a job shards into contiguous channel ranges with all
filter / delay / modulator state staying on its owner.  The only communication the north star asks
for is moving frames: scatter input rows from a root rank and gather output rows back
(``torch.distributed`` over NCCL/NVLink on the GPU box; the same code runs over gloo on CPU, which is
how the N>1 logic is tested without GPUs).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous range [lo, hi) of `total` rows owned by `rank`; sizes differ by at most one and
    every boundary is a multiple of 64 channels when total is (one warp group never straddles ranks)."""
    unit = 64 if total % (64 * world) == 0 else 1
    n = total // unit
    lo = (n * rank) // world * unit
    hi = (n * (rank + 1)) // world * unit
    return lo, hi


def scatter_rows(full, total_rows, row_len, dtype, device, root=0):
    """Root holds `full` [total_rows, row_len]; every rank returns its [hi-lo, row_len] shard."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(total_rows, rank, world)
    mine = torch.empty((hi - lo, row_len), dtype=dtype, device=device)
    if world == 1:
        mine.copy_(full[lo:hi])
        return mine
    if rank == root:
        ops = []
        for r in range(world):
            rlo, rhi = shard_range(total_rows, r, world)
            if r == root:
                mine.copy_(full[rlo:rhi])
            else:
                ops.append(dist.P2POp(dist.isend, full[rlo:rhi], r))      # contiguous row blocks: no staging copy
        for q in dist.batch_isend_irecv(ops):                              # one grouped NCCL launch for all peers
            q.wait()
    else:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, root)]):
            q.wait()
    return mine


def gather_rows(mine, total_rows, root=0):
    """Inverse of :func:`scatter_rows`; returns the assembled tensor on root, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return mine
    if rank == root:
        full = torch.empty((total_rows, mine.shape[1]), dtype=mine.dtype, device=mine.device)
        ops = []
        for r in range(world):
            rlo, rhi = shard_range(total_rows, r, world)
            if r == root:
                full[rlo:rhi].copy_(mine)
            else:
                ops.append(dist.P2POp(dist.irecv, full[rlo:rhi], r))
        for q in dist.batch_isend_irecv(ops):
            q.wait()
        return full
    for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(), root)]):
        q.wait()
    return None
