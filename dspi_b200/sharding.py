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
    """Contiguous range [lo, hi) of `total` rows owned by `rank`: 64-row units (one warp group of K1 never straddles two
    ranks) dealt out evenly - the same arithmetic as dspi_eqx_shard_range() in the C library."""
    unit = 64
    units = (total + unit - 1) // unit
    lo = min(total, units * rank // world * unit)
    hi = min(total, units * (rank + 1) // world * unit)
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


def chunk_ranges(n_rows, n_chunks, unit=64):
    """Row ranges [(lo, hi), ...] of one shard: at most `n_chunks` pieces on `unit`-row boundaries."""
    per = -(-n_rows // max(1, n_chunks))
    per = max(unit, -(-per // unit) * unit)
    return [(lo, min(n_rows, lo + per)) for lo in range(0, n_rows, per)]


def pipelined_scatter_process_gather(full, total_rows, row_len, dtype, device, process_range, n_chunks=8, root=0,
                                     compute_stream=None):
    """Frames originate on `root` ([total_rows, row_len]); every rank processes its contiguous shard and the results
    land back in `full` on root - chunked and software-pipelined:

        step j:   one grouped NCCL launch carries chunk j root -> peers AND chunk j-2 peers -> root (both NVLink
                  directions busy), while every rank's kernel works on chunk j-1.

    ``process_range(shard, lo, hi)`` enqueues the processing of rows [lo, hi) of the rank's shard tensor, in place
    (asynchronous on ``compute_stream`` for CUDA tensors, synchronous on CPU).  Returns the rank's shard tensor
    (root: a view of its rows of `full`).  Sharding and chunking change no bit of the result."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(total_rows, rank, world)
    cuda = torch.device(device).type == "cuda"
    if rank == root:
        mine = full[lo:hi]
    else:
        mine = torch.empty((hi - lo, row_len), dtype=dtype, device=device)
    ranges = {r: chunk_ranges(shard_range(total_rows, r, world)[1] - shard_range(total_rows, r, world)[0], n_chunks) for r in range(world)}
    K = max(len(v) for v in ranges.values())
    comm = torch.cuda.Stream(device=device) if cuda else None
    comp = compute_stream
    ev_recv = [torch.cuda.Event() for _ in range(K)] if cuda else None
    ev_done = [torch.cuda.Event() for _ in range(K)] if cuda else None
    if cuda:
        comm.wait_stream(torch.cuda.current_stream(device))          # `full` / `mine` are ready

    def compute(j):
        mr = ranges[rank]
        if j >= len(mr):
            return
        if cuda:
            comp.wait_event(ev_recv[j])
        process_range(mine, mr[j][0], mr[j][1])
        if cuda:
            ev_done[j].record(comp)

    if world == 1 or rank == root:
        # the root's own rows need no transfer: its kernel runs beside the transfers
        if cuda:
            comp.wait_stream(torch.cuda.current_stream(device))
        for a, b in ranges[rank]:
            process_range(mine, a, b)
        if world == 1:
            if cuda:
                torch.cuda.current_stream(device).wait_stream(comp)
            return mine
    for j in range(K + 2):
        ops = []
        if rank == root:
            for r in range(world):
                if r == root:
                    continue
                base = shard_range(total_rows, r, world)[0]
                rr = ranges[r]
                if j < len(rr):
                    ops.append(dist.P2POp(dist.isend, full[base + rr[j][0]:base + rr[j][1]], r))
                if 0 <= j - 2 < len(rr):
                    ops.append(dist.P2POp(dist.irecv, full[base + rr[j - 2][0]:base + rr[j - 2][1]], r))
        else:
            mr = ranges[rank]
            if j < len(mr):
                ops.append(dist.P2POp(dist.irecv, mine[mr[j][0]:mr[j][1]], root))
            if 0 <= j - 2 < len(mr):
                if cuda:
                    comm.wait_event(ev_done[j - 2])                   # results of chunk j-2 are complete
                ops.append(dist.P2POp(dist.isend, mine[mr[j - 2][0]:mr[j - 2][1]], root))
        if ops:
            if cuda:
                with torch.cuda.stream(comm):
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                    if rank != root and j < len(ranges[rank]):
                        ev_recv[j].record(comm)
            else:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        if rank != root:
            compute(j)                                               # chunk j: its kernel overlaps the next step's transfers
    if cuda:
        cur = torch.cuda.current_stream(device)
        cur.wait_stream(comm)
        cur.wait_stream(comp)
    return mine


def native_scatter_gather(engine, device_index, root=0):
    """The native pipeline (dspi_b200/csrc/nccl_sg.cu: NCCL calls issued from the C library, no Python in the loop) for
    the ranks of the default process group.  The engine of rank r must hold the channels
    ``shard_range(total, r, world)``.  Returns an ``api.ScatterGather``; call ``.process(full.data_ptr() or 0, total, T)``
    on every rank."""
    from dspi_b200 import api
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [api.nccl_unique_id() if rank == root else None]
    dist.broadcast_object_list(box, src=root)
    return api.ScatterGather(engine, device_index, box[0], rank, world, root)
