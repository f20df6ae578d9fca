"""N>1 host logic on CPU (gloo, world_size 2): shard ranges, frame scatter / gather, and the
invariant that sharding does not change a bit of the result (the per-shard compute is done by the
oracle here — there is no GPU in this test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dspi_b200 import layouts as L, sharding, workloads as W


def test_shard_ranges_cover_and_align():
    for total in (65536, 524288, 32768, 100, 7):
        for world in (1, 2, 4, 8):
            edges = [sharding.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            if total % (64 * world) == 0:
                assert all(lo % 64 == 0 for lo, _ in edges)
                assert len({hi - lo for lo, hi in edges}) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Cn, T, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.orc import Oracle
    orc = Oracle()
    fs = 96000.0
    x = torch.from_numpy(W.inputs_f32(Cn, T)) if rank == 0 else None
    mine = sharding.scatter_rows(x, Cn, T, torch.float32, "cpu")
    lo, hi = sharding.shard_range(Cn, rank, world)
    params = W.eq_params("B", hi - lo, fs=fs, seed=4, ch0=lo)       # rows keyed by absolute channel index
    bq = np.zeros(params.shape, L.BIQUAD_F32)
    orc.eq_coeffs(False, params, bq, fs)
    y = mine.numpy().copy()
    orc.eq_many("f32f", bq, y, 10, 96)
    full = sharding.gather_rows(torch.from_numpy(y), Cn)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_process_gather_world2(tmp_path, oracle):
    Cn, T, fs = 128, 480, 96000.0
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), Cn, T, out), nprocs=2, join=True)
    got = np.load(out)
    params = W.eq_params("B", Cn, fs=fs, seed=4)
    bq = np.zeros(params.shape, L.BIQUAD_F32)
    oracle.eq_coeffs(False, params, bq, fs)
    want = W.inputs_f32(Cn, T)
    oracle.eq_many("f32f", bq, want, 10, 96)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _worker_pipelined(rank, world, port, Cn, T, out_path, n_chunks):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.orc import Oracle
    orc = Oracle()
    fs = 96000.0
    lo, hi = sharding.shard_range(Cn, rank, world)
    params = W.eq_params("B", hi - lo, fs=fs, seed=4, ch0=lo)
    bq = np.zeros(params.shape, L.BIQUAD_F32)
    orc.eq_coeffs(False, params, bq, fs)

    def process_range(shard, a, b):                              # in place on rows [a, b) of the rank's shard
        y = shard[a:b].numpy()
        orc.eq_many("f32f", bq[a:b], y, 10, 96)

    x = torch.from_numpy(W.inputs_f32(Cn, T)) if rank == 0 else None
    sharding.pipelined_scatter_process_gather(x, Cn, T, torch.float32, "cpu", process_range, n_chunks=n_chunks)
    if rank == 0:
        np.save(out_path, x.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_scatter_process_gather_world2(tmp_path, oracle):
    """The chunked, software-pipelined form (transfer of chunk j, kernel of chunk j-1 and return of chunk j-2 in one step)
    gives the bits of the unsharded run - for shards that split into 1, 3 and more chunks than rows allow."""
    Cn, T, fs = 512, 192, 96000.0
    params = W.eq_params("B", Cn, fs=fs, seed=4)
    bq = np.zeros(params.shape, L.BIQUAD_F32)
    oracle.eq_coeffs(False, params, bq, fs)
    want = W.inputs_f32(Cn, T)
    oracle.eq_many("f32f", bq, want, 10, 96)
    for n_chunks in (1, 3, 16):
        out = str(tmp_path / f"piped{n_chunks}.npy")
        mp.spawn(_worker_pipelined, args=(2, _free_port(), Cn, T, out, n_chunks), nprocs=2, join=True)
        assert np.array_equal(np.load(out).view(np.uint32), want.view(np.uint32)), n_chunks


def test_chunk_ranges():
    for n, k in ((65536, 8), (256, 3), (64, 8), (100, 4), (8192, 1)):
        r = sharding.chunk_ranges(n, k)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert len(r) <= max(1, k) and all(lo % 64 == 0 for lo, _ in r)
