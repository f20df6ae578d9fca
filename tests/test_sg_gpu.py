"""One rank per GPU, frames on rank 0 (dspi_sg_*, the native NCCL scatter / process / gather pipeline): the block that comes
back to the root is byte-identical to one engine over all channels on one GPU (SURVEY §8e invariant, on hardware).  Spawns
one process per GPU; skips on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FS = 96000.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Cn, T, n_chunks, out_path):
    import torch.distributed as dist
    from dspi_b200 import api, sharding, workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lo, hi = sharding.shard_range(Cn, rank, world)
    params = W.eq_params("mixed", Cn, fs=FS, seed=9)[lo:hi].copy()
    eng = api.EqEngine("f32f", hi - lo, device=rank)
    eng.upload(api.compute_coefficients(params, fs=FS))
    sg = sharding.native_scatter_gather(eng, rank)
    full = torch.from_numpy(W.inputs_f32(Cn, T)).cuda() if rank == 0 else None
    for _ in range(2):                                     # two passes: state carries on every rank
        sg.process(full.data_ptr() if rank == 0 else 0, Cn, T, n_chunks)
    if rank == 0:
        np.save(out_path, full.cpu().numpy())
    dist.barrier()
    sg.close()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_chunks", [(2, 1), (2, 5), (4, 8), (8, 8)])
def test_native_pipeline_equals_single_engine(tmp_path, world, n_chunks):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from dspi_b200 import api, workloads as W
    Cn, T = 64 * 45, 384
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), Cn, T, n_chunks, out), nprocs=world, join=True)
    got = np.load(out)
    eng = api.EqEngine("f32f", Cn)
    try:
        eng.upload(api.compute_coefficients(W.eq_params("mixed", Cn, fs=FS, seed=9), fs=FS))
        buf = torch.from_numpy(W.inputs_f32(Cn, T)).cuda()
        for _ in range(2):
            eng.process_device(buf.data_ptr(), T, T)
        eng.sync()
        want = buf.cpu().numpy()
    finally:
        eng.close()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
