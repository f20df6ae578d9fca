#!/usr/bin/env python
"""Frames on rank 0 -> shards -> kernels -> back to rank 0 over NCCL: the plain grouped scatter / gather of round 1 against
the chunked, software-pipelined form, for several chunk counts.  Launch with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/nccl_sg_bench.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dspi_b200 import api, sharding, workloads as W          # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
Cn, T, FS = 65536, 6144, 96000.0
total = Cn * world
dev = torch.device("cuda", local_rank)
params = W.eq_params_fast("A", Cn, fs=FS, seed=1, ch0=rank * Cn)
eng = api.EqEngine("f32f", Cn, device=local_rank)
eng.upload(api.compute_coefficients(params, fs=FS))
eng.kernel_info()
comp = torch.cuda.ExternalStream(eng.stream, device=dev)
full = (torch.rand((total, T), dtype=torch.float32, device=dev) - 0.5) if rank == 0 else None


def process_range(shard, a, b):
    eng.process_device_range(shard[a:b].data_ptr(), T, T, a, b - a)


def plain():
    mine = sharding.scatter_rows(full, total, T, torch.float32, dev)
    torch.cuda.current_stream().synchronize()
    eng.process_device(mine.data_ptr(), T, T)
    eng.sync()
    sharding.gather_rows(mine, total)


def piped(k):
    sharding.pipelined_scatter_process_gather(full, total, T, torch.float32, dev, process_range, n_chunks=k, compute_stream=comp)
    torch.cuda.current_stream().synchronize()


def timeit(fn, reps=4):
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item()) / reps


sg = sharding.native_scatter_gather(eng, local_rank)


def native(k):
    sg.process(full.data_ptr() if rank == 0 else 0, total, T, k)


res = {"world": world, "bytes_each_way": (total - Cn) * T * 4}
res["plain_ms"] = timeit(plain) * 1e3
ks = [int(v) for v in os.environ.get("SG_CHUNKS", "1,2,4,8,16").split(",")]
if os.environ.get("SG_TORCH", "1") == "1":
    for k in (2, 4):
        res[f"piped_{k}_ms"] = timeit(lambda: piped(k)) * 1e3
for k in ks:
    res[f"native_{k}_ms"] = timeit(lambda: native(k)) * 1e3
res["native_auto_ms"] = timeit(lambda: native(0)) * 1e3
if rank == 0:
    for k, v in list(res.items()):
        if k.endswith("_ms"):
            res[k.replace("_ms", "_Gsamples_s")] = total * T / (v * 1e-3) / 1e9
    print(json.dumps(res))
sg.close()
eng.close()
dist.destroy_process_group()
