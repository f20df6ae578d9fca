#!/usr/bin/env python
"""BASELINE config 3 alone: N instances through the whole chain, device-resident, CUDA events.
    python scripts/chain_bench.py [--instances 8192] [--packets 16] [--fpp 96] [--reps 3]
Used under ncu for the per-kernel breakdown (profiles/*chain*)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                              # noqa: E402
from dspi_b200 import api, workloads as W                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=8192)
ap.add_argument("--packets", type=int, default=16)
ap.add_argument("--fpp", type=int, default=96)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--arith", default="f32f")
ap.add_argument("--no-sub", action="store_true", help="disable the sub output: no modulator work (isolates the other stages)")
a = ap.parse_args()
N, F, fs = a.instances, a.packets * a.fpp, 96000.0
q28 = a.arith == "q28"
if q28:
    P, bq = W.chain_config3_q28(N, fs=fs)
    eng = api.ChainEngineQ28(N, max_frames=F)
else:
    P, bq = W.chain_config3(N, fs=fs, seed=1)
    eng = api.ChainEngine(a.arith, N, max_frames=F)
n_out = 5 if q28 else 9
if a.no_sub:
    P["matrix"]["outputs"]["enabled"][:, n_out - 1] = 0
eng.set_params(P)
eng.upload_biquads(bq)
pcm = torch.randint(0, 256, (N, F * 6), dtype=torch.uint8, device="cuda")
spdif = torch.empty((N, 2 if q28 else 4, F, 2), dtype=torch.int32, device="cuda")
pdm = torch.empty((N, F, 8), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
eng.process_device(pcm.data_ptr(), 24, a.packets, a.fpp, spdif.data_ptr(), pdm.data_ptr())
eng.sync()
if q28:                                                   # no stream accessor on the Q28 chain: host clock around synchronised calls
    import time
    t0 = time.perf_counter()
    for _ in range(a.reps):
        eng.process_device(pcm.data_ptr(), 24, a.packets, a.fpp, spdif.data_ptr(), pdm.data_ptr())
    eng.sync()
    ms = (time.perf_counter() - t0) * 1e3 / a.reps
else:
    st = torch.cuda.ExternalStream(eng.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.reps):
        eng.process_device(pcm.data_ptr(), 24, a.packets, a.fpp, spdif.data_ptr(), pdm.data_ptr())
    e1.record(st)
    eng.sync()
    ms = e0.elapsed_time(e1) / a.reps
print(json.dumps({"instances": N, "frames": F, "sm_partition": eng.sm_partition(), "ms_per_step": ms, "instance_frames_per_s": N * F / (ms * 1e-3),
                  "arith": a.arith, "output_channel_samples_per_s": N * n_out * F / (ms * 1e-3), "realtime_factor": (F / fs) / (ms * 1e-3)}))
