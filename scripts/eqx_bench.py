#!/usr/bin/env python
"""One process, several GPUs (dspi_eqx_*): a block resident on device 0 is pulled / processed / pushed back by the peers over
NVLink in staged chunks.  Checks the result against one engine over all channels at full size, then times it.
    python scripts/eqx_bench.py --devices 2 --channels-per-gpu 65536"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dspi_b200 import api, workloads as W          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=2)
ap.add_argument("--channels-per-gpu", type=int, default=65536)
ap.add_argument("--frames", type=int, default=6144)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
FS, n_dev, T = 96000.0, a.devices, a.frames
Cn = a.channels_per_gpu * n_dev
bq = api.compute_coefficients(W.eq_params_fast("A", Cn, fs=FS, seed=2), q28=False, fs=FS)
torch.cuda.set_device(0)
x = torch.rand((Cn, T), dtype=torch.float32, device="cuda") - 0.5
ref = x.clone()
one = api.EqEngine("f32f", Cn)
one.upload(bq)
one.process_device(ref.data_ptr(), T, T)
one.sync()
one.close()
grp = api.EqGroup("f32f", Cn, list(range(n_dev)))
grp.upload(bq)
y = x.clone()
grp.process_root(y.data_ptr(), T)
torch.cuda.synchronize()
same = bool(torch.equal(y.view(torch.int32), ref.view(torch.int32)))
for _ in range(2):
    grp.process_root(y.data_ptr(), T)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    grp.process_root(y.data_ptr(), T)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / a.steps
print(json.dumps({"devices": n_dev, "channels": Cn, "frames": T, "identical_to_single_engine": same, "ms_per_step": ms,
                  "G_samples_s": Cn * T / (ms * 1e-3) / 1e9, "bytes_over_nvlink_each_way": (Cn - a.channels_per_gpu) * T * 4,
                  "chunk_env": os.environ.get("DSPI_HOST_CHUNK_MB", "default")}))
grp.close()
