#!/usr/bin/env python
"""K2 (Q28 cascade) throughput on one GPU, optionally with a fraction of the (channel, band) filters flat, i.e. bypassed
(dsp_process_rp2040.S:246-248) — the mixed-bypass shape a batch of unrelated instances has.
    python scripts/k2_bench.py --channels 32768 --frames 6144 --bypass-frac 0.3"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dspi_b200 import api, workloads as W          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=32768)
ap.add_argument("--frames", type=int, default=6144)
ap.add_argument("--bypass-frac", type=float, default=0.0)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
FS = 96000.0
p = W.eq_params_fast("A", a.channels, fs=FS, seed=4)
if a.bypass_frac > 0:
    flat = np.random.default_rng(7).random((a.channels, 10)) < a.bypass_frac
    g = p["gain_db"][:, :10]
    g[flat] = 0.0
    p["gain_db"][:, :10] = g
bq = api.compute_coefficients(p, q28=True, fs=FS)
eng = api.EqEngine("q28", a.channels)
eng.upload(bq)
bufs = [torch.randint(-2**27, 2**27, (a.channels, a.frames), dtype=torch.int32, device="cuda") for _ in range(3)]
st = torch.cuda.ExternalStream(eng.stream)
for i in range(3):
    eng.process_device(bufs[i].data_ptr(), a.frames, a.frames)
eng.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(a.steps):
    eng.process_device(bufs[i % 3].data_ptr(), a.frames, a.frames)
e1.record(st)
eng.sync()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"channels": a.channels, "frames": a.frames, "bypass_frac": a.bypass_frac, "bypassed": float(np.mean(bq["bypass"][:, :10] != 0)),
                  "ms": ms, "G_samples_s": a.channels * a.frames / (ms * 1e-3) / 1e9, "plain_env": os.environ.get("DSPI_K2_PLAIN", "1")}))
eng.close()
