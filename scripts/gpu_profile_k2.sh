#!/bin/bash
# ncu evidence for K2 after the straight-line blocks (config 4: 32768 channels x 6144), and the per-stage launch list of the
# Q28 chain (partition off: ncu cannot attach to green-context launches).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:eq_q28 -s 3 -c 1 -f -o gpurun_out/r2_q28 \
    python bench.py --arith q28 --channels 32768 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_q28_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_q28_launches.csv \
    python bench.py --arith q28 --channels 32768 --steps 4 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_q28_launches.log 2>&1
DSPI_PDM_SMS=0 bash scripts/gpu_chain_launches.sh r2_chainq q28
ls -la gpurun_out/r2_q28* gpurun_out/r2_chainq*
