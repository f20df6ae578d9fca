#!/bin/bash
# Runs on the GPU box (under gpurun).  Usage: scripts/gpu_profile.sh <tag> [variant] [arith]
# Produces gpurun_out/<tag>_launches.csv (every launch with its device time) and
# gpurun_out/<tag>.ncu-rep (full-set capture of the cascade kernel, 1 launch).
set -u
TAG=${1:-prof}; VAR=${2:-A}; AR=${3:-f32f}
mkdir -p gpurun_out
KERN=eq_f32; [ "$AR" = "q28" ] && KERN=eq_q28_kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --variant $VAR --arith $AR --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --steps 3 --warmup 3 --variant $VAR --arith $AR --no-e2e --no-cpu > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/${TAG}* 
