#!/bin/bash
# float chain (config 3): full ncu captures of one launch of each stage kernel
TAG=${1:-chain}
mkdir -p gpurun_out
for k in chain_pre_kernel chain_post_kernel chain_mix_kernel chain_outpost_kernel; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${TAG}_${k} \
    python scripts/chain_bench.py --packets 64 --reps 1 > gpurun_out/${TAG}_${k}.log 2>&1
done
ls -la gpurun_out/${TAG}_*
