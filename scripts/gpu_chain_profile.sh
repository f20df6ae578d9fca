#!/bin/bash
# chain (config 3) per-kernel breakdown: launch list + full captures of one front / out / pdm launch
TAG=${1:-chain}
mkdir -p gpurun_out
python scripts/chain_bench.py --packets 64 --reps 2 > gpurun_out/${TAG}_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/chain_bench.py --packets 64 --reps 1 > gpurun_out/${TAG}_launches.log 2>&1
for k in front out pdm; do
ncu --set full --clock-control none --import-source on -k regex:chain_${k}_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_${k} \
    python scripts/chain_bench.py --packets 64 --reps 1 > gpurun_out/${TAG}_${k}_full.log 2>&1
done
cat gpurun_out/${TAG}_bench.log
ls -la gpurun_out/${TAG}*
