#!/bin/bash
# chain (config 3): parity tests, timings, launch list; optional full capture of one kernel ($2 = front|out|pdm)
TAG=${1:-chainq}; FULL=${2:-}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_chainq_gpu.py -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python scripts/chain_bench.py --packets 64 --reps 3 2>&1 | tee gpurun_out/${TAG}_bench.log
python scripts/chain_bench.py --packets 16 --reps 3 2>&1 | tee -a gpurun_out/${TAG}_bench.log
python scripts/chain_bench.py --packets 64 --reps 3 --arith q28 2>&1 | tee -a gpurun_out/${TAG}_bench.log
for AR in f32f q28; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_${AR}_launches.csv \
    python scripts/chain_bench.py --packets 64 --reps 1 --arith $AR > gpurun_out/${TAG}_${AR}_launches.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_${AR}_launches.csv')) if len(r)>10 and r[0]!='ID']
for r in rows[-6:]: print(r[4].split('(')[0][-36:], r[8], r[-1])
PY
done
if [ -n "$FULL" ]; then
ncu --set full --clock-control none --import-source on -k regex:chain_${FULL}_kernel -s 9 -c 1 -f -o gpurun_out/${TAG}_${FULL} \
    python scripts/chain_bench.py --packets 64 --reps 1 > gpurun_out/${TAG}_${FULL}_full.log 2>&1
fi
