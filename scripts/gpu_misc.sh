#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bulk_gpu.py -q 2>&1 | grep -E "AssertionError|passed|failed|Error" | cut -c1-1500 | head -12
timeout 900 python -m pytest tests/test_chainq_gpu.py tests/test_chain_gpu.py -x -q 2>&1 | tail -5
python scripts/chain_bench.py --packets 64 --reps 3 --arith q28 2>&1 | tail -2
python scripts/chain_bench.py --packets 64 --reps 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1k_q28_launches.csv python scripts/chain_bench.py --packets 64 --reps 1 --arith q28 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r1k_q28_launches.csv')) if len(r)>10 and r[0]!='ID']
for r in rows[-8:]: print(r[4].split('(')[0][-36:], r[8], r[-1])
PY
