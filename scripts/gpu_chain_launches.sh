#!/bin/bash
# per-kernel device time of one config-3 call (8192 instances x 64 packets) under the modulator partition: ncu launch list,
# summed per kernel.  ncu serialises the launches, so the sums are the WORK of each stage on the SMs it may use, not the
# overlapped wall time (scripts/chain_bench.py without ncu gives that).
TAG=${1:-r2_chain}; AR=${2:-f32f}
mkdir -p gpurun_out
python scripts/chain_bench.py --packets 64 --reps 3 --arith $AR 2>&1 | tail -1 | cut -c1-160 | tee gpurun_out/${TAG}_${AR}_bench.txt
ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__cycles_active.avg --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_${AR}_launches.csv \
    python scripts/chain_bench.py --packets 64 --reps 1 --arith $AR > gpurun_out/${TAG}_${AR}_launches.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/${TAG}_${AR}_launches.csv')) if len(r) > 10 and r[0] != 'ID']
t = collections.OrderedDict()
per_id = collections.defaultdict(dict)
for r in rows:
    per_id[r[0]]['name'] = r[4].split('(')[0].split('<')[0].split('::')[-1]
    if 'gpu__time_duration' in r[-3]:
        v = float(r[-1].replace(',', ''))
        per_id[r[0]]['t'] = v * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(r[-2], 1e-6)
ids = sorted(per_id, key=int)
# the second call of the script is the timed-shape one: keep the last len/2 launches
half = ids[len(ids) // 2:]
for i in half:
    d = per_id[i]
    if 't' in d:
        a = t.setdefault(d['name'], [0, 0.0]); a[0] += 1; a[1] += d['t']
tot = sum(v[1] for v in t.values())
with open('gpurun_out/${TAG}_${AR}_stage_times.txt', 'w') as f:
    for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
        line = f"{k:40s} launches {v[0]:4d}  total {v[1]:8.3f} ms  ({100 * v[1] / tot:4.1f} %)"
        print(line); f.write(line + "\n")
    f.write(f"sum {tot:.3f} ms\n"); print("sum", round(tot, 3), "ms")
PY
