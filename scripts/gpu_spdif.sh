#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spdif_gpu.py -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/spdif_bench.json 2> gpurun_out/spdif_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/spdif_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d["other_configs"].get("spdif_encode_32768streams"))); print(d["other_configs"].get("error"))
PY
ncu --set full --clock-control none --import-source on -k regex:spdif_encode_kernel -s 3 -c 1 -f -o gpurun_out/r1_spdif \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1_spdif_full.log 2>&1
ls -la gpurun_out/r1_spdif*
