#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/val_tests.log 2>&1; tail -4 gpurun_out/val_tests.log
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/val_smoke.log 2>&1; tail -5 gpurun_out/val_smoke.log
( time python bench.py ) > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; tail -c 3000 gpurun_out/val_bench.json; tail -3 gpurun_out/val_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/val_bench_ref.json 2> gpurun_out/val_bench_ref.err; tail -c 800 gpurun_out/val_bench_ref.json; tail -3 gpurun_out/val_bench_ref.err
