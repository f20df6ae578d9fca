#!/bin/bash
# run-time specialised K1: parity tests, then B / A throughput with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eq_gpu.py -x -q > gpurun_out/jit_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/jit_tests.log
tail -5 gpurun_out/jit_tests.log
for cfg in "B f32f" "B f32s" "A f32f"; do
  set -- $cfg
  timeout 300 python bench.py --variant $1 --arith $2 --no-cpu --no-extras --no-e2e --steps 10 --warmup 3 > gpurun_out/jit_bench_$1_$2.json 2> gpurun_out/jit_bench_$1_$2.err
  tail -c 600 gpurun_out/jit_bench_$1_$2.json; echo
done
DSPI_JIT=0 timeout 300 python bench.py --variant B --arith f32f --no-cpu --no-extras --no-e2e --steps 10 --warmup 3 > gpurun_out/jit_bench_B_f32f_nojit.json 2>&1
tail -c 600 gpurun_out/jit_bench_B_f32f_nojit.json
