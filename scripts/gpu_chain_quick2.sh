#!/bin/bash
# chain parity tests, then config 3 on both chains
mkdir -p gpurun_out; rm -f gpurun_out/chain_quick2.txt
python -m pytest tests/test_chain_gpu.py tests/test_chain_ref_gpu.py tests/test_dynamics_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/chain_quick2_tests.txt
for ar in f32f q28 f32f q28; do
  python scripts/chain_bench.py --packets 64 --reps 3 --arith $ar 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/chain_quick2.txt
done
