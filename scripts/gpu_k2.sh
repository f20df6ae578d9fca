#!/bin/bash
# K2 (Q28 cascade, config 4: 32768 channels x 6144 frames): parity tests, then throughput with and without the
# straight-line blocks (DSPI_K2_PLAIN=0 takes the per-band branches of round 1), all bands on and 30 % of them flat.
mkdir -p gpurun_out; rm -f gpurun_out/k2_bench.txt
python -m pytest tests/test_eq_gpu.py tests/test_chain_ref_gpu.py -m gpu -x -q -k "q28 or Q28" 2>&1 | tail -2 | tee gpurun_out/k2_tests.txt
for plain in 1 0; do
  for frac in 0 0.3; do
    DSPI_K2_PLAIN=$plain python scripts/k2_bench.py --bypass-frac $frac 2>&1 | tail -1 | tee -a gpurun_out/k2_bench.txt
  done
done
python scripts/k2_bench.py --channels 37888 2>&1 | tail -1 | tee -a gpurun_out/k2_bench.txt
python scripts/chain_bench.py --packets 64 --reps 3 --arith q28 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/k2_bench.txt
DSPI_K2_PLAIN=0 python scripts/chain_bench.py --packets 64 --reps 3 --arith q28 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/k2_bench.txt
