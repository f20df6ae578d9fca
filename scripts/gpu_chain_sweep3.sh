#!/bin/bash
# modulator micro-benchmark (integer vs fp32 formulation) and the number of equal slices of a chain call
mkdir -p gpurun_out
./scripts/_bin/pdm_ubench > gpurun_out/r2_pdm_ubench2.txt 2>&1
grep -E "warps/SM  4|warps/SM  8" gpurun_out/r2_pdm_ubench2.txt
OUT=gpurun_out/r2_chain_sweep3.txt
: > $OUT
for pk in 16 64; do
  for uni in 0 8 12 16; do
    echo -n "packets=$pk DSPI_UNIFORM_SLICES=$uni f32f: " >> $OUT
    DSPI_UNIFORM_SLICES=$uni python scripts/chain_bench.py --packets $pk --arith f32f --reps 4 2>&1 | tail -1 | cut -c1-150 >> $OUT
  done
done
cat $OUT
