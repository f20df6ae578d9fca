#!/bin/bash
# config 3: slice plan (uniform vs short-ended) x modulator partition
mkdir -p gpurun_out
OUT=gpurun_out/r2_chain_sweep2.txt
: > $OUT
for pk in 16 64; do
  for uni in 1 0; do
    for sms in 64 72; do
      for ar in f32f q28; do
        echo -n "packets=$pk uniform_slices=$uni DSPI_PDM_SMS=$sms $ar: " >> $OUT
        DSPI_UNIFORM_SLICES=$uni DSPI_PDM_SMS=$sms python scripts/chain_bench.py --packets $pk --arith $ar --reps 4 2>&1 | tail -1 | cut -c1-200 >> $OUT
      done
    done
  done
done
cat $OUT
