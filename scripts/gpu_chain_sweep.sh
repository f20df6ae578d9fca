#!/bin/bash
# config 3 (8192 instances) against the size of the modulator's SM partition: DSPI_PDM_SMS = 0 (no partition) .. 80
mkdir -p gpurun_out
OUT=gpurun_out/r2_chain_sweep.txt
: > $OUT
for pk in 16 64; do
  for sms in 0 32 48 56 64 72 80; do
    for ar in f32f q28; do
      echo -n "packets=$pk DSPI_PDM_SMS=$sms $ar: " >> $OUT
      DSPI_PDM_SMS=$sms python scripts/chain_bench.py --packets $pk --arith $ar --reps 4 2>&1 | tail -1 >> $OUT
    done
  done
done
echo -n "no sub, packets=64 f32f: " >> $OUT; python scripts/chain_bench.py --packets 64 --no-sub --reps 4 2>&1 | tail -1 >> $OUT
cat $OUT
