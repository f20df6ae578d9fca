#!/bin/bash
# config 3 (8192 instances x 64 packets) under the placements of the streaming stages (DSPI_CHAIN_PLACE, chain_streams.cuh):
# bit 0 rings slice by slice, bit 1 rings on s_post, bit 2 mix / output stage on the modulator's SMs.  Parity tests run under 7.
# The switch exists in commit 1cc5253 only (every placement other than 0 was slower, profiles/r2_chain_placement_sweep.txt); check that commit out to rerun.
mkdir -p gpurun_out; rm -f gpurun_out/chain_place.txt
DSPI_CHAIN_PLACE=7 python -m pytest tests/test_chain_gpu.py tests/test_chain_ref_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/chain_place_tests.txt
for place in 0 1 3 4 5 7; do
  for ar in f32f q28; do
    echo -n "DSPI_CHAIN_PLACE=$place $ar: " | tee -a gpurun_out/chain_place.txt
    DSPI_CHAIN_PLACE=$place python scripts/chain_bench.py --packets 64 --reps 3 --arith $ar 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/chain_place.txt
  done
done
