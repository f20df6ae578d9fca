#!/bin/bash
# Round-2 ncu evidence: launch list of the default bench command, one full capture of K1 (headline), and full captures of
# the chain's modulator / front / output stage kernels (partition off: ncu cannot attach to green-context launches).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_A_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_A_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eq_f32 -s 3 -c 1 -f -o gpurun_out/r2_A \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_A_full.log 2>&1
for k in chain_pdm_kernel chain_pre_kernel chain_outpost_kernel; do
DSPI_PDM_SMS=0 ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -f -o gpurun_out/r2_${k} \
    python scripts/chain_bench.py --packets 64 --reps 1 > gpurun_out/r2_${k}.log 2>&1
done
DSPI_PDM_SMS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_chain_f32f_launches.csv \
    python scripts/chain_bench.py --packets 64 --reps 1 > gpurun_out/r2_chain_f32f_launches.log 2>&1
ls -la gpurun_out/r2_*ncu-rep gpurun_out/r2_*launches.csv
