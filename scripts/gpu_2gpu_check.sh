#!/bin/bash
# 2 GPUs: N-vs-1 identity tests of the group engine and the NCCL pipeline, then the root-resident block at full size
mkdir -p gpurun_out; rm -f gpurun_out/eqx_2gpu.txt
python -m pytest tests/test_eqx_gpu.py tests/test_sg_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/eqx_2gpu_tests.txt
python scripts/eqx_bench.py --devices 2 2>&1 | tail -1 | tee -a gpurun_out/eqx_2gpu.txt
DSPI_HOST_CHUNK_MB=32 python scripts/eqx_bench.py --devices 2 2>&1 | tail -1 | tee -a gpurun_out/eqx_2gpu.txt
