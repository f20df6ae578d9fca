#!/bin/bash
# e2e (pinned host -> H2D -> K1 -> D2H through dspi_eq_process_host) against the staging chunk size
mkdir -p gpurun_out; rm -f gpurun_out/e2e_chunks.txt
python -m pytest tests/test_eq_gpu.py tests/test_state_gpu.py tests/test_coeff_gpu.py -m gpu -x -q -k "not full_size" 2>&1 | tail -2 | tee gpurun_out/e2e_chunks_tests.txt
for mb in ${CHUNKS:-96 64 48 32 24 16}; do
  echo -n "DSPI_HOST_CHUNK_MB=$mb: " | tee -a gpurun_out/e2e_chunks.txt
  DSPI_HOST_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null | tail -1 | \
    python -c 'import sys, json; d = json.loads(sys.stdin.read()); print(d["e2e"]["value"] / 1e9, "G samples/s e2e;", d["value"] / 1e9, "G device-resident")' | tee -a gpurun_out/e2e_chunks.txt
done
