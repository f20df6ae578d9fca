#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on the cascade + chain tests,
# racecheck on the kernels with shared-memory rings (K1/K2 TMA ring, chain stages).  Summaries -> gpurun_out/.
mkdir -p gpurun_out
SEL='not full_size and not specialised and not kernel_choice'
( time timeout 1500 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 \
    python -m pytest tests/test_eq_gpu.py tests/test_chain_gpu.py tests/test_chainq_gpu.py tests/test_spdif_gpu.py tests/test_state_gpu.py tests/test_chain_ref_gpu.py tests/test_dynamics_gpu.py tests/test_eqx_gpu.py -x -q -m gpu -k "$SEL and not config3 and not quirks" ) > gpurun_out/san_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/san_memcheck.log | tail -8
( time timeout 1500 compute-sanitizer --tool racecheck --target-processes all --print-limit 20 \
    python -m pytest tests/test_eq_gpu.py tests/test_chain_gpu.py tests/test_chainq_gpu.py tests/test_chain_ref_gpu.py tests/test_dynamics_gpu.py -x -q -m gpu -k "$SEL and (matches_oracle or bit_exact or ragged or envelope or dynamics or committed_reference) and not 16" ) > gpurun_out/san_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|error" gpurun_out/san_racecheck.log | tail -8
