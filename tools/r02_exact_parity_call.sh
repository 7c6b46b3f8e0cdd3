#!/bin/bash
# r02: shared pgd_math.h -> bit-exact parity tests of both existing layouts + A/B timing (run on the GPU box)
mkdir -p gpurun_out
export PGDRIVE_B200_TEST_V2=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_v2.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02b_exact_tests.log
tail -30 gpurun_out/r02b_exact_tests.log
for layout in 0 1; do
  for actions in uniform forward; do
    LAYOUT=$layout ACTIONS=$actions timeout 120 python tools/quick_bench.py 2>&1 | tail -1
  done
done | tee gpurun_out/r02b_ab.log
