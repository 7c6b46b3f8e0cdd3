#!/bin/bash
# First GPU call for the one-thread-per-environment layout: parity tests, A/B timing against
# the cooperative kernel on both policies, one ncu capture, one bench line.  Run on the GPU box:
#   gpurun --timeout 900 -- 'bash tools/v2_first_gpu_call.sh'
# Everything lands in gpurun_out/v2_*.
mkdir -p gpurun_out
export PGDRIVE_B200_TEST_V2=1
timeout 300 python -m pytest tests/test_gpu_step_v2.py -m gpu -x -q -s > gpurun_out/v2_tests.log 2>&1
tail -15 gpurun_out/v2_tests.log
for layout in 0 1; do
  for actions in uniform forward; do
    LAYOUT=$layout ACTIONS=$actions timeout 120 python tools/quick_bench.py 2>&1 | tail -1
  done
done | tee gpurun_out/v2_ab.log
LAYOUT=1 STEPS=30 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_v2_kernel \
  -s 60 -c 1 -f -o gpurun_out/prof_v2_step python tools/quick_bench.py > gpurun_out/v2_ncu.log 2>&1
timeout 300 python bench.py --layout per_env --no-cpu > gpurun_out/v2_bench.json 2> gpurun_out/v2_bench.err
tail -c 600 gpurun_out/v2_bench.json
