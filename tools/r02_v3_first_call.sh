#!/bin/bash
# r02c: first GPU call of the role-per-warp layout (layout 2): parity, A/B against layouts 0 / 1, one ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step_v3.py -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/r02c_v3_tests.log
tail -25 gpurun_out/r02c_v3_tests.log
for layout in 0 2; do
  for actions in uniform forward; do
    LAYOUT=$layout ACTIONS=$actions timeout 120 python tools/quick_bench.py 2>&1 | tail -1
  done
done | tee gpurun_out/r02c_ab.log
LAYOUT=2 STEPS=30 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_v3_kernel \
  -s 60 -c 1 -f -o gpurun_out/prof_r02c_v3_uniform python tools/quick_bench.py > gpurun_out/r02c_ncu.log 2>&1
LAYOUT=2 STEPS=30 ACTIONS=forward timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_v3_kernel \
  -s 60 -c 1 -f -o gpurun_out/prof_r02c_v3_forward python tools/quick_bench.py >> gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
