#!/bin/bash
# On the GPU box (1 GPU): quick timing of the default library and the phase-clock build under both policies, then one
# ncu --set full capture of the step kernel in the steady state of each policy.  TAG names the files under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-profile}
for a in uniform forward; do
  ACTIONS=$a python tools/quick_bench.py 2>&1 | tail -1
  PGDRIVE_B200_LIB=pgdrive_b200/csrc/libvar_clk.so ACTIONS=$a python tools/quick_bench.py 2>&1 | tail -1
done | tee gpurun_out/${TAG}_quick.log
WARM=2048 STEPS=30 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_kernel \
  -s 2078 -c 1 -f -o gpurun_out/prof_${TAG}_uniform python tools/quick_bench.py > gpurun_out/${TAG}_ncu.log 2>&1
WARM=512 STEPS=30 ACTIONS=forward timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_kernel \
  -s 542 -c 1 -f -o gpurun_out/prof_${TAG}_forward python tools/quick_bench.py >> gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
