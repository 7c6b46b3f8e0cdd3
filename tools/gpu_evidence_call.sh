#!/bin/bash
# On the GPU box (1 GPU): the whole -m gpu suite, one bench.py line, the ncu launch list of a short bench.py run and
# --set full captures of the step kernel under both policies.  TAG names the files under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-evidence}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_gpu_tests.log
tail -6 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
tail -c 1500 gpurun_out/${TAG}_bench_1gpu.json; tail -3 gpurun_out/${TAG}_bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_launch_run.log 2>&1
WARM=2048 STEPS=30 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_kernel \
  -s 2078 -c 1 -f -o gpurun_out/prof_${TAG}_uniform python tools/quick_bench.py > gpurun_out/${TAG}_ncu.log 2>&1
WARM=512 STEPS=30 ACTIONS=forward timeout 300 ncu --set full --clock-control none --import-source on -k regex:pgd_step_kernel \
  -s 542 -c 1 -f -o gpurun_out/prof_${TAG}_forward python tools/quick_bench.py >> gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
