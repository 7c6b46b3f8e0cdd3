#!/bin/bash
# role-per-warp layout: parity, then warps-per-CTA x occupancy variants (+ the phase-clock build)
mkdir -p gpurun_out
TAG=${TAG:-r02e}
timeout 600 python -m pytest tests/test_gpu_step_v3.py -m gpu -x -q -s 2>&1 | tail -12 > gpurun_out/${TAG}_v3_tests.log
tail -12 gpurun_out/${TAG}_v3_tests.log
timeout 900 python tools/kernel_variants.py bench 2>&1 | tee gpurun_out/${TAG}_variants.log
