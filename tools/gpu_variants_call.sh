#!/bin/bash
# On the GPU box: step-kernel parity tests, then tools/kernel_variants.py bench (default library + every libvar_*.so).
# TAG names the logs under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-variants}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
timeout 900 python tools/kernel_variants.py bench 2>&1 | tee gpurun_out/${TAG}_variants.log
