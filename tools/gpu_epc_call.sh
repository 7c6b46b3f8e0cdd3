#!/bin/bash
# On the GPU box: environments per CTA sweep (PGDRIVE_B200_ENVS_PER_CTA overrides the launcher's choice), then the
# phase-clock build.
mkdir -p gpurun_out
TAG=${TAG:-epc}
for a in uniform forward; do
  echo "auto: $(ACTIONS=$a python tools/quick_bench.py 2>&1 | tail -1)"
  for e in 32 30 28 26 24 20 16; do
    echo "envs per CTA $e: $(PGDRIVE_B200_ENVS_PER_CTA=$e ACTIONS=$a python tools/quick_bench.py 2>&1 | tail -1)"
  done
  PGDRIVE_B200_LIB=pgdrive_b200/csrc/libvar_clk.so ACTIONS=$a python tools/quick_bench.py 2>&1 | tail -1
done | tee gpurun_out/${TAG}_sweep.log
