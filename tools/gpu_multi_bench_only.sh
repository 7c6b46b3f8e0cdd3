#!/bin/bash
# On a multi-GPU box: bench.py at N ranks with the given gather modes (no separate peer check: bench.py's own
# gather_check compares the consumer's checksum with the producers').  N, TAG, MODES, STEPS as in gpu_multi_call.sh.
mkdir -p gpurun_out
N=${N:-2}; TAG=${TAG:-multi}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for mode in ${MODES:-auto}; do
  NCCL_DEBUG=INFO timeout 900 $RUN bench.py --gpus $N --steps ${STEPS:-128} --warmup 8 --gather $mode \
    > gpurun_out/${TAG}_bench_${N}gpu_${mode}.json 2> gpurun_out/${TAG}_bench_${N}gpu_${mode}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${N}gpu_${mode}.json"))
    print("$mode N=$N: value %.1f M  ms/step %.4f  sim_only %.1f M  e2e %.1f M  gather_check %s  [%s]" % (
        d["value"] / 1e6, d["ms_per_step"], d["sim_only"]["value"] / 1e6, d["e2e"]["value"] / 1e6, d["gather_check"]["ok"],
        d["details"]["collective"][:60]))
except Exception as e:
    print("$mode N=$N failed:", e)
PY
  tail -2 gpurun_out/${TAG}_bench_${N}gpu_${mode}.err | cut -c1-300
done
