#!/bin/bash
# On a multi-GPU box: bit-exact check of both gather modes, then bench.py at N ranks with each gather mode.
# N = number of GPUs (gpurun --gpus N), TAG names the files under gpurun_out/.
mkdir -p gpurun_out
N=${N:-2}; TAG=${TAG:-multi}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tools/mgpu_peer_check.py 2>&1 | grep -i "gather" | tee gpurun_out/${TAG}_peer_check_${N}gpu.log
for mode in ${MODES:-peer copy}; do
  NCCL_DEBUG=INFO PGDRIVE_B200_LIB=${LIB:-} timeout 900 $RUN bench.py --gpus $N --steps ${STEPS:-128} --warmup 8 --gather $mode \
    > gpurun_out/${TAG}_bench_${N}gpu_${mode}.json 2> gpurun_out/${TAG}_bench_${N}gpu_${mode}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${N}gpu_${mode}.json"))
    print("$mode N=$N: value %.1f M  ms/step %.4f  sim_only %.1f M  e2e %.1f M (per rank %.1f M)  gather_check %s" % (
        d["value"] / 1e6, d["ms_per_step"], d["sim_only"]["value"] / 1e6, d["e2e"]["value"] / 1e6,
        d["e2e"]["per_rank"]["value"] / 1e6, d["gather_check"]["ok"]))
except Exception as e:
    print("$mode N=$N failed:", e)
PY
  grep -c "NCCL INFO" gpurun_out/${TAG}_bench_${N}gpu_${mode}.err | sed 's/^/NCCL INFO lines on stderr: /'
  tail -2 gpurun_out/${TAG}_bench_${N}gpu_${mode}.err | cut -c1-300
done
