#!/bin/bash
# On a multi-GPU box: (optionally) the bit-exact gather check, then bench.py --headline-only for a list of
# "gather-mode[:extra bench flags]" specs, e.g. SPECS="copy:--balance_equal copy sparse:--rank0-envs_40960".
# ("_" stands for a space inside a spec.)  N = GPUs, TAG names the files under gpurun_out/.
mkdir -p gpurun_out
N=${N:-2}; TAG=${TAG:-sweep}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ -n "${CHECK:-}" ]; then
  timeout 600 $RUN tools/mgpu_peer_check.py 2>&1 | grep -i "gather\|error\|Traceback" | tee gpurun_out/${TAG}_peer_check_${N}gpu.log
fi
for spec in ${SPECS:-auto}; do
  mode=${spec%%:*}; extra=""; [ "$spec" != "$mode" ] && extra=$(echo "${spec#*:}" | tr '_' ' ')
  timeout 600 $RUN bench.py --gpus $N --steps ${STEPS:-128} --warmup 8 --gather $mode --headline-only $extra \
    2> gpurun_out/${TAG}_${N}gpu_last.err | tee -a gpurun_out/${TAG}_headline_${N}gpu.jsonl
  grep -i "error\|Traceback" -A5 gpurun_out/${TAG}_${N}gpu_last.err | tail -12
done
