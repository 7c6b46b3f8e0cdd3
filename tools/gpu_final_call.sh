#!/bin/bash
# On the GPU box (1 GPU): the whole -m gpu suite, smoke(), one bench.py line, and the ncu launch list of a short bench.py
# run.  TAG names the files under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-final}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_1gpu.json"))
print("value %.1f M  kernel %.4f ms  frac %.3f  driving %.1f M  e2e %.1f M (d2h %d B)  cpu %.2f M  clocks %s" % (
    d["value"] / 1e6, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["driving"]["value"] / 1e6,
    d["e2e"]["value"] / 1e6, d["e2e"]["d2h_bytes_per_step"], (d["cpu_baseline"] or {}).get("value", 0) / 1e6, d["clocks"]))
PY
tail -2 gpurun_out/${TAG}_bench_1gpu.err
PGDRIVE_B200_BENCH_PREROLL=32 timeout ${NCU_TIMEOUT:-110} ncu --metrics gpu__time_duration.sum --clock-control none \
  -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_launch_run.log 2>&1
tail -1 gpurun_out/${TAG}_launch_run.log | cut -c1-200
