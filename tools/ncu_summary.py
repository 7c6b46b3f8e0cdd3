"""Summarise an ncu report (captured with --set full --import-source on) into profiles/<name>.md + .json.

    python tools/ncu_summary.py gpurun_out/prof_r1a.ncu-rep profiles/r01_step_kernel_a
"""
import collections
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]


def ncu(rep, *args):
    out = subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    return out.stdout


def main(rep, dest):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    metrics = {}
    for i, h in enumerate(hdr):
        if h in WANT:
            metrics[h] = dict(unit=units[i], per_launch=[r[i] for r in rows[2:]])
    kernels = [r[name_col] for r in rows[2:]]
    # per source line aggregation
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    lines = collections.defaultdict(lambda: [0, 0, 0, ""])
    tot = [0, 0]
    if len(src) > 3:
        h2 = src[2]
        i_inst, i_samp, i_thr = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index(
            "Thread Instructions Executed")
        for r in src[3:]:
            try:
                ln, inst, smp, thr = int(r[0]), int(r[i_inst]), int(r[i_samp]), int(r[i_thr])
            except (ValueError, IndexError):
                continue
            a = lines[ln]
            a[0] += inst
            a[1] += smp
            a[2] += thr
            a[3] = r[1].strip()
            tot[0] += inst
            tot[1] += smp
    top = sorted(lines.items(), key=lambda kv: -kv[1][1])[:25]
    summary = dict(report=rep, kernels=kernels, metrics=metrics,
                   top_lines=[dict(line=ln, inst_pct=100.0 * a[0] / max(tot[0], 1), stall_sample_pct=100.0 * a[1] / max(tot[1], 1),
                                   threads_per_inst=a[2] / max(a[0], 1), source=a[3][:100]) for ln, a in top])
    with open(dest + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(dest + ".md", "w") as f:
        f.write("# ncu summary: %s\n\nkernels: %s\n\n| metric | unit | per launch |\n|---|---|---|\n" % (rep, sorted(set(kernels))))
        for k in WANT:
            if k in metrics:
                f.write("| %s | %s | %s |\n" % (k, metrics[k]["unit"], ", ".join(metrics[k]["per_launch"])))
        f.write("\n## hottest source lines (by warp-stall samples)\n\n| line | inst % | samples % | threads/inst | source |\n|---|---|---|---|---|\n")
        for t in summary["top_lines"]:
            f.write("| %d | %.1f | %.1f | %.1f | `%s` |\n" % (t["line"], t["inst_pct"], t["stall_sample_pct"],
                                                            t["threads_per_inst"], t["source"].replace("|", "\\|")))
    print("wrote", dest + ".md")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
