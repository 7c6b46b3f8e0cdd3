"""Attribute executed instructions / stall samples of an ncu report to the kernel's phases (by '// ---- phase' markers)."""
import collections, csv, io, re, subprocess, sys
rep, cu = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "pgdrive_b200/csrc/pgd_step.cu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[2]
i_inst, i_samp, i_thr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
# phase map from the '// ---- phase X' markers of the source file (must be the profiled revision)
marks = []
first = None
for n, line in enumerate(open(cu), 1):
    m = re.search(r"// ---- (phase \w)", line)
    if m:
        marks.append((n, m.group(1)))
    if first is None and "pgd_step_kernel(DevTables" in line:
        first = n
marks.sort()
def phase_of(ln):
    if not marks or ln < first:
        return "helpers"
    cur = "prologue"
    for l0, name in marks:
        if ln >= l0:
            cur = name
    return cur
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
stalls = collections.defaultdict(collections.Counter)
cur = None
for r in rows[3:]:
    if not r:
        continue
    if r[0] == "Line No":
        break
    if r[0].isdigit():
        cur = int(r[0])
        continue
    if cur is None or len(r) <= i_thr or not r[2].startswith("0x"):
        continue
    ph = phase_of(cur)
    a = agg[ph]
    a[0] += int(r[i_inst]); a[1] += int(r[i_samp]); a[2] += int(r[i_thr]); a[3] += 1
    for i, h in stall_cols:
        stalls[ph][h] += int(r[i])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("%-10s %8s %8s %8s %6s  top stalls" % ("phase", "inst%", "samples%", "thr/inst", "sass"))
for ph, a in sorted(agg.items()):
    top = ", ".join("%s %.0f%%" % (h.replace("stall_", ""), 100.0 * v / max(1, sum(stalls[ph].values()))) for h, v in stalls[ph].most_common(3))
    print("%-10s %8.1f %8.1f %8.1f %6d  %s" % (ph, 100.0 * a[0] / ti, 100.0 * a[1] / ts, a[2] / max(a[0], 1), a[3], top))
