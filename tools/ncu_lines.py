"""Top source lines of an ncu report by warp-stall samples EXCLUDING barrier waits (who makes the others wait?).
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [N]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[2]
i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
cur = None
fname = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        cur = None
        continue
    if r[0] in ("Line No", "Function Name"):
        continue
    if r[0].isdigit():
        cur = (fname, int(r[0]))
        agg[cur][3] = r[1].strip()
        continue
    if cur is None or len(r) <= i_samp or not r[2].startswith("0x"):
        continue
    a = agg[cur]
    a[0] += int(r[i_inst]); a[1] += int(r[i_samp])
    for i, h in stall_cols:
        a[2][h.replace("stall_", "")] += int(r[i])
tot = sum(a[1] for a in agg.values())
totb = sum(a[2]["barrier"] for a in agg.values())
print("total samples %d, barrier %d (%.0f%%)" % (tot, totb, 100.0 * totb / max(tot, 1)))
key = lambda kv: -(kv[1][1] - kv[1][2]["barrier"])
for ln, a in sorted(agg.items(), key=key)[:topn]:
    nb = a[1] - a[2]["barrier"]
    top = ", ".join("%s %d" % (h, v) for h, v in a[2].most_common(3) if v)
    print("%-18s %5d  %5.1f%%  inst %8d  %-60s %s" % (ln[0][-18:], ln[1], 100.0 * nb / max(tot - totb, 1), a[0], a[3][:60], top))
