"""Executed warp instructions / stall samples / shared-memory wavefronts per CUDA source line of one kernel:
   python tools/line_hist.py report.ncu-rep regex:k_name [min_pct]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = defaultdict(lambda: [0, 0, 0, 0, ""])   # (file, line) -> inst, samples, wavefronts, ideal, text
fname, hdr, seen_fn = "", None, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        seen_fn += 1
        continue
    if r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr is None or len(r) <= ie:
        continue
    if not r[0].isdigit():
        continue     # SASS rows under a source line: the line row already holds their sum
    try:
        e = int(r[ie] or 0)
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg[key]
    def num(x):
        return int(x) if x.isdigit() else 0
    a[0] += e
    a[1] += num(r[isamp])
    a[2] += num(r[iw])
    a[3] += num(r[iwi])
    a[4] = r[1]
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
totw = sum(a[2] for a in agg.values())
print("total warp-inst %d  samples %d  smem wavefronts %d" % (tot, tots, totw))
for (f, ln), a in sorted(agg.items()):
    if a[0] >= tot * min_pct / 100 or a[1] >= tots * min_pct / 100:
        print("%5.1f%% i %5.1f%% s  wf %9d/%9d  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[1] / max(tots, 1), a[2], a[3], f, ln,
                                                               a[4].strip()[:110]))
