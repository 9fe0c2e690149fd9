"""Opcode histogram (executed warp instructions, stall samples) of one kernel from an ncu report:
   python tools/sass_hist.py report.ncu-rep regex:k_name [top]"""
import csv
import subprocess
import sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may match: use the first block
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[start]
ia, isrc, ie, iss = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[start + 1:end]:
    if len(r) <= ie:
        continue
    try:
        data.append((r[ia], r[isrc], int(r[ie] or 0), int(r[iss] or 0)))
    except ValueError:
        pass
tot = sum(d[2] for d in data)
print("total warp-inst", tot, "sass lines", len(data))
c, s = Counter(), Counter()
for a, src, e, sm in data:
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    c[op] += e
    s[op] += sm
for op, e in c.most_common(top):
    print("%-10s %12d %5.1f%%  samples %d" % (op, e, 100.0 * e / max(tot, 1), s[op]))
