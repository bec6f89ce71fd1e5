#!/usr/bin/env python
"""Top SASS instructions by stall samples from an ncu report (source page).
  python tools/ncu_hot.py gpurun_out/x.ncu-rep [N]"""
import csv
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(int(r[ci[s]]) for r in body) for s in stalls}
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]]))[:top]
for i in sorted(idx):
    r = body[i]
    s = int(r[ci["# Samples"]])
    main = max(stalls, key=lambda k: int(r[ci[k]]))
    print("%5d %5.1f%%  #%-5d %-12s %s" % (s, 100.0 * s / max(tot, 1), i, main[6:], r[ci["Source"]].strip()[:110]))
