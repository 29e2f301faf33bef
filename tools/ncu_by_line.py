"""Aggregate an .ncu-rep's warp-stall samples and executed instructions by CUDA source line (needs -lineinfo and
--import-source on).  python tools/ncu_by_line.py report.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, ""])   # samples, inst, shared wavefronts, ideal, text
path, hdr = "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr):
        continue
    f = lambda k: float(r[ix[k]]) if r[ix[k]].replace(".", "", 1).isdigit() else 0.0
    key = (path, int(r[0]) if r[0].isdigit() else -1)
    a = agg[key]
    a[0] += f("# Samples"); a[1] += f("Instructions Executed"); a[2] += f("L1 Wavefronts Shared"); a[3] += f("L1 Wavefronts Shared Ideal")
    a[4] = r[1]
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
for (p, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / tot:5.1f}% smp {100 * a[1] / toti:5.1f}% inst  smem wf {a[2]:10.0f}/{a[3]:10.0f}  {p}:{ln}  {a[4].strip()[:90]}")
