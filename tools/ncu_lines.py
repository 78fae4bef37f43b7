#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples from an ncu report (needs -lineinfo):
   python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
for r in rows:
    if len(r) >= 8 and r[0].isdigit() and r[2] == '-':
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[6]), int(r[7])))
        except ValueError:
            pass
agg = collections.OrderedDict()
for ln, src, samp, inst in lines:
    a = agg.setdefault(ln, [src, 0, 0]); a[1] += samp; a[2] += inst
tot_i = sum(a[2] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
for ln, (src, samp, inst) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{ln:5d} inst {inst:11d} {100*inst/tot_i:5.1f}%  samples {100*samp/max(1,tot_s):5.1f}%  {src[:110]}")
