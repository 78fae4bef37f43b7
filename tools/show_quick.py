#!/usr/bin/env python
"""Print the essentials of the logs tools/r2_quick.sh leaves in gpurun_out/."""
import csv, collections, glob, json, os, sys
d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
for f in ("pytest.log", "smoke.log"):
    p = os.path.join(d, f)
    if os.path.isfile(p):
        print(f, "|", " | ".join(l.strip() for l in open(p).read().strip().splitlines()[-3:])[:300])
for p in sorted(glob.glob(os.path.join(d, "bench_*.log")) + glob.glob(os.path.join(d, "exp_*.log")) + glob.glob(os.path.join(d, "ab_*.log"))):
    for l in open(p):
        if l.startswith("{"):
            j = json.loads(l); r = j.get("roofline") or {}; k = r.get("kernel") or {}
            e = j.get("e2e") or {}
            print(f"{os.path.basename(p):34s} {j['value']:10.0f} {j['unit']:9s} ms/step {j['ms_per_step']:.4f} frac {r.get('frac', 0):.3f} "
                  f"k1 {k.get('ms') or 0:.4f} launches {j.get('gpu_launches')} e2e {e.get('value')} clk {(j.get('clocks') or {}).get('sm_mhz')} {(j.get('clocks') or {}).get('reasons')}")
        elif "Error" in l or "error" in l:
            print(os.path.basename(p), l.strip()[:200])
p = os.path.join(d, "launches_custom.csv")
if os.path.isfile(p):
    rows = list(csv.reader(open(p))); hdr = None; out = collections.OrderedDict()
    for r in rows:
        if len(r) > 10 and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r)); out.setdefault((rec["ID"], rec["Kernel Name"][:48], rec["Grid Size"]), {})[rec["Metric Name"]] = rec["Metric Value"]
    for k, v in out.items():
        print(k[1], k[2], "us", float(v.get("gpu__time_duration.sum", 0)) / 1e3, "rdMB", round(float(v.get("dram__bytes_read.sum", 0)) / 1e6, 1),
              "wrMB", round(float(v.get("dram__bytes_write.sum", 0)) / 1e6, 1), "Minst", round(float(v.get("smsp__inst_executed.sum", 0)) / 1e6, 2), "ipc", v.get("sm__inst_issued.avg.per_cycle_active"))
