#!/usr/bin/env python
"""Top stall sites of one kernel from an ncu report's source page (SASS view).
usage: tools/ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:                                    # first matching launch only (the report may hold several)
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0] != "Address":
        data.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"total samples {tot}; instructions {len(data)}")
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {s:7d} {100.0*s/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} "
          + " ".join(f"{n}:{c}" for c, n in reasons if c))
