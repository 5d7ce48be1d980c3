#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>` launch list into launches, total time
and share per kernel.   usage: tools/ncu_launch_list.py X.csv "<command>" > profiles/<name>.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    t = tot[r[ix["Kernel Name"]]]
    t[0] += 1; t[1] += us
total = sum(t[1] for t in tot.values())
print(f"launch list of `{sys.argv[2]}` (ncu --metrics gpu__time_duration.sum --clock-control none;")
print("per-launch times under ncu are cold-cache and serialised: read the SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for name, (n, us) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{name[:70]:70s} {n:8d} {us:12.1f} {100 * us / total:6.1f}%")
