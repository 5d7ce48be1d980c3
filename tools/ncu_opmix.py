#!/usr/bin/env python
"""Executed-instruction mix of one kernel from an ncu report (source page): opcode -> warp-instructions.
usage: tools/ncu_opmix.py report.ncu-rep kernel_regex"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter()
for r in rows[2:]:
    if len(r) != len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    mix[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
tot = sum(mix.values())
print(f"total warp-instructions {tot}")
for op, n in mix.most_common(28):
    print(f"  {op:12s} {n:12d} {100.0*n/tot:5.1f}%")
