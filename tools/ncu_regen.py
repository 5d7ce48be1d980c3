#!/usr/bin/env python
"""Regenerate, for the CURRENT build, the two ncu-derived inputs of bench.py's roofline block (run on a GPU box):

    python tools/ncu_regen.py [--rows 1000000 --cols 100000 --k 32] [--tag r02]

1. `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum` over
   `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity` at the given shape: DRAM bytes and executed warp
   instructions per launch of the two pass kernels -> profiles/dram_traffic.json (key rows x cols x k : engine), which
   bench.py reads for `roofline.traffic` and `roofline.issue_slots` (both are then from THIS build, not a stale capture);
2. the raw CSV is kept as profiles/<tag>_ncu_pass_kernel_counters_<shape>.csv together with the block / grid shape of the
   profiled kernels, so the provenance can be checked (round 1's capture was of a mid-round kernel).
Numbers printed by a run under ncu are never bench values."""
import argparse
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--cols", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--from-csv", default=None, help="re-derive the table from a CSV of an earlier run instead of running ncu")
    a = ap.parse_args()
    shape = f"{a.rows}x{a.cols}x{a.k}"
    out_csv = ROOT / "profiles" / f"{a.tag}_ncu_pass_kernel_counters_{a.rows}x{a.cols}_k{a.k}.csv"
    metrics = "dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum"
    cmd = ["ncu", "--metrics", metrics, "--clock-control", "none", "-k", "regex:(h|w)_pass_tc_kernel|(h|w)_pass_kernel",
           "--csv", "--log-file", str(out_csv), sys.executable, str(ROOT / "bench.py"), "--rows", str(a.rows), "--cols", str(a.cols),
           "--k", str(a.k), "--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu", "--no-parity"]
    if a.from_csv:
        out_csv = Path(a.from_csv).resolve()
    else:
        print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
    rows = [r for r in csv.reader(io.StringIO("".join(l for l in out_csv.read_text().splitlines(True) if l.startswith('"'))))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}          # kernel kind -> list of launches {metric: value}
    for r in rows[1:]:
        name = r[ix["Kernel Name"]]
        kind = "h_pass" if "h_pass" in name else "w_pass"
        key = (r[ix["ID"]], kind, name, r[ix["Block Size"]], r[ix["Grid Size"]])
        per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    agg = {}
    for (lid, kind, name, block, grid), m in per.items():
        if kind == "h_pass" and ", false>" in name:           # the loss-only instantiation (CD = false) is not the hot kernel
            continue
        agg.setdefault(kind, []).append(dict(m, block=block, grid=grid, name=name))
    entry = {}
    for kind, launches in agg.items():
        # several instantiations of a pass may be launched per iteration (the H pass launches its FLIP = false and FLIP =
        # true kernels; the one that does not match the device-side flag returns at once): keep the one that did the work
        by_name = {}
        for l in launches:
            by_name.setdefault(l["name"], []).append(l)
        launches = max(by_name.values(), key=lambda ls: sum(l["gpu__time_duration.sum"] for l in ls))
        n = len(launches)
        entry[kind + "_ms_under_ncu"] = sum(l["gpu__time_duration.sum"] for l in launches) / n * 1e-6
        entry[kind] = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches) / n
        entry[kind + "_warp_instructions"] = sum(l["smsp__inst_executed.sum"] for l in launches) / n
        entry[kind + "_kernel"] = {"name": launches[0]["name"], "block": launches[0]["block"], "grid": launches[0]["grid"], "launches": n}
    wpr = (a.cols + 1023) // 1024 * 32
    entry["h_pass_algorithmic"] = a.rows * wpr * 4                 # P plane, one pass
    entry["w_pass_algorithmic"] = 2 * a.rows * wpr * 4             # P and M planes
    entry["source"] = str(out_csv.relative_to(ROOT)) if str(out_csv).startswith(str(ROOT)) else out_csv.name
    path = ROOT / "profiles" / "dram_traffic.json"
    try:
        table = json.loads(path.read_text())
    except Exception:
        table = {}
    table["_comment"] = ("per launch of the pass kernels, regenerated per build by tools/ncu_regen.py: dram__bytes_read.sum + "
                         "dram__bytes_write.sum (h_pass, w_pass) and smsp__inst_executed.sum (*_warp_instructions); key = "
                         "rows_per_gpu x cols x k : engine")
    engine = "tensor" if "tc_kernel" in entry.get("h_pass_kernel", {}).get("name", "") else "simt"
    table[f"{shape}:{engine}"] = entry
    path.write_text(json.dumps(table, indent=1) + "\n")
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
