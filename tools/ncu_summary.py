#!/usr/bin/env python
"""Turn ncu artefacts into the tables of profiles/*.md.

    python tools/ncu_summary.py launches <launch_list.csv>        # per-kernel launch table
    python tools/ncu_summary.py full <report.ncu-rep> [kernel]     # key metrics of one capture

The launch list comes from
    ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file X.csv <cmd>
and the report from
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o X <cmd>
(both only after the same command has exited 0 without ncu).
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v / 1e3 if unit.startswith("ns") else v * (1e3 if unit.startswith("ms") else 1.0)
        name = re.sub(r"^.*?(\w+_kernel)", r"\1", re.sub(r"\(.*", "", r[ik]))
        rows.append((name, us))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | avg us | share of GPU time |")
    print("|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / c:.1f} | {100 * t / total:.1f} % |")


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = re.sub(r"^.*?(\w+_kernel)", r"\1", re.sub(r"\(.*", "", vals[hdr.index("Kernel Name")]))
    print(f"kernel: `{kname}`\n")
    print("| metric | value |")
    print("|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {vals[i]} {units[i]} |")
    print("\nstall reasons (warps per issue slot, > 0.15):\n")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            v = float(vals[i].replace(",", "") or 0)
            if v > 0.15:
                name = h.split("issue_stalled_")[1].split("_per_issue")[0]
                print(f"* {name}: {v:.2f}")


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
