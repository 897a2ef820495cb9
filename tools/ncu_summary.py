#!/usr/bin/env python
"""Summarise ncu output for profiles/: `ncu_summary.py launches <csv>` or `ncu_summary.py full <ncu-rep>`."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        ns = float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.1f} | {100 * ns / tot:.1f}% |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in KEYS:
            if k in idx:
                print(f"- {k} = {r[idx[k]]} {units[idx[k]]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
