#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per launch with the metrics the roofline needs."""
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[0]
units = rows[1]
data = rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def find(sub):
    return [n for n in hdr if sub in n]


want = [
    ("dur_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smemKB", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("blk", "launch__block_size"),
    ("waves", "launch__waves_per_multiprocessor"),
    ("ipc", "sm__inst_executed.avg.per_cycle_active"),
]


def conv(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    return x


def scale(name, x, u):
    if not isinstance(x, float):
        return x
    if name == "dur_us":
        return x / 1e3 if u in ("ns", "nsecond") else (x * 1e3 if u in ("ms", "msecond") else x)
    if name.endswith("_MB"):
        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        return x * f
    if name == "smemKB":
        f = {"byte": 1e-3, "Kbyte": 1.0}.get(u, 1e-3)
        return x * f
    return x


print("id  " + " ".join(f"{n:>10}" for n, _ in want) + "  kernel")
for r in data:
    vals = []
    for n, m in want:
        if m in col:
            vals.append(scale(n, conv(r[col[m]], units[col[m]]), units[col[m]]))
        else:
            vals.append("-")
    name = r[col["Kernel Name"]][:70]
    print(f"{r[col['ID']]:>3} " + " ".join(f"{v:10.2f}" if isinstance(v, float) else f"{v:>10}" for v in vals) + "  " + name)
