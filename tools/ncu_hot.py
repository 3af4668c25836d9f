#!/usr/bin/env python
"""Top stall locations of an `ncu --page source --csv` export (SASS view): python tools/ncu_hot.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isrc, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
data = [(int(r[ismp]), i, r[isrc]) for i, r in enumerate(rows[2:]) if len(r) > ismp and r[ismp].isdigit()]
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, i, src in sorted(data, reverse=True)[:n]:
    # context: previous instruction lines to identify the region
    print(f"{s:7d} {100.0*s/tot:5.1f}%  line {i:5d}  {src.strip()}")
