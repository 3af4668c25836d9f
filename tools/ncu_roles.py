#!/usr/bin/env python
"""Stall samples of a tc2_kernel source-page export, split by warp role (regions are found from marker instructions)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; isrc = hdr.index("Source"); ismp = hdr.index("# Samples")
d = [(i, r[isrc].strip(), int(r[ismp])) for i, r in enumerate(rows[2:]) if len(r) > ismp and r[ismp].isdigit()]
d = d[:len(d) // 2] if len(d) > 2 and d[0][1] == d[len(d) // 2][1] else d
first_ldtm = next(i for i, s, _ in d if s.startswith("LDTM"))
first_mma = next(i for i, s, _ in d if "UTCHMMA" in s)
# producer region: between the last STTM/arrive of the epilogue and the first UTCHMMA; find the first LDG after the epilogue's last LDTM
last_ldtm = max(i for i, s, _ in d if s.startswith("LDTM"))
arr = [i for i, s, _ in d if s.startswith("SYNCS.ARRIVE") and i > last_ldtm]
prod_start = arr[0] + 1 if arr else last_ldtm
# MMA region start: the try_wait preceding the first UTCHMMA by the largest index below it
waits = [i for i, s, _ in d if "TRYWAIT" in s and i < first_mma]
prod_arrive = [i for i, s, _ in d if s.startswith("SYNCS.ARRIVE") and prod_start < i < first_mma]
mma_start = (prod_arrive[-1] + 1) if prod_arrive else first_mma
def reg(a, b): return sum(s for i, _, s in d if a <= i < b)
tot = sum(s for _, _, s in d)
print(f"total {tot}: setup {reg(0, first_ldtm - 80)}  epilogue {reg(first_ldtm - 80, prod_start)}  producers {reg(prod_start, mma_start)}  mma+tail {reg(mma_start, 10**9)}")
for name, a, b in (("epilogue", first_ldtm - 80, prod_start), ("producers", prod_start, mma_start), ("mma", mma_start, 10**9)):
    print("---", name)
    for i, s, n in sorted([x for x in d if a <= x[0] < b], key=lambda x: -x[2])[:8]:
        print(f"   {n:6d}  line {i:5d}  {s[:100]}")
