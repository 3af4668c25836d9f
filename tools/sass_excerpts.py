#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of the shipped library, the counts of the mnemonics that prove which
hardware path it uses, plus the first occurrences in context.  Usage: python tools/sass_excerpts.py [lib.so] > out.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vae-gam_b200", "vaegam", "libvaegam_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "USETMAXREG", "LDGSTS", "HMMA", "LDSM", "SYNCS", "UTCBAR"]
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : \S+\n", "\n" + sass)[1:]
print("SASS evidence, libvaegam_sm100.so (cuobjdump -sass; final round-2 build; tools/sass_excerpts.py).  Counts per kernel:")
print("  UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,")
print("  USETMAXREG = setmaxnreg, LDGSTS = cp.async, HMMA = mma.sync, LDSM = ldmatrix, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit\n")
shown = set()
for name, body in zip(names, blocks):
    if not re.search(r"tc2_kernel|tc_gather_kernel|wgrad_mma_kernel|recon_(fwd|bwd)_kernel", name):
        continue
    c = collections.Counter()
    for line in body.split("\n"):
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in KEYS:
                c[op] += 1
    print(name.replace("vg::", "")[:150])
    print("    " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    base = re.sub(r"<.*", "", name)
    for k in ("UTMALDG", "UTCHMMA", "USETMAXREG"):
        if c[k] and (base, k) not in shown:
            shown.add((base, k))
            ln = next(l for l in body.split("\n") if re.search(r"\b" + k, l))
            print("      e.g. " + ln.strip()[:140])
