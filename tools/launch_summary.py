#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/launch_summary.py <launches.csv> [steps] [top]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
n = 0
for row in r:
    v = float(row[mi].replace(',', ''))
    u = row[ui]
    v = v / 1000.0 if u == 'ns' else (v * 1000.0 if u == 'ms' else v)
    name = re.sub(r'\(.*', '', row[ki]).replace('void ', '')[:64]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v); n += 1
tot = sum(a[1] for a in agg.values())
print(f'total {tot / 1000:.3f} ms over {n} launches = {tot / steps / 1000:.3f} ms/step (cold-cache, serialised)')
print('  ms/step  share  n/step   max us  kernel')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{a[1] / steps / 1000:9.4f} {100 * a[1] / tot:5.1f}% {a[0] / steps:7.1f} {a[2]:8.1f}  {k}')
