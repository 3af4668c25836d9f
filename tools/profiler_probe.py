#!/usr/bin/env python
"""What a process sees when it runs under Nsight Compute: environment keys and mapped libraries (used to keep
whole-step CUDA-graph capture off under kernel-replay profiling).  ncu -c 1 python tools/profiler_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-gam_b200"))
import torch
torch.zeros(4, device="cuda").add_(1).sum().item()
print("ENV", sorted(k for k in os.environ if any(t in k.upper() for t in ("INJECT", "NSIGHT", "PROFILER", "NSYS", "NCU"))))
libs = set()
for line in open("/proc/self/maps"):
    p = line.split()[-1]
    if any(t in p.lower() for t in ("nsight", "injection", "nvperf", "ncu", "nsys")):
        libs.add(p)
print("MAPS", sorted(libs))
from vaegam.step import _under_profiler
print("under_profiler", _under_profiler())
