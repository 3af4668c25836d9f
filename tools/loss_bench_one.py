#!/usr/bin/env python
"""One variant of the fused loss pass for ncu: python tools/loss_bench_one.py <variant> <B> [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from vaegam import native  # noqa: E402
lib = native.load()
lib.vg_recon_tune(int(sys.argv[1]))
print(bench.fused_loss_roofline(torch.device("cuda", 0), B=int(sys.argv[2]), iters=int(sys.argv[3]) if len(sys.argv) > 3 else 2))
