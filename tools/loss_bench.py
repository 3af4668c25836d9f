#!/usr/bin/env python
"""Fused reconstruction/likelihood/GLM pass alone: python tools/loss_bench.py [iters] [B ...]
Prints one line per (batch, backward CTAs/SM variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from vaegam import native  # noqa: E402
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batches = [int(a) for a in sys.argv[2:]] or [32, 128, 512]
lib = native.load()
for B in batches:
    for cps in (1, 2, 3):
        lib.vg_recon_tune(cps)
        print(B, cps, bench.fused_loss_roofline(torch.device("cuda", 0), B=B, iters=iters), flush=True)
lib.vg_recon_tune(-1)
