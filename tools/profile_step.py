#!/usr/bin/env python
"""Runs a few whole training steps (forward + backward + fused Adam) of the native path on one GPU,
for use under `ncu` (tools/ is not product code).  Usage: python tools/profile_step.py [steps] [batch]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets sys.path for the package)
import torch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
from vaegam import dp  # noqa: E402

device = torch.device("cuda", 0)
model = bench.build_model(tempfile.mkdtemp(prefix="prof_"))
coh, vols, covs, sidx = bench.make_cohort_tensors(0, device)
reducer = dp.GradientAllReduce(model._flat, model.optimizer)
for i in range(steps):
    idx = torch.arange(i * B, (i + 1) * B, device=device) % vols.shape[0]
    dp.train_step(model, reducer, sidx[idx], covs[idx], vols[idx])
torch.cuda.synchronize()
model.check_status()
print("ok")
