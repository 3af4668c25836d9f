#!/usr/bin/env python
"""Eager steps with per-operation event timing: prints every operation whose slowest instance is far above its
median (tools/ is not product code).  Usage: python tools/anomaly_probe.py [steps]"""
import os
import sys
import tempfile

os.environ.setdefault("VAEGAM_CUDA_GRAPH", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
from vaegam import native  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
device = torch.device("cuda", 0)
model = bench.build_model(tempfile.mkdtemp(prefix="prob_"))
coh, vols, covs, sidx = bench.make_cohort_tensors(0, device)
B = bench.BATCH


def step(i):
    idx = torch.arange(i * B, (i + 1) * B, device=device) % vols.shape[0]
    loss = model.forward(sidx[idx], covs[idx], vols[idx], 'train', train_mode=False)
    model.optimizer.zero_grad()
    loss.backward()
    model.optimizer.step()


for i in range(3):
    step(i)
torch.cuda.synchronize()
native.profile(True)
for i in range(steps):
    step(i)
torch.cuda.synchronize()
rec = native.profile_collect()
native.profile(False)
for op, v in sorted(rec.items(), key=lambda kv: -max(kv[1])):
    v = np.asarray(v)
    med = float(np.median(v))
    flag = "  <-- outlier" if v.max() > 3 * med + 0.05 else ""
    print(f"{op:18s} n={len(v):4d} median {med:8.4f} max {v.max():8.4f} at {int(v.argmax())}{flag}")
