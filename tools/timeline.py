#!/usr/bin/env python
"""Kernel timeline of whole-step CUDA-graph replays through torch.profiler (CUPTI): start / end / stream of every
kernel of one replay, the idle gaps, and how much of the step each stream is busy.  tools/ is not product code.
Usage: python tools/timeline.py [batch] > gpurun_out/timeline.txt"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from vaegam import dp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH
device = torch.device("cuda", 0)
model = bench.build_model(tempfile.mkdtemp(prefix="tl_"))
coh, vols, covs, sidx = bench.make_cohort_tensors(0, device)
reducer = dp.GradientAllReduce(model._flat, model.optimizer)


def step(i):
    idx = torch.arange(i * B, (i + 1) * B, device=device) % vols.shape[0]
    dp.train_step(model, reducer, sidx[idx], covs[idx], vols[idx])


for i in range(8):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(8 + i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
# keep the middle replay: between the 2nd and 3rd "adam" kernels
adam = [i for i, e in enumerate(ev) if "adam_kernel<float>" in e.name]
lo, hi = (adam[0] + 3, adam[1] + 3) if len(adam) >= 2 else (0, len(ev))
sel = ev[lo:hi]
t0 = sel[0].time_range.start
print(f"{len(sel)} device activities in one step, span {(sel[-1].time_range.end - t0) / 1000:.3f} ms")
busy_end = t0
gap_total = 0.0
print("  start_us   dur_us  gap_us  name")
for e in sel:
    s, t = e.time_range.start, e.time_range.end
    gap = max(0.0, s - busy_end)
    gap_total += gap
    name = e.name.replace("void vg::", "").replace("vg::", "")[:70]
    print(f"{(s - t0):10.1f} {(t - s):8.1f} {gap:7.1f}  {name}")
    busy_end = max(busy_end, t)
print(f"idle (no kernel running anywhere) {gap_total / 1000:.3f} ms")
