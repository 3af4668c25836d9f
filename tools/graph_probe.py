#!/usr/bin/env python
"""Per-step time of VAE.train_batch on device-resident data: eager vs CUDA graph, with and without a host sync per step."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
dev = torch.device("cuda", 0)
coh, vols, covs, sidx = bench.make_cohort_tensors(0, dev)
B = 32
for graph in (False, True):
    model = bench.build_model(tempfile.mkdtemp(prefix="probe_"))
    model.use_cuda_graph = graph
    for sync in (False, True):
        for phase in ("warm", "timed"):
            n = 5 if phase == "warm" else 40
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for i in range(n):
                idx = torch.arange(i * B, (i + 1) * B, device=dev) % vols.shape[0]
                loss = model.train_batch(sidx[idx], covs[idx], vols[idx])
                if sync:
                    loss.item()
            e1.record()
            torch.cuda.synchronize()
            if phase == "timed":
                print(f"graph={graph} sync_each_step={sync}: {e0.elapsed_time(e1) / n:.3f} ms/step (wall {1e3 * (time.perf_counter() - t0) / n:.3f})", flush=True)
