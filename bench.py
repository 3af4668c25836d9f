#!/usr/bin/env python
"""Headline benchmark: VAE-GAM training volumes/s (fwd + bwd + Adam step) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # native sm_100a path (this repo)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host CPUs

Workload (BASELINE.json configs[1]): multi-subject synthetic checkerboard cohort, 14 subjects x 98
volumes of 41x49x35 (1372 volumes), batch 32, default num_inducing_pts=6 / gp_kl_scale=10, HRF on.
One "step" = one minibatch through forward, backward and the Adam update.  Every rank trains on
its own 14-subject cohort (weak scaling); gradients are all-reduced with NCCL.

Prints ONE JSON line (rank 0).  `value` = volumes/s with the cohort resident in HBM; `e2e` = the same
through the public API (`VAE.forward` / `loss.backward()` / `optimizer.step()`) with the batch copied
from pinned host memory and the loss read back every step.  `roofline` describes the dominant
operation, timed live with CUDA events on the launching stream (vg_profile_*); `kernels` lists
every operation's share.  `cpu_baseline` is the oracle port of the reference step
(oracle/ref_port.py, PyTorch fp32 CPU kernels, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vae-gam_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

V = 41 * 49 * 35
BATCH = 32
N_SUBJECTS = 14
WORKLOAD = "configs[1]: multi-subject synthetic checkerboard, 14 subj x 98 vol (41x49x35), B=32, m=6, gp_kl_scale=10"

# algorithmic work per operation and per image: (flops, bytes) for one image of the
# layer's batch; conv bytes = fp32 input + output tensors, flops = 2*M*N*K
_CONV = {  # name: (cin, cout, taps, out voxels, in voxels)
    "conv1": (1, 8, 27, 60489, 70315), "conv2": (8, 8, 27, 6992, 60489), "conv3": (8, 16, 27, 4998, 6992),
    "conv4": (16, 16, 27, 480, 4998), "conv5": (16, 16, 27, 192, 480),
    "convt1": (16, 16, 27, 560, 240), "convt2": (16, 16, 27, 4704, 560), "convt3": (16, 8, 27, 6624, 4704),
    "convt4": (8, 8, 45, 60489, 6624), "convt5": (8, 1, 27, 70315, 60489),
}


def conv_work(name, kind="fwd"):
    """(flops, bytes) per image for one pass (fwd, dgrad or wgrad) of a conv layer.  Algorithmic bytes (fp32
    tensors): fwd x + y; wgrad x + dy; dgrad dy + dx + the saved tensor its fused epilogue reads
    (ReLU mask or BatchNorm-backward statistics) — conv1.dgrad only produces the bn1 statistics (dy + x)."""
    cin, cout, taps, vo, vi = _CONV[name]
    if name.startswith("convt"):          # gather form: MACs = input voxels * taps * cin * cout
        flops = 2.0 * vi * taps * cin * cout
    else:
        flops = 2.0 * vo * taps * cin * cout
    nbytes = 4.0 * (vi * cin + vo * cout)
    if kind == "dgrad" and name != "conv1":
        nbytes += 4.0 * vi * cin
    return flops, nbytes


def op_work(op, B):
    """Algorithmic (flops, bytes) of one recorded operation at minibatch B; None if not modelled."""
    layer, _, kind = op.partition(".")
    if layer in _CONV:
        n = B if layer.startswith("conv") and not layer.startswith("convt") else 9 * B
        f, b = conv_work(layer, kind)
        return f * n, b * n
    if op == "recon_loss.fwd":
        return 30.0 * B * V, 4.0 * (10 * B * V + 9 * V)
    if op == "recon_loss.bwd":
        return 60.0 * B * V, 4.0 * (19 * B * V + 10 * V)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fused_loss_roofline(device, B=128, iters=20):
    """The fused reconstruction/likelihood/GLM pass timed alone at B=128 (maps 324 MB, larger than
    the 126 MB L2, so every launch streams from HBM): algorithmic bytes / CUDA-event time."""
    import ctypes as C
    from vaegam import native
    lib = native.load()
    VP = native.VP
    g = torch.Generator(device=device).manual_seed(0)
    maps = torch.rand(9, B, VP, device=device, generator=g)
    gains = torch.randn(8, B, device=device, generator=g)
    x = torch.rand(B, V, device=device, generator=g)
    eps = torch.full((VP,), -2.3, device=device)
    glm = torch.rand(8, VP, device=device, generator=g)
    logp = torch.empty(B, device=device); norms = torch.empty(8, B, device=device)
    dpre = torch.empty(9, B, VP, device=device); dg = torch.empty(8, B, device=device); deps = torch.empty(VP, device=device)
    nbytes = int(lib.vg_recon_workspace_bytes(B, V)); ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    st = native.stream_ptr()
    fwd = lambda: native.check(lib.vg_recon_loss_fwd(native.ptr(maps), native.ptr(gains), native.ptr(x), native.ptr(eps),
                                                     native.ptr(glm), B, V, native.ptr(logp), native.ptr(norms), None, None,
                                                     native.ptr(ws), nbytes, st))
    bwd = lambda: native.check(lib.vg_recon_loss_bwd(native.ptr(maps), native.ptr(gains), native.ptr(x), native.ptr(eps),
                                                     native.ptr(glm), native.ptr(norms), B, V, 1.0, native.ptr(dpre),
                                                     native.ptr(dg), native.ptr(deps), native.ptr(ws), nbytes, st))
    out = {}
    for name, fn, nb in (("fwd", fwd, 4.0 * (10 * B * V + 9 * V)), ("bwd", bwd, 4.0 * (19 * B * V + 10 * V))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"ms": round(ms, 4), "gbs": round(nb / (ms * 1e-3) / 1e9, 1), "bytes": nb}
    return out


def make_cohort_tensors(rank, device):
    from vaegam import synthetic as syn
    coh = syn.make_cohort(N_SUBJECTS, "checker", seed=rank)
    return coh, coh.volumes().to(device), torch.from_numpy(coh.covariates()).to(device), \
        torch.from_numpy(coh.subject_index()).to(device)


def build_model(workdir, seed=1):
    import vae_reg_GP
    from vaegam import synthetic as syn
    tr, te, glm, _ = syn.write_experiment(workdir, n_subjects=2, config="checker", glm="uniform")
    torch.manual_seed(seed)
    model = vae_reg_GP.VAE(save_dir=workdir, glm_maps=glm, csv_files=[tr, te])
    model.writer = vae_reg_GP._NullWriter()
    return model


def cpu_baseline(steps=3, warmup=1, threads=None):
    """The oracle port of the reference training step on the host CPUs (volumes/s)."""
    import tempfile
    from oracle import ref_port as rp
    from vaegam import synthetic as syn
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    work = tempfile.mkdtemp(prefix="cpu_base_")
    model = build_model_cpu(work)
    P = rp.cast_params(rp.params_from_module(model), torch.float32, requires_grad=True)
    coh = syn.make_cohort(1, "checker", seed=0)
    x = coh.volumes(rows=range(BATCH)).float()
    cov = torch.from_numpy(coh.covariates()[:BATCH])
    st = {}
    times = []
    for i in range(warmup + steps):
        noise = rp.draw_noise(BATCH, seed=100 + i)
        t0 = time.perf_counter()
        rp.training_step_cpu(P, st, x, cov, noise)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return BATCH / float(np.mean(times)), threads, float(np.mean(times))


def build_model_cpu(workdir):
    import vae_reg_GP
    from vaegam import synthetic as syn
    tr, te, glm, _ = syn.write_experiment(workdir, n_subjects=2, config="checker", glm="uniform")
    torch.manual_seed(1)
    m = vae_reg_GP.VAE(save_dir=workdir, glm_maps=glm, csv_files=[tr, te], device_name="cpu")
    return m


def run_reference(args):
    """`--impl reference`: the reference's own algorithm on the host cores (oracle port; the Python
    reference itself cannot travel to the GPU box).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t_start = time.time()
    val, threads, sec = cpu_baseline(steps=max(1, args.steps), warmup=max(0, args.warmup))
    line = {
        "impl": "reference", "metric": "training volumes/sec (fwd+bwd+step)", "value": val, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch": BATCH},
        "cpu_baseline": {"value": val, "unit": "volumes/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of B={BATCH} (oracle/ref_port.py: torch fp32 CPU conv/BN/linear "
                                   f"+ closed forms, autograd backward, Adam)"},
        "e2e": {"value": val, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t_start,
    }
    print(json.dumps(line), flush=True)


def finish(model, world):
    """Multi-rank exit: drop the captured CUDA graphs (they hold NCCL work) before the ranks part, then leave
    without running communicator destructors — tearing NCCL down under live graphs can block."""
    if world <= 1:
        return
    import gc
    import torch.distributed as dist
    model._graph_steps.clear()
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import tempfile
    import torch.distributed as dist
    from vaegam import dp, native
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    B = args.batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured" if peaks else "fallback"

    work = tempfile.mkdtemp(prefix=f"bench_r{rank}_")
    model = build_model(work)
    conv_mode = int(native.load().vg_get_conv_mode())
    coh, vols, covs, sidx = make_cohort_tensors(rank, device)
    n_items = vols.shape[0]
    reducer = dp.GradientAllReduce(model._flat, model.optimizer)
    reducer.broadcast_parameters()
    perm = torch.randperm(n_items, generator=torch.Generator().manual_seed(rank)).to(device)
    n_batches = n_items // B

    def batch_idx(i):
        j = i % n_batches
        return perm[j * B:(j + 1) * B]

    def resident_step(i):
        idx = batch_idx(i)
        return dp.train_step(model, reducer, sidx[idx], covs[idx], vols[idx])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput
    GRAPH_SETTLE = 5 if getattr(model, "use_cuda_graph", False) else 0
    for i in range(max(args.warmup, 3) + GRAPH_SETTLE):   # two eager steps, capture on the third, then a few untimed
        resident_step(i)                                  # replays (first launches upload the graph)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("VAEGAM_BENCH_NO_SMI") != "1":
        sampler.start()
    from vaegam.step import GraphStep
    n0 = native.launch_count() + GraphStep.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        resident_step(args.warmup + i)
    e1.record()
    sync_all()
    launches = native.launch_count() + GraphStep.replayed_launches - n0     # host-side launches + kernels inside graph replays
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    model.check_status()

    # ---------------- end to end through the public API, host buffers
    # A loader-style input pipeline: two pinned host slots + two device slots; the H2D copy of step i+1 is
    # issued on a copy stream while step i computes.  Every step's copy happens inside the timed region.
    host_vol = vols.cpu()
    host_cov, host_idx = covs.cpu(), sidx.cpu()
    perm_h = perm.cpu()
    copy_stream = torch.cuda.Stream()
    import queue
    import threading
    NHOST, NDEV = 3, 2
    hslots = [{"hx": torch.empty(B, 41, 49, 35).pin_memory(), "hc": torch.empty(B, 8).pin_memory(),
               "hi": torch.empty(B, dtype=torch.int64).pin_memory(), "copied": torch.cuda.Event()} for _ in range(NHOST)]
    dslots = [{"dx": torch.empty(B, 41, 49, 35, device=device), "dc": torch.empty(B, 8, device=device),
               "di": torch.empty(B, dtype=torch.int64, device=device),
               "ready": torch.cuda.Event(), "free": torch.cuda.Event()} for _ in range(NDEV)]
    for d in dslots:
        d["free"].record()

    def loader(first, count, out_q, free_q):
        """What a DataLoader worker does: gathers batch i of the epoch's permutation straight into a pinned slot."""
        for i in range(first, first + count):
            h = free_q.get()
            hslots[h]["copied"].synchronize()          # the H2D that last read this pinned slot has finished
            j = i % n_batches
            idx = perm_h[j * B:(j + 1) * B]
            torch.index_select(host_vol, 0, idx, out=hslots[h]["hx"])
            torch.index_select(host_cov, 0, idx, out=hslots[h]["hc"])
            torch.index_select(host_idx, 0, idx, out=hslots[h]["hi"])
            out_q.put((i, h))

    def e2e_run(first, count):
        out_q, free_q = queue.Queue(), queue.Queue()
        for h in range(NHOST):
            free_q.put(h)
        th = threading.Thread(target=loader, args=(first, count, out_q, free_q), daemon=True)
        th.start()

        def h2d(i):                      # pinned host slot -> device slot i % NDEV on the copy stream
            k, h = out_q.get()
            assert k == i
            hs, ds = hslots[h], dslots[i % NDEV]
            copy_stream.wait_event(ds["free"])          # the step that last read this device slot has finished
            with torch.cuda.stream(copy_stream):
                ds["dx"].copy_(hs["hx"], non_blocking=True)
                ds["dc"].copy_(hs["hc"], non_blocking=True)
                ds["di"].copy_(hs["hi"], non_blocking=True)
                ds["ready"].record(copy_stream)
                hs["copied"].record(copy_stream)
            free_q.put(h)

        h2d(first)
        out = 0.0
        for i in range(first, first + count):
            sl = dslots[i % NDEV]
            torch.cuda.current_stream().wait_event(sl["ready"])
            # what train_epoch runs per batch (+ the NCCL gradient all-reduce when world > 1)
            loss = model.train_batch(sl["di"], sl["dc"], sl["dx"], reducer=reducer if world > 1 else None)
            if i + 1 < first + count:
                h2d(i + 1)                               # next step's inputs travel while this step computes
            out = loss.item()                                                      # D2H, as train_epoch does
            sl["free"].record()
        th.join()
        return out

    e2e_run(0, 4)        # includes the two eager steps before the whole-step CUDA graph is captured
    sync_all()
    k2 = max(3, min(args.steps, 20))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(4, k2)
    f1.record()
    sync_all()
    e2e_ms = f0.elapsed_time(f1)   # device clock; the loop is host-synchronous (loss.item())

    # ---------------- max over ranks
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    # ---------------- live per-operation timing (separate pass, same steps)
    kernels, roof = {}, None
    if not args.no_profile and rank == 0:
        native.profile(True)
        ksteps = 3
        for i in range(ksteps):          # local steps only: the other ranks are not in this pass
            idx = batch_idx(i)
            loss = model.forward(sidx[idx], covs[idx], vols[idx], 'train', train_mode=False)
            model.optimizer.zero_grad()
            loss.backward()
            model.optimizer.step()
        torch.cuda.synchronize()
        rec = native.profile_collect()
        native.profile(False)
        total = sum(sum(v) for v in rec.values())
        rows = []
        for op, v in rec.items():
            per_step = sum(v) / ksteps
            w = op_work(op, B)
            row = {"op": op, "ms_per_step": round(per_step, 4), "share": round(sum(v) / total, 4)}
            if w:
                row["gflops"] = round(w[0] / (per_step * 1e-3) / 1e9, 1)
                row["gbs"] = round(w[1] / (per_step * 1e-3) / 1e9, 1)
            rows.append(row)
        rows.sort(key=lambda r: -r["ms_per_step"])
        kernels = {"ops": rows, "profiled_ms_per_step": round(total / ksteps, 3)}
        top = next((r for r in rows if "gbs" in r), None)
        if top:
            w = op_work(top["op"], B)
            traffic = None
            try:      # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel from the committed ncu capture
                tr = json.load(open(os.path.join(ROOT, "profiles", "r1_s5_ncu_traffic.json")))["ops"].get(top["op"])
                if tr and B == BATCH:
                    traffic = tr["dram_read_bytes"] + tr["dram_write_bytes"]
            except Exception:
                pass
            roof = {"kernel": top["op"], "bound": "hbm", "achieved": top["gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(top["gbs"] / hbm_peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes": w[1], "ms": top["ms_per_step"], "share_of_step": top["share"],
                    "tensor_frac": round(top["gflops"] / 1e3 / tf_peak, 5)}
        rl = [r for r in rows if r["op"].startswith("recon_loss")]
        kernels["fused_loss"] = [{"op": r["op"], "gbs": r.get("gbs"), "frac_hbm": round(r.get("gbs", 0) / hbm_peak, 4)} for r in rl]
        for bb in (128, 512):       # alone, inputs (0.36 / 1.4 GB) larger than L2: the stage's own roofline figure
            fl = fused_loss_roofline(device, B=bb)
            kernels[f"fused_loss_b{bb}_alone"] = {k: dict(v, frac_hbm=round(v["gbs"] / hbm_peak, 4)) for k, v in fl.items()}

    if rank != 0:
        finish(model, world)
        return
    total_vols = args.steps * B * world
    value = total_vols / (ms * 1e-3)
    line = {
        "metric": "training volumes/sec (fwd+bwd+step)", "value": value, "unit": "volumes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if conv_mode == 1 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world} (one minibatch per rank, NCCL grad all-reduce)",
                   "launch": (("whole step (fwd + bwd + all-reduce + Adam) replayed as one CUDA graph; captured during "
                               f"warm-up, {GRAPH_SETTLE} extra untimed replays before the timed region")
                              if any(gs.graph is not None for gs in model._graph_steps.values())
                              else "eager launches (CUDA graph off, not captured, or running under a profiler)"),
                   "l2": "inputs cycle through a 386 MB HBM-resident cohort and the step's 1.8 GB activation "
                         "working set, both larger than the 126 MB L2",
                   "gain_stage": "fp64", "other_stages": "fp32",
                   "conv": ("bf16 operands / fp32 accumulate: tcgen05+TMEM implicit GEMM (fwd, dgrad), mma.sync (wgrad)"
                            if conv_mode == 1 else "fp32 CUDA-core direct convolution (check mode)")},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": k2 * B * world / (e2e_ms * 1e-3), "unit": "volumes/s",
                "h2d_bytes_per_step": int(B * V * 4 + B * 8 * 4 + B * 8), "d2h_bytes_per_step": 4,
                "steps": k2, "api": "VAE.train_batch (the per-batch body of train_epoch: fwd + bwd [+ NCCL all-reduce] + Adam as one "
                       "CUDA-graph launch after two eager steps) + loss.item()",
                "input_pipeline": "loader thread gathers each batch into pinned host slots; H2D of step i+1 on a copy "
                                  "stream during step i"},
        "roofline": roof, "kernels": kernels,
        # layer-by-layer fp32 activation traffic of a whole step, ~52.7 MB forward x 3 per volume
        "step_hbm_ceiling": {"algorithmic_bytes_per_volume": 158.1e6, "peak_gbs": hbm_peak,
                             "ceiling_volumes_per_s_per_gpu": round(hbm_peak * 1e9 / 158.1e6, 1),
                             "frac": round(value / world / (hbm_peak * 1e9 / 158.1e6), 4)},
    }
    if not args.no_cpu_baseline and world == 1:
        val, threads, sec = cpu_baseline(steps=3, warmup=1)
        line["cpu_baseline"] = {"value": val, "unit": "volumes/s", "cores": threads, "kind": "port",
                                "sample": f"3 steps of B={BATCH} after 1 warm-up ({sec:.2f} s/step)"}
    print(json.dumps(line), flush=True)
    finish(model, world)


if __name__ == "__main__":
    main()
