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
# BASELINE.json configs[3] ("scaled cohort"): per-GPU B=128, m=8 (m <= 8, F7), 64 subjects = 6 272 volumes
WORKLOAD4 = "configs[3]: scaled cohort, 64 synthetic subj x 98 vol (41x49x35), B=128 per GPU, m=8, gp_kl_scale=10"
CONFIGS = {2: dict(workload=WORKLOAD, batch=32, m=6, subjects=14),
           4: dict(workload=WORKLOAD4, batch=128, m=8, subjects=64)}
ARITH_NAMES = {0: "f32", 1: "bf16", 2: "bf16+f32enc"}


def make_config(workload, B, world):
    """The `config` object — identical in the native and the reference arm (same keys, same values)."""
    return {"workload": workload, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": "inputs cycle through an HBM-resident cohort (>= 386 MB) and the step's >= 1.8 GB activation working "
                  "set, both larger than the 126 MB L2"}

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
    nbytes = 4.0 * (vi * cin + vo * cout)          # fp32 input + output tensors of the layer
    return flops, nbytes


def conv_extra_bytes(name, kind):
    """Bytes a fused epilogue reads ON TOP of §8(d)(ii): the saved activation of a data gradient (ReLU mask or
    BatchNorm-backward statistics).  Reported separately ("extended") — the §8(d) figure is the headline."""
    cin, cout, taps, vo, vi = _CONV[name]
    return 4.0 * vi * cin if (kind == "dgrad" and name != "conv1") else 0.0


def absorbed_layer_bytes(op, B):
    """Layer-by-layer fp32 bytes (the accounting of SURVEY §8(d)(iii)) of the reference layers ONE launch of `op`
    replaces, where that is more than the convolution pass itself.  convt5.dgrad also executes bnt5's BatchNorm +
    ReLU backward in its epilogue: dgrad reads dy (1 ch) and writes dX (8 ch); the absorbed
    BatchNorm backward would then read dX and the saved activation and write its own dX (3 x 8 ch)."""
    if op == "convt5.dgrad":
        cin, cout, taps, vo, vi = _CONV["convt5"]
        return 9 * B * 4.0 * (vo * cout + vi * cin + 3 * vi * cin)
    return None


def op_work(op, B, extended=False):
    """Algorithmic (flops, bytes) of one recorded operation at minibatch B; None if not modelled."""
    layer, _, kind = op.partition(".")
    if layer in _CONV:
        n = B if layer.startswith("conv") and not layer.startswith("convt") else 9 * B
        f, b = conv_work(layer, kind)
        if extended:
            b += conv_extra_bytes(layer, kind)
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


def make_cohort_tensors(rank, device, subjects=N_SUBJECTS):
    """HBM-resident cohort: (volumes, covariates, subject index).  Up to 14 subjects are generated one by one
    (vaegam.synthetic); a larger cohort (configs[3]: 64 subjects) tiles those 14 anatomies / designs with fresh
    device-side acquisition noise per volume, so every volume is distinct without minutes of host-side generation."""
    from vaegam import synthetic as syn
    coh = syn.make_cohort(min(subjects, N_SUBJECTS), "checker", seed=rank)
    vols = coh.volumes().to(device)
    covs = torch.from_numpy(coh.covariates().copy()).to(device)
    sidx = torch.from_numpy(coh.subject_index().copy()).to(device)
    if subjects > N_SUBJECTS:
        reps = (subjects + N_SUBJECTS - 1) // N_SUBJECTS
        g = torch.Generator(device=device).manual_seed(1234 + rank)
        n_keep = subjects * 98
        big = vols.repeat(reps, 1, 1, 1)[:n_keep]
        big = (big + 0.02 * torch.randn(big.shape, device=device, generator=g)).clamp_(0, 1)
        covs = covs.repeat(reps, 1)[:n_keep]
        sidx = (sidx.repeat(reps) + N_SUBJECTS * torch.arange(reps, device=device).repeat_interleave(sidx.numel()))[:n_keep]
        vols = big
    return coh, vols, covs, sidx


def build_model(workdir, seed=1, m=6, device_name="auto"):
    import vae_reg_GP
    from vaegam import synthetic as syn
    tr, te, glm, _ = syn.write_experiment(workdir, n_subjects=2, config="checker", glm="uniform")
    torch.manual_seed(seed)
    model = vae_reg_GP.VAE(save_dir=workdir, glm_maps=glm, csv_files=[tr, te], num_inducing_pts=m, device_name=device_name)
    model.writer = vae_reg_GP._NullWriter()
    return model


def cpu_baseline(steps=3, warmup=1, threads=None, B=BATCH, m=6):
    """The oracle port of the reference training step on the host CPUs (volumes/s)."""
    import tempfile
    from oracle import ref_port as rp
    from vaegam import synthetic as syn
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    work = tempfile.mkdtemp(prefix="cpu_base_")
    model = build_model(work, m=m, device_name="cpu")
    P = rp.cast_params(rp.params_from_module(model), torch.float32, requires_grad=True)
    coh = syn.make_cohort(2, "checker", seed=0)
    x = coh.volumes(rows=range(B)).float()
    cov = torch.from_numpy(coh.covariates()[:B].copy())
    st = {}
    times = []
    for i in range(warmup + steps):
        noise = rp.draw_noise(B, seed=100 + i)
        t0 = time.perf_counter()
        rp.training_step_cpu(P, st, x, cov, noise)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return B / float(np.mean(times)), threads, float(np.mean(times))


def stock_torch_b200(device, steps=4, warmup=2, B=BATCH):
    """Diagnostic (SURVEY §8d last row): the reference's algorithm through STOCK PyTorch (cuDNN / cuBLAS /
    cuSOLVER) on this same B200 — the practical bar.  Uses the unmodified reference staged under baseline/_ref
    when it is there (forward(train_mode=False) + backward + Adam, logging off, as SURVEY §8d prescribes), else
    the oracle port moved to the GPU.  TF32 off (fp32) and on."""
    import tempfile
    from oracle import ref_port as rp
    from vaegam import synthetic as syn
    out = {}
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    coh = syn.make_cohort(2, "checker", seed=0)
    x = coh.volumes(rows=range(B)).float().to(device)
    cov = torch.from_numpy(coh.covariates()[:B].copy()).to(device)
    ids = torch.from_numpy(coh.subject_index()[:B].copy()).to(device)
    work = tempfile.mkdtemp(prefix="stock_")
    tr, te, glm, _ = syn.write_experiment(work, n_subjects=2, config="checker", glm="uniform")
    # The reference inverts Ku in fp32 without jitter (gp.py:107); with z-scored motion covariates the
    # inducing grid is finer than the length scale and cuSOLVER's fp32 inverse leaves the gain covariance non-PD at
    # initialisation.  For this TIMING diagnostic the motion covariates are spread 3x wider (grid step > length scale),
    # which conditions Ku without changing a single operation of the step.
    import pandas as pd
    SPREAD = 3.0
    for path in (tr, te):
        df = pd.read_csv(path, index_col=0)
        df[["x", "y", "z", "rot_x", "rot_y", "rot_z"]] *= SPREAD
        df.to_csv(path)
    cov = cov.clone()
    cov[:, 1:7] *= SPREAD
    ref_model = None
    kind = "oracle port (oracle/ref_port.py) on cuda"
    if os.path.isfile(os.path.join(ref_dir, "vae_reg_GP.py")):
        try:
            from oracle import ref_loader
            ref_loader.REFERENCE_DIR = ref_dir
            ref_vae, _, _ = ref_loader.load_reference()
            torch.manual_seed(1)
            ref_model = ref_vae.VAE(save_dir=work, glm_maps=glm, csv_files=[tr, te])
            ref_model.writer = ref_loader.NullWriter()
            kind = "unmodified reference (baseline/_ref) VAE.forward(train_mode=False) + backward + Adam.step"
        except Exception as e:      # missing optional dependency etc.: fall back to the port
            out["reference_import_error"] = repr(e)[:200]
            ref_model = None
    if ref_model is None:
        model = build_model(work, device_name="cpu")
        P = rp.cast_params({k: v.to(device) for k, v in rp.params_from_module(model).items()}, torch.float32,
                           requires_grad=True)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        # torch's defaults on this stack: cuDNN convolutions may use TF32, matmuls may not.  The matmul switch stays off:
        # TF32 in the reference's A (S - Ku) A^T makes its gain covariance non-positive-definite.
        for name, tf32 in (("fp32", False), ("tf32_conv_torch_default", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            st, times = {}, []
            try:
              for i in range(warmup + steps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if ref_model is not None:
                    loss = ref_model.forward(ids, cov, x, 'train', train_mode=False)
                    ref_model.optimizer.zero_grad()
                    loss.backward()
                    ref_model.optimizer.step()
                else:
                    with torch.device(device):
                        rp.training_step_cpu(P, st, x, cov, rp.draw_noise(B, seed=100 + i))
                torch.cuda.synchronize()
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
              out[name] = {"value": round(B / float(np.mean(times)), 1), "unit": "volumes/s",
                           "ms_per_step": round(1e3 * float(np.mean(times)), 2)}
            except Exception as e:       # e.g. the reference's own fp32 GP algebra losing positive definiteness
                out[name] = {"error": repr(e)[:160], "completed_steps": len(times)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["what"] = kind
    out["steps"] = steps
    return out


def run_reference(args):
    """`--impl reference`: the reference's own algorithm on the host cores (oracle port; the Python
    reference itself cannot travel to the GPU box).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t_start = time.time()
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    val, threads, sec = cpu_baseline(steps=max(1, args.steps), warmup=max(0, args.warmup), B=B, m=cfg["m"])
    line = {
        "impl": "reference", "metric": "training volumes/sec (fwd+bwd+step)", "value": val, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(cfg["workload"], B, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": "volumes/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of B={B} (oracle/ref_port.py: torch fp32 CPU conv/BN/linear "
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


def timed_steps(model, reducer, vols, covs, sidx, B, steps, warmup, rank, world, sampler=None):
    """W untimed + K timed resident steps (inputs already in HBM); returns (ms, launches, clocks)."""
    import torch.distributed as dist
    from vaegam import dp, native
    from vaegam.step import GraphStep
    n_items = vols.shape[0]
    perm = torch.randperm(n_items, generator=torch.Generator().manual_seed(rank)).to(vols.device)
    n_batches = n_items // B

    def step(i):
        j = i % n_batches
        idx = perm[j * B:(j + 1) * B]
        return dp.train_step(model, reducer, sidx[idx], covs[idx], vols[idx])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    settle = 5 if getattr(model, "use_cuda_graph", False) else 0
    for i in range(max(warmup, 3) + settle):    # two eager steps, capture on the third, then a few untimed replays
        step(i)
    sync_all()
    if sampler is not None:
        sampler.start()
    n0 = native.launch_count() + GraphStep.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    sync_all()
    launches = native.launch_count() + GraphStep.replayed_launches - n0
    clocks = sampler.stop() if sampler is not None else None
    model.check_status()
    return e0.elapsed_time(e1), launches, clocks, settle


def max_over_ranks(values, device, world):
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def params_identical(model, device, world):
    """Data-parallel invariant: every rank holds bit-identical parameters after the timed steps."""
    import torch.distributed as dist
    if world == 1:
        return None
    same = torch.ones(1, dtype=torch.int32, device=device)
    for buf in (model._flat.flat32, model._flat.flat64):
        ref = buf.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(ref, buf):
            same.zero_()
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    return bool(int(same.item()))


def ncu_facts(op):
    """dram traffic and tensor-pipe % of an operation from this round's committed ncu capture (profiles/);
    an ncu capture can never come from the timed run itself, so the source file is named."""
    for name in ("r2_ncu_traffic.json", "r1_s5_ncu_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            tr = d["ops"].get(op)
            if tr:
                return {"traffic": tr["dram_read_bytes"] + tr["dram_write_bytes"], "tensor_pipe_pct": tr.get("tensor_pipe_pct"),
                        "source": f"profiles/{name} ({d.get('build', 'ncu --set full capture')})", "stale": name.startswith("r1_")}
        except Exception:
            continue
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration: 2 = configs[1] (the metric's workload), 4 = configs[3] (scaled cohort)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the diagnostic sub-measurements")
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
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured" if peaks else "fallback"

    work = tempfile.mkdtemp(prefix=f"bench_r{rank}_")
    model = build_model(work, m=cfg["m"])
    conv_mode = int(native.load().vg_get_conv_mode())        # the library default the model's calls resolve to
    coh, vols, covs, sidx = make_cohort_tensors(rank, device, cfg["subjects"])
    reducer = dp.GradientAllReduce(model._flat, model.optimizer)
    reducer.broadcast_parameters()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput
    sampler = ClockSampler(local) if (rank == 0 and os.environ.get("VAEGAM_BENCH_NO_SMI") != "1") else None
    ms, launches, clocks, settle = timed_steps(model, reducer, vols, covs, sidx, B, args.steps, args.warmup, rank, world, sampler)
    n_items = vols.shape[0]
    perm = torch.randperm(n_items, generator=torch.Generator().manual_seed(rank)).to(device)
    n_batches = n_items // B

    def batch_idx(i):
        j = i % n_batches
        return perm[j * B:(j + 1) * B]

    # ---------------- end to end through the public API, host buffers
    # A loader-style input pipeline: pinned host slots + two device slots; the H2D copy of step i+1 is issued on a
    # copy stream while step i computes, and every step's loss travels back through a pinned buffer (D2H per step;
    # the host reads the value of step i-1 while step i runs, as a training loop that logs one step late does).
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
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]

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
            k = i % 2
            loss_host[k].copy_(loss, non_blocking=True)                           # D2H of this step's loss
            loss_done[k].record()
            sl["free"].record()
            if i + 1 < first + count:
                h2d(i + 1)                               # next step's inputs travel while this step computes
            if i > first:                                # read the previous step's loss (already on the host)
                loss_done[1 - k].synchronize()
                out += float(loss_host[1 - k])
        loss_done[(first + count - 1) % 2].synchronize()
        out += float(loss_host[(first + count - 1) % 2])
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
    e2e_ms = f0.elapsed_time(f1)   # device clock; every loss value has reached the host before f1's sync returns
    identical = params_identical(model, device, world)
    graphed = any(gs.graph is not None for gs in model._graph_steps.values())     # before the sub-measurements drop the graphs

    # ---------------- max over ranks
    ms, e2e_ms = max_over_ranks([ms, e2e_ms], device, world)

    # ---------------- sub-measurements every rank takes part in: the other arithmetic modes, BASELINE configs[3]
    extras = {}
    if not args.no_extras:
        for name, arith in (("fp32_check_mode", "fp32"), ("bf16_mode", "bf16")):
            model.arith = arith
            t_ms, _, _, _ = timed_steps(model, reducer, vols, covs, sidx, B, 10, 3, rank, world)
            (t_ms,) = max_over_ranks([t_ms], device, world)
            extras[name] = {"value": round(10 * B * world / (t_ms * 1e-3), 1), "unit": "volumes/s",
                            "ms_per_step": round(t_ms / 10, 4), "arith": arith, "steps": 10}
        model.arith = "default"
        if args.config != 4:
            c4 = CONFIGS[4]
            del host_vol
            model._graph_steps.clear()
            work4 = tempfile.mkdtemp(prefix=f"bench4_r{rank}_")
            m4 = build_model(work4, m=c4["m"])
            red4 = dp.GradientAllReduce(m4._flat, m4.optimizer)
            red4.broadcast_parameters()
            _, v4, cv4, si4 = make_cohort_tensors(rank, device, c4["subjects"])
            t_ms, _, _, _ = timed_steps(m4, red4, v4, cv4, si4, c4["batch"], 10, 3, rank, world)
            (t_ms,) = max_over_ranks([t_ms], device, world)
            extras["config4"] = {"workload": c4["workload"], "value": round(10 * c4["batch"] * world / (t_ms * 1e-3), 1),
                                 "unit": "volumes/s", "ms_per_step": round(t_ms / 10, 4), "batch_per_gpu": c4["batch"],
                                 "global_batch": c4["batch"] * world, "steps": 10, "volumes_resident_per_gpu": int(v4.shape[0]),
                                 "dp_params_identical": params_identical(m4, device, world)}
            m4._graph_steps.clear()
            del m4, red4, v4, cv4, si4
            torch.cuda.empty_cache()

    # ---------------- live per-operation timing (separate pass, same steps)
    kernels, roof = {}, None
    if not args.no_profile and rank == 0:
        native.profile(True)
        ksteps = 5
        for i in range(ksteps):          # local steps only: the other ranks are not in this pass
            idx = batch_idx(i)
            loss = model.forward(sidx[idx], covs[idx], vols[idx], 'train', train_mode=False)
            model.optimizer.zero_grad()
            loss.backward()
            model.optimizer.step()
        torch.cuda.synchronize()
        rec = native.profile_collect()
        native.profile(False)
        # per-step time of an operation = the MEDIAN over the profiled steps of that step's instances (an eager pass
        # shares the box with the host: one stalled launch must not pass for a kernel's duration)
        def step_median(v):
            per = len(v) // ksteps
            if per < 1 or per * ksteps != len(v):
                return sum(v) / ksteps
            return float(np.median([sum(v[i * per:(i + 1) * per]) for i in range(ksteps)]))
        med = {op: step_median(v) for op, v in rec.items()}
        total = sum(med.values())
        rows = []
        for op, v in rec.items():
            per_step = med[op]
            w = op_work(op, B)
            row = {"op": op, "ms_per_step": round(per_step, 4), "share": round(per_step / total, 4)}
            if w:
                row["gflops"] = round(w[0] / (per_step * 1e-3) / 1e9, 1)
                row["gbs"] = round(w[1] / (per_step * 1e-3) / 1e9, 1)
                row["frac_hbm"] = round(row["gbs"] / hbm_peak, 4)
                if op.split(".")[0] in _CONV:
                    row["tensor_frac"] = round(row["gflops"] / 1e3 / tf_peak, 5)     # algorithmic FLOP/s / measured bf16 peak
                    nf = ncu_facts(op)
                    if nf and nf.get("tensor_pipe_pct") is not None:
                        row["tensor_pipe_pct_ncu"] = nf["tensor_pipe_pct"]
            rows.append(row)
        rows.sort(key=lambda r: -r["ms_per_step"])
        kernels = {"ops": rows, "profiled_ms_per_step": round(total, 3), "profiled_steps": ksteps, "statistic": "median over steps"}
        top = next((r for r in rows if "gbs" in r), None)
        if top:
            w = op_work(top["op"], B)
            wx = op_work(top["op"], B, extended=True)
            nf = ncu_facts(top["op"]) if B == BATCH else None
            ext_gbs = wx[1] / (top["ms_per_step"] * 1e-3) / 1e9
            roof = {"kernel": top["op"], "bound": "hbm", "achieved": top["gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(top["gbs"] / hbm_peak, 4),
                    "traffic": nf["traffic"] if nf else None, "traffic_source": nf["source"] if nf else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes": w[1], "algorithmic_bytes_rule": "SURVEY 8(d)(ii): fp32 input + output tensors of the layer pass",
                    "extended": {"bytes": wx[1], "achieved": round(ext_gbs, 1), "frac": round(ext_gbs / hbm_peak, 4),
                                 "rule": "8(d)(ii) + the saved activation a fused data-gradient epilogue reads (ReLU mask / BatchNorm-backward sums)"},
                    "ms": top["ms_per_step"], "share_of_step": top["share"],
                    "tensor_frac": round(top["gflops"] / 1e3 / tf_peak, 5),
                    "tensor_pipe_pct_ncu": nf.get("tensor_pipe_pct") if nf else None}
            ab = absorbed_layer_bytes(top["op"], B)
            if ab:
                ab_gbs = ab / (top["ms_per_step"] * 1e-3) / 1e9
                roof["layers_absorbed"] = {
                    "bytes": ab, "achieved": round(ab_gbs, 1), "frac": round(ab_gbs / hbm_peak, 4),
                    "rule": "fp32 layer-by-layer bytes (8(d)(iii) accounting) of the two reference layers this launch executes: "
                            "the convolution data gradient AND bnt5's BatchNorm+ReLU backward fused into its epilogue; "
                            "not the headline frac, which stays the strict 8(d)(ii) figure of the convolution alone"}
        rl = [r for r in rows if r["op"].startswith("recon_loss")]
        kernels["fused_loss"] = [{"op": r["op"], "gbs": r.get("gbs"), "frac_hbm": round(r.get("gbs", 0) / hbm_peak, 4)} for r in rl]
        for bb in (128, 512):       # alone, inputs (0.36 / 1.4 GB) larger than L2: the stage's own roofline figure
            fl = fused_loss_roofline(device, B=bb)
            kernels[f"fused_loss_b{bb}_alone"] = {k: dict(v, frac_hbm=round(v["gbs"] / hbm_peak, 4)) for k, v in fl.items()}
        if not args.no_extras:
            # BASELINE configs[4] (--recons_only): forward with the 10 map outputs + their D2H, no file writes
            nrec = 10
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.no_grad():
                for i in range(nrec):
                    idx = batch_idx(i)
                    model.forward(sidx[idx], covs[idx], vols[idx], 'reconstruction', return_latent_rec=True, train_mode=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            extras["config5_recons"] = {"workload": "configs[4]: forward(return_latent_rec=True): z + 10 maps per volume copied to the host, file writes excluded",
                                        "value": round(nrec * B / dt, 1), "unit": "volumes/s", "d2h_bytes_per_step": int(10 * B * V * 4 + B * 32 * 4)}
            try:
                extras["stock_torch_b200"] = stock_torch_b200(device)
            except Exception as e:
                extras["stock_torch_b200"] = {"error": repr(e)[:300]}

    if rank != 0:
        finish(model, world)
        return
    total_vols = args.steps * B * world
    value = total_vols / (ms * 1e-3)
    line = {
        "metric": "training volumes/sec (fwd+bwd+step)", "value": value, "unit": "volumes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": ARITH_NAMES.get(conv_mode, "bf16"), "data": "synthetic",
        "config": make_config(cfg["workload"], B, world),
        "impl_notes": {
            "launch": (("whole step (fwd + bwd + bucketed all-reduce + Adam) replayed as one CUDA graph; captured during "
                        f"warm-up, {settle} extra untimed replays before the timed region") if graphed
                       else "eager launches (CUDA graph off, not captured, or running under a profiler)"),
            "gradient_allreduce": ("3 buckets on a comm stream, each issued as soon as its backward phase is enqueued "
                                   "(decoder -> encoder FC -> encoder convolutions); joined before Adam") if world > 1 else "none (1 GPU)",
            "gain_stage": "fp64", "other_stages": "fp32",
            "conv": {0: "fp32 CUDA-core direct convolution (check mode)",
                     1: "bf16 operands / fp32 accumulate: tcgen05+TMEM implicit GEMM (fwd, dgrad), mma.sync (wgrad)",
                     2: "mixed: bf16 operands / fp32 accumulate on tcgen05+TMEM (decoder fwd, every dgrad), mma.sync bf16 (wgrad); "
                        "the encoder's five forward convolutions in fp32 (they set the gradient accuracy, tools/precision_study.py)"}[conv_mode]},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": k2 * B * world / (e2e_ms * 1e-3), "unit": "volumes/s",
                "h2d_bytes_per_step": int(B * V * 4 + B * 8 * 4 + B * 8), "d2h_bytes_per_step": 4,
                "steps": k2, "api": "VAE.train_batch (the per-batch body of train_epoch: fwd + bwd [+ NCCL all-reduce] + Adam as one "
                       "CUDA-graph launch after two eager steps); each step's loss copied to pinned host memory and read there",
                "input_pipeline": "loader thread gathers each batch into pinned host slots; H2D of step i+1 on a copy "
                                  "stream during step i"},
        "roofline": roof, "kernels": kernels,
        # layer-by-layer fp32 activation traffic of a whole step, ~52.7 MB forward x 3 per volume
        "step_hbm_ceiling": {"algorithmic_bytes_per_volume": 158.1e6, "peak_gbs": hbm_peak,
                             "ceiling_volumes_per_s_per_gpu": round(hbm_peak * 1e9 / 158.1e6, 1),
                             "frac": round(value / world / (hbm_peak * 1e9 / 158.1e6), 4)},
    }
    if identical is not None:
        line["dp_params_identical"] = identical
    line.update(extras)
    if not args.no_cpu_baseline and world == 1:
        val, threads, sec = cpu_baseline(steps=3, warmup=1, B=B, m=cfg["m"])
        line["cpu_baseline"] = {"value": val, "unit": "volumes/s", "cores": threads, "kind": "port",
                                "sample": f"3 steps of B={B} after 1 warm-up ({sec:.2f} s/step)"}
    print(json.dumps(line), flush=True)
    finish(model, world)


if __name__ == "__main__":
    main()
