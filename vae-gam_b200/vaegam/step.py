"""Host side of the native VAE-GAM step: buffer tables, the autograd bridge and the flat
parameter / gradient / Adam-state storage.

`StepEngine` owns, per minibatch size, one workspace and one `VgStepIO` table and calls
`vg_step_fwd` / `vg_step_bwd` (include/vaegam.h), which chain every kernel of
vae_reg_GP.py:307-413 (+ autograd backward, :427-428) on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import native
from .native import V, VP

GP_KEYS = ["task", "x", "y", "z", "xrot", "yrot", "zrot", "sex"]            # vae_reg_GP.py:68
IMG_KEYS = ["base", "task", "x_mot", "y_mot", "z_mot", "pitch_mot", "roll_mot", "yaw_mot", "sex",
            "full_rec"]                                                      # vae_reg_GP.py:308-309
NUM_LATENTS = 32


def param_order() -> List[str]:
    """The reference's named_parameters() order (ctor order, vae_reg_GP.py:54-218)."""
    names = ["epsilon", "sa_task", "logstd_task"]
    for k in GP_KEYS[1:7]:
        names += [f"qu_m_{k}", f"qu_S_{k}", f"logkvar_{k}", f"logls_{k}", f"sa_{k}", f"logstd_{k}"]
    names += ["sa_sex", "logstd_sex"]
    for l in ("conv1", "conv2", "conv3", "conv4", "conv5", "bn1", "bn3", "bn5", "fc1", "fc2", "fc31", "fc32",
              "fc33", "fc41", "fc42", "fc43", "fc5", "fc6", "fc7", "fc8", "convt1", "convt2", "convt3",
              "convt4", "convt5", "bnt1", "bnt3", "bnt5"):
        names += [l + ".weight", l + ".bias"]
    return names


PARAM_ORDER = param_order()
assert len(PARAM_ORDER) == native.VG_NUM_PARAMS


def hrf_taps() -> np.ndarray:
    """Double-gamma HRF at np.arange(0, 20, 1.4), peak-normalised to 0.6 (utils.py:22-36,
    vae_reg_GP.py:292); gamma.pdf(t, a) = t^(a-1) e^-t / Gamma(a)."""
    t = np.arange(0, 20, 1.4)
    v = t ** 5 * np.exp(-t) / math.factorial(5) - 0.35 * t ** 11 * np.exp(-t) / math.factorial(11)
    return v / v.max() * 0.6


class FlatParams:
    """All parameters in one fp32 buffer (+ one fp64 buffer for epsilon): the layout the fused
    Adam kernel and the NCCL gradient all-reduce work on.  `nn.Parameter.data` / `.grad` are
    views, so state_dict(), checkpoints and user code see ordinary tensors."""

    def __init__(self, named_params: Dict[str, torch.nn.Parameter], device):
        self.names = PARAM_ORDER
        missing = [n for n in self.names if n not in named_params]
        if missing:
            raise ValueError(f"parameters missing from module: {missing}")
        self.params = [named_params[n] for n in self.names]
        self.device = device
        self.version = 0
        self.rebuild()

    def rebuild(self):
        """(Re)pack the current parameter values into the flat buffers."""
        self.version += 1          # pointer tables cached by StepEngine are rebuilt when this changes
        p32 = [p for p in self.params if p.dtype == torch.float32]
        p64 = [p for p in self.params if p.dtype == torch.float64]
        assert len(p32) + len(p64) == len(self.params)
        self.n32 = sum(p.numel() for p in p32)
        self.n64 = sum(p.numel() for p in p64)
        dev = self.device
        self.flat32 = torch.empty(self.n32, dtype=torch.float32, device=dev)
        self.flat64 = torch.empty(self.n64, dtype=torch.float64, device=dev)
        self.grad32 = torch.zeros(self.n32, dtype=torch.float32, device=dev)
        self.grad64 = torch.zeros(self.n64, dtype=torch.float64, device=dev)
        self.slices = {}
        o32 = o64 = 0
        for n, p in zip(self.names, self.params):
            k = p.numel()
            if p.dtype == torch.float32:
                buf, gbuf, off = self.flat32, self.grad32, o32
                o32 += k
            else:
                buf, gbuf, off = self.flat64, self.grad64, o64
                o64 += k
            view = buf[off:off + k].view(p.shape)
            view.copy_(p.data.to(dev))
            p.data = view
            p.grad = None
            self.slices[n] = (p.dtype, off, k)

    def grad_view(self, name):
        dt, off, k = self.slices[name]
        buf = self.grad32 if dt == torch.float32 else self.grad64
        p = self.params[self.names.index(name)]
        return buf[off:off + k].view(p.shape)

    def is_packed(self) -> bool:
        """True while every Parameter still aliases the flat buffer (load_state replaces
        Parameter objects; the model re-packs after that)."""
        for n, p in zip(self.names, self.params):
            dt, off, k = self.slices[n]
            buf = self.flat32 if dt == torch.float32 else self.flat64
            if p.data_ptr() != buf.data_ptr() + off * buf.element_size():
                return False
        return True


class StepBuffers:
    """Workspace + IO table for one minibatch size."""

    def __init__(self, engine: "StepEngine", B: int):
        lib = native.load()
        dev = engine.device
        self.B = B
        self.cfg = native.VgStepConfig()
        self.cfg.b, self.cfg.m = B, engine.m
        self.cfg.neural_covariates = int(engine.neural_covariates)
        self.cfg.want_maps = 0
        self.cfg.gp_kl_scale = float(engine.gp_kl_scale)
        self.cfg.glm_reg_scale = float(engine.glm_reg_scale)
        self.cfg.arith = int(engine.arith)
        self.ws_bytes = int(lib.vg_step_workspace_bytes(C.byref(self.cfg)))
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.scalars = torch.zeros(8, dtype=torch.float64, device=dev)
        self.z = torch.empty(B, NUM_LATENTS, **f32)
        self.maps = torch.zeros(9, B, VP, **f32)
        self.g = torch.empty(8, B, **f32)
        self.beta_mean = torch.empty(8, B, **f32)
        self.beta_var = torch.empty(8, B, **f32)
        self.status = torch.zeros(16, dtype=torch.int32, device=dev)
        self.cons = None
        self.x_rec = None
        self.io = native.VgStepIO()
        io = self.io
        io.out_scalars, io.z, io.maps, io.g = (native.ptr(t) for t in (self.scalars, self.z, self.maps, self.g))
        io.beta_mean, io.beta_var, io.status = native.ptr(self.beta_mean), native.ptr(self.beta_var), native.ptr(self.status)
        io.glm_t, io.taps = native.ptr(engine.glm_t), native.ptr(engine.taps)
        for k in range(6):
            io.xu[k] = native.ptr(engine.xu[k])
        # keep input tensors alive between forward and backward
        self.live = {}

    def ensure_map_outputs(self, dev):
        if self.cons is None:
            self.cons = torch.empty(8, self.B, V, dtype=torch.float32, device=dev)
            self.x_rec = torch.empty(self.B, V, dtype=torch.float32, device=dev)
            self.io.cons, self.io.x_rec = native.ptr(self.cons), native.ptr(self.x_rec)


class StepEngine:
    """Runs the native step for a VAE module (vae_reg_GP.VAE in this package)."""

    def __init__(self, flat: FlatParams, xu: List[torch.Tensor], glm_maps: torch.Tensor, m: int,
                 gp_kl_scale: float, glm_reg_scale: float, neural_covariates: bool, device, arith: int = 0):
        native.require_cuda()
        native.load()
        self.flat = flat
        self.arith = int(arith)     # VG_ARITH_* carried by every call (0 = the library's process default)
        self.device = device
        self.m = int(m)
        self.gp_kl_scale = float(gp_kl_scale)
        self.glm_reg_scale = float(glm_reg_scale)
        self.neural_covariates = bool(neural_covariates)
        self.xu = [t.detach().to(device=device, dtype=torch.float32).contiguous() for t in xu]
        # glm_maps: (V, 9) fp64, col 0 = pandas index (vae_reg_GP.py:58-59, :388) -> (8, VP) fp32
        glm_t = torch.zeros(8, VP, dtype=torch.float32, device=device)
        glm_t[:, :V] = glm_maps[:, 1:9].to(device=device, dtype=torch.float32).t()
        self.glm_t = glm_t.contiguous()
        self.taps = torch.from_numpy(hrf_taps()).to(device=device, dtype=torch.float64)
        self._bufs: Dict[int, StepBuffers] = {}

    def buffers(self, B: int) -> StepBuffers:
        sb = self._bufs.get(B)
        if sb is None:
            sb = StepBuffers(self, B)
            self._bufs[B] = sb
        return sb

    def _bind_params(self, sb: StepBuffers, with_grads: bool):
        """Parameter / gradient pointer tables of the IO struct.  The 2 x 97 pointers only change when the
        flat buffers are re-packed (FlatParams.version) or a Parameter stops aliasing them, so the tables
        are rebuilt only then: the first and last parameter's addresses are the cheap per-step check."""
        ps = self.flat.params
        key = (self.flat.version, ps[0].data_ptr(), ps[-1].data_ptr(), self.flat.grad32.data_ptr(),
               self.flat.grad64.data_ptr())
        if getattr(sb, "bound_key", None) == key:
            return
        for i, p in enumerate(ps):
            sb.io.params[i] = p.data_ptr()
        for i, n in enumerate(self.flat.names):
            dt, off, k = self.flat.slices[n]
            buf = self.flat.grad32 if dt == torch.float32 else self.flat.grad64
            sb.io.grads[i] = buf.data_ptr() + off * buf.element_size()
        sb.bound_key = key

    def draw_noise(self, B: int, generator=None):
        """Same draws, same order as the reference on this device: eps_W (B,1), eps_D (B,32)
        (lowrank_multivariate_normal.py:214-223), then one (B,) per covariate
        (multivariate_normal.py:251-254, vae_reg_GP.py:369)."""
        dev = self.device
        n = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev).normal_(generator=generator)
        eps_w = n(B, 1)
        eps_d = n(B, NUM_LATENTS)
        eps_g = torch.stack([n(B) for _ in range(8)])
        return {"eps_w": eps_w, "eps_d": eps_d, "eps_g": eps_g}

    def forward(self, x: torch.Tensor, covariates: torch.Tensor, noise: dict, want_maps: bool) -> StepBuffers:
        lib = native.load()
        B = x.shape[0]
        sb = self.buffers(B)
        dev = self.device
        x = x.detach().to(device=dev, dtype=torch.float32).reshape(B, V).contiguous()
        cov = covariates.detach().to(device=dev, dtype=torch.float32).contiguous()
        eps_w = noise["eps_w"].detach().to(device=dev, dtype=torch.float32).contiguous()
        eps_d = noise["eps_d"].detach().to(device=dev, dtype=torch.float32).contiguous()
        eps_g = noise["eps_g"].detach().to(device=dev, dtype=torch.float32).contiguous()
        assert cov.shape == (B, 8) and eps_w.shape == (B, 1) and eps_d.shape == (B, NUM_LATENTS) and eps_g.shape == (8, B)
        sb.live = {"x": x, "cov": cov, "eps_w": eps_w, "eps_d": eps_d, "eps_g": eps_g}
        io = sb.io
        io.x, io.covariates = native.ptr(x), native.ptr(cov)
        io.eps_w, io.eps_d, io.eps_g = native.ptr(eps_w), native.ptr(eps_d), native.ptr(eps_g)
        sb.cfg.want_maps = int(want_maps)
        sb.cfg.gp_kl_scale = float(self.gp_kl_scale)
        sb.cfg.glm_reg_scale = float(self.glm_reg_scale)
        sb.cfg.arith = int(self.arith)
        self.flat.last_status = sb.status       # the fused Adam skips its update when these flags are set
        if want_maps:
            sb.ensure_map_outputs(dev)
        self._bind_params(sb, with_grads=False)
        native.check(lib.vg_step_fwd(C.byref(sb.cfg), C.byref(io), native.ptr(sb.workspace), sb.ws_bytes,
                                     native.stream_ptr()), "vg_step_fwd")
        return sb

    def backward(self, sb: StepBuffers, reducer=None):
        """Writes d(tot)/d(param) for every parameter into the flat gradient buffers
        (overwriting them).  With a multi-rank `reducer` (vaegam.dp.GradientAllReduce) the backward runs in
        its three phases and each phase's gradient bucket is all-reduced on the reducer's stream while the
        next phase computes; the caller joins with `reducer.finish()` before the optimizer step."""
        lib = native.load()
        self.flat.grad32.zero_()
        self.flat.grad64.zero_()
        self._bind_params(sb, with_grads=True)
        if reducer is None or reducer.world == 1 or not reducer.overlapped:
            native.check(lib.vg_step_bwd(C.byref(sb.cfg), C.byref(sb.io), native.ptr(sb.workspace), sb.ws_bytes,
                                         native.stream_ptr()), "vg_step_bwd")
            return
        for phase in range(native.VG_BWD_PHASES):
            native.check(lib.vg_step_bwd_phase(C.byref(sb.cfg), C.byref(sb.io), native.ptr(sb.workspace), sb.ws_bytes,
                                               phase, reducer.stream.cuda_stream, native.stream_ptr()),
                         "vg_step_bwd_phase")
            reducer.reduce_phase(phase)


class _StepFn(torch.autograd.Function):
    """tot_loss = VAE-GAM objective; backward hands autograd one gradient view per parameter."""

    @staticmethod
    def forward(ctx, engine: StepEngine, sb_holder: list, x, covariates, noise, want_maps, *params):
        sb = engine.forward(x, covariates, noise, want_maps)
        ctx.engine = engine
        ctx.sb = sb
        sb_holder.append(sb)
        return sb.scalars[:1].to(torch.float32)            # shape (1,), fp32 like the reference

    @staticmethod
    def backward(ctx, grad_out):
        engine, sb = ctx.engine, ctx.sb
        flat = engine.flat
        # Normal training (`optimizer.zero_grad(); loss.backward()`): every .grad is None and autograd
        # simply adopts the views returned below — no copies.  If the caller kept gradients from an
        # earlier backward (accumulation), those ALIAS the flat buffer this call is about to
        # overwrite, so the new gradient is produced in a scratch copy and autograd adds it.
        first = flat.params[0]
        accumulating = first.grad is not None and first.grad.data_ptr() == flat.grad_view(flat.names[0]).data_ptr()
        if accumulating:
            keep32, keep64 = flat.grad32.clone(), flat.grad64.clone()
        engine.backward(sb)
        go = grad_out.reshape(())
        flat.grad32.mul_(go.to(torch.float32))
        flat.grad64.mul_(go.to(torch.float64))
        if accumulating:
            new32, new64 = flat.grad32.clone(), flat.grad64.clone()
            flat.grad32.copy_(keep32)
            flat.grad64.copy_(keep64)
            grads = []
            for n, p in zip(flat.names, flat.params):
                dt, off, k = flat.slices[n]
                grads.append((new32 if dt == torch.float32 else new64)[off:off + k].view(p.shape))
            grads = tuple(grads)
        else:
            grads = tuple(flat.grad_view(n) for n in flat.names)
        return (None, None, None, None, None, None) + grads


def run_step(engine: StepEngine, x, covariates, noise=None, want_maps=False):
    """Returns (tot_loss tensor with grad_fn, StepBuffers)."""
    if noise is None:
        noise = engine.draw_noise(x.shape[0])
    holder = []
    tot = _StepFn.apply(engine, holder, x, covariates, noise, want_maps, *engine.flat.params)
    return tot, holder[0]


def _under_profiler() -> bool:
    """Nsight Compute / Systems inject a library into the process; kernel-replay profiling cannot follow
    launches made during stream capture, so the whole-step graph is not used under them.  (Image-level
    variables such as NV_CUDA_NSIGHT_COMPUTE_VERSION say nothing: only injection hooks count.)"""
    import os
    if any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "CUDA_INJECTION32_PATH", "NV_NSIGHT_INJECTION_PORT_BASE",
                                     "NV_NSIGHT_INJECTION_TRANSPORT_TYPE", "NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NVTX_INJECTION64_PATH",
                                     "NSYS_PROFILING_SESSION_ID")):
        return True
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                low = line.lower()
                if "injection" in low and (".so" in low) and ("nsight" in low or "cuda-injection" in low or "nsys" in low
                                                              or "target" in low):
                    return True
    except OSError:
        pass
    return False


class GraphStep:
    """One whole training step — vg_step_fwd, gradient zeroing, vg_step_bwd, fused Adam — captured once
    per minibatch size as a CUDA graph and replayed.  The native calls are allocation-free,
    sync-free and fork / join their helper streams with events only, so the ~130 launches of a step become
    one `cudaGraphLaunch`; inputs are copied into static buffers and the noise is drawn eagerly into static
    buffers (one launch for all ten arrays; the eager `forward` keeps the reference's ten calls and their order).  The graph holds raw pointers into the flat parameter / gradient / Adam
    buffers: it is rebuilt when they are re-packed (FlatParams.version)."""

    WARMUP = 2      # eager steps before capture: every lazily created stream / event / attribute exists by then
    replayed_launches = 0   # kernels launched through graph replays (vg_launch_count only sees host-side launches)

    def __init__(self, engine: StepEngine, optimizer: "FlatAdam", B: int, reducer=None):
        dev = engine.device
        self.engine, self.opt, self.B = engine, optimizer, B
        self.reducer = reducer          # vaegam.dp.GradientAllReduce (or None): NCCL all-reduce inside the graph
        self.failed = _under_profiler() # True: stay on the eager path
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = torch.zeros(B, V, **f32)
        self.cov = torch.zeros(B, 8, **f32)
        # one flat buffer, three views: the training fast path draws all of a step's noise with ONE launch
        self.noise_flat = torch.zeros(B * (1 + NUM_LATENTS + 8), **f32)
        self.noise = {"eps_w": self.noise_flat[:B].view(B, 1),
                      "eps_d": self.noise_flat[B:B * (1 + NUM_LATENTS)].view(B, NUM_LATENTS),
                      "eps_g": self.noise_flat[B * (1 + NUM_LATENTS):].view(8, B)}
        self.loss = torch.zeros(1, **f32)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.calls = 0
        self.version = engine.flat.version
        self.sb: Optional[StepBuffers] = None
        self.baked = None

    def signature(self):
        """Every host-side scalar the captured kernels received by value: a change (learning-rate schedule,
        loss weights, gradient scale) needs a new capture."""
        g = self.opt.param_groups[0]
        e = self.engine
        return (e.gp_kl_scale, e.glm_reg_scale, e.neural_covariates, e.m, e.arith, float(g["lr"]), tuple(g["betas"]),
                float(g["eps"]), float(self.opt.grad_scale))

    def _body(self):
        eng = self.engine
        sb = eng.forward(self.x, self.cov, self.noise, False)
        red = self.reducer
        if red is not None and red.world > 1 and red.overlapped:
            eng.backward(sb, red)       # bucketed all-reduce on the reducer's stream, overlapped with the backward
            red.finish()                # join before Adam; 1/world is folded into Adam's grad_scale
        else:
            eng.backward(sb)
            if red is not None:
                red.inline()            # sum over ranks on this stream
        self.opt.launch(sb.status)
        self.loss.copy_(sb.scalars[:1])
        self.sb = sb

    def draw(self, generator=None):
        """i.i.d. N(0,1) for eps_W, eps_D and the eight gain draws in one launch.  (The eager `forward` draws them
        with the reference's ten calls, in its order — `StepEngine.draw_noise`; the values of a given seed differ,
        the distribution does not.)"""
        self.noise_flat.normal_(generator=generator)

    def run(self, x, covariates, noise=None) -> torch.Tensor:
        B = self.B
        self.x.copy_(x.reshape(B, V), non_blocking=True)
        self.cov.copy_(covariates, non_blocking=True)
        if noise is None:
            self.draw()
        else:
            for k in ("eps_w", "eps_d", "eps_g"):
                self.noise[k].copy_(noise[k])
        self.calls += 1
        if self.graph is not None and self.baked != self.signature():
            self.graph = None           # a by-value kernel argument changed: capture again
        if self.graph is None and not self.failed and self.calls > self.WARMUP:
            try:
                g = torch.cuda.CUDAGraph()
                n0 = native.launch_count()
                with torch.cuda.graph(g):
                    self._body()
                self.launches_per_replay = native.launch_count() - n0
                self.graph = g          # capture only records; the replay below is this call's step
                self.baked = self.signature()
            except RuntimeError as e:   # e.g. a collective that cannot be captured: stay on the eager path
                import warnings
                warnings.warn(f"whole-step CUDA graph capture failed, running eagerly: {e}")
                self.failed = True
                torch.cuda.synchronize()
        if self.graph is not None:
            self.graph.replay()
            GraphStep.replayed_launches += self.launches_per_replay
        else:
            self._body()
        self.opt._host_steps += 1
        return self.loss.clone()        # the static buffer is overwritten by the next step


class FlatAdam(torch.optim.Adam):
    """torch.optim.Adam whose step() is ONE fused native kernel over the flat buffers
    (vg_adam_step).  State tensors (`exp_avg`, `exp_avg_sq`, `step`) are views of flat
    buffers, so state_dict()/load_state_dict() keep the reference's checkpoint layout
    (vae_reg_GP.py:457,480)."""

    def __init__(self, flat: FlatParams, lr=1e-3):
        super().__init__(flat.params, lr=lr)
        self.flat = flat
        self.grad_scale = 1.0
        self._alloc_state()

    def _alloc_state(self):
        f = self.flat
        dev = f.device
        self.m32 = torch.zeros(f.n32, dtype=torch.float32, device=dev)
        self.v32 = torch.zeros(f.n32, dtype=torch.float32, device=dev)
        self.m64 = torch.zeros(f.n64, dtype=torch.float64, device=dev)
        self.v64 = torch.zeros(f.n64, dtype=torch.float64, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._host_steps = 0

    def _materialise_state(self):
        """Expose the flat moments as per-parameter optimizer state (views)."""
        f = self.flat
        for n, p in zip(f.names, f.params):
            dt, off, k = f.slices[n]
            m, v = (self.m32, self.v32) if dt == torch.float32 else (self.m64, self.v64)
            self.state[p] = {"step": torch.tensor(float(self._host_steps)),
                             "exp_avg": m[off:off + k].view(p.shape),
                             "exp_avg_sq": v[off:off + k].view(p.shape)}

    def state_dict(self):
        if self._host_steps > 0:
            self._materialise_state()
        return super().state_dict()

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        f = self.flat
        steps = 0
        for n, p in zip(f.names, f.params):
            st = self.state.get(p)
            if not st:
                continue
            dt, off, k = f.slices[n]
            m, v = (self.m32, self.v32) if dt == torch.float32 else (self.m64, self.v64)
            m[off:off + k].copy_(st["exp_avg"].reshape(-1).to(m.device))
            v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1).to(v.device))
            steps = max(steps, int(float(st["step"])))
        self._host_steps = steps
        self.step_count.fill_(steps)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.flat.params:
            p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closure is not supported")
        f = self.flat
        # gradients normally ARE views of the flat gradient buffer (see _StepFn.backward);
        # anything else (user-modified .grad) is gathered into it first.
        for n, p in zip(f.names, f.params):
            gv = f.grad_view(n)
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
        self.launch(getattr(f, "last_status", None))
        self._host_steps += 1

    @torch.no_grad()
    def launch(self, status=None):
        """The fused Adam kernel over the flat buffers as they are (graph-capturable: the step count lives on
        the device).  `status`: the step's 16-int device flags; when any of the 8 gain flags is set (a
        covariance was not positive definite) the kernel leaves parameters, moments and step count untouched —
        the reference raises inside forward before backward()/step() (vae_reg_GP.py:368, gp.py:51)."""
        f = self.flat
        lib = native.load()
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        native.check(lib.vg_adam_step(native.ptr(f.flat32), native.ptr(f.grad32), native.ptr(self.m32),
                                      native.ptr(self.v32), f.n32, native.ptr(f.flat64), native.ptr(f.grad64),
                                      native.ptr(self.m64), native.ptr(self.v64), f.n64, float(g["lr"]), float(b1),
                                      float(b2), float(g["eps"]), float(self.grad_scale),
                                      native.ptr(self.step_count), native.ptr(status), 8 if status is not None else 0,
                                      native.stream_ptr()), "vg_adam_step")
