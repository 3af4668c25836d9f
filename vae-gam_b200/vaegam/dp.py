"""Data-parallel training over minibatches: one process per GPU, NCCL over
NVLink/NVSwitch.  The independent unit is a MINIBATCH (BatchNorm statistics, the HRF filter
over the batch index and the B x B gain covariance couple the volumes of a batch), so each
rank runs the full step on its own batch, BatchNorm is not synchronised, and the only
exchange is the gradient all-reduce: 6.5 MB per step (1 494 109 fp32 + 70 315 fp64) taken
straight from the flat gradient buffers — three contiguous buckets, no bucketing copies — and
overlapped with the backward (GradientAllReduce).

N-GPU DP with local batch B equals the AVERAGE of N independent reference steps of batch B
(the objective is not a per-sample mean: glm_reg ~ B * sum_b, GP KL is per batch).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_indices(n_items: int, rank: int, world: int, epoch: int = 0, shuffle: bool = True, seed: int = 0):
    """Disjoint, equally sized shards of range(n_items) (DistributedSampler-style, drop tail)."""
    g = torch.Generator().manual_seed(seed + epoch)
    perm = torch.randperm(n_items, generator=g) if shuffle else torch.arange(n_items)
    per = n_items // world
    return perm[rank * per:(rank + 1) * per]


class GradientAllReduce:
    """Sum-all-reduce of the flat gradient buffers; the 1/world factor is folded into the fused
    Adam kernel (`grad_scale`).

    Overlap with the backward: the flat fp32 gradient is cut into three contiguous buckets
    in the order the backward completes them — the flat layout is the reference's parameter order, so the
    decoder's parameters (fc5.. bnt5) are its tail, the encoder's fully connected layers (fc1..fc43) its
    middle, and the gain parameters + encoder convolutions its head:

        phase 0 (objective + decoder backward)   -> epsilon (fp64 buffer) + flat32[fc5.weight:]     3.3 + 0.56 MB
        phase 1 (latent + encoder FC backward)   -> flat32[fc1.weight : fc5.weight]                2.6 MB
        phase 2 (encoder convolutions, gains)    -> flat32[: fc1.weight]                           0.1 MB

    `vg_step_bwd_phase` makes `self.stream` wait for the producers of a phase (events only, helper stream
    included) and `reduce_phase` enqueues that bucket's `ncclAllReduce` there, so it runs while the next phase
    computes; only the small last bucket is exposed.  The same sequence is what a whole-step CUDA graph
    captures (`vaegam.step.GraphStep`).  `overlap=False` (or a CPU/gloo group) gives the plain
    reduce-after-backward."""

    def __init__(self, flat, optimizer=None, group=None, overlap: bool = True):
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.optimizer = optimizer
        if optimizer is not None:
            optimizer.grad_scale = 1.0 / self.world
        self.cuda = flat.grad32.is_cuda
        self.stream = torch.cuda.Stream() if (self.cuda and overlap) else None
        self.overlapped = self.stream is not None
        self._pending = None
        self._buckets_for = None

    def buckets(self):
        """[[tensors of phase 0], [phase 1], [phase 2]] — views of the flat gradient buffers (rebuilt when the
        flat buffers were re-packed)."""
        f = self.flat
        key = (f.version, f.grad32.data_ptr(), f.grad64.data_ptr())
        if self._buckets_for != key:
            o_fc1, o_fc5 = f.slices["fc1.weight"][1], f.slices["fc5.weight"][1]
            assert 0 < o_fc1 < o_fc5 < f.n32
            self._buckets = [[f.grad64, f.grad32[o_fc5:]], [f.grad32[o_fc1:o_fc5]], [f.grad32[:o_fc1]]]
            self._buckets_for = key
        return self._buckets

    def broadcast_parameters(self, src: int = 0):
        if self.world == 1:
            return
        dist.broadcast(self.flat.flat32, src, group=self.group)
        dist.broadcast(self.flat.flat64, src, group=self.group)

    def reduce_phase(self, phase: int):
        """All-reduce the bucket that backward phase `phase` completed, on `self.stream` (which
        vg_step_bwd_phase has already made wait for the bucket's producers)."""
        if self.world == 1:
            return
        with torch.cuda.stream(self.stream):
            for t in self.buckets()[phase]:
                dist.all_reduce(t, group=self.group)
        self._pending = True

    def start(self):
        if self.world == 1:
            return
        bufs = (self.flat.grad32, self.flat.grad64)
        if self.stream is None:
            for b in bufs:
                dist.all_reduce(b, group=self.group)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for b in bufs:
                dist.all_reduce(b, group=self.group)
        self._pending = True

    def inline(self):
        """All-reduce on the CURRENT stream (no side stream, no events)."""
        if self.world == 1:
            return
        for b in (self.flat.grad32, self.flat.grad64):
            dist.all_reduce(b, group=self.group)

    def finish(self):
        """The current stream waits for every all-reduce enqueued on the reducer's stream."""
        if self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
            self._pending = None

    def __call__(self):
        self.start()
        self.finish()


def train_step(model, reducer: GradientAllReduce, ids, covariates, x, noise=None):
    """forward + backward + gradient all-reduce + fused Adam; returns the local loss tensor.  With
    `model.use_cuda_graph` the whole sequence (all-reduce included) is one CUDA-graph launch per step."""
    if getattr(model, "use_cuda_graph", False):
        return model.train_batch(ids, covariates, x, _noise=noise, reducer=reducer)
    loss = model.forward(ids, covariates, x, 'train', train_mode=False, _noise=noise)
    model.optimizer.zero_grad()
    loss.backward()
    reducer()
    model.optimizer.step()
    return loss
