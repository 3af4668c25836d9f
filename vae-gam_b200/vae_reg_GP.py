"""Drop-in replacement for the reference's `vae_reg_GP.py` (dannyfa/VAE-GAM) on B200.

Same module name, same `VAE` class surface (ctor kwargs, parameter names, `forward`,
`encode`, `decode`, `train_epoch`, `test_epoch`, `train_loop`, `save_state`, `load_state`,
`project_latent`, `reconstruct`, `plot_GPs`; reference vae_reg_GP.py:35-715), same checkpoint
layout — but the arithmetic of the training step runs as hand-written sm_100a CUDA kernels
behind the C ABI in include/vaegam.h (libvaegam_sm100.so).  The nn.Module layers created
here are parameter CONTAINERS only (they give the reference's init RNG order, state_dict
keys and checkpoint layout); their PyTorch forward is never called.

There is no CPU path: the constructor works anywhere (so checkpoints can be inspected), but
`forward/encode/decode` raise without a CUDA device and the built library.

Deliberate differences from the reference:
  * gains are computed in fp64 (the reference's fp32 `torch.inverse(Ku)` is ill-conditioned);
  * TensorBoard image/figure logging inside `forward` is decimated: `self.log_every`
    (default 0 = off; the reference logs every step, costing >1 s/step);
  * the 10 per-step D2H map copies happen only when `return_latent_rec=True`;
  * `load_state` re-registers the loaded parameters with the optimizer (the reference
    silently freezes epsilon and all GP parameters after a resume).
"""
from __future__ import annotations

import datetime
import itertools
import os

import numpy as np
import pandas as pd
import torch
from torch import nn

import gp
import utils
from vaegam import native
from vaegam.step import FlatAdam, FlatParams, GP_KEYS, GraphStep, IMG_KEYS, StepEngine, run_step

IMG_SHAPE = (41, 49, 35)
IMG_DIM = int(np.prod(IMG_SHAPE))
_GP_COVS = GP_KEYS[1:7]


class _NullWriter:
    def __getattr__(self, name):
        return lambda *a, **k: None


def _make_writer(log_dir):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=log_dir)
    except Exception:  # tensorboard missing: logging is optional
        return _NullWriter()


class VAE(nn.Module):
    def __init__(self, nf=8, save_dir='', lr=1e-3, num_covariates=8, num_latents=32, device_name="auto",
                 num_inducing_pts=6, gp_kl_scale=10.0, glm_maps='', glm_reg_scale=1.0, csv_files='',
                 neural_covariates=True):
        super().__init__()
        if nf != 8 or num_covariates != 8 or num_latents != 32:
            raise ValueError("the sm_100a kernels are specialised for nf=8, num_covariates=8, num_latents=32 "
                             "(the only configuration the reference's CLI can produce)")
        self.nf, self.save_dir, self.lr = nf, save_dir, lr
        self.num_covariates, self.num_latents = num_covariates, num_latents
        self.neural_covariates = neural_covariates
        self.z_dim = num_latents + num_covariates + 1
        assert device_name != "cuda" or torch.cuda.is_available()
        if device_name == "auto":
            device_name = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device_name)
        if self.save_dir != '' and not os.path.exists(self.save_dir):
            os.makedirs(self.save_dir)
        dev = self.device
        # per-voxel log-precision map, fp64 like the reference (vae_reg_GP.py:54-56)
        self.epsilon = nn.Parameter(torch.full(IMG_SHAPE, -np.log(10), dtype=torch.float64, device=dev))
        self.glm_maps = torch.from_numpy(pd.read_csv(glm_maps).to_numpy()).to(dev)      # (V, 9): index + 8 maps
        if self.glm_maps.shape != (IMG_DIM, num_covariates + 1):
            raise ValueError(f"glm_maps must have {IMG_DIM} rows and index + 8 columns, got {tuple(self.glm_maps.shape)}")
        self.glm_reg_scale = glm_reg_scale
        self.inducing_pts = num_inducing_pts
        self.gp_kl_scale = torch.as_tensor(gp_kl_scale).to(dev)
        self.max_ls = torch.as_tensor(3.0).to(dev)
        xu_ranges = utils.get_xu_ranges(csv_files)
        # gain parameters, created in the reference's order so a given torch seed yields the
        # same initial values (vae_reg_GP.py:68-172): binary covariates get (sa, logstd) only
        self.gp_params = {k: {} for k in GP_KEYS}
        m = self.inducing_pts

        def lin_pair(key):
            sa = nn.Parameter(torch.normal(1, 1, size=(1, 1)).to(dev))
            ls = nn.Parameter(torch.normal(0, 1, size=(1, 1)).to(dev))
            setattr(self, "sa_" + key, sa)
            self.gp_params[key]['sa'] = sa
            setattr(self, "logstd_" + key, ls)
            self.gp_params[key]['logstd'] = ls

        lin_pair('task')
        for j, key in enumerate(_GP_COVS):
            xu = torch.linspace(xu_ranges[j][0], xu_ranges[j][1], m).to(dev)
            setattr(self, "xu_" + key, xu)
            self.gp_params[key]['xu'] = xu
            qm = nn.Parameter(torch.normal(0.0, 1.0, size=[1, m]).to(dev))
            setattr(self, "qu_m_" + key, qm)
            self.gp_params[key]['qu_m'] = qm
            qs = nn.Parameter(2 * torch.eye(m).to(dev))
            setattr(self, "qu_S_" + key, qs)
            self.gp_params[key]['qu_S'] = qs
            lk = nn.Parameter(torch.as_tensor(0.0).to(dev))
            setattr(self, "logkvar_" + key, lk)
            self.gp_params[key]['logkvar'] = lk
            ll = nn.Parameter(torch.as_tensor(0.0).to(dev))
            setattr(self, "logls_" + key, ll)
            self.gp_params[key]['log_ls'] = ll
            lin_pair(key)
        lin_pair('sex')
        self._build_network()
        self.epoch = 0
        self.loss = {'train': {}, 'test': {}}
        ts = datetime.datetime.now().date()
        self.writer = _make_writer(os.path.join(self.save_dir, 'run', ts.strftime('%m_%d_%Y')))
        self.to(dev)
        # TensorBoard hooks inside forward(): every `log_every` training steps (0 = never)
        self.log_every = int(os.environ.get("VAEGAM_TB_EVERY", "0"))
        self._step_counter = 0
        # arithmetic of the convolutions, carried by every native call: "default" (the library's process default,
        # VAEGAM_CONV_MODE / vg_set_conv_mode: mixed unless set), "fp32", "bf16" or "mixed" (include/vaegam.h)
        self.arith = os.environ.get("VAEGAM_ARITH", "default")
        # whole-step CUDA graphs in train_batch / train_epoch (VAEGAM_CUDA_GRAPH=0 switches them off)
        self.use_cuda_graph = os.environ.get("VAEGAM_CUDA_GRAPH", "1") != "0"
        self._graph_steps = {}
        self._engine = None
        self._flat = None
        self._last = None
        self._pack()

    # ------------------------------------------------------------------ structure
    def _build_network(self):
        nf = self.nf
        # encoder (parameter containers; forward is native)
        self.conv1 = nn.Conv3d(1, nf, 3, 1)
        self.conv2 = nn.Conv3d(nf, nf, 3, 2)
        self.conv3 = nn.Conv3d(nf, 2 * nf, 3, 1)
        self.conv4 = nn.Conv3d(2 * nf, 2 * nf, 3, 2)
        self.conv5 = nn.Conv3d(2 * nf, 2 * nf, 3, 1)
        self.bn1 = nn.BatchNorm3d(1, track_running_stats=False)
        self.bn3 = nn.BatchNorm3d(nf, track_running_stats=False)
        self.bn5 = nn.BatchNorm3d(2 * nf, track_running_stats=False)
        self.fc1 = nn.Linear(2 * nf * 6 * 8 * 4, 200)
        self.fc2 = nn.Linear(200, 100)
        for name in ("fc31", "fc32", "fc33"):
            setattr(self, name, nn.Linear(100, 50))
        for name in ("fc41", "fc42", "fc43"):
            setattr(self, name, nn.Linear(50, self.num_latents))
        # decoder
        self.fc5 = nn.Linear(self.z_dim, 50)
        self.fc6 = nn.Linear(50, 100)
        self.fc7 = nn.Linear(100, 200)
        self.fc8 = nn.Linear(200, 2 * nf * 6 * 8 * 5)
        self.convt1 = nn.ConvTranspose3d(2 * nf, 2 * nf, 3, 1)
        self.convt2 = nn.ConvTranspose3d(2 * nf, 2 * nf, 3, 2, padding=(1, 0, 1), output_padding=(1, 0, 1))
        self.convt3 = nn.ConvTranspose3d(2 * nf, nf, 3, 1)
        self.convt4 = nn.ConvTranspose3d(nf, nf, (5, 3, 3), 2)
        self.convt5 = nn.ConvTranspose3d(nf, 1, 3, 1)
        self.bnt1 = nn.BatchNorm3d(2 * nf, track_running_stats=False)
        self.bnt3 = nn.BatchNorm3d(2 * nf, track_running_stats=False)
        self.bnt5 = nn.BatchNorm3d(nf, track_running_stats=False)

    _LAYER_NAMES = ['fc1', 'fc2', 'fc31', 'fc32', 'fc33', 'fc41', 'fc42', 'fc43', 'fc5', 'fc6', 'fc7', 'fc8',
                    'bn1', 'bn3', 'bn5', 'bnt1', 'bnt3', 'bnt5', 'conv1', 'conv2', 'conv3', 'conv4', 'conv5',
                    'convt1', 'convt2', 'convt3', 'convt4', 'convt5']

    def _get_layers(self):
        """name -> layer, the 28 per-layer state_dict keys of the checkpoint (vae_reg_GP.py:220-234)."""
        return {n: getattr(self, n) for n in self._LAYER_NAMES}

    def _pack(self):
        """Move all parameters into the flat buffers and (re)build optimizer + native engine."""
        named = dict(self.named_parameters())
        old_opt = getattr(self, "optimizer", None)
        self._flat = FlatParams(named, self.device)
        new_opt = FlatAdam(self._flat, lr=self.lr)
        same_layout = old_opt is not None and getattr(old_opt, "flat", None) is not None and \
            old_opt.flat.slices == self._flat.slices
        if same_layout and getattr(old_opt, "_host_steps", 0) > 0:      # moments only survive an identical layout
            new_opt.m32.copy_(old_opt.m32); new_opt.v32.copy_(old_opt.v32)
            new_opt.m64.copy_(old_opt.m64); new_opt.v64.copy_(old_opt.v64)
            new_opt.step_count.copy_(old_opt.step_count)
            new_opt._host_steps = old_opt._host_steps
        self.optimizer = new_opt
        self._engine = None
        self._graph_steps = {}

    def _get_engine(self) -> StepEngine:
        native.require_cuda()
        if self.device.type != "cuda":
            raise native.NativeError("model is on CPU; the VAE-GAM hot path needs a CUDA device")
        if self._flat is None or not self._flat.is_packed():
            self._pack()
        if self._engine is None:
            xu = [self.gp_params[k]['xu'] for k in _GP_COVS]
            self._engine = StepEngine(self._flat, xu, self.glm_maps, self.inducing_pts, self._as_float("gp_kl_scale"),
                                      self._as_float("glm_reg_scale"), self.neural_covariates, self.device)
        self._engine.arith = native.ARITH_BY_NAME[self.arith] if isinstance(self.arith, str) else int(self.arith)
        self._engine.gp_kl_scale = self._as_float("gp_kl_scale")
        self._engine.glm_reg_scale = self._as_float("glm_reg_scale")
        self._engine.neural_covariates = bool(self.neural_covariates)
        return self._engine

    def _as_float(self, attr):
        """Host value of a loss weight.  `gp_kl_scale` is a device tensor like the reference's (vae_reg_GP.py:64)
        and float() of it synchronises the device, so the conversion is cached per object identity: it is
        redone only when the attribute is re-assigned (ctor, load_state, user code)."""
        obj = getattr(self, attr)
        cache = self.__dict__.setdefault("_float_cache", {})
        hit = cache.get(attr)
        if hit is None or hit[0] is not obj:
            hit = (obj, float(obj))
            cache[attr] = hit
        return hit[1]

    # ------------------------------------------------------------------ public pieces
    def encode(self, x):
        """x -> (mu (B,32), u (B,32,1), d (B,32)); no autograd (use forward for training)."""
        import ctypes as C
        eng = self._get_engine()
        x = x.detach().to(self.device, torch.float32).reshape(-1, IMG_DIM).contiguous()
        B = x.shape[0]
        sb = eng.buffers(B)
        eng._bind_params(sb, with_grads=False)
        sb.io.x = native.ptr(x)
        heads = torch.empty(3, B, self.num_latents, dtype=torch.float32, device=self.device)
        native.check(native.load().vg_encode_fwd(C.byref(sb.cfg), C.byref(sb.io), native.ptr(heads),
                                                 native.ptr(sb.workspace), sb.ws_bytes, native.stream_ptr()),
                     "vg_encode_fwd")
        return heads[0], heads[1].unsqueeze(-1), torch.exp(heads[2])

    def decode(self, z):
        """z (n, 41) -> (n, 70315); the n rows form one BatchNorm batch, as in the reference."""
        import ctypes as C
        eng = self._get_engine()
        z = z.detach().to(self.device, torch.float32).reshape(-1, self.z_dim).contiguous()
        n = z.shape[0]
        lib = native.load()
        nbytes = int(lib.vg_decode_workspace_bytes(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        sb = eng.buffers(1)
        eng._bind_params(sb, with_grads=False)
        out = torch.empty(n, IMG_DIM, dtype=torch.float32, device=self.device)
        native.check(lib.vg_decode_fwd(C.byref(sb.io), native.ptr(z), n, int(eng.arith), native.ptr(out), native.ptr(ws),
                                       nbytes, native.stream_ptr()), "vg_decode_fwd")
        return out

    def calc_linW_KL(self, sa, std):
        """KL(N(sa, std^2) || N(1, 0.5^2)) (reference vae_reg_GP.py:266-281)."""
        return torch.log(0.5 / std) + (std ** 2 + (sa - 1.0) ** 2) / (2 * 0.25) - 0.5

    def do_hrf_conv(self, covariate_vals):
        """Causal 15-tap HRF filter over the batch index (reference vae_reg_GP.py:283-305)."""
        taps = torch.as_tensor(utils.hrf(np.arange(0, 20, 1.4)), dtype=covariate_vals.dtype,
                               device=covariate_vals.device)
        n = covariate_vals.shape[0]
        padded = torch.cat([covariate_vals.new_zeros(taps.numel() - 1), covariate_vals])
        return torch.nn.functional.conv1d(padded.view(1, 1, -1), taps.flip(0).view(1, 1, -1)).view(-1)[:n]

    # ------------------------------------------------------------------ the hot path
    def forward(self, ids, covariates, x, log_type, return_latent_rec=False, train_mode=True, _noise=None):
        """Objective of vae_reg_GP.py:307-413.  Returns tot_loss (shape (1,), fp32, differentiable
        w.r.t. all 97 parameters) or (tot_loss, z ndarray (B,32), imgs dict of (B,70315) float32).
        `_noise` = {'eps_w','eps_d','eps_g'} injects the random draws (parity tests)."""
        eng = self._get_engine()
        B = ids.shape[0]
        log_now = bool(train_mode) and self.log_every > 0 and (self._step_counter % self.log_every == 0)
        if train_mode:
            self._step_counter += 1
        want_maps = bool(return_latent_rec) or log_now
        tot, sb = run_step(eng, x.reshape(B, -1), covariates, _noise, want_maps)
        self._last = sb
        imgs = None
        if want_maps:
            imgs = {'base': sb.maps[0, :, :IMG_DIM].cpu().numpy()}
            cons = sb.cons.cpu().numpy()
            for i in range(1, self.num_covariates + 1):
                imgs[IMG_KEYS[i]] = cons[i - 1]
            imgs['full_rec'] = sb.x_rec.cpu().numpy()
        if log_now:
            self._log_step(sb, imgs, covariates, B, log_type)
        if return_latent_rec:
            return tot, sb.z.cpu().numpy(), imgs
        return tot

    def _log_step(self, sb, imgs, covariates, B, log_type):
        """The reference's in-forward TensorBoard hooks (vae_reg_GP.py:333-337,372,384-386,396-398)."""
        for name, key in (('base_map', 'base'), ('task_map', 'task'), ('full_reconstruction', 'full_rec')):
            for slc in (12, 15, 18):
                utils.log_map(self.writer, IMG_SHAPE, imgs[key], slc, name, B, log_type)
        mean, var = sb.beta_mean.cpu(), sb.beta_var.cpu()
        for i, key in enumerate(GP_KEYS):
            utils.log_beta(self.writer, covariates[:, i].detach().cpu(), mean[i], torch.diag(var[i]), key, log_type)

    def check_status(self):
        """Raise if the last step met a non-positive-definite covariance (the reference raises
        ValueError from torch's constraint check at vae_reg_GP.py:368 / gp.py:51).  Reads a
        16-int device flag, so it synchronises; train_epoch calls it right after loss.item()."""
        if self._last is None:
            return
        st = self._last.status.cpu().tolist()
        for i, s in enumerate(st[:8]):
            if s:
                what = "qu_S" if s >= 1000 else "gain covariance"
                raise ValueError(f"covariate '{GP_KEYS[i]}': {what} is not positive definite "
                                 f"(pivot {s % 1000})")

    def _batch(self, sample):
        dev = self.device
        return sample['subjid'].to(dev), sample['covariates'].to(dev), sample['volume'].to(dev)

    def train_batch(self, ids, covariates, x, _noise=None, reducer=None):
        """forward + backward + Adam step for one minibatch (the body of the reference's train_epoch loop,
        vae_reg_GP.py:421-429) as ONE CUDA-graph launch per step after two eager warm-up steps; returns the
        loss tensor (shape (1,), no autograd graph).  Falls back to the eager three-call sequence when graphs
        are switched off (`self.use_cuda_graph = False` / VAEGAM_CUDA_GRAPH=0) or when this step logs to
        TensorBoard.  `reducer` (vaegam.dp.GradientAllReduce) puts the NCCL gradient all-reduce between
        backward and Adam — inside the graph too."""
        eng = self._get_engine()
        B = ids.shape[0]
        log_now = self.log_every > 0 and (self._step_counter % self.log_every == 0)
        if not self.use_cuda_graph or log_now:
            loss = self.forward(ids, covariates, x, 'train', train_mode=True, _noise=_noise)
            self.optimizer.zero_grad()
            loss.backward()
            if reducer is not None:
                reducer()
            self.optimizer.step()
            return loss.detach()
        self._step_counter += 1
        gs = self._graph_steps.get(B)
        if gs is None or gs.version != self._flat.version or gs.engine is not eng or gs.opt is not self.optimizer \
                or gs.reducer is not reducer:
            gs = GraphStep(eng, self.optimizer, B, reducer)
            self._graph_steps[B] = gs
        gs.opt.grad_scale = self.optimizer.grad_scale
        loss = gs.run(x.to(self.device), covariates.to(self.device), _noise)
        self._last = gs.sb
        for p in self._flat.params:         # gradients live in the flat buffer; nothing is left in .grad
            p.grad = None
        return loss

    def train_epoch(self, train_loader):
        self.train()
        train_loss = 0.0
        for sample in train_loader:
            ids, covariates, x = self._batch(sample)
            loss = self.train_batch(ids, covariates, x)
            train_loss += loss.item()
            self.check_status()
        train_loss /= len(train_loader.dataset)
        print('Epoch: {} Average loss: {:.4f}'.format(self.epoch, train_loss))
        self.epoch += 1
        return train_loss

    def test_epoch(self, test_loader):
        self.eval()
        test_loss = 0.0
        with torch.no_grad():
            for sample in test_loader:
                ids, covariates, x = self._batch(sample)
                test_loss += self.forward(ids, covariates, x, 'test', train_mode=False).item()
        test_loss /= len(test_loader.dataset)
        print('Test loss: {:.4f}'.format(test_loss))
        return test_loss

    # ------------------------------------------------------------------ checkpoints
    def save_state(self, filename):
        """Same dict layout as the reference (vae_reg_GP.py:452-471)."""
        state = {name: layer.state_dict() for name, layer in self._get_layers().items()}
        state.update({
            'optimizer_state': self.optimizer.state_dict(), 'loss': self.loss, 'z_dim': self.z_dim,
            'epoch': self.epoch, 'lr': self.lr, 'save_dir': self.save_dir, 'epsilon': self.epsilon,
            'glm_reg_scale': self.glm_reg_scale, 'gp_kl_scale': self.gp_kl_scale,
            'inducing_pts': self.inducing_pts, 'gp_params': self.gp_params})
        torch.save(state, os.path.join(self.save_dir, filename))

    def load_state(self, filename):
        """Reads checkpoints written by this class or by the reference (vae_reg_GP.py:473-539)."""
        try:
            ckpt = torch.load(filename, map_location=self.device)
        except Exception:
            ckpt = torch.load(filename, map_location=self.device, weights_only=False)
        assert ckpt['z_dim'] == self.z_dim
        for name, layer in self._get_layers().items():
            layer.load_state_dict(ckpt[name])
        self.loss, self.epoch = ckpt['loss'], ckpt['epoch']
        self.glm_reg_scale, self.gp_kl_scale = ckpt['glm_reg_scale'], ckpt['gp_kl_scale']
        self.inducing_pts = ckpt['inducing_pts']
        replaced = False
        with torch.no_grad():
            self.epsilon.copy_(ckpt['epsilon'].to(self.device))
            for key, entry in ckpt['gp_params'].items():
                for pname, val in entry.items():
                    if pname == 'xu':
                        xu = val.detach().to(self.device)
                        setattr(self, "xu_" + key, xu)
                        self.gp_params[key]['xu'] = xu
                        continue
                    attr = {'log_ls': 'logls_'}.get(pname, pname + '_') + key
                    cur = getattr(self, attr)
                    new = val.detach().to(self.device)
                    if cur.shape != new.shape:       # checkpoint trained with another num_inducing_pts
                        cur = nn.Parameter(new.clone())
                        setattr(self, attr, cur)
                        self.gp_params[key][pname] = cur
                        replaced = True
                    else:
                        cur.copy_(new)
        # values were copied INTO the registered parameters (they keep aliasing the flat buffer when
        # shapes match), so the optimizer keeps owning them — unlike the reference
        # A replaced Parameter (other num_inducing_pts) changes the flat layout: FlatParams still lists the OLD
        # objects, so is_packed() cannot see it — re-pack from named_parameters() unconditionally in that case.
        if replaced or not self._flat.is_packed():
            self._pack()
        self.optimizer.load_state_dict(ckpt['optimizer_state'])
        self._engine = None
        self._graph_steps = {}

    # ------------------------------------------------------------------ callers of the hot path
    def project_latent(self, loaders_dict, save_dir, title=None, split=98):
        """Encode the unshuffled train set; UMAP scatter when umap/matplotlib exist, and always
        a CSV of the latent means (reference vae_reg_GP.py:542-583)."""
        loader = loaders_dict['UnShuffled_train']
        latent = np.zeros((len(loader.dataset), self.num_latents))
        j = 0
        with torch.no_grad():
            for sample in loader:
                mu, _, _ = self.encode(sample['volume'].to(self.device))
                latent[j:j + len(mu)] = mu.cpu().numpy()
                j += len(mu)
        stem = os.path.join(save_dir, str(self.epoch).zfill(3))
        np.savetxt(stem + '_latent_means.csv', latent, delimiter=',')
        try:
            from umap import UMAP
            import matplotlib.pyplot as plt
        except ImportError:
            return latent
        proj = UMAP(n_components=2, n_neighbors=20, min_dist=0.1, metric='euclidean',
                    random_state=42).fit_transform(latent)
        colors = itertools.cycle(['b', 'g', 'r', 'c', 'm', 'y', 'k', 'orange', 'blueviolet', 'hotpink', 'lime',
                                  'skyblue', 'teal', 'sienna'])
        for i in range(0, len(latent), split):
            plt.scatter(proj[i:i + split, 0], proj[i:i + split, 1], color=next(colors), s=1.0, alpha=0.6)
            plt.axis('off')
        if title is not None:
            plt.title(title)
        plt.savefig(stem + '_temp.pdf')
        return latent

    def reconstruct(self, loader, ref_niis, save_dirs):
        """Per-volume NIfTI maps for every key of `imgs` (reference vae_reg_GP.py:585-620), same file tree.

        B200-first: the 10 maps of a batch are the optional outputs of the fused
        reconstruction pass and stay on the device as one (10, B, V) block; per-subject sums are
        accumulated there (fp64 index_add) so that `build_model_recons.mk_avg_maps` does not have to
        read the 10 x N files back; one asynchronous D2H per batch into alternating pinned buffers,
        and the files of batch i are written by a small thread pool while batch i+1 computes."""
        from vaegam.nib_compat import nib
        from concurrent.futures import ThreadPoolExecutor
        eng = self._get_engine()
        n_subj = len(save_dirs)
        sums = torch.zeros(n_subj, len(IMG_KEYS), IMG_DIM, dtype=torch.float64, device=self.device)
        counts = torch.zeros(n_subj, dtype=torch.float64, device=self.device)
        ref_cache = {}
        pinned = [None, None]
        pending = [[], []]
        copied = [torch.cuda.Event(), torch.cuda.Event()]

        def write_volume(arr10, s, vol, ready):
            ready.synchronize()                   # this batch's D2H has landed (waited for here, off the main thread)
            vol_dir = os.path.join(save_dirs[s], 'vol_{}'.format(vol))
            os.makedirs(vol_dir, exist_ok=True)
            aff, hdr = ref_cache[s]
            for k, key in enumerate(IMG_KEYS):
                nib.save(nib.Nifti1Image(arr10[k].reshape(IMG_SHAPE), aff, hdr),
                         os.path.join(vol_dir, 'recon_{}.nii'.format(key)))

        with torch.no_grad(), ThreadPoolExecutor(max_workers=4) as pool:
            for it, sample in enumerate(loader):
                ids, covariates, x = self._batch(sample)
                B = ids.shape[0]
                vol_num, subjidx = sample['vol_num'].tolist(), sample['subjid'].tolist()
                for s in set(subjidx):
                    if s not in ref_cache:
                        ref = nib.load(ref_niis[s])
                        ref_cache[s] = (ref.affine, ref.header)
                _, sb = run_step(eng, x.reshape(B, -1), covariates, None, True)
                self._last = sb
                stack = torch.cat([sb.maps[0, :, :IMG_DIM].unsqueeze(0), sb.cons, sb.x_rec.unsqueeze(0)])   # (10,B,V)
                per_vol = stack.permute(1, 0, 2)                                                          # (B,10,V)
                sums.index_add_(0, ids.to(self.device), per_vol.double())
                counts.index_add_(0, ids.to(self.device), torch.ones(B, dtype=torch.float64, device=self.device))
                slot = it % 2
                for f in pending[slot]:
                    f.result()                        # the files that read this pinned buffer are written
                if pinned[slot] is None or pinned[slot].shape[0] < B:
                    pinned[slot] = torch.empty(max(B, loader.batch_size or B), len(IMG_KEYS), IMG_DIM).pin_memory()
                host = pinned[slot][:B]
                host.copy_(per_vol, non_blocking=True)
                copied[slot] = torch.cuda.Event()
                copied[slot].record()
                arr = host.numpy()
                pending[slot] = [pool.submit(write_volume, arr[b], subjidx[b], vol_num[b], copied[slot]) for b in range(B)]
            for fs in pending:
                for f in fs:
                    f.result()
        self.check_status()
        self._recon_avg = {"epoch": self.epoch, "save_dirs": list(save_dirs), "sums": sums, "counts": counts}

    def plot_GPs(self, csv_file='', save_dir=''):
        """Posterior gain mean / variance over all rows of the CSV, one sorted CSV (+ PDF when
        matplotlib exists) per motion covariate (reference vae_reg_GP.py:622-689).  Only
        diag(Sigma) is needed, so the O(N^2) matrix is never formed."""
        plot_dir = os.path.join(save_dir, str(self.epoch).zfill(3) + '_GP_plots')
        os.makedirs(plot_dir, exist_ok=True)
        data = pd.read_csv(csv_file)
        allcov = torch.from_numpy(data[['x', 'y', 'z', 'rot_x', 'rot_y', 'rot_z']].to_numpy())
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            plt = None
        for j, key in enumerate(_GP_COVS):
            prm = self.gp_params[key]
            kvar = prm['logkvar'].exp() + 0.1
            ls = self.max_ls * torch.sigmoid(prm['log_ls'].exp() + 0.5)
            regressor = gp.GP(prm['xu'], kvar, ls, prm['qu_m'], prm['qu_S'])
            xq = allcov[:, j].to(self.device)
            f_bar, var = regressor.evaluate_posterior_diag(xq)
            s2 = prm['logstd'][0].exp() ** 2
            mean = prm['sa'][0] * xq + f_bar
            var = s2 * xq ** 2 + var
            frame = pd.DataFrame({"xq": allcov[:, j].numpy(), "mean": mean.detach().cpu().numpy(),
                                  "vars": var.detach().cpu().numpy()}).sort_values(by=["xq"])
            frame.to_csv(os.path.join(plot_dir, '{}_GP_{}_full.csv'.format(str(self.epoch).zfill(3), key)))
            if plt is not None:
                plt.clf()
                plt.plot(frame["xq"], frame["mean"], c='darkblue', alpha=0.5, label='Beta posterior mean')
                two = 2 * np.sqrt(frame["vars"])
                plt.fill_between(frame["xq"], frame["mean"] - two, frame["mean"] + two, color='lightblue',
                                 alpha=0.3, label='2 sigma')
                plt.legend(loc='best')
                plt.title('GP Plot {}_full_set'.format(key))
                plt.savefig(os.path.join(plot_dir, 'GP_{}_full_set.pdf'.format(key)))

    def train_loop(self, loaders, epochs=100, test_freq=2, save_freq=10, save_dir=''):
        print("=" * 40)
        print("Training: epochs", self.epoch, "to", self.epoch + epochs - 1)
        print("Training set:", len(loaders['Shuffled_train'].dataset))
        print("Test set:", len(loaders['test'].dataset))
        print("=" * 40)
        for epoch in range(self.epoch, self.epoch + epochs):
            loss = self.train_epoch(loaders['Shuffled_train'])
            self.loss['train'][epoch] = loss
            self.writer.add_scalar("Loss/Train", loss, self.epoch)
            utils.log_qu_plots(self.epoch, self.gp_params, self.writer, 'train')
            utils.log_qkappa_plots(self.gp_params, self.writer, 'train')
            self.writer.flush()
            if (test_freq is not None) and (epoch % test_freq == 0):
                self.loss['test'][epoch] = self.test_epoch(loaders['test'])
            if (save_freq is not None) and (epoch % save_freq == 0) and (epoch > 0):
                # the reference joins save_dir twice; an absolute path keeps both behaviours
                self.save_state(os.path.join(os.path.abspath(save_dir), "checkpoint_" + str(epoch).zfill(3) + '.tar'))
        self.writer.close()
