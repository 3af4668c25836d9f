"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the VAE-GAM training step.

Parity status: PINNED.  `tests/test_oracle_vs_golden.py` checks this port against
golden vectors produced by the unmodified reference (`tests/golden/make_golden.py`,
run in the build container where `/root/reference` is mounted), and
`tests/test_host_logic.py` re-checks initial values, utils and checkpoint interchange live against the
reference whenever it is present.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module.  The product (`vae-gam_b200/`) never does.

What it restates (reference file:line):
  encode          vae_reg_GP.py:236-252   (layers :189-204)
  decode          vae_reg_GP.py:254-264   (layers :207-218)
  latent sample   vae_reg_GP.py:321-325   + torch lowrank_multivariate_normal.py:214-223
  latent KL       vae_reg_GP.py:400       + torch kl.py:342-372 (rank-1 closed form)
  linear-gain KL  vae_reg_GP.py:266-281
  gains           vae_reg_GP.py:345-369   + gp.py:67-110,113-136 ; sample: multivariate_normal.py:251-254
  GP KL           gp.py:41-65
  HRF FIR         vae_reg_GP.py:283-305   + utils.py:22-36
  objective       vae_reg_GP.py:380-410

The arithmetic of the reference lives in PyTorch (third-party; image has torch
2.11.0).  This port therefore uses the same library primitives for the dense
layers (conv3d / conv_transpose3d / batch_norm / linear) — so that its CPU
timing is representative of the reference's CPU path — and closed forms for the
`torch.distributions` objects.  `oracle/np_oracle.py` restates those primitives
independently in numpy.  Noise is injected explicitly; `draw_noise` reproduces
the reference's RNG order.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

IMG_SHAPE = (41, 49, 35)
IMG_DIM = 41 * 49 * 35
GP_KEYS = ["task", "x", "y", "z", "xrot", "yrot", "zrot", "sex"]      # vae_reg_GP.py:68
IMG_KEYS = ["base", "task", "x_mot", "y_mot", "z_mot", "pitch_mot", "roll_mot",
            "yaw_mot", "sex", "full_rec"]                                  # vae_reg_GP.py:308-309
NUM_LATENTS = 32
K = 8


def has_gp(i: int) -> bool:
    """Covariate i (1-based) has a GP term: vae_reg_GP.py:352."""
    return 1 < i < 8


def hrf_taps(dtype=torch.float64) -> torch.Tensor:
    """utils.py:22-36 evaluated at np.arange(0, 20, 1.4) (vae_reg_GP.py:292)."""
    t = torch.arange(0, 15, dtype=torch.float64) * 1.4
    peak = t ** 5 * torch.exp(-t) / math.factorial(5)
    under = t ** 11 * torch.exp(-t) / math.factorial(11)
    v = peak - 0.35 * under
    return (v / v.max() * 0.6).to(dtype)


def draw_noise(B: int, seed: Optional[int] = None, generator=None, dtype=torch.float32):
    """Noise in the reference's draw order: eps_W (B,1), eps_D (B,32), eps_1..8 (B,)."""
    if seed is not None:
        torch.manual_seed(seed)
    n = lambda *s: torch.empty(*s, dtype=dtype).normal_(generator=generator)
    eps_w = n(B, 1)
    eps_d = n(B, NUM_LATENTS)
    eps_g = torch.stack([n(B) for _ in range(K)])
    return {"eps_w": eps_w, "eps_d": eps_d, "eps_g": eps_g}


def params_from_module(model) -> Dict[str, torch.Tensor]:
    """Flat dict (reference state-dict names) + constants from a VAE-like module."""
    P = {n: p.detach().clone() for n, p in model.named_parameters()}
    for key in GP_KEYS[1:7]:
        P["xu_" + key] = model.gp_params[key]["xu"].detach().clone()
    P["glm_maps"] = model.glm_maps.detach().clone()
    return P


def cast_params(P, dtype, requires_grad=False):
    out = {}
    for k, v in P.items():
        t = v.detach().to(dtype) if v.is_floating_point() else v.detach()
        if requires_grad and not (k.startswith("xu_") or k == "glm_maps"):
            t = t.clone().requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------- stage 1
def bn(x, P, name):
    """BatchNorm3d(track_running_stats=False): batch statistics always (Appendix B)."""
    return F.batch_norm(x, None, None, P[name + ".weight"], P[name + ".bias"], True, 0.0, 1e-5)


def encode(P, x):
    B = x.shape[0]
    h = x.reshape(B, 1, *IMG_SHAPE)
    h = F.relu(F.conv3d(bn(h, P, "bn1"), P["conv1.weight"], P["conv1.bias"], 1))
    h = F.relu(F.conv3d(h, P["conv2.weight"], P["conv2.bias"], 2))
    h = F.relu(F.conv3d(bn(h, P, "bn3"), P["conv3.weight"], P["conv3.bias"], 1))
    h = F.relu(F.conv3d(h, P["conv4.weight"], P["conv4.bias"], 2))
    h = F.relu(F.conv3d(bn(h, P, "bn5"), P["conv5.weight"], P["conv5.bias"], 1))
    h = h.reshape(B, -1)
    lin = lambda t, n: F.linear(t, P[n + ".weight"], P[n + ".bias"])
    h = F.relu(lin(h, "fc1"))
    h = F.relu(lin(h, "fc2"))
    mu = lin(F.relu(lin(h, "fc31")), "fc41")
    u = lin(F.relu(lin(h, "fc32")), "fc42")
    d = torch.exp(lin(F.relu(lin(h, "fc33")), "fc43"))
    return mu, u, d


def decode(P, zcat):
    lin = lambda t, n: F.linear(t, P[n + ".weight"], P[n + ".bias"])
    h = F.relu(lin(zcat, "fc5"))
    h = F.relu(lin(h, "fc6"))
    h = F.relu(lin(h, "fc7"))
    h = F.relu(lin(h, "fc8"))
    h = h.reshape(-1, 16, 6, 8, 5)
    h = F.relu(F.conv_transpose3d(bn(h, P, "bnt1"), P["convt1.weight"], P["convt1.bias"], 1))
    h = F.relu(F.conv_transpose3d(h, P["convt2.weight"], P["convt2.bias"], 2,
                                  padding=(1, 0, 1), output_padding=(1, 0, 1)))
    h = F.relu(F.conv_transpose3d(bn(h, P, "bnt3"), P["convt3.weight"], P["convt3.bias"], 1))
    h = F.relu(F.conv_transpose3d(h, P["convt4.weight"], P["convt4.bias"], 2))
    h = F.conv_transpose3d(bn(h, P, "bnt5"), P["convt5.weight"], P["convt5.bias"], 1)
    return torch.sigmoid(h.reshape(-1, IMG_DIM))


def latent_sample_kl(mu, u, d, eps_w, eps_d):
    """E7-E9.  Returns z (B,32), KLz (B,), d after jitter."""
    if bool((d < 1e-6).any()):                       # vae_reg_GP.py:321-323 (all elements)
        d = d + 1e-6
    z = mu + u * eps_w + d.sqrt() * eps_d            # rank-1 factor: u (B,32) times eps_w (B,1)
    su = (u * u / d).sum(-1)
    klz = 0.5 * (-torch.log1p(su) - d.log().sum(-1) + d.sum(-1) + (u * u).sum(-1)
                 + (mu * mu).sum(-1) - mu.shape[-1])
    return z, klz, d


# --------------------------------------------------------------------------- stage 3
def lin_w_kl(sa, logstd):
    """KL(N(sa, s^2) || N(1, 0.5^2)), s = exp(logstd): vae_reg_GP.py:266-281."""
    s = torch.exp(logstd)
    return torch.log(0.5 / s) + (s * s + (sa - 1.0) ** 2) / (2 * 0.25) - 0.5


def rbf(d, k_var, ls):
    """gp.py:121-136."""
    return k_var * torch.exp(-(d / (math.sqrt(2.0) * ls)) ** 2)


def gp_posterior(xu, k_var, ls, qu_m, qu_s, xq):
    """gp.py:67-110.  xu (m,), qu_m (m,), qu_s (m,m) used unsymmetrised, xq (B,)."""
    m = xu.shape[0]
    step = xu[1] - xu[0]
    kk = torch.arange(m, dtype=xq.dtype)
    knu = rbf((xu[0] - xq)[None, :] + kk[:, None] * step, k_var, ls)       # (m,B)  gp.py:92-95
    knn = rbf(xq[None, :] - xq[:, None], k_var, ls)                         # (B,B)  gp.py:97-102
    ku = rbf((kk[:, None] - kk[None, :]).abs() * step, k_var, ls)           # (m,m)  gp.py:104-105
    a = knu.T @ torch.linalg.inv(ku)                                        # gp.py:107
    f_bar = a @ qu_m
    sigma = knn + a @ (qu_s - ku) @ a.T
    return f_bar, sigma


def gp_kl(qu_m, qu_s):
    """KL(N(qu_m, qu_S) || N(0, 10 I)), gp.py:41-65.  MultivariateNormal takes the
    Cholesky factor of qu_S, i.e. only its lower triangle enters."""
    m = qu_m.shape[0]
    l = torch.linalg.cholesky(qu_s)          # reads the lower triangle only
    tr = (l * l).sum()                       # trace(L L^T)
    logdet = 2 * torch.log(torch.diagonal(l)).sum()
    return 0.5 * (tr / 10.0 + (qu_m * qu_m).sum() / 10.0 - m + m * math.log(10.0) - logdet)


def hrf_fir(g, taps):
    """Causal 15-tap FIR over the batch index: vae_reg_GP.py:283-305."""
    B = g.shape[0]
    out = torch.zeros_like(g)
    for s in range(min(B, taps.shape[0])):
        out[s:] = out[s:] + taps[s] * g[:B - s]
    return out


def gains(P, cov, eps_g, neural_covariates=True):
    """Stage 3 for all 8 covariates.  Returns g (8,B), gp_kl_loss (scalar), aux."""
    dtype = cov.dtype
    B = cov.shape[0]
    eye = torch.eye(B, dtype=dtype)
    taps = hrf_taps(dtype)
    g_all, means, covs = [], [], []
    kl_sum = torch.zeros((), dtype=dtype)
    for i in range(1, K + 1):
        key = GP_KEYS[i - 1]
        xq = cov[:, i - 1]
        sa = P["sa_" + key].reshape(())
        logstd = P["logstd_" + key].reshape(())
        kl_sum = kl_sum + lin_w_kl(sa, logstd)
        mean = sa * xq
        c = torch.exp(logstd) ** 2 * xq ** 2 * eye
        if has_gp(i):
            k_var = torch.exp(P["logkvar_" + key]) + 0.1
            ls = 3.0 * torch.sigmoid(torch.exp(P["logls_" + key]) + 0.5)
            qm = P["qu_m_" + key].reshape(-1)
            qs = P["qu_S_" + key]
            f_bar, sigma = gp_posterior(P["xu_" + key].to(dtype), k_var, ls, qm, qs, xq)
            mean = mean + f_bar
            c = c + sigma
            kl_sum = kl_sum + gp_kl(qm, qs)
        L = torch.linalg.cholesky(c + 1e-5 * eye)                 # vae_reg_GP.py:368
        g = mean + L @ eps_g[i - 1]
        if neural_covariates and i < (K - 6):                     # vae_reg_GP.py:377
            g = hrf_fir(g, taps)
        g_all.append(g)
        means.append(mean)
        covs.append(c)
    return torch.stack(g_all), kl_sum, {"mean": torch.stack(means), "cov": torch.stack(covs)}


# --------------------------------------------------------------------------- stage 4
def recon_loss(maps, g, x, epsilon, glm, glm_reg_scale):
    """R1-R4 given the 9 decoder maps (9,B,V), gains (8,B), x (B,V), epsilon (V), glm (V,8).

    Returns dict with x_rec, cons (8,B,V), logp (B,), glm_reg (scalar, = B * sum_i sum_b ||cons_ib - G_i||)."""
    B = x.shape[0]
    cons = g[:, :, None] * maps[1:]
    x_rec = maps[0] + cons.sum(0)
    diffn = torch.linalg.vector_norm(cons - glm.T[:, None, :], dim=-1)      # (8,B)
    glm_reg = B * diffn.sum()                                                # vae_reg_GP.py:388-389
    w = torch.exp(2 * epsilon)
    r = x - x_rec
    logp = (-0.5 * r * r * w + epsilon - 0.5 * math.log(2 * math.pi)).sum(-1)
    return {"x_rec": x_rec, "cons": cons, "logp": logp, "glm_reg": glm_reg, "glm_norms": diffn}


# --------------------------------------------------------------------------- whole step
def step(P, x, cov, noise, gp_kl_scale=10.0, glm_reg_scale=1.0, neural_covariates=True,
         keep_maps=True, g_override=None):
    """One forward pass of vae_reg_GP.py:307-413 with injected noise.  P, x, cov, noise
    must share one dtype (fp32 to mimic the reference, fp64 for the truth).

    `g_override` (8,B) replaces the VALUE of the sampled gains (post-HRF) while keeping
    this port's differentiable path: the reference's fp32 `torch.inverse(Ku)` makes its
    own gains deviate from the fp64 truth by up to ~3e-2, so stage-local
    parity of everything downstream is checked with the reference's gains injected."""
    dtype = x.dtype
    B = x.shape[0]
    mu, u, d = encode(P, x)
    z, klz, d = latent_sample_kl(mu, u, d, noise["eps_w"].to(dtype), noise["eps_d"].to(dtype))
    oh = torch.eye(K + 1, dtype=dtype)
    zcat = torch.cat([z[None].expand(K + 1, B, -1), oh[:, None, :].expand(K + 1, B, -1)], -1)
    maps = torch.stack([decode(P, zcat[j]) for j in range(K + 1)])           # 9 separate BN batches
    g, gp_kl_loss, aux = gains(P, cov, noise["eps_g"].to(dtype), neural_covariates)
    if g_override is not None:
        g = g + (g_override.to(dtype) - g).detach()
    eps = P["epsilon"].reshape(-1).to(dtype)                                  # .float() at :402
    glm = P["glm_maps"][:, 1:].to(dtype)                                      # col 0 = pandas index
    rl = recon_loss(maps, g, x.reshape(B, -1), eps, glm, glm_reg_scale)
    neg_elbo = -(rl["logp"] - klz).mean()
    tot = neg_elbo + gp_kl_scale * gp_kl_loss + glm_reg_scale * rl["glm_reg"]
    out = {"tot": tot, "neg_elbo": neg_elbo, "gp_kl": gp_kl_loss, "glm_reg": rl["glm_reg"],
           "logp": rl["logp"], "klz": klz, "mu": mu, "u": u, "d": d, "z": z, "g": g,
           "beta_mean": aux["mean"], "beta_cov": aux["cov"], "glm_norms": rl["glm_norms"]}
    if keep_maps:
        out["maps"] = maps
        out["cons"] = rl["cons"]
        out["x_rec"] = rl["x_rec"]
    return out


def imgs_from(out) -> Dict[str, torch.Tensor]:
    """The reference's `imgs` dict (vae_reg_GP.py:331,391,392)."""
    d = {"base": out["maps"][0]}
    for i in range(1, K + 1):
        d[IMG_KEYS[i]] = out["cons"][i - 1]
    d["full_rec"] = out["x_rec"]
    return d


def training_step_cpu(P, opt_state, x, cov, noise, lr=1e-3, **kw):
    """fwd + bwd + Adam (vae_reg_GP.py:425-429) on leaf tensors in P (requires_grad)."""
    leaves = [v for v in P.values() if v.requires_grad]
    for v in leaves:
        v.grad = None
    out = step(P, x, cov, noise, keep_maps=False, **kw)
    out["tot"].backward()
    if opt_state.get("opt") is None:
        opt_state["opt"] = torch.optim.Adam(leaves, lr=lr)
    opt_state["opt"].step()
    return float(out["tot"])
