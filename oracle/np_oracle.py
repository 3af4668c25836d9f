"""TEST INFRASTRUCTURE ONLY — numpy fp64 restatement of the VAE-GAM forward step.

Independent of torch's kernels: every primitive the reference gets from PyTorch
(conv3d, conv_transpose3d, batch-stat batchnorm, linear, the distribution closed
forms, cholesky / inverse) is written out here with numpy loops / einsum, so that
`oracle/ref_port.py` (which calls the torch primitives) and the CUDA kernels are
both checked against arithmetic that shares no code with either.

Parity status: PINNED through `tests/test_oracle_vs_golden.py` (golden vectors from
the unmodified reference) — see ref_port.py for the reference file:line map; the
layer semantics follow SURVEY.md Appendix B:
  * Conv3d (vae_reg_GP.py:189-193): no padding, out = floor((in-3)/s)+1.
  * ConvTranspose3d (vae_reg_GP.py:211-215), weight (Cin,Cout,kD,kH,kW), gather form
      out[n,co,o] = b[co] + sum_ci sum_{k: (o+p-k) % s == 0, 0 <= (o+p-k)/s < In} x[n,ci,(o+p-k)/s] w[ci,co,k]
      Out = (In-1) s - 2p + k + output_padding.
  * BatchNorm3d(track_running_stats=False) (:194-196,216-218): biased batch variance, eps 1e-5.
Only small batches are practical (pure numpy): B <= 4 runs in seconds.
"""
from __future__ import annotations

import math

import numpy as np

IMG_SHAPE = (41, 49, 35)
IMG_DIM = 41 * 49 * 35
GP_KEYS = ["task", "x", "y", "z", "xrot", "yrot", "zrot", "sex"]


def conv3d(x, w, b, stride):
    """x (N,Ci,D,H,W), w (Co,Ci,kd,kh,kw)."""
    n, ci, D, H, W = x.shape
    co, _, kd, kh, kw = w.shape
    od, oh, ow = (D - kd) // stride + 1, (H - kh) // stride + 1, (W - kw) // stride + 1
    out = np.zeros((n, co, od, oh, ow))
    for a in range(kd):
        for bb in range(kh):
            for c in range(kw):
                xs = x[:, :, a:a + stride * (od - 1) + 1:stride,
                       bb:bb + stride * (oh - 1) + 1:stride,
                       c:c + stride * (ow - 1) + 1:stride]
                out += np.einsum("nidhw,oi->nodhw", xs, w[:, :, a, bb, c])
    return out + b[None, :, None, None, None]


def conv_transpose3d(x, w, b, stride, padding=(0, 0, 0), output_padding=(0, 0, 0)):
    """x (N,Ci,D,H,W), w (Ci,Co,kd,kh,kw); scatter form (equivalent to the gather form above)."""
    n, ci, D, H, W = x.shape
    _, co, kd, kh, kw = w.shape
    full = [(D - 1) * stride + kd, (H - 1) * stride + kh, (W - 1) * stride + kw]
    # allocate room for output_padding beyond the scatter footprint
    ext = [f + op for f, op in zip(full, output_padding)]
    buf = np.zeros((n, co, *ext))
    for a in range(kd):
        for bb in range(kh):
            for c in range(kw):
                contrib = np.einsum("nidhw,io->nodhw", x, w[:, :, a, bb, c])
                buf[:, :, a:a + stride * (D - 1) + 1:stride,
                    bb:bb + stride * (H - 1) + 1:stride,
                    c:c + stride * (W - 1) + 1:stride] += contrib
    pd, ph, pw = padding
    od = (D - 1) * stride - 2 * pd + kd + output_padding[0]
    oh = (H - 1) * stride - 2 * ph + kh + output_padding[1]
    ow = (W - 1) * stride - 2 * pw + kw + output_padding[2]
    out = buf[:, :, pd:pd + od, ph:ph + oh, pw:pw + ow]
    return out + b[None, :, None, None, None]


def batchnorm(x, gamma, beta, eps=1e-5):
    mean = x.mean(axis=(0, 2, 3, 4), keepdims=True)
    var = x.var(axis=(0, 2, 3, 4), keepdims=True)          # biased
    return (x - mean) / np.sqrt(var + eps) * gamma[None, :, None, None, None] + beta[None, :, None, None, None]


def relu(x):
    return np.maximum(x, 0.0)


def linear(x, w, b):
    return x @ w.T + b


def encode(P, x):
    B = x.shape[0]
    h = x.reshape(B, 1, *IMG_SHAPE)
    h = relu(conv3d(batchnorm(h, P["bn1.weight"], P["bn1.bias"]), P["conv1.weight"], P["conv1.bias"], 1))
    h = relu(conv3d(h, P["conv2.weight"], P["conv2.bias"], 2))
    h = relu(conv3d(batchnorm(h, P["bn3.weight"], P["bn3.bias"]), P["conv3.weight"], P["conv3.bias"], 1))
    h = relu(conv3d(h, P["conv4.weight"], P["conv4.bias"], 2))
    h = relu(conv3d(batchnorm(h, P["bn5.weight"], P["bn5.bias"]), P["conv5.weight"], P["conv5.bias"], 1))
    h = h.reshape(B, -1)
    h = relu(linear(h, P["fc1.weight"], P["fc1.bias"]))
    h = relu(linear(h, P["fc2.weight"], P["fc2.bias"]))
    mu = linear(relu(linear(h, P["fc31.weight"], P["fc31.bias"])), P["fc41.weight"], P["fc41.bias"])
    u = linear(relu(linear(h, P["fc32.weight"], P["fc32.bias"])), P["fc42.weight"], P["fc42.bias"])
    d = np.exp(linear(relu(linear(h, P["fc33.weight"], P["fc33.bias"])), P["fc43.weight"], P["fc43.bias"]))
    return mu, u, d


def decode(P, zcat):
    h = zcat
    for n in ("fc5", "fc6", "fc7", "fc8"):
        h = relu(linear(h, P[n + ".weight"], P[n + ".bias"]))
    h = h.reshape(-1, 16, 6, 8, 5)
    h = relu(conv_transpose3d(batchnorm(h, P["bnt1.weight"], P["bnt1.bias"]), P["convt1.weight"], P["convt1.bias"], 1))
    h = relu(conv_transpose3d(h, P["convt2.weight"], P["convt2.bias"], 2, (1, 0, 1), (1, 0, 1)))
    h = relu(conv_transpose3d(batchnorm(h, P["bnt3.weight"], P["bnt3.bias"]), P["convt3.weight"], P["convt3.bias"], 1))
    h = relu(conv_transpose3d(h, P["convt4.weight"], P["convt4.bias"], 2))
    h = conv_transpose3d(batchnorm(h, P["bnt5.weight"], P["bnt5.bias"]), P["convt5.weight"], P["convt5.bias"], 1)
    return 1.0 / (1.0 + np.exp(-h.reshape(-1, IMG_DIM)))


def latent_sample_kl(mu, u, d, eps_w, eps_d):
    if (d < 1e-6).any():
        d = d + 1e-6
    z = mu + u * eps_w + np.sqrt(d) * eps_d
    su = (u * u / d).sum(-1)
    klz = 0.5 * (-np.log1p(su) - np.log(d).sum(-1) + d.sum(-1) + (u * u).sum(-1) + (mu * mu).sum(-1) - mu.shape[-1])
    return z, klz, d


def cholesky_lower(a):
    """Textbook Cholesky-Banachiewicz on the lower triangle."""
    n = a.shape[0]
    l = np.zeros_like(a)
    for i in range(n):
        for j in range(i + 1):
            s = a[i, j] - (l[i, :j] * l[j, :j]).sum()
            if i == j:
                if s <= 0:
                    raise ValueError("matrix not positive definite")
                l[i, j] = math.sqrt(s)
            else:
                l[i, j] = s / l[j, j]
    return l


def gauss_jordan_inverse(a):
    n = a.shape[0]
    m = np.concatenate([a.copy(), np.eye(n)], 1)
    for c in range(n):
        p = c + np.argmax(np.abs(m[c:, c]))
        m[[c, p]] = m[[p, c]]
        m[c] /= m[c, c]
        for r in range(n):
            if r != c:
                m[r] -= m[r, c] * m[c]
    return m[:, n:]


def rbf(d, k_var, ls):
    return k_var * np.exp(-(d / (math.sqrt(2.0) * ls)) ** 2)


def hrf_taps():
    t = np.arange(0, 20, 1.4)
    v = t ** 5 * np.exp(-t) / math.factorial(5) - 0.35 * t ** 11 * np.exp(-t) / math.factorial(11)
    return v / v.max() * 0.6


def gains(P, cov, eps_g, neural=True):
    B = cov.shape[0]
    eye = np.eye(B)
    taps = hrf_taps()
    gs, kl = [], 0.0
    means, covs = [], []
    for i in range(1, 9):
        key = GP_KEYS[i - 1]
        xq = cov[:, i - 1]
        sa = float(P["sa_" + key].reshape(-1)[0])
        s = math.exp(float(P["logstd_" + key].reshape(-1)[0]))
        kl += math.log(0.5 / s) + (s * s + (sa - 1) ** 2) / 0.5 - 0.5
        mean = sa * xq
        c = s * s * xq * xq * eye
        if 1 < i < 8:
            k_var = math.exp(float(P["logkvar_" + key])) + 0.1
            ls = 3.0 / (1.0 + math.exp(-(math.exp(float(P["logls_" + key])) + 0.5)))
            xu = P["xu_" + key]
            m = xu.shape[0]
            qm, qs = P["qu_m_" + key].reshape(-1), P["qu_S_" + key]
            step = xu[1] - xu[0]
            kk = np.arange(m)
            knu = rbf((xu[0] - xq)[None, :] + kk[:, None] * step, k_var, ls)
            knn = rbf(xq[None, :] - xq[:, None], k_var, ls)
            ku = rbf(np.abs(kk[:, None] - kk[None, :]) * step, k_var, ls)
            a = knu.T @ gauss_jordan_inverse(ku)
            mean = mean + a @ qm
            c = c + knn + a @ (qs - ku) @ a.T
            l = cholesky_lower(qs)
            kl += 0.5 * ((l * l).sum() / 10 + (qm * qm).sum() / 10 - m + m * math.log(10.0)
                         - 2 * np.log(np.diag(l)).sum())
        L = cholesky_lower(c + 1e-5 * eye)
        g = mean + L @ eps_g[i - 1]
        if neural and i < 2:
            o = np.zeros_like(g)
            for t in range(B):
                for s_ in range(min(t, 14) + 1):
                    o[t] += taps[s_] * g[t - s_]
            g = o
        gs.append(g)
        means.append(mean)
        covs.append(c)
    return np.stack(gs), kl, np.stack(means), np.stack(covs)


def step(P, x, cov, noise, gp_kl_scale=10.0, glm_reg_scale=1.0, neural=True, g_override=None):
    """P: dict name -> float64 ndarray (reference state-dict names + xu_*, glm_maps (V,9))."""
    B = x.shape[0]
    mu, u, d = encode(P, x)
    z, klz, d = latent_sample_kl(mu, u, d, noise["eps_w"], noise["eps_d"])
    maps = []
    for j in range(9):
        oh = np.zeros((B, 9))
        oh[:, j] = 1
        maps.append(decode(P, np.concatenate([z, oh], 1)))
    maps = np.stack(maps)
    g, gpkl, bmean, bcov = gains(P, cov, noise["eps_g"], neural)
    if g_override is not None:
        g = g_override
    eps = P["epsilon"].reshape(-1)
    glm = P["glm_maps"][:, 1:]
    cons = g[:, :, None] * maps[1:]
    x_rec = maps[0] + cons.sum(0)
    norms = np.sqrt(((cons - glm.T[:, None, :]) ** 2).sum(-1))
    glm_reg = B * norms.sum()
    r = x.reshape(B, -1) - x_rec
    logp = (-0.5 * r * r * np.exp(2 * eps) + eps - 0.5 * math.log(2 * math.pi)).sum(-1)
    neg_elbo = -(logp - klz).mean()
    tot = neg_elbo + gp_kl_scale * gpkl + glm_reg_scale * glm_reg
    return {"tot": tot, "neg_elbo": neg_elbo, "gp_kl": gpkl, "glm_reg": glm_reg, "logp": logp, "klz": klz,
            "mu": mu, "u": u, "d": d, "z": z, "g": g, "beta_mean": bmean, "beta_cov": bcov,
            "maps": maps, "cons": cons, "x_rec": x_rec}
