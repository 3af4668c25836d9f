"""The gain-stage algebra (csrc/gp_core.h — the exact source the CUDA kernel compiles) run on the
host with a 1-thread team (tests/cpu_emul) against the fp64 oracle with autograd.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import ref_port as rp

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "cpu_emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL, "libgp_emul.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(EMUL, "gp_emul.cpp")], check=True)
    lib = ctypes.CDLL(so)
    lib.emul_gain_ws_doubles.restype = ctypes.c_size_t
    return lib


def fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_params(B, m, seed):
    gen = torch.Generator().manual_seed(seed)
    cov = torch.randn(B, 8, generator=gen) * 1.2
    eps = torch.randn(8, B, generator=gen)
    P = {}
    for i, key in enumerate(rp.GP_KEYS):
        P["sa_" + key] = torch.normal(1, 1, size=(1, 1), generator=gen)
        P["logstd_" + key] = torch.normal(0, 1, size=(1, 1), generator=gen) * 0.5
        if rp.has_gp(i + 1):
            P["qu_m_" + key] = torch.randn(1, m, generator=gen)
            S = torch.randn(m, m, generator=gen) * 0.3
            P["qu_S_" + key] = (2 * torch.eye(m) + S @ S.T).float()
            P["logkvar_" + key] = torch.tensor(0.1 * i)
            P["logls_" + key] = torch.tensor(-0.2 + 0.1 * i)
            P["xu_" + key] = torch.linspace(-3.3, 3.4, m)
    return cov, eps, P


@pytest.mark.parametrize("B,m,neural", [(1, 6, True), (4, 6, True), (32, 6, True), (32, 6, False), (7, 4, True), (40, 8, True)])
def test_gain_forward_backward(emul, B, m, neural):
    cov, eps, P = make_params(B, m, 10 * B + m)
    Pd = {k: v.double().clone().requires_grad_(not k.startswith("xu_")) for k, v in P.items()}
    g, kl, aux = rp.gains(Pd, cov.double(), eps.double(), neural)
    dg = torch.randn(8, B, generator=torch.Generator().manual_seed(1)).double()
    kl_scale = 10.0
    ((g * dg).sum() + kl_scale * kl).backward()
    taps = rp.hrf_taps().numpy().copy()
    covn = cov.numpy().copy()
    ws = np.zeros(emul.emul_gain_ws_doubles(B, m))
    kl_total = 0.0
    for i, key in enumerate(rp.GP_KEYS):
        hg, hrf = int(rp.has_gp(i + 1)), int(neural and i == 0)
        z1 = np.zeros(1, np.float32)
        arrs = [P["sa_" + key].numpy().reshape(-1).copy(), P["logstd_" + key].numpy().reshape(-1).copy()]
        if hg:
            arrs += [P["qu_m_" + key].numpy().reshape(-1).copy(), P["qu_S_" + key].numpy().copy(),
                     P["logkvar_" + key].numpy().reshape(1).copy(), P["logls_" + key].numpy().reshape(1).copy(),
                     P["xu_" + key].numpy().copy()]
        else:
            arrs += [z1] * 5
        e = eps[i].numpy().copy()
        gout, klo = np.zeros(B, np.float32), np.zeros(2)
        bm, bv, st = np.zeros(B, np.float32), np.zeros(B, np.float32), np.zeros(1, np.int32)
        emul.emul_gain_fwd(fp(covn), 8, i, fp(e), *[fp(a) for a in arrs], fp(taps), hg, hrf, B, m, fp(ws), fp(gout),
                           fp(klo), fp(bm), fp(bv), fp(st))
        assert st[0] == 0
        scale = max(1.0, float(g[i].abs().max()))
        assert np.abs(gout - g[i].detach().numpy()).max() < 1e-6 * scale
        assert np.abs(bm - aux["mean"][i].detach().numpy()).max() < 1e-6 * scale
        kl_total += klo.sum()
        outs = [np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(m, np.float32), np.zeros((m, m), np.float32),
                np.zeros(1, np.float32), np.zeros(1, np.float32)]
        dgi = dg[i].float().numpy().copy()
        emul.emul_gain_bwd(fp(covn), 8, i, fp(e), *[fp(a) for a in arrs], fp(taps), hg, hrf, B, m, fp(ws), fp(dgi),
                           ctypes.c_double(kl_scale), *[fp(a) for a in outs])
        names = ["sa_", "logstd_"] + (["qu_m_", "qu_S_", "logkvar_", "logls_"] if hg else [])
        for nme, o in zip(names, outs):
            ref = Pd[nme + key].grad.numpy().reshape(o.shape)
            if nme == "logkvar_":   # ~0 by cancellation (A is independent of k_var); absolute scale of its terms
                assert np.abs(o - ref).max() < (1e-4 if m < 8 else 1e-1) * max(1.0, float(np.abs(Pd["qu_S_" + key].grad.numpy()).max()))
            else:
                # m = 8: cond(Ku) ~ 1e8, so two fp64 evaluation orders already differ at ~1e-4
                tol = 2e-6 if m < 8 else 1e-3
                assert np.abs(o - ref).max() <= tol * max(1e-3, np.abs(ref).max()), (key, nme)
    assert abs(kl_total - float(kl)) < 1e-9 * max(1.0, abs(float(kl)))


def test_gain_flags_non_positive_definite(emul):
    B, m = 4, 6
    cov, eps, P = make_params(B, m, 5)
    key = "x"
    bad_S = (-torch.eye(m)).numpy().astype(np.float32).copy()
    arrs = [P["sa_" + key].numpy().reshape(-1).copy(), P["logstd_" + key].numpy().reshape(-1).copy(),
            P["qu_m_" + key].numpy().reshape(-1).copy(), bad_S, P["logkvar_" + key].numpy().reshape(1).copy(),
            P["logls_" + key].numpy().reshape(1).copy(), P["xu_" + key].numpy().copy()]
    ws = np.zeros(emul.emul_gain_ws_doubles(B, m))
    gout, klo, st = np.zeros(B, np.float32), np.zeros(2), np.zeros(1, np.int32)
    emul.emul_gain_fwd(fp(cov.numpy().copy()), 8, 1, fp(eps[1].numpy().copy()), *[fp(a) for a in arrs],
                       fp(rp.hrf_taps().numpy().copy()), 1, 0, B, m, fp(ws), fp(gout), fp(klo), None, None, fp(st))
    assert st[0] != 0
