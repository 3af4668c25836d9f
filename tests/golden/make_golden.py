"""Generate golden vectors from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py            # writes tests/golden/case_*.npz

The reference (`/root/reference`, read-only) is imported through `oracle/ref_loader.py`
(inert stubs for matplotlib / nibabel / umap, CPU version of `gp._striped_matrix`) and
INSTRUMENTED, not edited: `MultivariateNormal` in the reference module's namespace is
replaced by a recording subclass so the per-covariate gain mean / covariance / sample
(vae_reg_GP.py:368-369) can be captured, `encode` and `do_hrf_conv` are wrapped to record
their outputs.  The individual loss terms are recovered by re-running `forward` under the
same seed with (gp_kl_scale, glm_reg_scale) in {(0,0),(1,0),(0,1)}.

Everything regenerable is stored as a recipe (seeds, sizes), not as data:
  params   = reference ctor under torch.manual_seed(param_seed)   (checksums stored)
  x        = torch.rand(B,41,49,35, generator=Generator().manual_seed(x_seed))
  cov      = first B rows of vaegam.synthetic.make_cohort(2, config, seed=0)
  noise    = oracle.ref_port.draw_noise(B, seed=noise_seed)  (the reference's RNG order)
Stored outputs: total + terms, mu/u/d/z, gain mean/cov/sample, strided samples of the 10
maps, and for every parameter gradient its L2 norm and a strided sample.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vae-gam_b200"))

from oracle.ref_loader import NullWriter, load_reference  # noqa: E402
from vaegam import synthetic as syn  # noqa: E402

CASES = {
    # name: B, m, neural, glm kind, gp_kl_scale, glm_reg_scale, config, map stride, seeds
    "b2_m6_neural": dict(B=2, m=6, neural=True, glm="uniform", gp_kl_scale=10.0, glm_reg_scale=1.0,
                         config="checker", stride=53, param_seed=1, x_seed=3, noise_seed=11),
    "b4_m6_control": dict(B=4, m=6, neural=False, glm="zeros", gp_kl_scale=10.0, glm_reg_scale=0.0,
                          config="control", stride=53, param_seed=2, x_seed=4, noise_seed=12),
    "b4_m4_neural": dict(B=4, m=4, neural=True, glm="uniform", gp_kl_scale=10.0, glm_reg_scale=1.0,
                         config="checker", stride=53, param_seed=3, x_seed=5, noise_seed=13),
    "b32_m6_neural": dict(B=32, m=6, neural=True, glm="uniform", gp_kl_scale=10.0, glm_reg_scale=1.0,
                          config="checker", stride=211, param_seed=1, x_seed=6, noise_seed=14),
}
GRAD_SAMPLE = 64


def sample_flat(t, n=GRAD_SAMPLE):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].double().numpy()


def build_inputs(c, workdir):
    tr, te, glm, coh = syn.write_experiment(workdir, n_subjects=2, config=c["config"], glm=c["glm"])
    x = torch.rand(c["B"], 41, 49, 35, generator=torch.Generator().manual_seed(c["x_seed"]))
    cov = torch.from_numpy(coh.covariates()[: c["B"]].copy())
    ids = torch.from_numpy(coh.subject_index()[: c["B"]].copy())
    return tr, te, glm, x, cov, ids


def run_case(name, c, ref_vae):
    work = tempfile.mkdtemp(prefix="golden_")
    tr, te, glm, x, cov, ids = build_inputs(c, work)
    rec = {"mean": [], "cov": [], "g": [], "hrf": [], "enc": []}

    import torch.distributions as td

    class RecMVN(td.MultivariateNormal):
        def __init__(self, loc, covariance_matrix=None, **kw):
            super().__init__(loc, covariance_matrix, **kw)
            rec["mean"].append(loc.detach().clone())
            rec["cov"].append(covariance_matrix.detach().clone())

        def rsample(self, *a, **k):
            s = super().rsample(*a, **k)
            rec["g"].append(s.detach().clone())
            return s

    ref_vae.MultivariateNormal = RecMVN
    try:
        torch.manual_seed(c["param_seed"])
        m = ref_vae.VAE(save_dir=work, glm_maps=glm, csv_files=[tr, te], num_inducing_pts=c["m"],
                        gp_kl_scale=c["gp_kl_scale"], glm_reg_scale=c["glm_reg_scale"],
                        neural_covariates=c["neural"])
        m.writer = NullWriter()
        enc0, hrf0 = m.encode, m.do_hrf_conv

        def enc(xx):
            o = enc0(xx)
            rec["enc"].append([t.detach().clone() for t in o])
            return o

        def hrf(v):
            o = hrf0(v)
            rec["hrf"].append(o.detach().clone())
            return o

        m.encode, m.do_hrf_conv = enc, hrf

        def fwd(gp_s, glm_s, want_maps=False):
            for k in rec:
                rec[k].clear()
            m.gp_kl_scale = torch.as_tensor(gp_s)
            m.glm_reg_scale = glm_s
            torch.manual_seed(c["noise_seed"])
            return m.forward(ids, cov, x, "train", return_latent_rec=want_maps, train_mode=False)

        with torch.no_grad():
            t00 = float(fwd(0.0, 0.0))
            t10 = float(fwd(1.0, 0.0))
            t01 = float(fwd(0.0, 1.0))
        m.zero_grad()
        loss, z, imgs = fwd(c["gp_kl_scale"], c["glm_reg_scale"], want_maps=True)
        loss.backward()
    finally:
        ref_vae.MultivariateNormal = td.MultivariateNormal

    out = {
        "tot": float(loss), "neg_elbo": t00, "gp_kl": t10 - t00, "glm_reg": t01 - t00,
        "z": z.astype(np.float64),
        "mu": rec["enc"][0][0].double().numpy(), "u": rec["enc"][0][1].squeeze(-1).double().numpy(),
        "d": rec["enc"][0][2].double().numpy(),
        "beta_mean": torch.stack(rec["mean"]).double().numpy(),
        "beta_cov_jittered": torch.stack(rec["cov"]).double().numpy(),   # includes + 1e-5 I
        "g_pre_hrf": torch.stack(rec["g"]).double().numpy(),
    }
    g = torch.stack(rec["g"]).clone()
    if rec["hrf"]:
        g[0] = rec["hrf"][0]
    out["g"] = g.double().numpy()
    for k, v in imgs.items():
        out["map_" + k] = v[:, :: c["stride"]].astype(np.float32)
    names = []
    for n, p in m.named_parameters():
        names.append(n)
        out["gradnorm_" + n] = float(p.grad.double().norm())
        out["gradsample_" + n] = sample_flat(p.grad)
        out["paramsum_" + n] = np.array([float(p.detach().double().sum()),
                                         float(p.detach().double().abs().sum())])
    out["param_names"] = np.array(names)
    out["recipe"] = np.array(repr(c))
    path = os.path.join(HERE, f"case_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: tot={out['tot']:.6g} neg_elbo={t00:.6g} gp_kl={out['gp_kl']:.6g} "
          f"glm_reg={out['glm_reg']:.6g} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    ref_vae, _, _ = load_reference()
    torch.set_num_threads(os.cpu_count())
    for name, c in CASES.items():
        run_case(name, c, ref_vae)


if __name__ == "__main__":
    main()
