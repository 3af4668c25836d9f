"""Shared test helpers: golden-case recipes -> inputs, product model construction."""
import ast
import os
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")
CASES = ["b2_m6_neural", "b4_m6_control", "b4_m4_neural", "b32_m6_neural"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"), allow_pickle=False)
    recipe = ast.literal_eval(str(g["recipe"]))
    return g, recipe


def build_case(recipe, device_name="auto"):
    """Product VAE + inputs for a golden recipe (same seeds as tests/golden/make_golden.py)."""
    import vae_reg_GP
    from vaegam import synthetic as syn
    work = tempfile.mkdtemp(prefix="case_")
    tr, te, glm, coh = syn.write_experiment(work, n_subjects=2, config=recipe["config"], glm=recipe["glm"])
    B = recipe["B"]
    x = torch.rand(B, 41, 49, 35, generator=torch.Generator().manual_seed(recipe["x_seed"]))
    cov = torch.from_numpy(coh.covariates()[:B].copy())
    ids = torch.from_numpy(coh.subject_index()[:B].copy())
    torch.manual_seed(recipe["param_seed"])
    model = vae_reg_GP.VAE(save_dir=work, glm_maps=glm, csv_files=[tr, te], num_inducing_pts=recipe["m"],
                           gp_kl_scale=recipe["gp_kl_scale"], glm_reg_scale=recipe["glm_reg_scale"],
                           neural_covariates=recipe["neural"], device_name=device_name)
    return model, x, cov, ids


def check_param_sums(model, g):
    for n, p in model.named_parameters():
        s = g["paramsum_" + n]
        got = float(p.detach().double().sum().cpu())
        assert abs(got - s[0]) <= 1e-9 * max(1.0, abs(s[0])), f"initial value of {n} differs from the reference"


def sample_flat(t, n=64):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx.to(f.device)].double().cpu().numpy()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def nifti_experiment(tmp_path, n_subjects=2, n_vols=5, with_test_csv=False):
    """A tiny cohort backed by real 4-D NIfTI files (BASELINE config 5 needs reference images for affine / header)."""
    import pandas as pd
    from vaegam import synthetic as syn
    from vaegam.nib_compat import nib
    coh = syn.make_cohort(n_subjects, "checker", seed=3, n_vols=n_vols)
    vols = coh.volumes().numpy() * 3284.5                                     # the loader divides by 3284.5
    tab = coh.table.copy()
    sidx = coh.subject_index()
    paths = []
    for s, name in enumerate(tab["subjid"].unique().tolist()):
        v4 = np.moveaxis(vols[sidx == s], 0, -1).astype(np.float32)           # (41,49,35,T)
        path = str(tmp_path / f"{name}.nii.gz")
        nib.save(nib.Nifti1Image(v4, np.diag([3.0, 3.0, 3.5, 1.0])), path)
        paths.append(path)
    tab["nii_path"] = [paths[s] for s in sidx]
    tab["volume #"] = np.concatenate([np.arange(n_vols)] * n_subjects)
    csv = str(tmp_path / "train.csv")
    tab.to_csv(csv)
    glm = str(tmp_path / "glm.csv")
    pd.DataFrame(syn.glm_maps_uniform(), columns=syn.GLM_COLS).to_csv(glm)
    return csv, glm, coh
