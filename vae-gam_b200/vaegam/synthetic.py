"""Synthetic 4-D fMRI cohorts of the reference's data_dims shape (41,49,35,98).

There is no network and no real data in this image, so every experiment named
in BASELINE.json runs on data produced here.  The generator
restates, in its own words, the pieces of the reference's offline scripts that
define what the trainer sees:

* CSV schema written by ``pre_proc_vaefmri.py:126-132`` (pandas index column,
  then ``subjid, "volume #", nii_path, task, x, y, z, rot_x, rot_y, rot_z, sex``)
  and read positionally by ``DataClass_GP.py:32-46``;
* block designs of ``utils.py:75-111`` (20 s blocks, TR 1.4 s, volume times
  ``(1..T)*TR``): checkerboard starts with rest, the control design with task;
* z-scoring of the six motion columns over all rows with population std
  (``utils.py:113-123``);
* GLM-map CSV of ``get_beta_map_regularizer.py:105-107``: index column plus
  eight columns ``task,x,y,z,xrot,yrot,zrot,sex`` with 70 315 rows;
* the control glyph placement ``[15:25, 34:47, 9:22]`` of
  ``add_control_signal.py:119-123`` (a fixed 13x13 "3" bitmap stands in for the
  MNIST digit, which cannot be downloaded here).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import pandas as pd
import torch

IMG_SHAPE = (41, 49, 35)
IMG_DIM = 41 * 49 * 35
N_VOLS = 98
TR = 1.4
MOTION_COLS = ["x", "y", "z", "rot_x", "rot_y", "rot_z"]
GLM_COLS = ["task", "x", "y", "z", "xrot", "yrot", "zrot", "sex"]
INTENSITY_SCALE = 3284.5  # DataClass_GP.py:49

_GLYPH3 = [
    "..#########..",
    ".###########.",
    ".##.......###",
    "..........###",
    ".........###.",
    "....#######..",
    "....#######..",
    ".........###.",
    "..........###",
    "..........###",
    ".##.......###",
    ".###########.",
    "..#########..",
]


def block_design(n_vols: int = N_VOLS, control: bool = False) -> np.ndarray:
    """Binary task regressor per volume (utils.py:75-111)."""
    block = (np.arange(1, n_vols + 1) * TR) // 20
    odd = (block.astype(np.int64) % 2) == 1
    task = odd if not control else ~odd
    return task.astype(np.int64)


def hrf_taps() -> np.ndarray:
    """15 double-gamma taps at TR resolution (utils.py:22-36, vae_reg_GP.py:292).

    gamma.pdf(t, a) = t^(a-1) e^-t / Gamma(a); peak a=6, undershoot a=12 weighted
    0.35; normalised so the maximum is 0.6.
    """
    from math import factorial

    t = np.arange(0, 20, TR)
    peak = t ** 5 * np.exp(-t) / factorial(5)
    under = t ** 11 * np.exp(-t) / factorial(11)
    v = peak - 0.35 * under
    return v / v.max() * 0.6


def make_covariate_table(n_subjects: int, n_vols: int = N_VOLS, seed: int = 0,
                         control: bool = False, prefix: str = "sub-S") -> pd.DataFrame:
    """Per-volume table in the reference's CSV schema (without the index column)."""
    rng = np.random.default_rng(seed)
    task = block_design(n_vols, control)
    rows = []
    for s in range(n_subjects):
        e = rng.standard_normal((n_vols, 6))
        mot = np.zeros_like(e)
        mot[0] = e[0]
        for t in range(1, n_vols):  # AR(1), rho = 0.9, unit stationary variance
            mot[t] = 0.9 * mot[t - 1] + np.sqrt(1 - 0.81) * e[t]
        subj = f"{prefix}{s:04d}"
        nii = f"synthetic://{subj}.nii"
        for t in range(n_vols):
            rows.append((subj, t, nii, int(task[t]), *mot[t], s % 2))
    df = pd.DataFrame(rows, columns=["subjid", "volume #", "nii_path", "task",
                                     *MOTION_COLS, "sex"])
    for c in MOTION_COLS:  # utils.zscore: population std over all rows
        df[c] = (df[c] - df[c].mean()) / df[c].std(ddof=0)
    return df


def _ellipsoid_mask() -> np.ndarray:
    g = np.meshgrid(*[np.linspace(-1, 1, n) for n in IMG_SHAPE], indexing="ij")
    return ((g[0] ** 2 + g[1] ** 2 + g[2] ** 2) <= 1.0).astype(np.float32)


def control_glyph_map() -> np.ndarray:
    """Spatial support of the control signal, flat (V,)."""
    bmp = np.array([[c == "#" for c in r] for r in _GLYPH3], dtype=np.float32)
    vol = np.zeros(IMG_SHAPE, np.float32)
    vol[15:25, 34:47, 9:22] = bmp[None, :, :]
    return vol.reshape(-1)


def v1_blob_map() -> np.ndarray:
    g = np.meshgrid(*[np.arange(n, dtype=np.float32) for n in IMG_SHAPE], indexing="ij")
    d2 = ((g[0] - 20) / 6) ** 2 + ((g[1] - 8) / 5) ** 2 + ((g[2] - 14) / 5) ** 2
    return np.exp(-0.5 * d2).reshape(-1).astype(np.float32)


@dataclass
class Cohort:
    """A synthetic cohort: covariate table + a deterministic volume synthesiser."""
    table: pd.DataFrame
    control: bool
    signal: str            # "glyph" | "blob" | "none"
    intensity: float
    seed: int

    def __len__(self):
        return len(self.table)

    def subject_index(self) -> np.ndarray:
        subj = self.table["subjid"]
        uniq = subj.unique().tolist()
        return subj.map({s: i for i, s in enumerate(uniq)}).to_numpy(np.int64)

    def covariates(self) -> np.ndarray:
        return self.table[["task", *MOTION_COLS, "sex"]].to_numpy(np.float32)

    def volumes(self, device="cpu", rows=None) -> torch.Tensor:
        """(N,41,49,35) fp32 in [0,1] ("already divided by 3284.5")."""
        idx = np.arange(len(self.table)) if rows is None else np.asarray(rows)
        sidx = self.subject_index()
        mask = torch.from_numpy(_ellipsoid_mask()).reshape(-1)
        if self.signal == "glyph":
            smap = torch.from_numpy(control_glyph_map())
        elif self.signal == "blob":
            smap = torch.from_numpy(v1_blob_map())
        else:
            smap = torch.zeros(IMG_DIM)
        task = torch.from_numpy(self.table["task"].to_numpy(np.float32))
        if self.signal == "blob":  # HRF-shaped response per subject run
            taps = torch.from_numpy(hrf_taps().astype(np.float32))
            tv = task.clone()
            starts = [0] + [i for i in range(1, len(sidx)) if sidx[i] != sidx[i - 1]] + [len(sidx)]
            for s0, s1 in zip(starts[:-1], starts[1:]):      # one run per subject
                seg = task[s0:s1]
                full = torch.zeros(len(seg))
                for k in range(min(len(taps), len(seg))):
                    full[k:] += taps[k] * seg[:len(seg) - k]
                tv[s0:s1] = full
            task = tv
        out = torch.empty((len(idx), IMG_DIM), dtype=torch.float32)
        anat_cache = {}
        for o, r in enumerate(idx):
            s = int(sidx[r])
            if s not in anat_cache:
                g = torch.Generator().manual_seed(1000 + s + 7919 * self.seed)
                anat_cache[s] = 0.35 * mask * (1 + 0.1 * torch.randn(IMG_DIM, generator=g))
            g = torch.Generator().manual_seed(50_000 + int(r) + 7919 * self.seed)
            v = anat_cache[s] + 0.02 * torch.randn(IMG_DIM, generator=g)
            v = v + self.intensity * task[r] * smap
            out[o] = v.clamp_(0, 1)
        return out.reshape(-1, *IMG_SHAPE).to(device)


def make_cohort(n_subjects: int, config: str = "checker", seed: int = 0,
                n_vols: int = N_VOLS) -> Cohort:
    """config: 'control' (BASELINE config 1), 'checker' (2), 'v1' (3), 'cohort' (4)."""
    control = config == "control"
    tab = make_covariate_table(n_subjects, n_vols, seed, control)
    if control:
        return Cohort(tab, True, "glyph", 1000.0 / INTENSITY_SCALE, seed)
    return Cohort(tab, False, "blob", 0.05, seed)


def glm_maps_uniform(seed: int = 7) -> np.ndarray:
    """(V,8) U(0,1) maps (parity micro-inputs and config 2)."""
    return np.random.default_rng(seed).random((IMG_DIM, 8))


def glm_maps_lsq(cohort: Cohort, max_rows: int = 392, device="cpu", block: int = 98) -> np.ndarray:
    """Least-squares beta maps, max-scaled (get_beta_map_regularizer.py:94-103,
    utils.py:170-178): beta = (G'G)^-1 G' Y with G = [task, 6 motion]; sex map
    = mean(sex==1) - mean(sex==0).  Streams the volumes in blocks (vaegam.glm_maps)."""
    from . import glm_maps
    n = min(len(cohort), max_rows)
    cov = cohort.covariates()[:n].astype(np.float64)
    sex = cov[:, 7] > 0.5
    sums = {True: None, False: None}

    def blocks():
        for r0 in range(0, n, block):
            rows = np.arange(r0, min(n, r0 + block))
            y = cohort.volumes(device=device, rows=rows).reshape(len(rows), -1)
            for flag in (True, False):
                sel = torch.from_numpy(sex[rows] == flag).to(y.device)
                if bool(sel.any()):
                    part = y[sel].double().sum(0)
                    sums[flag] = part if sums[flag] is None else sums[flag] + part
            yield y, torch.from_numpy(cov[rows, :7])

    beta_only = glm_maps.lsq_beta_maps(blocks(), None, scale=False)      # (V, 8), sex column zero
    if sex.any() and (~sex).any():
        smap = (sums[True] / int(sex.sum()) - sums[False] / int((~sex).sum())).cpu().numpy()
    else:
        smap = np.zeros(IMG_DIM)
    maps = beta_only.T.copy()
    maps[7] = smap
    for i in range(maps.shape[0]):
        mx = maps[i].max()
        if mx != 0:
            maps[i] = maps[i] / mx
    return maps.T.copy()


def write_experiment(out_dir: str, n_subjects: int = 2, config: str = "checker",
                     seed: int = 0, glm: str = "uniform", test_subjects: int = 1):
    """Write train/test CSVs and the GLM-map CSV; returns (train_csv, test_csv, glm_csv, cohort)."""
    os.makedirs(out_dir, exist_ok=True)
    cohort = make_cohort(n_subjects, config, seed)
    test = make_cohort(test_subjects, config, seed + 1)
    # the reference z-scores train and test together in one table; keep one scale
    train_csv = os.path.join(out_dir, "train.csv")
    test_csv = os.path.join(out_dir, "test.csv")
    glm_csv = os.path.join(out_dir, "glm_maps.csv")
    cohort.table.to_csv(train_csv)
    test.table.to_csv(test_csv)
    if glm == "zeros":
        maps = np.zeros((IMG_DIM, 8))
    elif glm == "lsq":
        maps = glm_maps_lsq(cohort)
    else:
        maps = glm_maps_uniform()
    pd.DataFrame(maps, columns=GLM_COLS).to_csv(glm_csv)
    return train_csv, test_csv, glm_csv, cohort
