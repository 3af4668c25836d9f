"""Least-squares GLM beta maps for the map regulariser (reference
get_beta_map_regularizer.py:94-107 and utils.scale_beta_maps, utils.py:169-178).

The reference builds `filtered_data` (V, S*T) and the stacked design matrix gamma (S*T, 7) in host
memory and evaluates  beta = (gamma' gamma)^-1 gamma' Y  with numpy.  Here the volumes stream through
in row blocks on whatever device they live on (a 64-subject cohort is 1.8 GB of fp32 volumes): only
gamma' Y (7, V) and gamma' gamma (7, 7) are accumulated, in fp64, and the 7x7 system is solved once.
The arithmetic is a plain library GEMM (torch.matmul) — an offline preprocessing step, not the
training hot path.  Output: the reference's CSV (pandas index column + task, x, y, z, xrot, yrot,
zrot, sex), which `vae_reg_GP.VAE(glm_maps=...)` reads.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import numpy as np
import pandas as pd
import torch

GLM_COLS = ["task", "x", "y", "z", "xrot", "yrot", "zrot", "sex"]


def lsq_beta_maps(blocks: Iterable[Tuple[torch.Tensor, torch.Tensor]], sex_map: Optional[torch.Tensor] = None,
                  scale: bool = True) -> np.ndarray:
    """blocks: iterable of (volumes (n, V) or (n, 41, 49, 35), gamma (n, 7) = [task, 6 motion regressors]).
    sex_map (V,): the higher-level sex contrast map (reference: an FSL cope image); zeros when None.
    Returns (V, 8) float64, each map divided by its maximum when `scale` (utils.scale_beta_maps)."""
    gty = gtg = None
    for vols, gamma in blocks:
        y = vols.reshape(vols.shape[0], -1).to(torch.float64)
        g = gamma.to(device=y.device, dtype=torch.float64)
        if g.shape != (y.shape[0], 7):
            raise ValueError(f"gamma block must be (n, 7), got {tuple(g.shape)} for {y.shape[0]} volumes")
        if gty is None:
            gty = torch.zeros(7, y.shape[1], dtype=torch.float64, device=y.device)
            gtg = torch.zeros(7, 7, dtype=torch.float64, device=y.device)
        gty += g.t() @ y
        gtg += g.t() @ g
    if gty is None:
        raise ValueError("no data")
    beta = torch.linalg.inv(gtg) @ gty                                   # (7, V), get_beta_map_regularizer.py:94-96
    sex = torch.zeros(1, beta.shape[1], dtype=torch.float64, device=beta.device) if sex_map is None \
        else sex_map.reshape(1, -1).to(device=beta.device, dtype=torch.float64)
    maps = torch.cat([beta, sex], 0)                                     # :99-101
    if scale:                                                            # :103, utils.py:169-178 (max scaling)
        mx = maps.max(dim=1, keepdim=True).values
        maps = torch.where(mx != 0, maps / torch.where(mx != 0, mx, torch.ones_like(mx)), maps)
    return maps.t().contiguous().cpu().numpy()


def write_glm_csv(path: str, maps: np.ndarray) -> str:
    """(V, 8) -> the reference's `scld_GLM_beta_maps.csv` layout (get_beta_map_regularizer.py:105-107)."""
    if maps.ndim != 2 or maps.shape[1] != 8:
        raise ValueError(f"maps must be (V, 8), got {maps.shape}")
    pd.DataFrame(maps, columns=GLM_COLS).to_csv(path)
    return path
