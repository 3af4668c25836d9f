"""Drop-in replacement for the reference's `utils.py`: same function names and semantics
(reference utils.py:22-389).  Only `hrf` and `get_xu_ranges` feed the hot path (as host-side
constants); the rest are CLI / preprocessing helpers and TensorBoard loggers.  matplotlib is
optional here: figure loggers become no-ops when it is not installed."""
import argparse
import re
from copy import deepcopy
from math import factorial

import numpy as np
import pandas as pd

try:
    import matplotlib
    matplotlib.use('agg')
    import matplotlib.pyplot as plt
except ImportError:
    plt = None

MOTION = ['x', 'y', 'z', 'rot_x', 'rot_y', 'rot_z']


def hrf(times):
    """Double-gamma HRF (peak shape 6, undershoot shape 12 weighted 0.35), scaled so that its
    maximum is 0.6.  gamma.pdf(t, a) = t^(a-1) exp(-t) / (a-1)!  (reference utils.py:22-36)."""
    t = np.asarray(times, dtype=np.float64)
    v = t ** 5 * np.exp(-t) / factorial(5) - 0.35 * t ** 11 * np.exp(-t) / factorial(11)
    return v / np.max(v) * 0.6


def get_xu_ranges(csv_files, eps=1e-3):
    """[min - eps, max + eps] of each motion column over train + test CSV (reference utils.py:39-56)."""
    frames = [pd.read_csv(csv_files[0]), pd.read_csv(csv_files[1])]
    return [[min(f[c].min() for f in frames) - eps, max(f[c].max() for f in frames) + eps] for c in MOTION]


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def _blocks(vol_times, first_block_on):
    odd = (np.asarray(vol_times) // 20).astype(np.int64) % 2 == 1
    return (odd ^ first_block_on).astype(np.int64)


def stimulus_to_neural(vol_times):
    """20 s blocks, rest first (reference utils.py:75-91)."""
    return _blocks(vol_times, False)


def control_stimulus_to_neural(vol_times):
    """20 s blocks, stimulus first (reference utils.py:93-111)."""
    return _blocks(vol_times, True)


def zscore(df):
    """z-score the six motion columns over all rows, population std (reference utils.py:113-123)."""
    for col in MOTION:
        df[col] = (df[col] - df[col].mean()) / df[col].std(ddof=0)
    return df


def mk_spherical_mask(size, radius):
    """size^3 array with an L1 ball of the given radius at the centre (reference utils.py:126-150)."""
    c = size // 2
    g = np.abs(np.arange(size) - c)
    dist = g[:, None, None] + g[None, :, None] + g[None, None, :]
    return (dist <= radius).astype(np.float64)


def read_design_mat(mat_file_path):
    """FSL design.mat: numbers start on line 6, tab separated (reference utils.py:152-167)."""
    with open(mat_file_path) as f:
        rows = f.readlines()[5:]
    return np.array([[float(v) for v in re.split(r'\t+', r.rstrip())] for r in rows])


def scale_beta_maps(beta_maps):
    """Divide each map by its maximum (reference utils.py:169-178)."""
    for i in range(beta_maps.shape[0]):
        beta_maps[i, :] = beta_maps[i, :] / np.amax(beta_maps[i, :].flatten())
    return beta_maps


# ----------------------------------------------------------------------------- TensorBoard
def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def log_qu_plots(epoch, gp_params, writer, log_type):
    """q(u) mean +/- 2 sigma for the six motion covariates (reference utils.py:182-273); works for
    any number of inducing points."""
    if plt is None:
        return
    keys = ['x', 'y', 'z', 'xrot', 'yrot', 'zrot']
    fig, axs = plt.subplots(3, 2, figsize=(15, 15))
    for ax, key in zip(axs.reshape(-1), keys):
        m = _np(gp_params[key]['qu_m']).reshape(-1)
        two = 2 * np.sqrt(np.diag(_np(gp_params[key]['qu_S'])))
        xu = _np(gp_params[key]['xu'])
        ax.plot(xu, m, c='darkblue', alpha=0.5, label='q(u) posterior mean')
        ax.fill_between(xu, m - two, m + two, color='lightblue', alpha=0.3, label='2 sigma')
        ax.legend(loc='best')
        ax.set_title('q(u) {} covariate at epoch {}'.format(key, epoch))
    writer.add_figure("q(u)_{}".format(log_type), fig)


def log_qkappa_plots(gp_params, writer, log_type):
    """q(kappa) densities for all covariates (reference utils.py:275-345)."""
    if plt is None:
        return
    fig, axs = plt.subplots(3, 3, figsize=(15, 15))
    for ax, key in zip(axs.reshape(-1), gp_params.keys()):
        mu = float(_np(gp_params[key]['sa']).reshape(-1)[0])
        sd = float(np.exp(_np(gp_params[key]['logstd']).reshape(-1)[0]))
        xs = np.linspace(mu - 2.326 * sd, mu + 2.326 * sd, 100)
        ax.plot(xs, np.exp(-0.5 * ((xs - mu) / sd) ** 2) / (sd * np.sqrt(2 * np.pi)), lw=2, alpha=0.5)
        ax.set_title('{} q(k)'.format(key))
    writer.add_figure("q(k)_{}".format(log_type), fig)


def log_beta(writer, xq, beta_mean, beta_cov, covariate_name, log_type):
    """Gain posterior mean +/- 2 sigma against the covariate (reference utils.py:347-371)."""
    if plt is None:
        return
    frame = pd.DataFrame({'xq': _np(xq), 'mean': _np(beta_mean),
                          'two_sig': 2 * np.sqrt(np.diag(_np(beta_cov)))}).sort_values(by=["xq"])
    fig = plt.figure()
    plt.plot(frame['xq'], frame['mean'], c='darkblue', alpha=0.5, label='Beta posterior mean')
    plt.fill_between(frame['xq'], frame['mean'] - frame['two_sig'], frame['mean'] + frame['two_sig'],
                     color='lightblue', alpha=0.3, label='2 sigma')
    plt.legend(loc='best')
    plt.title('Beta_{}'.format(covariate_name))
    writer.add_figure("Beta/{}_{}".format(covariate_name, log_type), fig)


def log_map(writer, img_shape, map, slice, map_name, batch_size, log_type):
    """One sagittal slice per batch element as a TensorBoard image (reference utils.py:373-389)."""
    vols = np.asarray(map).reshape((batch_size, img_shape[0], img_shape[1], img_shape[2]))
    for i in range(batch_size):
        writer.add_image('{}_{}_{}/{}'.format(map_name, log_type, slice, i), np.rot90(vols[i, slice, :, :]),
                         dataformats='HW')
