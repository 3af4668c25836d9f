"""Drop-in replacement for the reference's `build_model_recons.py` (same two entry points and
the same output tree; reference build_model_recons.py:15-116):
  reconstructions/{ckpt}_model_recons/<subj>/vol_<n>/recon_<key>.nii
  reconstructions/{ckpt}_avg_model_recons/<subj>/<key>_avg.nii   and   .../<key>_avg.nii
"""
import os

import numpy as np
import pandas as pd

from vaegam.nib_compat import nib

_ALL_MAPS = ['base', 'task', 'full_rec', 'x_mot', 'y_mot', 'z_mot', 'pitch_mot', 'roll_mot', 'yaw_mot', 'sex']


def mk_single_volumes(loader, model, csv_file, save_dir):
    dset = pd.read_csv(csv_file)
    subjs = dset.subjid.unique().tolist()
    ref_niis = dset.nii_path.unique().tolist()
    ckpt_num = str(model.epoch).zfill(3)
    subj_dirs = []
    for s in subjs:
        d = os.path.join(save_dir, 'reconstructions', '{}_model_recons'.format(ckpt_num), s)
        os.makedirs(d)
        subj_dirs.append(d)
    model.reconstruct(loader, ref_niis, subj_dirs)


def mk_avg_maps(csv_file, model, save_dir, mk_motion_maps=False):
    ckpt_num = str(model.epoch).zfill(3)
    single = os.path.join(save_dir, 'reconstructions', '{}_model_recons'.format(ckpt_num))
    avg = os.path.join(save_dir, 'reconstructions', '{}_avg_model_recons'.format(ckpt_num))
    os.makedirs(avg, exist_ok=True)
    dset = pd.read_csv(csv_file)
    ref_niis = dset.nii_path.unique().tolist()
    subjs = dset.subjid.unique().tolist()
    maps = _ALL_MAPS if mk_motion_maps else [_ALL_MAPS[i] for i in (0, 1, 2, 9)]
    dev = getattr(model, "_recon_avg", None)
    if dev is not None and dev["epoch"] == model.epoch and \
            dev["save_dirs"] == [os.path.join(single, s) for s in subjs] and float(dev["counts"].min()) > 0:
        # per-subject sums accumulated on the device by VAE.reconstruct: no file is read back
        from vaegam.step import IMG_KEYS
        means = (dev["sums"] / dev["counts"][:, None, None]).cpu().numpy()          # (S, 10, V) fp64
        for key in maps:
            k = IMG_KEYS.index(key)
            for i, s in enumerate(subjs):
                out = os.path.join(avg, s)
                os.makedirs(out, exist_ok=True)
                _save_map(means[i, k].reshape(41, 49, 35), ref_niis[i], out, key)
            _save_map(means[:, k].mean(0).reshape(41, 49, 35), ref_niis[0], avg, key)
        return
    for key in maps:
        grand = np.zeros((41, 49, 35), np.float64)
        for i, s in enumerate(subjs):
            sdir = os.path.join(single, s)
            acc, count = np.zeros((41, 49, 35), np.float64), 0
            for vol in os.listdir(sdir):
                acc += np.asarray(nib.load(os.path.join(sdir, vol, 'recon_{}.nii'.format(key))).dataobj)
                count += 1
            acc /= count
            out = os.path.join(avg, s)
            os.makedirs(out, exist_ok=True)
            _save_map(acc, ref_niis[i], out, key)
            grand += acc
        _save_map(grand / len(subjs), ref_niis[0], avg, key)


def _save_map(map, reference, save_dir, ext):
    ref = nib.load(reference)
    nib.save(nib.Nifti1Image(map, ref.affine, ref.header), os.path.join(save_dir, '{}_avg.nii'.format(ext)))
