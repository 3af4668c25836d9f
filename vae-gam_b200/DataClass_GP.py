"""Drop-in replacement for the reference's `DataClass_GP.py` (same names and sample contract,
reference DataClass_GP.py:11-89): CSV row -> {'covariates' fp32 (8,), 'volume' fp32 (41,49,35)
already divided by 3284.5, 'subjid' int64, 'vol_num' fp64}.

Differences that do not change the contract: the 4-D image of a subject is decoded once and
cached (the reference re-decodes the whole file for every sample, :48); `synthetic://` paths
(written by vaegam.synthetic) are generated instead of read; `.nii/.nii.gz` files are read with
the bundled NIfTI-1 reader.
"""
import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader, Dataset

import nibabel as nib

INTENSITY_MAX = 3284.5   # global scale used by the reference (DataClass_GP.py:49)


class FMRIDataset(Dataset):
    def __init__(self, csv_file, transform=None):
        self.df = pd.read_csv(csv_file)
        self.transform = transform
        self._subjects = self.df.subjid.unique().tolist()
        self._cache = {}
        self._synthetic = None

    def __len__(self):
        return len(self.df)

    def _synthetic_volume(self, idx):
        if self._synthetic is None:
            from vaegam import synthetic as syn
            tab = self.df.drop(columns=[self.df.columns[0]]) if self.df.columns[0].startswith("Unnamed") else self.df
            control = bool(tab["task"].iloc[0] == 1)        # control design starts with a task block
            self._synthetic = syn.Cohort(tab.reset_index(drop=True), control, "glyph" if control else "blob",
                                         1000.0 / INTENSITY_MAX if control else 0.05, 0)
        return self._synthetic.volumes(rows=[idx])[0].numpy()

    def __getitem__(self, idx):
        row = self.df.iloc[idx]
        subj, vol_num, nii = row.iloc[1], row.iloc[2], row.iloc[3]
        if str(nii).startswith("synthetic://"):
            scld_vol = self._synthetic_volume(idx)
        else:
            if nii not in self._cache:
                self._cache = {nii: np.asarray(nib.load(nii).dataobj)}     # keep one subject resident
            scld_vol = np.true_divide(self._cache[nii][:, :, :, int(vol_num)], INTENSITY_MAX).reshape(41, 49, 35)
        sample = {'subj_idx': self._subjects.index(subj), 'subj': subj, 'volume': scld_vol, 'vol_num': vol_num,
                  'task': row.iloc[4], 'trans_x': row.iloc[5], 'trans_y': row.iloc[6], 'trans_z': row.iloc[7],
                  'rot_x': row.iloc[8], 'rot_y': row.iloc[9], 'rot_z': row.iloc[10], 'sex': row.iloc[11]}
        return self.transform(sample) if self.transform else sample


class ToTensor(object):
    "Converts a sample's arrays to the tensors the model consumes."

    def __call__(self, sample):
        covars = np.array([sample[k] for k in ('task', 'trans_x', 'trans_y', 'trans_z', 'rot_x', 'rot_y', 'rot_z',
                                               'sex')], dtype=np.float64)
        return {'covariates': torch.from_numpy(covars).float(),
                'volume': torch.from_numpy(np.ascontiguousarray(sample['volume'])).float(),
                'subjid': torch.tensor(sample['subj_idx'], dtype=torch.int64),
                'vol_num': torch.tensor(sample['vol_num'], dtype=torch.float64)}


def setup_data_loaders(batch_size=32, shuffle=(True, False, False), train_csv='', test_csv=''):
    train = FMRIDataset(csv_file=train_csv, transform=ToTensor())
    test = FMRIDataset(csv_file=test_csv, transform=ToTensor())
    mk = lambda ds, sh: DataLoader(ds, batch_size=batch_size, shuffle=sh, num_workers=0)
    return {'Shuffled_train': mk(train, shuffle[0]), 'UnShuffled_train': mk(train, shuffle[1]),
            'test': mk(test, shuffle[2])}
