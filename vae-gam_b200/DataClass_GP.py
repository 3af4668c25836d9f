"""Drop-in replacement for the reference's `DataClass_GP.py` (same names and sample contract,
reference DataClass_GP.py:11-89): CSV row -> {'covariates' fp32 (8,), 'volume' fp32 (41,49,35)
already divided by 3284.5, 'subjid' int64, 'vol_num' fp64}.

Differences that do not change the contract: the 4-D image of a subject is decoded once and
cached (the reference re-decodes the whole file for every sample, :48); `synthetic://` paths
(written by vaegam.synthetic) are generated instead of read; `.nii/.nii.gz` files are read with
nibabel when it is installed, else with the bundled NIfTI-1 reader.

B200-first loader: `setup_data_loaders(..., resident=True)` (or VAEGAM_RESIDENT_LOADER=1;
default "auto" = whenever CUDA is available and the cohort fits the memory budget) returns `ResidentLoader`s:
every subject file is decoded ONCE, the whole cohort (281 KB per volume: 1 372 volumes = 386 MB, the 64-subject
cohort 1.76 GB of the 180 GB) lives in HBM, and a batch is one device-side gather — no per-sample Python, no
H2D copy per step.  Same iteration contract (dict batches, `len(loader.dataset)`, `loader.batch_size`,
last batch short), same shuffling RNG stream as `torch.utils.data.DataLoader(shuffle=True)`, optional
DistributedSampler-style rank shards.
"""
import os

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader, Dataset

from vaegam.nib_compat import nib

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


class ResidentCohort:
    """All samples of an FMRIDataset as device tensors: volume (N,41,49,35) fp32 (already / 3284.5),
    covariates (N,8) fp32, subjid (N,) int64, vol_num (N,) fp64 — the fields of the reference's batch dict
    (DataClass_GP.py:61-71)."""

    def __init__(self, dataset: "FMRIDataset", device):
        df = dataset.df
        n = len(df)
        self.dataset, self.device = dataset, torch.device(device)
        cols = df.columns
        cov = df[[cols[i] for i in range(4, 12)]].to_numpy(dtype=np.float64)
        self.covariates = torch.from_numpy(cov).float().to(self.device)
        subj = df[cols[1]].tolist()
        self.subjid = torch.tensor([dataset._subjects.index(s) for s in subj], dtype=torch.int64, device=self.device)
        self.vol_num = torch.from_numpy(df[cols[2]].to_numpy(dtype=np.float64)).to(self.device)
        self.volume = torch.empty(n, 41, 49, 35, dtype=torch.float32, device=self.device)
        paths = df[cols[3]].astype(str).tolist()
        vnum = df[cols[2]].to_numpy()
        if n and paths[0].startswith("synthetic://"):
            for lo in range(0, n, 256):           # generated in chunks, never the whole cohort on the host
                hi = min(n, lo + 256)
                dataset._synthetic_volume(lo)     # builds the Cohort object
                self.volume[lo:hi] = dataset._synthetic.volumes(rows=range(lo, hi)).to(self.device)
        else:
            by_file = {}
            for i, p in enumerate(paths):
                by_file.setdefault(p, []).append(i)
            for p, rows in by_file.items():       # one decode per subject file (the reference: one per SAMPLE, :48)
                img = np.asarray(nib.load(p).dataobj)
                vols = np.moveaxis(img[:, :, :, [int(vnum[i]) for i in rows]], -1, 0)
                t = torch.from_numpy(np.ascontiguousarray(vols, dtype=np.float32)).reshape(len(rows), 41, 49, 35)
                self.volume[torch.as_tensor(rows, device=self.device)] = (t / INTENSITY_MAX).to(self.device)

    @staticmethod
    def nbytes(n):
        return n * (41 * 49 * 35 * 4 + 8 * 4 + 16)


class ResidentLoader:
    """DataLoader-shaped iterable over a ResidentCohort.  Shuffling consumes the global CPU RNG exactly like
    torch's RandomSampler (one int64 seed per epoch, then `randperm(n, generator=seeded)`), so a given
    `torch.manual_seed` yields the batches the reference's DataLoader would.  `rank`/`world` take this rank's
    equal share of every epoch's permutation (DistributedSampler-style, tail dropped)."""

    def __init__(self, cohort: ResidentCohort, batch_size=32, shuffle=False, rank=0, world=1):
        self.cohort, self.dataset = cohort, cohort.dataset
        self.batch_size, self.shuffle = int(batch_size), bool(shuffle)
        self.rank, self.world = int(rank), int(world)

    def _order(self):
        n = len(self.dataset)
        # torch's DataLoader draws a base seed from the global RNG for every iterator it creates (even with
        # num_workers=0), then RandomSampler draws its own seed: consume both, in that order
        torch.empty((), dtype=torch.int64).random_()
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
        else:
            perm = torch.arange(n)
        if self.world > 1:
            per = n // self.world
            perm = perm[self.rank * per:(self.rank + 1) * per]
        return perm.to(self.cohort.device)

    def __len__(self):
        n = len(self.dataset) // self.world if self.world > 1 else len(self.dataset)
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        c = self.cohort
        order = self._order()
        for lo in range(0, order.numel(), self.batch_size):
            idx = order[lo:lo + self.batch_size]
            yield {'covariates': c.covariates[idx], 'volume': c.volume[idx], 'subjid': c.subjid[idx],
                   'vol_num': c.vol_num[idx]}


def _want_resident(resident, n_samples):
    if resident is None:
        env = os.environ.get("VAEGAM_RESIDENT_LOADER", "auto").lower()
        resident = {"1": True, "true": True, "0": False, "false": False}.get(env, "auto")
    if resident == "auto":
        if not torch.cuda.is_available():
            return False
        free, _ = torch.cuda.mem_get_info()
        return ResidentCohort.nbytes(n_samples) < 0.25 * free
    return bool(resident)


def setup_data_loaders(batch_size=32, shuffle=(True, False, False), train_csv='', test_csv='', resident=None,
                       device=None, rank=0, world=1):
    """Reference signature (DataClass_GP.py:73) + keyword-only extras: `resident` (True / False / None = env or
    auto), `device`, and `rank` / `world` shards for data-parallel training (resident loaders only)."""
    train = FMRIDataset(csv_file=train_csv, transform=ToTensor())
    test = FMRIDataset(csv_file=test_csv, transform=ToTensor())
    if _want_resident(resident, len(train) + len(test)):
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ctrain, ctest = ResidentCohort(train, dev), ResidentCohort(test, dev)
        return {'Shuffled_train': ResidentLoader(ctrain, batch_size, shuffle[0], rank, world),
                'UnShuffled_train': ResidentLoader(ctrain, batch_size, shuffle[1]),
                'test': ResidentLoader(ctest, batch_size, shuffle[2])}
    mk = lambda ds, sh: DataLoader(ds, batch_size=batch_size, shuffle=sh, num_workers=0)
    return {'Shuffled_train': mk(train, shuffle[0]), 'UnShuffled_train': mk(train, shuffle[1]),
            'test': mk(test, shuffle[2])}
