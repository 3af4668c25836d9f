"""Inert nibabel stand-in (oracle/ test infrastructure).

The reference only uses nibabel for file I/O (`DataClass_GP.py:48`,
`vae_reg_GP.py:618-620`, `build_model_recons.py:88,113-116`), none of which is
on the hot path the oracle restates.
"""


class Nifti1Image:
    def __init__(self, dataobj=None, affine=None, header=None):
        self.dataobj, self.affine, self.header = dataobj, affine, header


def load(path):
    raise RuntimeError("nibabel stub: file I/O is outside the oracle's scope")


def save(img, path):
    raise RuntimeError("nibabel stub: file I/O is outside the oracle's scope")
