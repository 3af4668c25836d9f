"""Inert compatibility shims for running the UNMODIFIED reference scripts against this package
(SURVEY F4/F5): `np.float` & friends (removed in NumPy 1.24; reference build_model_recons.py:74,85) and,
only when the real package is not installed, a `nibabel` module backed by `vaegam.nifti` (the reference
calls `nib.load(...).dataobj/.affine/.header`, `nib.Nifti1Image`, `nib.save`; DataClass_GP.py:48,
vae_reg_GP.py:618-620, build_model_recons.py:88,113-116).  Nothing here shadows an installed package."""
import importlib.util
import sys

import numpy as np


def install_numpy_aliases():
    for name, typ in (("float", float), ("int", int), ("bool", bool), ("object", object), ("complex", complex)):
        if name not in np.__dict__:          # attribute access would raise (and warn) on NumPy >= 1.24
            setattr(np, name, typ)


def install_nibabel():
    """Register vaegam.nifti as `nibabel` iff no real nibabel can be imported."""
    usable = lambda m: hasattr(m, "__version__") and hasattr(m, "load") and hasattr(m, "Nifti1Image")
    if "nibabel" in sys.modules and usable(sys.modules["nibabel"]):
        return sys.modules["nibabel"]
    try:
        found = importlib.util.find_spec("nibabel") is not None
    except (ImportError, ValueError):
        found = False
    if found:
        try:
            import nibabel
            if usable(nibabel):                  # an inert test stub (no __version__) does not count
                return nibabel
        except ImportError:
            pass
    from . import nifti
    sys.modules["nibabel"] = nifti
    return nifti


def install():
    install_numpy_aliases()
    return install_nibabel()
