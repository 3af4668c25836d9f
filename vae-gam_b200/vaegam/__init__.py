"""Internals of the B200-native VAE-GAM hot path: ctypes binding of libvaegam_sm100.so
(`native`), the step engine / autograd bridge / flat optimizer (`step`), data-parallel
training (`dp`), synthetic cohorts (`synthetic`) and a minimal NIfTI-1 codec (`nifti`).
The public surface is the set of drop-in modules one directory up."""
__version__ = "0.1.0"

from . import compat as _compat

_compat.install_numpy_aliases()      # `np.float` for the reference's build_model_recons.py
