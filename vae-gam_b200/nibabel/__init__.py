"""The subset of nibabel the reference calls, backed by vaegam.nifti (nibabel itself is not
installed here).  Import order matters only when the real nibabel is absent."""
from vaegam.nifti import Nifti1Header, Nifti1Image, load, save  # noqa: F401

__version__ = "0.vaegam"
