"""`nib`: the real nibabel when it is installed, otherwise the bundled NIfTI-1 subset (vaegam.nifti)."""
from .compat import install_nibabel

nib = install_nibabel()
