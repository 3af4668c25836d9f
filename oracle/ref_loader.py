"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference in the build container.

Only `tests/`, `tests/golden/make_golden.py` and ad-hoc validation may use this.
It needs `/root/reference`, which exists in the build container and NOT on the
GPU box, so nothing that runs on the GPU box may call `load_reference()`.

Three inert shims make the reference importable/runnable on CPU:
  1. stub `matplotlib`, `nibabel`, `umap` (oracle/shims; plotting / file I/O only);
  2. `gp._striped_matrix` without its hard-coded `.cuda()` (gp.py:113-119);
  3. `numpy.float = float` for build_model_recons.py:74,85.
The reference's arithmetic is untouched.
"""
import os
import sys

REFERENCE_DIR = os.environ.get("VAEGAM_REFERENCE_DIR", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "vae_reg_GP.py"))


def load_reference():
    """Return (vae_reg_GP, gp, utils) modules of the reference, shimmed for CPU."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    import numpy as np
    import torch

    if not hasattr(np, "float"):
        np.float = float
    for name in ("vae_reg_GP", "gp", "utils", "DataClass_GP", "build_model_recons"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_DIR):
            del sys.modules[name]  # a same-named drop-in module was imported first
    for p in (_SHIMS, REFERENCE_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_DIR)
    sys.path.insert(0, _SHIMS)
    try:
        import gp as ref_gp
        import utils as ref_utils
        import vae_reg_GP as ref_vae
    finally:
        sys.path.remove(_SHIMS)
        sys.path.remove(REFERENCE_DIR)
    # park them under private names so the drop-in modules can be imported later, and drop the
    # stubs from sys.modules (the reference modules keep their own references to them)
    for name in ("vae_reg_GP", "gp", "utils"):
        sys.modules["_reference_" + name] = sys.modules.pop(name)
    for name, mod in list(sys.modules.items()):
        if getattr(mod, "__file__", None) and str(mod.__file__).startswith(_SHIMS):
            del sys.modules[name]

    if not torch.cuda.is_available():
        def _striped_matrix_cpu(n):
            idx = torch.arange(n)
            return (idx[:, None] - idx[None, :]).abs().to(torch.float32)

        ref_gp._striped_matrix = _striped_matrix_cpu
    return ref_vae, ref_gp, ref_utils


class NullWriter:
    """Stands in for the TensorBoard SummaryWriter the ctor creates (vae_reg_GP.py:184)."""

    def __getattr__(self, name):
        return lambda *a, **k: None
