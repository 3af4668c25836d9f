#!/usr/bin/env python
"""Run an UNMODIFIED script of the reference tree (dannyfa/VAE-GAM) on the B200-native drop-in modules.

    python vae-gam_b200/run_reference.py /path/to/VAE-GAM/multsubj_reg_run_GP.py \
        --train_csv train.csv --test_csv test.csv --glm_maps glm.csv --save_dir out --epochs 3

Why a launcher: `python /path/to/VAE-GAM/multsubj_reg_run_GP.py` puts the SCRIPT's directory first on
`sys.path`, ahead of PYTHONPATH, so its bare imports (`import vae_reg_GP as vae_reg`, `import DataClass_GP as data`,
`import build_model_recons as recon`, `from utils import str2bool`; reference multsubj_reg_run_GP.py:14-17) would
find the reference's own PyTorch modules.  This launcher executes the script with `runpy.run_path`, which does not
touch `sys.path`, after putting THIS directory first — so the same unedited script trains through
libvaegam_sm100.so.  (`python -P script.py` with PYTHONPATH=vae-gam_b200 is the equivalent one-liner on
Python >= 3.11.)  It also installs the inert shims the reference needs on a current software stack: `np.float`
(build_model_recons.py:74,85) and, only if nibabel is not installed, the bundled NIfTI-1 subset under that name.

`--use-reference-recons` additionally lets the reference's OWN build_model_recons.py (also unmodified) drive
`VAE.reconstruct` instead of the drop-in module of the same name.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    use_ref_recons = "--use-reference-recons" in argv
    if use_ref_recons:
        argv.remove("--use-reference-recons")
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"run_reference: no such script: {script}")
    script_dir = os.path.dirname(script)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (HERE, script_dir)]
    sys.path.insert(0, HERE)
    from vaegam import compat
    compat.install()
    if use_ref_recons:       # the reference's own post-processing module, imported from where it lies
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_model_recons", os.path.join(script_dir, "build_model_recons.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["build_model_recons"] = mod
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
