"""Golden LOSS CURVE from the UNMODIFIED reference (run in the build container, CPU):

    python tests/golden/make_curve.py            # writes tests/golden/curve_config1.npz

BASELINE configs[0] in miniature: the single-subject control experiment (vaegam.synthetic 'control' cohort: 98
volumes, glyph signal, neural_covariates=False, zero GLM maps, glm_reg_scale = 0) trained for 3 epochs of
unshuffled minibatches of 32 (32, 32, 32, 2) with the reference's own loop body — forward(train_mode=False),
zero_grad, backward, Adam.step (vae_reg_GP.py:425-429).  Before every forward the global seed is set to
1000 * epoch + batch, so the noise of each step is oracle.ref_port.draw_noise(B, seed) and the product can be fed
the same draws.  Stored: the 12 step losses, the 3 epoch means (sum / 98, as train_epoch reports them) and the recipe.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vae-gam_b200"))

from oracle.ref_loader import NullWriter, load_reference  # noqa: E402
from vaegam import synthetic as syn  # noqa: E402

RECIPE = dict(config="control", n_subjects=1, batch=32, epochs=3, param_seed=7, gp_kl_scale=10.0, glm_reg_scale=0.0,
              neural=False, m=6, glm="zeros")


def main():
    ref_vae, _, _ = load_reference()
    torch.set_num_threads(os.cpu_count())
    work = tempfile.mkdtemp(prefix="curve_")
    r = RECIPE
    tr, te, glm, coh = syn.write_experiment(work, n_subjects=r["n_subjects"], config=r["config"], glm=r["glm"])
    x_all = coh.volumes()
    cov_all = torch.from_numpy(coh.covariates().copy())
    ids_all = torch.from_numpy(coh.subject_index().copy())
    torch.manual_seed(r["param_seed"])
    m = ref_vae.VAE(save_dir=work, glm_maps=glm, csv_files=[tr, te], num_inducing_pts=r["m"], gp_kl_scale=r["gp_kl_scale"],
                    glm_reg_scale=r["glm_reg_scale"], neural_covariates=r["neural"])
    m.writer = NullWriter()
    n = x_all.shape[0]
    steps, epochs = [], []
    for ep in range(r["epochs"]):
        tot = 0.0
        for bi, lo in enumerate(range(0, n, r["batch"])):
            sl = slice(lo, min(n, lo + r["batch"]))
            torch.manual_seed(1000 * ep + bi)
            loss = m.forward(ids_all[sl], cov_all[sl], x_all[sl], 'train', train_mode=False)
            m.optimizer.zero_grad()
            loss.backward()
            m.optimizer.step()
            steps.append(float(loss))
            tot += float(loss)
            print(f"epoch {ep} batch {bi}: {float(loss):.4f}", flush=True)
        epochs.append(tot / n)
    np.savez_compressed(os.path.join(HERE, "curve_config1.npz"), step_losses=np.array(steps), epoch_losses=np.array(epochs),
                        recipe=np.array(repr(r)))
    print("epochs:", epochs)


if __name__ == "__main__":
    main()
