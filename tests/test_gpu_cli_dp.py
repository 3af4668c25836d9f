"""GPU tests of the two outer layers of the hot path:

* the reference's CLI (`multsubj_reg_run_GP.py`) and post-processing module (`build_model_recons.py`),
  UNMODIFIED, driven through `vae-gam_b200/run_reference.py` on a two-subject cohort backed by NIfTI files:
  train_loop (2 epochs, TensorBoard hooks every step), checkpoint, then `--recons_only --from_ckpt`.  The
  reference sources are taken from `baseline/_ref/` (staged by `__graft_entry__.build()` in the build container,
  git-ignored, travels to the GPU box) or `/root/reference`; without either the package's own CLI of the same
  flags runs instead;
* data-parallel training on two real GPUs over NCCL: the all-reduced gradient equals the mean of the two ranks'
  fp64-oracle gradients, and parameters stay bit-identical across ranks after graph-replayed steps.
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import nifti_experiment

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vae-gam_b200")


def _reference_dir():
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(d, "multsubj_reg_run_GP.py")):
            return d
    return None


def test_reference_cli_runs_unchanged(tmp_path):
    csv, glm, coh = nifti_experiment(tmp_path, n_subjects=2, n_vols=6)
    ref = _reference_dir()
    launcher = [sys.executable, os.path.join(PKG, "run_reference.py")]
    if ref is not None:
        work = tmp_path / "VAE-GAM"
        work.mkdir()
        for f in os.listdir(ref):                       # the whole checkout: its own modules sit next to the script
            if f.endswith(".py"):
                shutil.copy(os.path.join(ref, f), str(work / f))
        script = str(work / "multsubj_reg_run_GP.py")
        launcher += ["--use-reference-recons"]          # the reference's build_model_recons.py too (np.float, nibabel)
        sha_before = {f: open(os.path.join(ref, f), "rb").read() for f in ("multsubj_reg_run_GP.py", "build_model_recons.py")}
        for f, blob in sha_before.items():
            assert open(str(work / f), "rb").read() == blob
    else:
        script = os.path.join(PKG, "multsubj_reg_run_GP.py")
    out1 = tmp_path / "run_train"
    env = dict(os.environ, VAEGAM_TB_EVERY="1", PYTHONPATH="")
    common = ["--train_csv", csv, "--test_csv", csv, "--glm_maps", glm, "--batch-size", "4", "--split", "6", "--seed", "2"]
    r = subprocess.run(launcher + [script] + common + ["--save_dir", str(out1), "--epochs", "2", "--save_freq", "1",
                                                       "--test_freq", "1"],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "Epoch: 1 Average loss" in r.stdout and "Test loss" in r.stdout
    ck = out1 / "checkpoint_001.tar"
    assert ck.is_file(), os.listdir(str(out1))
    # train_loop -> project_latent, plot_GPs, reconstructions and averages at epoch 2
    assert (out1 / "002_latent_means.csv").is_file() and (out1 / "002_GP_plots" / "002_GP_x_full.csv").is_file()
    subj = sorted(os.listdir(str(out1 / "reconstructions" / "002_model_recons")))
    assert len(subj) == 2
    vols = os.listdir(str(out1 / "reconstructions" / "002_model_recons" / subj[0]))
    assert len(vols) == 6 and len(os.listdir(str(out1 / "reconstructions" / "002_model_recons" / subj[0] / vols[0]))) == 10
    assert (out1 / "reconstructions" / "002_avg_model_recons" / "full_rec_avg.nii").is_file()
    # TensorBoard hooks ran inside forward (VAEGAM_TB_EVERY=1): an event file with image summaries exists
    events = [os.path.join(dp_, f) for dp_, _, fs in os.walk(str(out1 / "run")) for f in fs if "tfevents" in f]
    assert events and max(os.path.getsize(e) for e in events) > 10_000
    # ---- --recons_only from the checkpoint (BASELINE configs[4])
    out2 = tmp_path / "run_recons"
    r = subprocess.run(launcher + [script] + common + ["--save_dir", str(out2), "--recons_only", "True", "--from_ckpt", "True",
                                                       "--ckpt_path", str(ck)],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "Loading model state from" in r.stdout
    from vaegam.nib_compat import nib
    # the checkpoint written during epoch index 1 carries model.epoch == 2 (train_epoch has already advanced it,
    # reference vae_reg_GP.py:433,712-714), so the reconstruction tree is numbered 002 again
    avg = np.asarray(nib.load(str(out2 / "reconstructions" / "002_avg_model_recons" / "base_avg.nii")).dataobj)
    assert avg.shape == (41, 49, 35) and np.isfinite(avg).all() and 0 < avg.mean() < 1
    one = np.asarray(nib.load(str(out2 / "reconstructions" / "002_model_recons" / subj[0] / vols[0] / "recon_full_rec.nii")).dataobj)
    assert one.shape == (41, 49, 35) and np.isfinite(one).all()


def _dp_worker(rank, world, port, tmp, files, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import vae_reg_GP
    from oracle import ref_port as rp
    from vaegam import dp, synthetic as syn
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    tr, te, glm = files
    torch.manual_seed(1 + rank)                           # different initial values: the broadcast must fix that
    model = vae_reg_GP.VAE(save_dir=os.path.join(tmp, f"r{rank}"), glm_maps=glm, csv_files=[tr, te])
    model.writer = vae_reg_GP._NullWriter()
    model.arith = "fp32"                                  # tight comparison with the oracle's gradients
    reducer = dp.GradientAllReduce(model._flat, model.optimizer)
    reducer.broadcast_parameters()
    coh = syn.make_cohort(2, "checker", seed=0)
    B = 4
    rows = range(rank * B, (rank + 1) * B)                # each rank its own minibatch
    x = coh.volumes(rows=rows)
    cov = torch.from_numpy(coh.covariates()[rank * B:(rank + 1) * B].copy())
    ids = torch.zeros(B, dtype=torch.int64)
    noise = rp.draw_noise(B, seed=50 + rank)
    # ---- one eager step up to the reduced gradient (overlapped, bucketed all-reduce on the comm stream)
    eng = model._get_engine()
    sb = eng.forward(x.to(dev), cov.to(dev), noise, False)
    eng.backward(sb, reducer)
    reducer.finish()
    torch.cuda.synchronize()
    got32, got64 = model._flat.grad32.clone().cpu(), model._flat.grad64.clone().cpu()
    # ---- the two ranks' fp64 oracle gradients, averaged (both ranks compute both: no communication in the check)
    P0 = {k: v.cpu() for k, v in rp.params_from_module(model).items()}
    want = None
    for r in range(world):
        Pd = rp.cast_params(P0, torch.float64, requires_grad=True)
        xr = coh.volumes(rows=range(r * B, (r + 1) * B)).double()
        cr = torch.from_numpy(coh.covariates()[r * B:(r + 1) * B].copy()).double()
        out = rp.step(Pd, xr, cr, rp.draw_noise(B, seed=50 + r), 10.0, 1.0, True, keep_maps=False)
        out["tot"].backward()
        g = {n: Pd[n].grad.clone() for n, _ in model.named_parameters()}
        want = g if want is None else {n: want[n] + g[n] for n in g}
    worst = 0.0
    gmax = max(float(v.norm()) for v in want.values())
    for n, p in model.named_parameters():
        dt, off, k = model._flat.slices[n]
        mine = (got32 if dt == torch.float32 else got64)[off:off + k].double().view(p.shape)     # SUM over ranks
        if n.startswith(("logkvar_", "logls_")):
            continue
        err = float((mine - want[n]).norm() / (want[n].norm() + 1e-5 * gmax))
        worst = max(worst, err)
    # ---- graph-replayed training steps keep the replicas bit-identical
    for step in range(10):
        model.train_batch(ids.to(dev), cov.to(dev), x.to(dev), _noise=rp.draw_noise(B, seed=100 * step + rank), reducer=reducer)
    torch.cuda.synchronize()
    graphed = any(gs.graph is not None for gs in model._graph_steps.values())
    ref32 = model._flat.flat32.clone()
    dist.broadcast(ref32, 0)
    same = torch.tensor([int(torch.equal(ref32, model._flat.flat32))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret["worst"], ret["same"], ret["graphed"] = worst, int(same.item()), graphed
        ret["scale"] = float(model.optimizer.grad_scale)
    model._graph_steps.clear()
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)                                           # skip NCCL teardown under live graphs (as bench.py does)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_dp_two_gpus_allreduced_gradient_is_the_mean_and_replicas_stay_identical(tmp_path):
    import torch.multiprocessing as mp
    from vaegam import synthetic as syn
    tr, te, glm, _ = syn.write_experiment(str(tmp_path), n_subjects=2, config="checker", glm="uniform")
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, str(tmp_path), (tr, te, glm), ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret["scale"] == 0.5                    # Adam applies SUM * 1/world = the mean
    assert ret["worst"] < 5e-3, ret["worst"]      # summed gradient == sum of the ranks' oracle gradients (fp32 kernels)
    assert ret["same"] == 1 and ret["graphed"]
