"""Host-side behaviour of the drop-in modules, CPU only: parameter naming / init order, checkpoint
layout and interchange, loader contract, synthetic CSV schema, NIfTI codec, utils parity, and
live comparison with the reference when it is mounted."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle.ref_loader import reference_available

HAVE_REF = reference_available()


@pytest.fixture(scope="module")
def experiment(tmp_path_factory):
    from vaegam import synthetic as syn
    d = str(tmp_path_factory.mktemp("exp"))
    tr, te, glm, coh = syn.write_experiment(d, n_subjects=2, config="checker")
    return d, tr, te, glm, coh


def make_model(experiment, seed=1, **kw):
    import vae_reg_GP
    d, tr, te, glm, coh = experiment
    torch.manual_seed(seed)
    return vae_reg_GP.VAE(save_dir=d, glm_maps=glm, csv_files=[tr, te], device_name="cpu", **kw)


def test_parameter_names_and_shapes(experiment):
    from vaegam.step import PARAM_ORDER
    m = make_model(experiment)
    names = [n for n, _ in m.named_parameters()]
    assert names == PARAM_ORDER and len(names) == 97
    assert sum(p.numel() for p in m.parameters()) == 1564424
    assert m.epsilon.dtype == torch.float64 and tuple(m.epsilon.shape) == (41, 49, 35)
    assert set(m.gp_params) == {'task', 'x', 'y', 'z', 'xrot', 'yrot', 'zrot', 'sex'}
    assert set(m.gp_params['task']) == {'sa', 'logstd'}
    assert set(m.gp_params['x']) == {'xu', 'qu_m', 'qu_S', 'logkvar', 'log_ls', 'sa', 'logstd'}
    assert m.gp_params['x']['qu_m'] is m.qu_m_x        # dict aliases the registered Parameters
    assert m._flat.is_packed()


def test_ctor_rejects_other_architectures(experiment):
    import vae_reg_GP
    d, tr, te, glm, coh = experiment
    with pytest.raises(ValueError):
        vae_reg_GP.VAE(nf=4, save_dir=d, glm_maps=glm, csv_files=[tr, te], device_name="cpu")


def test_checkpoint_round_trip_and_layout(experiment):
    d = experiment[0]
    m = make_model(experiment, seed=3)
    m.epoch = 7
    m.loss['train'][6] = 1.5
    m.save_state("ck_007.tar")
    ck = torch.load(os.path.join(d, "ck_007.tar"), weights_only=False)
    layer_keys = {'fc1', 'fc2', 'fc31', 'fc32', 'fc33', 'fc41', 'fc42', 'fc43', 'fc5', 'fc6', 'fc7', 'fc8', 'bn1', 'bn3',
                  'bn5', 'bnt1', 'bnt3', 'bnt5', 'conv1', 'conv2', 'conv3', 'conv4', 'conv5', 'convt1', 'convt2',
                  'convt3', 'convt4', 'convt5'}
    assert set(ck) == layer_keys | {'optimizer_state', 'loss', 'z_dim', 'epoch', 'lr', 'save_dir', 'epsilon',
                                    'glm_reg_scale', 'gp_kl_scale', 'inducing_pts', 'gp_params'}
    assert set(ck['bn1']) == {'weight', 'bias'} and ck['z_dim'] == 41
    m2 = make_model(experiment, seed=4)
    m2.load_state(os.path.join(d, "ck_007.tar"))
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), n
    assert m2.epoch == 7 and m2.loss['train'][6] == 1.5 and m2._flat.is_packed()


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_same_seed_gives_reference_initialisation_and_checkpoints_interchange(experiment):
    from oracle.ref_loader import NullWriter, load_reference
    d, tr, te, glm, coh = experiment
    ref_vae, _, _ = load_reference()
    torch.manual_seed(11)
    ref = ref_vae.VAE(save_dir=d, glm_maps=glm, csv_files=[tr, te])
    ref.writer = NullWriter()
    mine = make_model(experiment, seed=11)
    for (n, a), (n2, b) in zip(ref.named_parameters(), mine.named_parameters()):
        assert n == n2 and torch.equal(a.detach(), b.detach()), n
    # reference checkpoint -> this implementation
    ref.epoch = 5
    ref.save_state("ref_ck.tar")
    other = make_model(experiment, seed=12)
    other.load_state(os.path.join(d, "ref_ck.tar"))
    for (n, a), (_, b) in zip(ref.named_parameters(), other.named_parameters()):
        assert torch.equal(a.detach(), b.detach()), n
    # this implementation's checkpoint -> reference
    mine.save_state("mine_ck.tar")
    torch.manual_seed(13)
    ref2 = ref_vae.VAE(save_dir=d, glm_maps=glm, csv_files=[tr, te])
    ref2.load_state(os.path.join(d, "mine_ck.tar"))
    for (n, a), (_, b) in zip(mine.named_parameters(), ref2.named_parameters()):
        assert torch.equal(a.detach(), b.detach()), n


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_utils_match_reference():
    import utils as mine
    from oracle.ref_loader import load_reference
    _, _, ref = load_reference()
    t = np.arange(0, 20, 1.4)
    assert np.allclose(mine.hrf(t), ref.hrf(t), atol=1e-14)
    vt = np.arange(1, 99) * 1.4
    assert np.array_equal(mine.stimulus_to_neural(vt), ref.stimulus_to_neural(vt))
    assert np.array_equal(mine.control_stimulus_to_neural(vt), ref.control_stimulus_to_neural(vt))
    assert np.array_equal(mine.mk_spherical_mask(11, 3), ref.mk_spherical_mask(11, 3))
    for v in ("yes", "False", "1", "n"):
        assert mine.str2bool(v) == ref.str2bool(v)


def test_synthetic_schema_and_xu_ranges(experiment):
    import utils
    d, tr, te, glm, coh = experiment
    df = pd.read_csv(tr)
    assert list(df.columns[1:]) == ["subjid", "volume #", "nii_path", "task", "x", "y", "z", "rot_x", "rot_y", "rot_z", "sex"]
    assert len(df) == 2 * 98
    assert abs(df["x"].mean()) < 1e-9 and abs(df["x"].std(ddof=0) - 1) < 1e-9
    g = pd.read_csv(glm)
    assert g.shape == (70315, 9)
    r = utils.get_xu_ranges([tr, te])
    assert len(r) == 6 and all(lo < hi for lo, hi in r)


def test_loader_contract(experiment):
    import DataClass_GP as data
    d, tr, te, glm, coh = experiment
    loaders = data.setup_data_loaders(batch_size=5, train_csv=tr, test_csv=te)
    assert set(loaders) == {'Shuffled_train', 'UnShuffled_train', 'test'}
    batch = next(iter(loaders['UnShuffled_train']))
    assert batch['volume'].shape == (5, 41, 49, 35) and batch['volume'].dtype == torch.float32
    assert batch['covariates'].shape == (5, 8) and batch['covariates'].dtype == torch.float32
    assert batch['subjid'].dtype == torch.int64 and batch['vol_num'].dtype == torch.float64
    assert len(loaders['test'].dataset) == 98
    assert torch.allclose(batch['covariates'], torch.from_numpy(coh.covariates()[:5]))
    assert torch.allclose(batch['volume'], coh.volumes(rows=range(5)))


def test_loader_reads_nifti_files(tmp_path):
    import DataClass_GP as data
    import nibabel as nib
    rng = np.random.default_rng(0)
    vol4d = (rng.random((41, 49, 35, 3)) * 3000).astype(np.float32)
    path = str(tmp_path / "sub-A.nii.gz")
    nib.save(nib.Nifti1Image(vol4d, np.diag([3.0, 3.0, 3.0, 1.0])), path)
    rows = [("sub-A", t, path, t % 2, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 1) for t in range(3)]
    df = pd.DataFrame(rows, columns=["subjid", "volume #", "nii_path", "task", "x", "y", "z", "rot_x", "rot_y", "rot_z", "sex"])
    csv = str(tmp_path / "t.csv")
    df.to_csv(csv)
    ds = data.FMRIDataset(csv, transform=data.ToTensor())
    s = ds[2]
    assert torch.allclose(s['volume'], torch.from_numpy(vol4d[..., 2] / 3284.5))
    assert s['covariates'].tolist() == pytest.approx([0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 1])


def test_nifti_round_trip(tmp_path):
    from vaegam import nifti
    a = np.random.default_rng(1).standard_normal((41, 49, 35)).astype(np.float32)
    aff = np.array([[3., 0, 0, -60], [0, 3., 0, -70], [0, 0, 3.5, -50], [0, 0, 0, 1]])
    for name in ("a.nii", "a.nii.gz"):
        p = str(tmp_path / name)
        nifti.save(nifti.Nifti1Image(a, aff), p)
        b = nifti.load(p)
        assert np.array_equal(np.asarray(b.dataobj), a) and np.allclose(b.affine, aff)
    d = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    p = str(tmp_path / "d.nii")
    nifti.save(nifti.Nifti1Image(d, np.eye(4)), p)
    assert np.array_equal(np.asarray(nifti.load(p).dataobj), d)
    with pytest.raises(ValueError):
        open(str(tmp_path / "bad.nii"), "wb").write(b"\0" * 400)
        nifti.load(str(tmp_path / "bad.nii"))


def test_do_hrf_conv_and_linw_kl_match_oracle(experiment):
    from oracle import ref_port as rp
    m = make_model(experiment)
    g = torch.randn(40, dtype=torch.float64)
    assert torch.allclose(m.do_hrf_conv(g), rp.hrf_fir(g, rp.hrf_taps()), atol=1e-12)
    g5 = torch.randn(5, dtype=torch.float64)      # shorter than the 15 taps
    assert torch.allclose(m.do_hrf_conv(g5), rp.hrf_fir(g5, rp.hrf_taps()), atol=1e-12)
    sa, ls = torch.tensor(1.3), torch.tensor(-0.4)
    assert torch.allclose(m.calc_linW_KL(sa, ls.exp()), rp.lin_w_kl(sa, ls))
    ref = torch.distributions.kl.kl_divergence(torch.distributions.Normal(sa, ls.exp()), torch.distributions.Normal(1.0, 0.5))
    assert torch.allclose(m.calc_linW_KL(sa, ls.exp()), ref)


def test_forward_refuses_to_run_without_cuda(experiment):
    from vaegam import native
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = make_model(experiment)
    d, tr, te, glm, coh = experiment
    x = coh.volumes(rows=range(2))
    with pytest.raises(native.NativeError):
        m.forward(torch.zeros(2, dtype=torch.int64), torch.from_numpy(coh.covariates()[:2]), x, 'train', train_mode=False)
    with pytest.raises(native.NativeError):
        m.encode(x)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vae-gam_b200")
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp_, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp_, f)


def test_lsq_glm_maps_match_the_reference_formula(tmp_path):
    """vaegam.glm_maps (streamed G'Y / G'G accumulation) against the reference's dense numpy evaluation
    (get_beta_map_regularizer.py:94-107 + utils.scale_beta_maps), and the CSV it writes is what the VAE reads."""
    from vaegam import glm_maps, synthetic as syn
    rng = np.random.default_rng(0)
    n, v = 60, 501
    y = rng.random((n, v)).astype(np.float32)
    gamma = rng.standard_normal((n, 7))
    sex_map = rng.standard_normal(v)
    pinv = np.linalg.inv(gamma.T @ gamma) @ gamma.T                       # the reference's pseudo_inv
    want = np.concatenate([pinv @ y.astype(np.float64), sex_map[None]], 0)
    want = want / want.max(axis=1, keepdims=True)
    blocks = [(torch.from_numpy(y[i:i + 17]), torch.from_numpy(gamma[i:i + 17])) for i in range(0, n, 17)]
    got = glm_maps.lsq_beta_maps(blocks, torch.from_numpy(sex_map))
    assert got.shape == (v, 8) and np.allclose(got, want.T, rtol=1e-9, atol=1e-12)
    path = glm_maps.write_glm_csv(str(tmp_path / "scld_GLM_beta_maps.csv"), got)
    back = pd.read_csv(path).to_numpy()
    assert back.shape == (v, 9) and np.allclose(back[:, 1:], got)          # index column + 8 maps
    with pytest.raises(ValueError):
        glm_maps.lsq_beta_maps([(torch.zeros(3, 5), torch.zeros(3, 6))])
    # the synthetic cohort's maps go through the same code (BASELINE config 3)
    coh = syn.make_cohort(2, "v1", seed=1, n_vols=40)
    maps = syn.glm_maps_lsq(coh, block=17)
    assert maps.shape == (41 * 49 * 35, 8) and np.isfinite(maps).all() and np.allclose(maps.max(0)[:7], 1.0)


# ----------------------------------------------------------------------------- round 2 additions
def test_load_state_with_other_inducing_points_repacks(experiment):
    """a checkpoint trained with another num_inducing_pts replaces Parameter objects; the flat
    buffers, the optimizer and the engine must be rebuilt around the NEW objects (reference load_state:
    vae_reg_GP.py:473-539)."""
    d = experiment[0]
    small = make_model(experiment, seed=5, num_inducing_pts=4)
    small.save_state("ck_m4.tar")
    big = make_model(experiment, seed=6, num_inducing_pts=6)
    big.load_state(os.path.join(d, "ck_m4.tar"))
    assert big.inducing_pts == 4 and tuple(big.qu_m_x.shape) == (1, 4) and tuple(big.qu_S_x.shape) == (4, 4)
    named = dict(big.named_parameters())
    assert big._flat.is_packed()
    for n, p in zip(big._flat.names, big._flat.params):
        assert p is named[n], n                                   # the flat table lists the live objects
    owned = {id(p) for g in big.optimizer.param_groups for p in g["params"]}
    assert all(id(p) in owned for p in named.values())            # and the optimizer owns all of them
    assert big._flat.n32 == small._flat.n32 and big.optimizer.m32.numel() == big._flat.n32
    for (n, a), (_, b) in zip(small.named_parameters(), big.named_parameters()):
        assert torch.equal(a.detach(), b.detach()), n
    assert big.gp_params['x']['qu_m'] is big.qu_m_x


def test_numpy_aliases_and_nibabel_are_not_shadowed():
    """`np.float` exists after importing the package (reference build_model_recons.py:74,85) and `nibabel`
    resolves to a real installation when there is one, else to the bundled codec — never to a directory shim."""
    import importlib.util
    import sys
    import vaegam  # noqa: F401
    assert np.float is float and np.int is int
    from vaegam.nib_compat import nib
    spec_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vae-gam_b200", "nibabel")
    assert not os.path.exists(spec_dir)
    assert hasattr(nib, "load") and hasattr(nib, "Nifti1Image") and hasattr(nib, "save")
    assert sys.modules["nibabel"] is nib


def test_nifti_qform_only_header(tmp_path):
    """Files that carry only a qform (sform_code == 0) get their affine from the quaternion (NIfTI-1 method 2)."""
    import struct
    from vaegam import nifti
    a = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    p = str(tmp_path / "q.nii")
    nifti.save(nifti.Nifti1Image(a, np.eye(4)), p)
    raw = bytearray(open(p, "rb").read())
    # 90 degree rotation about z: quaternion (a, b, c, d) = (cos45, 0, 0, sin45); pixdim (2, 3, 4); qfac -1
    struct.pack_into("<8f", raw, 76, -1.0, 2.0, 3.0, 4.0, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<hh", raw, 252, 1, 0)                       # qform_code = 1, sform_code = 0
    struct.pack_into("<6f", raw, 256, 0.0, 0.0, float(np.sqrt(0.5)), 10.0, 20.0, 30.0)
    open(p, "wb").write(bytes(raw))
    img = nifti.load(p)
    want = np.array([[0, -3, 0, 10], [2, 0, 0, 20], [0, 0, -4, 30], [0, 0, 0, 1]], dtype=np.float64)
    assert np.allclose(img.affine, want, atol=1e-6)
    assert np.array_equal(np.asarray(img.dataobj), a)


def test_launcher_resolves_the_drop_in_modules(tmp_path, capsys):
    """run_reference.py executes an unmodified reference script with the drop-in directory FIRST on sys.path
    (python <script> would put the script's own directory first and import the reference's PyTorch modules)."""
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vae-gam_b200")
    src = "/root/reference" if HAVE_REF else pkg
    work = tmp_path / "VAE-GAM"
    work.mkdir()
    for f in ("multsubj_reg_run_GP.py", "build_model_recons.py", "vae_reg_GP.py", "utils.py"):
        shutil.copy(os.path.join(src, f), str(work / f))          # decoys next to the script, as in a checkout
    probe = ("import sys, runpy\n"
             "sys.argv = ['x', %r, '--help']\n"
             "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit as e:\n    assert e.code in (0, None), e.code\n"
             "import vae_reg_GP, DataClass_GP, build_model_recons, utils\n"
             "print('RESOLVED', vae_reg_GP.__file__, DataClass_GP.__file__, build_model_recons.__file__, utils.__file__)\n"
             % (str(work / "multsubj_reg_run_GP.py"), os.path.join(pkg, "run_reference.py")))
    env = dict(os.environ, PYTHONPATH="")
    out = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESOLVED")][0]
    for path in line.split()[1:]:
        assert os.path.dirname(os.path.abspath(path)) == pkg, line
    assert "--num_inducing_pts" in out.stdout                      # the reference's own argparse help was printed


def test_resident_loader_matches_dataloader_batches(experiment):
    """GPU-resident loader (here on the CPU device): same contract and, for a given torch seed, the same
    shuffled batches as torch's DataLoader — so switching it on does not change a run."""
    import DataClass_GP as data
    d, tr, te, glm, coh = experiment
    torch.manual_seed(21)
    std = data.setup_data_loaders(batch_size=40, train_csv=tr, test_csv=te, resident=False)
    a = [b for b in std['Shuffled_train']]
    torch.manual_seed(21)
    res = data.setup_data_loaders(batch_size=40, train_csv=tr, test_csv=te, resident=True, device="cpu")
    assert isinstance(res['Shuffled_train'], data.ResidentLoader)
    b = [x for x in res['Shuffled_train']]
    assert len(a) == len(b) == len(res['Shuffled_train']) == 5 and len(res['test'].dataset) == 98
    assert res['UnShuffled_train'].batch_size == 40
    for x, y in zip(a, b):
        for k in ('volume', 'covariates', 'subjid', 'vol_num'):
            assert x[k].dtype == y[k].dtype and x[k].shape == y[k].shape and torch.equal(x[k], y[k]), k
    assert b[-1]['volume'].shape[0] == 2 * 98 - 4 * 40            # ragged last batch
    # rank shards: disjoint, equal, and their union is one epoch's permutation
    r0 = data.setup_data_loaders(batch_size=49, train_csv=tr, test_csv=te, resident=True, device="cpu", rank=0, world=2)
    r1 = data.setup_data_loaders(batch_size=49, train_csv=tr, test_csv=te, resident=True, device="cpu", rank=1, world=2)
    torch.manual_seed(3)                                          # every rank seeds the same way (as torchrun ranks do)
    v0 = torch.cat([x['vol_num'] + 1000 * x['subjid'] for x in r0['Shuffled_train']])
    torch.manual_seed(3)
    v1 = torch.cat([x['vol_num'] + 1000 * x['subjid'] for x in r1['Shuffled_train']])
    assert v0.numel() == v1.numel() == 98 and len(set(v0.tolist()) & set(v1.tolist())) == 0


def test_gradient_buckets_partition_the_flat_buffers(experiment):
    """The three all-reduce buckets (vaegam.dp) are contiguous views that tile the flat gradient buffers exactly,
    in backward-completion order: decoder tail, encoder FC middle, gains + encoder convolutions head."""
    from vaegam import dp
    m = make_model(experiment)
    red = dp.GradientAllReduce(m._flat, m.optimizer, overlap=False)
    b = red.buckets()
    f = m._flat
    assert len(b) == 3 and b[0][0] is f.grad64
    tail, mid, head = b[0][1], b[1][0], b[2][0]
    assert head.data_ptr() == f.grad32.data_ptr()
    assert mid.data_ptr() == head.data_ptr() + 4 * head.numel()
    assert tail.data_ptr() == mid.data_ptr() + 4 * mid.numel()
    assert head.numel() + mid.numel() + tail.numel() == f.n32
    where = lambda n: f.slices[n][1]
    assert where("fc1.weight") == head.numel() and where("fc5.weight") == head.numel() + mid.numel()
    for n in ("conv1.weight", "bn5.bias", "sa_task", "qu_S_zrot"):
        assert where(n) < head.numel()
    for n in ("fc8.weight", "convt5.bias", "bnt1.weight"):
        assert where(n) >= head.numel() + mid.numel()
