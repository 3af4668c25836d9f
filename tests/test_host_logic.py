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
