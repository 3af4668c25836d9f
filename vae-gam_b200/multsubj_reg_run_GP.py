"""Trainer / reconstruction CLI with the reference's flags (reference multsubj_reg_run_GP.py:21-54,
67-94).  The reference's own script also runs unchanged with this directory on PYTHONPATH; this
copy exists because the reference tree is not shipped with the package."""
import argparse
import os
import time

import torch

import DataClass_GP as data
import build_model_recons as recon
import vae_reg_GP as vae_reg
from utils import str2bool


def build_parser():
    p = argparse.ArgumentParser(description='user args for vae_gam model')
    p.add_argument('--train_csv', type=str, default='')
    p.add_argument('--test_csv', type=str, default='')
    p.add_argument('--save_dir', type=str, default='')
    p.add_argument('--batch-size', type=int, default=32)
    p.add_argument('--epochs', type=int, default=300)
    p.add_argument('--seed', type=int, default=1)
    p.add_argument('--save_freq', type=int, default=100)
    p.add_argument('--test_freq', type=int, default=200)
    p.add_argument('--split', type=int, default=98)
    p.add_argument('--glm_reg_scale', type=float, default=1.0)
    p.add_argument('--glm_maps', type=str, default='')
    p.add_argument('--num_inducing_pts', type=int, default=6)
    p.add_argument('--gp_kl_scale', type=float, default=10.0)
    p.add_argument('--from_ckpt', type=str2bool, nargs='?', const=True, default=False)
    p.add_argument('--ckpt_path', type=str, default='')
    p.add_argument('--recons_only', type=str2bool, nargs='?', const=True, default=False)
    p.add_argument('--neural_covariates', type=str2bool, nargs='?', const=True, default=True)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    torch.manual_seed(args.seed)
    if args.save_dir == '':
        args.save_dir = os.getcwd()
    os.makedirs(args.save_dir, exist_ok=True)
    t0 = time.time()
    loaders = data.setup_data_loaders(batch_size=args.batch_size, train_csv=args.train_csv, test_csv=args.test_csv)
    model = vae_reg.VAE(num_inducing_pts=args.num_inducing_pts, gp_kl_scale=args.gp_kl_scale,
                        glm_reg_scale=args.glm_reg_scale, glm_maps=args.glm_maps, save_dir=args.save_dir,
                        csv_files=[args.train_csv, args.test_csv], neural_covariates=args.neural_covariates)
    if args.from_ckpt:
        assert os.path.exists(args.ckpt_path), 'checkpoint file does not exist'
        print('=' * 40)
        print('Loading model state from: {}'.format(args.ckpt_path))
        model.load_state(filename=args.ckpt_path)
    if not args.recons_only:
        model.train_loop(loaders, epochs=args.epochs, test_freq=args.test_freq, save_freq=args.save_freq,
                         save_dir=args.save_dir)
    else:
        assert args.from_ckpt, 'To choose recons_only option, --from_ckpt needs to be TRUE.'
    model.project_latent(loaders, title="Latent Space plot", split=args.split, save_dir=args.save_dir)
    model.plot_GPs(csv_file=args.train_csv, save_dir=args.save_dir)
    recon.mk_single_volumes(loaders['UnShuffled_train'], model, args.train_csv, args.save_dir)
    recon.mk_avg_maps(args.train_csv, model, args.save_dir, mk_motion_maps=True)
    print('Total model runtime (seconds): {}'.format(time.time() - t0))


if __name__ == "__main__":
    main()
