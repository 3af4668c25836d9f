"""Drop-in replacement for the reference's `gp.py`: 1-D sparse GP (grid inducing points, RBF
kernel) for the covariate gain functions.

Same class and function names as the reference (gp.py:13-136).  In the training step the
whole gain stage (this posterior for the six motion covariates, the Cholesky sample, the HRF
filter and both KL terms) runs fused in `vg_gain_fwd/bwd`; this module is the stand-alone
surface: `GP.evaluate_posterior` is used with all rows of the dataset by `VAE.plot_GPs`
(reference vae_reg_GP.py:655-660) and calls `vg_gp_posterior` (fp64 internally).
"""
import ctypes as C

import numpy as np
import torch

from vaegam import native


class GP():
    """1D Gaussian process with inducing points on an even grid and a Gaussian kernel."""

    def __init__(self, Xu, k_var, ls, qu_m, qu_S):
        assert len(Xu) > 1
        self.device = Xu.device
        self.n = Xu.shape[0]
        self.step = Xu[1] - Xu[0]
        self.Xu, self.k_var, self.ls, self.qu_m, self.qu_S = Xu, k_var, ls, qu_m, qu_S

    def _f32(self, t):
        return torch.as_tensor(t).detach().to(self.device, torch.float32).contiguous()

    def _posterior(self, X_q, want_sigma):
        native.require_cuda()
        lib = native.load()
        xq = self._f32(X_q).reshape(-1)
        nq = xq.shape[0]
        f_bar = torch.empty(nq, dtype=torch.float32, device=self.device)
        var = torch.empty(nq, dtype=torch.float32, device=self.device)
        sigma = torch.empty(nq, nq, dtype=torch.float32, device=self.device) if want_sigma else None
        a_ws = torch.empty(nq, self.n, dtype=torch.float64, device=self.device) if want_sigma else None
        args = [self._f32(self.Xu), self._f32(self.k_var).reshape(1), self._f32(self.ls).reshape(1),
                self._f32(self.qu_m).reshape(-1), self._f32(self.qu_S)]
        native.check(lib.vg_gp_posterior(native.ptr(args[0]), self.n, native.ptr(args[1]), native.ptr(args[2]),
                                         native.ptr(args[3]), native.ptr(args[4]), native.ptr(xq), nq,
                                         native.ptr(f_bar), native.ptr(var), native.ptr(sigma), native.ptr(a_ws),
                                         native.stream_ptr()), "vg_gp_posterior")
        torch.cuda.current_stream().synchronize()   # temporaries above die with this frame
        return f_bar, var, sigma

    def evaluate_posterior(self, X_q):
        """q(f) at the query points: (f_bar (n_q,), Sigma (n_q, n_q)) — reference gp.py:67-110."""
        f_bar, _, sigma = self._posterior(X_q, True)
        return f_bar, sigma

    def evaluate_posterior_diag(self, X_q):
        """(f_bar, diag(Sigma)) without forming the n_q x n_q matrix."""
        f_bar, var, _ = self._posterior(X_q, False)
        return f_bar, var

    def compute_GP_kl(self, num_inducing_pts, i=None, xq=None, save_dir=None):
        """KL(N(qu_m, qu_S) || N(0, 10 I)) (reference gp.py:41-65); differentiable torch closed form."""
        m = num_inducing_pts
        S = self.qu_S.double()
        L = torch.linalg.cholesky(S)
        qm = self.qu_m.double().reshape(-1)
        kl = 0.5 * (torch.diagonal(S).sum() / 10 + (qm * qm).sum() / 10 - m + m * np.log(10.0)
                    - 2 * torch.log(torch.diagonal(L)).sum())
        return kl.to(self.qu_m.dtype).reshape(1)


def _striped_matrix(n):
    """n-by-n matrix of |i - j| (reference gp.py:113-119), on the current default device."""
    idx = torch.arange(n, dtype=torch.float32, device="cuda" if torch.cuda.is_available() else "cpu")
    return (idx[:, None] - idx[None, :]).abs()


def _distance_to_kernel(dist_mat, k_var, ls, scale_factor=1.0):
    """Gaussian kernel of a (signed) distance matrix (reference gp.py:121-136)."""
    return k_var * torch.exp(-torch.pow(scale_factor / np.sqrt(2) / ls * dist_mat, 2))
