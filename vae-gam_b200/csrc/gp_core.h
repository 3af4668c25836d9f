// Gain stage for ONE covariate, written against a small "team"
// abstraction so that the very same source runs
//   * on the GPU, one thread block (one warp for B <= 32) per covariate  — gp_gain.cu
//   * on the host with a 1-thread team, for the CPU unit tests          — tests/cpu_emul
// Arithmetic is fp64 throughout (the reference's fp32 `torch.inverse(Ku)` is
// ill-conditioned; its own result is only accurate to ~1e-2).
//
// Reference call sites restated here:
//   vae_reg_GP.py:345-351   linear gain mean / diagonal covariance
//   vae_reg_GP.py:266-281   KL(N(sa, s^2) || N(1, 0.5^2))
//   vae_reg_GP.py:354-359   k_var = exp(logkvar)+0.1, ls = 3*sigmoid(exp(log_ls)+0.5)
//   gp.py:67-110,113-136    Knu, Knn, Ku, A = Knu^T inv(Ku), f_bar, Sigma
//   gp.py:41-65             KL(N(qu_m, qu_S) || N(0, 10 I))   (qu_S through its lower triangle)
//   vae_reg_GP.py:368-369   L = chol(cov + 1e-5 I), g = mean + L eps   (multivariate_normal.py:251-254)
//   vae_reg_GP.py:283-305   causal 15-tap HRF FIR over the batch index
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define VG_HD __host__ __device__
#define VG_TEAM_FN __device__      // team templates are device-only under nvcc
#else
#define VG_HD
#define VG_TEAM_FN
#endif

namespace vg {

constexpr int kHrfTaps = 15;
constexpr int kMaxInducing = 16;

struct GainOne {
  const float* cov;      // (B, cov_stride); this covariate is column cov_col
  int cov_stride, cov_col;
  const float* eps;      // (B)
  const float* sa;
  const float* logstd;
  const float* qu_m;     // (m)
  const float* qu_S;     // (m,m)
  const float* logkvar;
  const float* logls;
  const float* xu;       // (m)
  const double* taps;    // (15)
  int has_gp, hrf, B, m;
  double* ws;            // workspace, gain_ws_doubles(B, m) doubles
};

struct GainWs {
  double *xq, *mean, *gpre, *C, *A, *Knu, *Kinv, *Ku, *M, *Sinv, *AM, *X, *W, *Abar, *Knubar, *tvec, *gbar,
      *small;  // small: 4 * m*m scratch + 16
};

VG_HD inline size_t gain_ws_doubles(int B, int m) {
  return (size_t)5 * B + (size_t)2 * B * B + (size_t)5 * B * m + (size_t)8 * m * m + 32;
}

VG_HD inline GainWs gain_ws_carve(double* p, int B, int m) {
  GainWs w;
  w.xq = p; p += B;
  w.mean = p; p += B;
  w.gpre = p; p += B;
  w.tvec = p; p += B;
  w.gbar = p; p += B;
  w.C = p; p += (size_t)B * B;
  w.X = p; p += (size_t)B * B;
  w.A = p; p += (size_t)B * m;
  w.Knu = p; p += (size_t)B * m;
  w.AM = p; p += (size_t)B * m;
  w.W = p; p += (size_t)B * m;
  w.Abar = p; p += (size_t)B * m;   // Knubar aliases AM in the backward pass
  w.Knubar = w.AM;
  w.Kinv = p; p += m * m;
  w.Ku = p; p += m * m;
  w.M = p; p += m * m;
  w.Sinv = p; p += m * m;
  w.small = p;
  return w;
}

VG_HD inline double rbf_k(double d, double kvar, double ls) { return kvar * exp(-(d * d) / (2.0 * ls * ls)); }

// Gauss-Jordan inverse with partial pivoting of an n x n matrix (n <= kMaxInducing); serial.
VG_HD inline void small_inverse(const double* a, double* inv, int n, double* scratch /* n*2n */) {
  const int w = 2 * n;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      scratch[i * w + j] = a[i * n + j];
      scratch[i * w + n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = fabs(scratch[c * w + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(scratch[r * w + c]) > best) { best = fabs(scratch[r * w + c]); p = r; }
    if (p != c)
      for (int j = 0; j < w; ++j) { double t = scratch[c * w + j]; scratch[c * w + j] = scratch[p * w + j]; scratch[p * w + j] = t; }
    const double d = 1.0 / scratch[c * w + c];
    for (int j = 0; j < w; ++j) scratch[c * w + j] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = scratch[r * w + c];
      if (f != 0.0)
        for (int j = 0; j < w; ++j) scratch[r * w + j] -= f * scratch[c * w + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = scratch[i * w + n + j];
}

// Serial Cholesky of the lower triangle of an n x n matrix; returns 0 or 1-based failed pivot.
VG_HD inline int small_cholesky(const double* a, double* l, int n) {
  int bad = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = a[i * n + j];
      for (int k = 0; k < j; ++k) s -= l[i * n + k] * l[j * n + k];
      if (i == j) {
        if (!(s > 0.0)) { if (!bad) bad = i + 1; s = 1e-300; }
        l[i * n + i] = sqrt(s);
      } else {
        l[i * n + j] = s / l[j * n + j];
      }
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) l[i * n + j] = 0.0;
  return bad;
}

struct GainHyper {
  double sa, s, kvar, ls, step, xu0, elk, ell, sig;  // elk = exp(logkvar), ell = exp(logls), sig = sigmoid(ell+0.5)
};

VG_HD inline GainHyper gain_hyper(const GainOne& in) {
  GainHyper h;
  h.sa = (double)in.sa[0];
  h.s = exp((double)in.logstd[0]);
  h.kvar = h.ls = h.step = h.xu0 = h.elk = h.ell = h.sig = 0.0;
  if (in.has_gp) {
    h.elk = exp((double)in.logkvar[0]);
    h.kvar = h.elk + 0.1;
    h.ell = exp((double)in.logls[0]);
    h.sig = 1.0 / (1.0 + exp(-(h.ell + 0.5)));
    h.ls = 3.0 * h.sig;
    h.xu0 = (double)in.xu[0];
    h.step = (double)in.xu[1] - (double)in.xu[0];
  }
  return h;
}

// In-place Cholesky of the lower triangle of C (B x B, row-major), rows shared by the team.
template <class Team>
VG_TEAM_FN inline int team_cholesky(Team& tm, double* C, int B) {
  int bad = 0;
  for (int j = 0; j < B; ++j) {
    tm.sync();
    double piv = C[(size_t)j * B + j];
    if (!(piv > 0.0)) { if (!bad) bad = j + 1; piv = 1e-300; }
    const double d = sqrt(piv);
    tm.sync();
    if (tm.rank() == 0) C[(size_t)j * B + j] = d;
    for (int a = j + 1 + tm.rank(); a < B; a += tm.size()) C[(size_t)a * B + j] /= d;
    tm.sync();
    for (int a = j + 1 + tm.rank(); a < B; a += tm.size()) {
      const double l = C[(size_t)a * B + j];
      for (int k = j + 1; k <= a; ++k) C[(size_t)a * B + k] -= l * C[(size_t)k * B + j];
    }
  }
  tm.sync();
  return bad;
}

// ------------------------------------------------------------------------------ forward
// Outputs (any may be null): g (B) fp32 post-HRF gains, kl (2) = {linear-weight KL, GP KL},
// beta_mean (B), beta_var (B) = diag(beta_cov) (before the 1e-5 jitter), status.
template <class Team>
VG_TEAM_FN inline void gain_forward(Team& tm, const GainOne& in, float* g, double* kl, float* beta_mean,
                               float* beta_var, int* status) {
  const int B = in.B, m = in.m;
  GainWs w = gain_ws_carve(in.ws, B, m);
  const GainHyper h = gain_hyper(in);
  int bad = 0;
  for (int a = tm.rank(); a < B; a += tm.size()) {
    const double x = (double)in.cov[(size_t)a * in.cov_stride + in.cov_col];
    w.xq[a] = x;
    w.mean[a] = h.sa * x;
  }
  double kl_gp = 0.0;
  if (in.has_gp) {
    if (tm.rank() == 0) {
      for (int p = 0; p < m; ++p)
        for (int q = 0; q < m; ++q) {
          const double d = fabs((double)(p - q)) * h.step;
          w.Ku[p * m + q] = rbf_k(d, h.kvar, h.ls);
          w.M[p * m + q] = (double)in.qu_S[p * m + q] - w.Ku[p * m + q];
        }
      small_inverse(w.Ku, w.Kinv, m, w.small);
      // GP KL: S enters through chol(lower triangle)
      double* Ls = w.small;                 // m*m
      double* Ssym = w.small + m * m;       // m*m, symmetric copy of the lower triangle
      for (int p = 0; p < m; ++p)
        for (int q = 0; q < m; ++q) Ssym[p * m + q] = (double)in.qu_S[(p >= q ? p : q) * m + (p >= q ? q : p)];
      const int bs = small_cholesky(Ssym, Ls, m);
      if (bs) bad = 1000 + bs;
      double tr = 0, qq = 0, ld = 0;
      for (int p = 0; p < m; ++p) {
        tr += Ssym[p * m + p];
        qq += (double)in.qu_m[p] * (double)in.qu_m[p];
        ld += log(Ls[p * m + p]);
      }
      kl_gp = 0.5 * (tr / 10.0 + qq / 10.0 - m + m * log(10.0) - 2.0 * ld);
      small_inverse(Ssym, w.Sinv, m, w.small + 2 * m * m);
    }
    tm.sync();
    for (int a = tm.rank(); a < B; a += tm.size()) {
      const double base = h.xu0 - w.xq[a];
      for (int p = 0; p < m; ++p) w.Knu[(size_t)p * B + a] = rbf_k(base + p * h.step, h.kvar, h.ls);
      double fb = 0.0;
      for (int p = 0; p < m; ++p) {
        double s = 0.0;
        for (int q = 0; q < m; ++q) s += w.Knu[(size_t)q * B + a] * w.Kinv[q * m + p];
        w.A[(size_t)a * m + p] = s;
        fb += s * (double)in.qu_m[p];
      }
      w.mean[a] += fb;
      for (int q = 0; q < m; ++q) {
        double s = 0.0;
        for (int p = 0; p < m; ++p) s += w.A[(size_t)a * m + p] * w.M[p * m + q];
        w.AM[(size_t)a * m + q] = s;
      }
    }
    tm.sync();
  }
  // covariance, lower triangle (row a owned by one thread)
  for (int a = tm.rank(); a < B; a += tm.size()) {
    for (int b = 0; b <= a; ++b) {
      double c = 0.0;
      if (in.has_gp) {
        c = rbf_k(w.xq[b] - w.xq[a], h.kvar, h.ls);
        for (int q = 0; q < m; ++q) c += w.AM[(size_t)a * m + q] * w.A[(size_t)b * m + q];
      }
      if (a == b) {
        c += h.s * h.s * w.xq[a] * w.xq[a];
        if (beta_var) beta_var[a] = (float)c;
        c += 1e-5;
      }
      w.C[(size_t)a * B + b] = c;
    }
    if (beta_mean) beta_mean[a] = (float)w.mean[a];
  }
  const int bc = team_cholesky(tm, w.C, B);
  if (bc && !bad) bad = bc;
  for (int a = tm.rank(); a < B; a += tm.size()) {
    double s = w.mean[a];
    for (int k = 0; k <= a; ++k) s += w.C[(size_t)a * B + k] * (double)in.eps[k];
    w.gpre[a] = s;
  }
  tm.sync();
  for (int t = tm.rank(); t < B; t += tm.size()) {
    double s;
    if (in.hrf) {
      s = 0.0;
      const int kmax = t < kHrfTaps - 1 ? t : kHrfTaps - 1;
      for (int k = 0; k <= kmax; ++k) s += in.taps[k] * w.gpre[t - k];
    } else {
      s = w.gpre[t];
    }
    if (g) g[t] = (float)s;
  }
  if (tm.rank() == 0) {
    if (kl) {
      kl[0] = log(0.5 / h.s) + (h.s * h.s + (h.sa - 1.0) * (h.sa - 1.0)) / 0.5 - 0.5;
      kl[1] = kl_gp;
    }
    if (status) *status = bad;
  }
  tm.sync();
}

// ------------------------------------------------------------------------------ backward
struct GainGradOut {   // fp32 accumulators (+=); null for absent parameters
  float *sa, *logstd, *qu_m, *qu_S, *logkvar, *logls;
};

// Needs the workspace left by gain_forward for the same covariate (xq, mean, L in C, A, Knu,
// Kinv, Ku, M, Sinv).  dg: dLoss/dg (post-HRF).  kl_scale weights both KL terms.
template <class Team>
VG_TEAM_FN inline void gain_backward(Team& tm, const GainOne& in, const float* dg, double kl_scale,
                                const GainGradOut& out) {
  const int B = in.B, m = in.m;
  GainWs w = gain_ws_carve(in.ws, B, m);
  const GainHyper h = gain_hyper(in);
  const double* L = w.C;
  // gbar = d/d g_pre (transpose of the causal FIR)
  for (int s = tm.rank(); s < B; s += tm.size()) {
    double v;
    if (in.hrf) {
      v = 0.0;
      for (int k = 0; k < kHrfTaps && s + k < B; ++k) v += in.taps[k] * (double)dg[s + k];
    } else {
      v = (double)dg[s];
    }
    w.gbar[s] = v;
  }
  tm.sync();
  // t = L^T gbar
  for (int a = tm.rank(); a < B; a += tm.size()) {
    double s = 0.0;
    for (int i = a; i < B; ++i) s += L[(size_t)i * B + a] * w.gbar[i];
    w.tvec[a] = s;
  }
  tm.sync();
  // X = Phi(t eps^T) L^{-1}  (lower triangular), one row per thread
  double* X = w.X;
  for (int a = tm.rank(); a < B; a += tm.size()) {
    for (int k = a; k >= 0; --k) {
      double p = w.tvec[a] * (double)in.eps[k];
      if (k == a) p *= 0.5;
      for (int j = k + 1; j <= a; ++j) p -= X[(size_t)a * B + j] * L[(size_t)j * B + k];
      X[(size_t)a * B + k] = p / L[(size_t)k * B + k];
    }
    for (int k = a + 1; k < B; ++k) X[(size_t)a * B + k] = 0.0;
  }
  tm.sync();
  // Y = L^{-T} X, one column per thread, in place in X
  for (int c = tm.rank(); c < B; c += tm.size()) {
    for (int r = B - 1; r >= 0; --r) {
      double v = X[(size_t)r * B + c];
      for (int j = r + 1; j < B; ++j) v -= L[(size_t)j * B + r] * X[(size_t)j * B + c];
      X[(size_t)r * B + c] = v / L[(size_t)r * B + r];
    }
  }
  tm.sync();
  // Cbar(a,b) = (Y(a,b) + Y(b,a)) / 2 is used on the fly below.
#define VG_CBAR(a_, b_) (0.5 * (X[(size_t)(a_) * B + (b_)] + X[(size_t)(b_) * B + (a_)]))

  // linear part
  double p_sa = 0.0, p_ls = 0.0;
  for (int a = tm.rank(); a < B; a += tm.size()) {
    p_sa += w.gbar[a] * w.xq[a];
    p_ls += VG_CBAR(a, a) * 2.0 * h.s * h.s * w.xq[a] * w.xq[a];
  }
  const double d_sa = tm.sum(p_sa) + kl_scale * 4.0 * (h.sa - 1.0);
  const double d_logstd = tm.sum(p_ls) + kl_scale * (4.0 * h.s * h.s - 1.0);
  if (tm.rank() == 0) {
    if (out.sa) out.sa[0] += (float)d_sa;
    if (out.logstd) out.logstd[0] += (float)d_logstd;
  }
  if (!in.has_gp) { tm.sync(); return; }

  // W = Cbar A  (B x m);  Abar = gbar qm^T + W (M + M^T)
  for (int a = tm.rank(); a < B; a += tm.size()) {
    for (int p = 0; p < m; ++p) {
      double s = 0.0;
      for (int b = 0; b < B; ++b) s += VG_CBAR(a, b) * w.A[(size_t)b * m + p];
      w.W[(size_t)a * m + p] = s;
    }
    for (int p = 0; p < m; ++p) {
      double s = w.gbar[a] * (double)in.qu_m[p];
      for (int q = 0; q < m; ++q) s += w.W[(size_t)a * m + q] * (w.M[p * m + q] + w.M[q * m + p]);
      w.Abar[(size_t)a * m + p] = s;
    }
  }
  tm.sync();
  // m x m reductions over the batch: AtW = A^T W, Kinvbar = Knu Abar, qmbar = A^T gbar
  double* AtW = w.small;                 // m*m
  double* Kinvbar = w.small + m * m;     // m*m
  double* Kubar = w.small + 2 * m * m;   // m*m
  double* T1 = w.small + 3 * m * m;      // m*m
  for (int p = 0; p < m; ++p) {
    double pq = 0.0;
    for (int a = tm.rank(); a < B; a += tm.size()) pq += w.A[(size_t)a * m + p] * w.gbar[a];
    const double dq = tm.sum(pq);
    if (tm.rank() == 0 && out.qu_m) out.qu_m[p] += (float)(dq + kl_scale * (double)in.qu_m[p] / 10.0);
    for (int q = 0; q < m; ++q) {
      double s1 = 0.0, s2 = 0.0;
      for (int a = tm.rank(); a < B; a += tm.size()) {
        s1 += w.A[(size_t)a * m + p] * w.W[(size_t)a * m + q];
        s2 += w.Knu[(size_t)p * B + a] * w.Abar[(size_t)a * m + q];
      }
      const double r1 = tm.sum(s1), r2 = tm.sum(s2);
      if (tm.rank() == 0) { AtW[p * m + q] = r1; Kinvbar[p * m + q] = r2; }
    }
  }
  tm.sync();
  if (tm.rank() == 0) {
    // dS = A^T Cbar A + kl_scale * 0.5 * (I/10 - S^{-1})
    if (out.qu_S)
      for (int p = 0; p < m; ++p)
        for (int q = 0; q < m; ++q)
          out.qu_S[p * m + q] += (float)(AtW[p * m + q] + kl_scale * 0.5 * ((p == q ? 0.1 : 0.0) - w.Sinv[p * m + q]));
    // Kubar = -AtW - Kinv^T Kinvbar Kinv^T
    for (int p = 0; p < m; ++p)
      for (int q = 0; q < m; ++q) {
        double s = 0.0;
        for (int r = 0; r < m; ++r) s += w.Kinv[r * m + p] * Kinvbar[r * m + q];
        T1[p * m + q] = s;
      }
    for (int p = 0; p < m; ++p)
      for (int q = 0; q < m; ++q) {
        double s = 0.0;
        for (int r = 0; r < m; ++r) s += T1[p * m + r] * w.Kinv[q * m + r];
        Kubar[p * m + q] = -AtW[p * m + q] - s;
      }
  }
  tm.sync();
  // Knubar(q,a) = sum_p Kinv(q,p) Abar(a,p); kernel hyper-parameter sums
  double p_kv = 0.0, p_l = 0.0;
  const double ls3 = h.ls * h.ls * h.ls;
  for (int a = tm.rank(); a < B; a += tm.size()) {
    const double base = h.xu0 - w.xq[a];
    for (int q = 0; q < m; ++q) {
      double kb = 0.0;
      for (int p = 0; p < m; ++p) kb += w.Kinv[q * m + p] * w.Abar[(size_t)a * m + p];
      const double d = base + q * h.step;
      const double kv = w.Knu[(size_t)q * B + a];
      p_kv += kb * kv / h.kvar;
      p_l += kb * kv * d * d / ls3;
    }
    for (int b = 0; b < B; ++b) {
      const double d = w.xq[b] - w.xq[a];
      const double kv = rbf_k(d, h.kvar, h.ls);
      const double cb = VG_CBAR(a, b);
      p_kv += cb * kv / h.kvar;
      p_l += cb * kv * d * d / ls3;
    }
  }
  if (tm.rank() == 0) {
    for (int p = 0; p < m; ++p)
      for (int q = 0; q < m; ++q) {
        const double d = fabs((double)(p - q)) * h.step;
        const double kv = w.Ku[p * m + q];
        p_kv += Kubar[p * m + q] * kv / h.kvar;
        p_l += Kubar[p * m + q] * kv * d * d / ls3;
      }
  }
  const double d_kvar = tm.sum(p_kv);
  const double d_ls = tm.sum(p_l);
  if (tm.rank() == 0) {
    if (out.logkvar) out.logkvar[0] += (float)(d_kvar * h.elk);
    if (out.logls) out.logls[0] += (float)(d_ls * 3.0 * h.sig * (1.0 - h.sig) * h.ell);
  }
  tm.sync();
#undef VG_CBAR
}

}  // namespace vg
