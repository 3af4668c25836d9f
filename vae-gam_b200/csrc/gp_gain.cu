// Gain stage on the GPU: one thread block per covariate GP (a single warp for B <= 32, as
// the north star asks; four warps for larger minibatches), all K covariates in one launch.
// The algebra lives in gp_core.h (shared with the CPU emulation used by the unit tests).
#include "common.cuh"
#include "gp_core.h"

namespace vg {

struct BlockTeam {
  double* red;
  __device__ int rank() const { return threadIdx.x; }
  __device__ int size() const { return blockDim.x; }
  __device__ void sync() const { __syncthreads(); }
  __device__ double sum(double v) const {
    v = warp_sum(v);
    if (blockDim.x == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
  }
};

__device__ __forceinline__ GainOne make_one(const VgGainParams& p, const float* cov, const float* eps,
                                            const double* taps, int i, int b, int m, double* ws) {
  GainOne in;
  in.cov = cov; in.cov_stride = 8; in.cov_col = i;
  in.eps = eps + (size_t)i * b;
  in.sa = p.sa[i]; in.logstd = p.logstd[i];
  in.qu_m = p.qu_m[i]; in.qu_S = p.qu_S[i]; in.logkvar = p.logkvar[i]; in.logls = p.logls[i]; in.xu = p.xu[i];
  in.taps = taps;
  in.has_gp = p.has_gp[i]; in.hrf = p.hrf[i]; in.B = b; in.m = m;
  in.ws = ws + (size_t)i * gain_ws_doubles(b, m);
  return in;
}

__global__ void gain_fwd_kernel(const __grid_constant__ VgGainParams p, const float* cov, const float* eps,
                                const double* taps, int b, int m, float* g, double* kl_terms, float* beta_mean,
                                float* beta_var, int* status, double* ws) {
  __shared__ double red[32];
  BlockTeam tm{red};
  const int i = blockIdx.x;
  const GainOne in = make_one(p, cov, eps, taps, i, b, m, ws);
  gain_forward(tm, in, g + (size_t)i * b, kl_terms + 2 * i, beta_mean ? beta_mean + (size_t)i * b : nullptr,
               beta_var ? beta_var + (size_t)i * b : nullptr, status ? status + i : nullptr);
}

__global__ void gain_bwd_kernel(const __grid_constant__ VgGainParams p, const __grid_constant__ VgGainGrads gr,
                                const float* cov, const float* eps, const double* taps, const float* dg,
                                double kl_scale, int b, int m, double* ws) {
  __shared__ double red[32];
  BlockTeam tm{red};
  const int i = blockIdx.x;
  const GainOne in = make_one(p, cov, eps, taps, i, b, m, ws);
  GainGradOut out{gr.sa[i], gr.logstd[i], gr.qu_m[i], gr.qu_S[i], gr.logkvar[i], gr.logls[i]};
  gain_backward(tm, in, dg + (size_t)i * b, kl_scale, out);
}

// gp.GP.evaluate_posterior for arbitrary query points (plot_GPs uses N = all rows).
__global__ void gp_posterior_kernel(const float* xu, int m, const float* k_var, const float* ls_p,
                                    const float* qu_m, const float* qu_S, const float* xq, int nq, float* f_bar,
                                    float* var, float* sigma, double* a_ws) {
  __shared__ double Kinv[kMaxInducing * kMaxInducing];
  __shared__ double M[kMaxInducing * kMaxInducing];
  __shared__ double scratch[2 * kMaxInducing * kMaxInducing];
  __shared__ double Ku[kMaxInducing * kMaxInducing];
  const double kvar = (double)k_var[0], ls = (double)ls_p[0];
  const double step = (double)xu[1] - (double)xu[0], xu0 = (double)xu[0];
  if (threadIdx.x == 0) {
    for (int p = 0; p < m; ++p)
      for (int q = 0; q < m; ++q) {
        Ku[p * m + q] = rbf_k(fabs((double)(p - q)) * step, kvar, ls);
        M[p * m + q] = (double)qu_S[p * m + q] - Ku[p * m + q];
      }
    small_inverse(Ku, Kinv, m, scratch);
  }
  __syncthreads();
  // each block recomputes the tiny inverse; A rows go to a_ws (nq, m) for the optional Sigma
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < nq; a += gridDim.x * blockDim.x) {
    double knu[kMaxInducing], av[kMaxInducing];
    const double base = xu0 - (double)xq[a];
    for (int p = 0; p < m; ++p) knu[p] = rbf_k(base + p * step, kvar, ls);
    double fb = 0.0;
    for (int p = 0; p < m; ++p) {
      double s = 0.0;
      for (int q = 0; q < m; ++q) s += knu[q] * Kinv[q * m + p];
      av[p] = s;
      fb += s * (double)qu_m[p];
      if (a_ws) a_ws[(size_t)a * m + p] = s;
    }
    double v = kvar;
    for (int p = 0; p < m; ++p)
      for (int q = 0; q < m; ++q) v += av[p] * M[p * m + q] * av[q];
    f_bar[a] = (float)fb;
    if (var) var[a] = (float)v;
  }
}

__global__ void gp_sigma_kernel(const float* xu, int m, const float* k_var, const float* ls_p, const float* qu_S,
                                const float* xq, int nq, const double* a_ws, float* sigma) {
  __shared__ double M[kMaxInducing * kMaxInducing];
  const double kvar = (double)k_var[0], ls = (double)ls_p[0];
  const double step = (double)xu[1] - (double)xu[0];
  if (threadIdx.x == 0)
    for (int p = 0; p < m; ++p)
      for (int q = 0; q < m; ++q)
        M[p * m + q] = (double)qu_S[p * m + q] - rbf_k(fabs((double)(p - q)) * step, kvar, ls);
  __syncthreads();
  const long long total = (long long)nq * nq;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(t / nq), b = (int)(t % nq);
    double s = rbf_k((double)xq[b] - (double)xq[a], kvar, ls);
    for (int p = 0; p < m; ++p) {
      double am = 0.0;
      for (int q = 0; q < m; ++q) am += M[p * m + q] * a_ws[(size_t)b * m + q];
      s += a_ws[(size_t)a * m + p] * am;
    }
    sigma[t] = (float)s;
  }
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_gain_workspace_bytes(int b, int m) { return 8 * gain_ws_doubles(b, m) * sizeof(double); }

static int gain_threads(int b) { return b <= 32 ? 32 : 128; }

extern "C" int vg_gain_fwd(const VgGainParams* p, const float* covariates, const float* eps, const double* taps,
                           int b, int m, float* g, double* kl_terms, float* beta_mean, float* beta_var,
                           int* status, void* workspace, size_t workspace_bytes, void* stream) {
  VG_CHECK_ARG(p && covariates && eps && taps && g && kl_terms, "null pointer");
  VG_CHECK_ARG(b > 0 && m >= 2 && m <= kMaxInducing, "need b > 0 and 2 <= m <= 16");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_gain_workspace_bytes(b, m), "workspace too small");
  gain_fwd_kernel<<<8, gain_threads(b), 0, as_stream(stream)>>>(*p, covariates, eps, taps, b, m, g, kl_terms,
                                                                 beta_mean, beta_var, status, (double*)workspace);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_gain_bwd(const VgGainParams* p, const VgGainGrads* grads, const float* covariates,
                           const float* eps, const double* taps, const float* dg, double kl_scale, int b, int m,
                           void* workspace, size_t workspace_bytes, void* stream) {
  VG_CHECK_ARG(p && grads && covariates && eps && taps && dg, "null pointer");
  VG_CHECK_ARG(b > 0 && m >= 2 && m <= kMaxInducing, "need b > 0 and 2 <= m <= 16");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_gain_workspace_bytes(b, m), "workspace too small");
  gain_bwd_kernel<<<8, gain_threads(b), 0, as_stream(stream)>>>(*p, *grads, covariates, eps, taps, dg, kl_scale,
                                                                 b, m, (double*)workspace);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_gp_posterior(const float* xu, int m, const float* k_var, const float* ls, const float* qu_m,
                               const float* qu_S, const float* xq, int nq, float* f_bar, float* var,
                               float* sigma, double* a_ws, void* stream) {
  VG_CHECK_ARG(xu && k_var && ls && qu_m && qu_S && xq && f_bar, "null pointer");
  VG_CHECK_ARG(nq > 0 && m >= 2 && m <= kMaxInducing, "need nq > 0 and 2 <= m <= 16");
  VG_CHECK_ARG(!sigma || a_ws, "sigma needs the (nq, m) fp64 scratch a_ws");
  cudaStream_t st = as_stream(stream);
  gp_posterior_kernel<<<cdiv(nq, 128), 128, 0, st>>>(xu, m, k_var, ls, qu_m, qu_S, xq, nq, f_bar, var, sigma, a_ws);
  VG_LAUNCH_CHECK();
  if (sigma) {
    long long total = (long long)nq * nq;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 8 * vg_sm_count()) blocks = 8 * vg_sm_count();
    gp_sigma_kernel<<<blocks, 256, 0, st>>>(xu, m, k_var, ls, qu_S, xq, nq, a_ws, sigma);
    VG_LAUNCH_CHECK();
  }
  return VG_OK;
}
