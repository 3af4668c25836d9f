// Host build of csrc/gp_core.h with a 1-thread team: lets the CPU test-suite check the gain
// stage's forward/backward algebra (the exact source the CUDA kernel compiles) against the
// oracle without a GPU.  TEST INFRASTRUCTURE — never loaded by the product.
#include "../../vae-gam_b200/csrc/gp_core.h"
#include <stdlib.h>

namespace {
struct SerialTeam {
  int rank() const { return 0; }
  int size() const { return 1; }
  void sync() const {}
  double sum(double v) const { return v; }
};
}  // namespace

extern "C" {

size_t emul_gain_ws_doubles(int B, int m) { return vg::gain_ws_doubles(B, m); }

// params: sa, logstd, qu_m (m), qu_S (m*m), logkvar, logls, xu (m) as float arrays
void emul_gain_fwd(const float* cov, int cov_stride, int cov_col, const float* eps, const float* sa,
                   const float* logstd, const float* qu_m, const float* qu_S, const float* logkvar,
                   const float* logls, const float* xu, const double* taps, int has_gp, int hrf, int B, int m,
                   double* ws, float* g, double* kl, float* beta_mean, float* beta_var, int* status) {
  vg::GainOne in{cov, cov_stride, cov_col, eps, sa, logstd, qu_m, qu_S, logkvar, logls, xu, taps, has_gp, hrf, B, m, ws};
  SerialTeam tm;
  vg::gain_forward(tm, in, g, kl, beta_mean, beta_var, status);
}

void emul_gain_bwd(const float* cov, int cov_stride, int cov_col, const float* eps, const float* sa,
                   const float* logstd, const float* qu_m, const float* qu_S, const float* logkvar,
                   const float* logls, const float* xu, const double* taps, int has_gp, int hrf, int B, int m,
                   double* ws, const float* dg, double kl_scale, float* d_sa, float* d_logstd, float* d_qu_m,
                   float* d_qu_S, float* d_logkvar, float* d_logls) {
  vg::GainOne in{cov, cov_stride, cov_col, eps, sa, logstd, qu_m, qu_S, logkvar, logls, xu, taps, has_gp, hrf, B, m, ws};
  SerialTeam tm;
  vg::GainGradOut out{d_sa, d_logstd, d_qu_m, d_qu_S, d_logkvar, d_logls};
  vg::gain_backward(tm, in, dg, kl_scale, out);
}
}
