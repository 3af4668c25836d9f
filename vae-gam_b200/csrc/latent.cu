// Latent sample + KL, fused:
//   vae_reg_GP.py:321-323  d-jitter (if ANY d < 1e-6, all d += 1e-6)
//   vae_reg_GP.py:324-325  LowRankMultivariateNormal(mu,u,d).rsample()
//                          = mu + u*eps_W + sqrt(d)*eps_D   (lowrank_multivariate_normal.py:214-223)
//   vae_reg_GP.py:400      KL(q(z|x) || N(0,I)), rank-1 closed form of torch kl.py:342-372
//   vae_reg_GP.py:326-330,339-343   the 9 decoder inputs [z | onehot(j)]
// One CTA; one warp per batch row (32 latents = 32 lanes).
#include "common.cuh"

namespace vg {

constexpr int L = 32;   // num_latents (vae_reg_GP.py:36)
constexpr int ZD = 41;  // z_dim = 32 + 8 + 1 (vae_reg_GP.py:45)

__global__ void __launch_bounds__(256)
latent_fwd_kernel(const float* __restrict__ heads, const float* __restrict__ eps_w,
                  const float* __restrict__ eps_d, int b, float* z, float* klz, float* d_out, float* zcat,
                  int* jitter_flag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* mu = heads;
  const float* u = heads + (size_t)b * L;
  const float* logd = heads + (size_t)2 * b * L;
  int small = 0;
  for (int r = warp; r < b; r += nwarps) small |= (expf(logd[r * L + lane]) < 1e-6f) ? 1 : 0;
  const int any_small = __syncthreads_or(small);
  const float jit = any_small ? 1e-6f : 0.f;
  if (threadIdx.x == 0 && jitter_flag) *jitter_flag = any_small ? 1 : 0;
  for (int r = warp; r < b; r += nwarps) {
    const int i = r * L + lane;
    const float m = mu[i], uu = u[i];
    const float d = expf(logd[i]) + jit;
    const float zz = m + uu * eps_w[r] + sqrtf(d) * eps_d[i];
    z[i] = zz;
    d_out[i] = d;
    const float su = warp_sum(uu * uu / d);
    const float s = warp_sum(-logf(d) + d + uu * uu + m * m);
    if (lane == 0) klz[r] = 0.5f * (-log1pf(su) + s - (float)L);
    for (int j = 0; j < 9; ++j) {
      float* row = zcat + ((size_t)j * b + r) * ZD;
      row[lane] = zz;
      if (lane < 9) row[L + lane] = (lane == j) ? 1.f : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
latent_bwd_kernel(const float* __restrict__ heads, const float* __restrict__ eps_w,
                  const float* __restrict__ eps_d, const float* __restrict__ d_used,
                  const float* __restrict__ dzcat, const float* __restrict__ dklz, int b, float* dheads) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* mu = heads;
  const float* u = heads + (size_t)b * L;
  const float* logd = heads + (size_t)2 * b * L;
  for (int r = warp; r < b; r += nwarps) {
    const int i = r * L + lane;
    float dz = 0.f;
    for (int j = 0; j < 9; ++j) dz += dzcat[((size_t)j * b + r) * ZD + lane];
    const float m = mu[i], uu = u[i], d = d_used[i];
    const float su = warp_sum(uu * uu / d);
    const float wk = dklz[r];
    const float inv1 = 1.f / (1.f + su);
    const float dmu = dz + wk * m;
    const float du = dz * eps_w[r] + wk * (uu - (uu / d) * inv1);
    const float dd = dz * eps_d[i] * 0.5f * rsqrtf(d) + wk * 0.5f * ((uu * uu) / (d * d) * inv1 - 1.f / d + 1.f);
    dheads[i] = dmu;
    dheads[(size_t)b * L + i] = du;
    dheads[(size_t)2 * b * L + i] = dd * expf(logd[i]);
  }
}

}  // namespace vg

using namespace vg;

extern "C" int vg_latent_fwd(const float* heads, const float* eps_w, const float* eps_d, int b, float* z,
                             float* klz, float* d_out, float* zcat, int* jitter_flag, void* stream) {
  VG_CHECK_ARG(heads && eps_w && eps_d && z && klz && d_out && zcat && b > 0, "bad arguments");
  latent_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(heads, eps_w, eps_d, b, z, klz, d_out, zcat, jitter_flag);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_latent_bwd(const float* heads, const float* eps_w, const float* eps_d, const float* d_used,
                             const float* dzcat, const float* dklz, int b, float* dheads, void* stream) {
  VG_CHECK_ARG(heads && eps_w && eps_d && d_used && dzcat && dklz && dheads && b > 0, "bad arguments");
  latent_bwd_kernel<<<1, 256, 0, as_stream(stream)>>>(heads, eps_w, eps_d, d_used, dzcat, dklz, b, dheads);
  VG_LAUNCH_CHECK();
  return VG_OK;
}
