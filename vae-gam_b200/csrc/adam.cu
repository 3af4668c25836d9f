// Fused Adam over the flat parameter buffer (torch.optim.Adam defaults: vae_reg_GP.py:179,
// step at :429).  Same update rule as torch's single-tensor Adam:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// The step counter lives on the device so the call is CUDA-graph replayable.
#include "common.cuh"

namespace vg {

template <typename T>
__global__ void __launch_bounds__(256)
adam_kernel(T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m, T* __restrict__ v, long long n,
            double lr, double b1, double b2, double eps, double gscale, const long long* step_count,
            const int32_t* __restrict__ skip_flags, int n_flags) {
  // a minibatch that met a non-positive-definite covariance (status flags) must not move anything
  for (int i = 0; i < n_flags; ++i)
    if (skip_flags[i] != 0) return;
  const long long t = *step_count + 1;
  const double bc1 = 1.0 - pow(b1, (double)t);
  const double bc2 = 1.0 - pow(b2, (double)t);
  const T step_size = (T)(lr / bc1);
  const T inv_sqrt_bc2 = (T)(1.0 / sqrt(bc2));
  const T c1 = (T)(1.0 - b1), c2 = (T)(1.0 - b2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T gi = g[i] * (T)gscale;
    const T mi = m[i] + c1 * (gi - m[i]);            // lerp, as torch does
    const T vi = (T)b2 * v[i] + c2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const T denom = sqrt(vi) * inv_sqrt_bc2 + (T)eps;
    p[i] -= step_size * (mi / denom);
  }
}

__global__ void bump_kernel(long long* c, const int32_t* __restrict__ skip_flags, int n_flags) {
  for (int i = 0; i < n_flags; ++i)
    if (skip_flags[i] != 0) return;
  *c += 1;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_adam_step(float* p32, const float* g32, float* m32, float* v32, long long n32, double* p64,
                            const double* g64, double* m64, double* v64, long long n64, double lr, double beta1,
                            double beta2, double eps, double grad_scale, long long* step_count,
                            const int32_t* skip_flags, int n_flags, void* stream) {
  VG_CHECK_ARG(step_count, "null step counter");
  VG_CHECK_ARG(n_flags >= 0 && n_flags <= 64 && (n_flags == 0 || skip_flags), "bad skip flags");
  if (!skip_flags) n_flags = 0;
  cudaStream_t st = as_stream(stream);
  if (n32 > 0) {
    VG_CHECK_ARG(p32 && g32 && m32 && v32, "null fp32 segment");
    int blocks = (int)((n32 + 255) / 256);
    if (blocks > 4 * vg_sm_count()) blocks = 4 * vg_sm_count();
    adam_kernel<float><<<blocks, 256, 0, st>>>(p32, g32, m32, v32, n32, lr, beta1, beta2, eps, grad_scale, step_count,
                                               skip_flags, n_flags);
    VG_LAUNCH_CHECK();
  }
  if (n64 > 0) {
    VG_CHECK_ARG(p64 && g64 && m64 && v64, "null fp64 segment");
    int blocks = (int)((n64 + 255) / 256);
    if (blocks > 4 * vg_sm_count()) blocks = 4 * vg_sm_count();
    adam_kernel<double><<<blocks, 256, 0, st>>>(p64, g64, m64, v64, n64, lr, beta1, beta2, eps, grad_scale, step_count,
                                                skip_flags, n_flags);
    VG_LAUNCH_CHECK();
  }
  bump_kernel<<<1, 1, 0, st>>>(step_count, skip_flags, n_flags);
  VG_LAUNCH_CHECK();
  return VG_OK;
}
