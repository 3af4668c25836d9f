// Batch-statistics BatchNorm helpers (nn.BatchNorm3d(track_running_stats=False),
// vae_reg_GP.py:194-196,216-218).  The normalisation itself is folded into the consumer
// convolution's operand load (conv.cu); what remains are the reductions that cannot be
// produced by a convolution epilogue, the coefficient finalisation, the backward apply
// and the layout changes at the conv <-> fully-connected seams.
#include "common.cuh"
#include "tc_common.cuh"

namespace vg {

// x: (n, spatial, c) channels-last.  grid = (blocks_per_image, n).
template <int C>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, long long spatial, int group_size, double* stats) {
  __shared__ double sred[2 * C];
  const int tid = threadIdx.x;
  if (tid < 2 * C) sred[tid] = 0.0;
  __syncthreads();
  const int n = blockIdx.y;
  const long long total = spatial * C;           // floats in this image
  const float* xi = x + (size_t)n * total;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  // each thread walks float4 lanes; with C in {1,8,16} and 256 threads the channel of a
  // lane is fixed per thread when the stride (in floats) is a multiple of C.
  if constexpr (C % 4 == 0) {
    const long long nv = total / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;   // multiple of C/4 (256 % 4 == 0)
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < nv; i += stride) {
      float4 v = ldg_stream(reinterpret_cast<const float4*>(xi) + i);
      s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
      s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
      s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
    }
    const int c0 = (int)((((long long)blockIdx.x * blockDim.x + tid) * 4) % C);
    // lanes that differ by a multiple of C / 4 hold the same four channels: reduce them in the warp first, so a warp
    // issues C / 4 shared-memory atomics per sum instead of 32 (they serialise: this was 25 us for a 4 MB tensor)
#pragma unroll
    for (int o = 16; o >= C / 4; o >>= 1)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
      }
    if ((tid & 31) < C / 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sred[2 * (c0 + j)], (double)s1[j]);
        atomicAdd(&sred[2 * (c0 + j) + 1], (double)s2[j]);
      }
    }
  } else {  // C == 1
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < total; i += stride) {
      const float v = __ldg(xi + i);
      s1[0] += v;
      s2[0] = fmaf(v, v, s2[0]);
    }
    const float a = warp_sum(s1[0]), b = warp_sum(s2[0]);
    if ((tid & 31) == 0) {
      atomicAdd(&sred[0], (double)a);
      atomicAdd(&sred[1], (double)b);
    }
  }
  __syncthreads();
  if (tid < 2 * C) atomicAdd(stats + (size_t)(n / group_size) * 2 * C + tid, sred[tid]);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int groups, int c, double count,
                                   float* scale, float* shift, float* istd, float* mistd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * c) return;
  const int ch = i % c;
  const double mean = stats[2 * i] / count;
  double var = stats[2 * i + 1] / count - mean * mean;   // biased variance
  if (var < 0) var = 0;
  const double is = 1.0 / sqrt(var + 1e-5);
  const double sc = (double)gamma[ch] * is;
  scale[i] = (float)sc;
  shift[i] = (float)((double)beta[ch] - mean * sc);
  if (istd) istd[i] = (float)is;
  if (mistd) mistd[i] = (float)(mean * is);
}

// dx = scale * (dy - m1 - xhat*m2) [* (x>0)].  A thread walks float4s of one image with a stride that is a
// multiple of C / 4, so it always meets the same 4 channels: their coefficients are computed once (the two
// fp64 divisions per channel used to be redone for every element) and 4 float4 pairs are in flight per pass.
template <int C>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                    const double* __restrict__ sums, const float* __restrict__ scale,
                    const float* __restrict__ istd, const float* __restrict__ mistd, int group_size,
                    long long spatial, double count, int relu_mask, float* dx) {
  const int n = blockIdx.y;
  const int grp = n / group_size;
  const long long total = spatial * C;
  const size_t base = (size_t)n * total;
  constexpr int W = (C % 4 == 0) ? 4 : 1;
  constexpr int U = 4;
  const long long nv = total / W;
  const long long stride = (long long)gridDim.x * blockDim.x;     // multiple of 256, hence of C / W
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float m1[W], m2[W], is[W], mis[W], sc[W];
  {
    const int c0 = (int)((i0 * W) % C);
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int gc = grp * C + c0 + j;
      m1[j] = (float)(sums[2 * gc] / count);
      m2[j] = (float)(sums[2 * gc + 1] / count);
      is[j] = istd[gc]; mis[j] = mistd[gc]; sc[j] = scale[gc];
    }
  }
  for (long long i = i0; i < nv; i += U * stride) {
    float4 a[U], b[U];
    float a1[U], b1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * stride;
      if (iu >= nv) break;
      if constexpr (W == 4) {
        a[u] = ldg_stream(reinterpret_cast<const float4*>(dy + base) + iu);
        b[u] = ldg_stream(reinterpret_cast<const float4*>(x + base) + iu);
      } else {
        a1[u] = dy[base + iu];
        b1[u] = x[base + iu];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * stride;
      if (iu >= nv) break;
      float d[4], xv[4], o[4];
      if constexpr (W == 4) {
        d[0] = a[u].x; d[1] = a[u].y; d[2] = a[u].z; d[3] = a[u].w;
        xv[0] = b[u].x; xv[1] = b[u].y; xv[2] = b[u].z; xv[3] = b[u].w;
      } else {
        d[0] = a1[u]; xv[0] = b1[u];
      }
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float xh = fmaf(xv[j], is[j], -mis[j]);
        float r = sc[j] * (d[j] - m1[j] - xh * m2[j]);
        if (relu_mask && !(xv[j] > 0.f)) r = 0.f;
        o[j] = r;
      }
      if constexpr (W == 4)
        stg_stream(reinterpret_cast<float4*>(dx + base) + iu, make_float4(o[0], o[1], o[2], o[3]));
      else
        dx[base + iu] = o[0];
    }
  }
}

// The same apply for tensors stored as bf16 (x and / or dx; dy is fp32): a thread item is 8 consecutive channels
// (one 16-byte bf16 word, two fp32 words), its stride a multiple of C / 8 items, so the 8 channels' coefficients
// stay in registers; two items in flight.
template <int C, bool X16, bool DX16>
__global__ void __launch_bounds__(256)
bn_bwd_apply8_kernel(const float* __restrict__ dy, const void* __restrict__ x_, const double* __restrict__ sums,
                     const float* __restrict__ scale, const float* __restrict__ istd, const float* __restrict__ mistd,
                     int group_size, long long spatial, double count, int relu_mask, void* dx_, float* dx_sum) {
  static_assert(C % 8 == 0, "8-channel items");
  __shared__ float s_sum[C];
  if (dx_sum && threadIdx.x < C) s_sum[threadIdx.x] = 0.f;
  float osum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int n = blockIdx.y;
  const int grp = n / group_size;
  const long long total = spatial * C;
  const size_t base = (size_t)n * total;
  constexpr int U = 2;
  const long long nv = total / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;     // multiple of 256, hence of C / 8
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float m1[8], m2[8], is[8], mis[8], sc[8];
  {
    const int c0 = (int)((i0 * 8) % C);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gc = grp * C + c0 + j;
      m1[j] = (float)(sums[2 * gc] / count);
      m2[j] = (float)(sums[2 * gc + 1] / count);
      is[j] = istd[gc]; mis[j] = mistd[gc]; sc[j] = scale[gc];
    }
  }
  const float* xf = static_cast<const float*>(x_);
  const __nv_bfloat16* xh = static_cast<const __nv_bfloat16*>(x_);
  for (long long i = i0; i < nv; i += U * stride) {
    float4 a[U][2], b[U][2];
    uint4 bq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * stride;
      if (iu >= nv) break;
      a[u][0] = ldg_stream(reinterpret_cast<const float4*>(dy + base) + 2 * iu);
      a[u][1] = ldg_stream(reinterpret_cast<const float4*>(dy + base) + 2 * iu + 1);
      if constexpr (X16) {
        bq[u] = ldg_u4(xh + base + 8 * iu);
      } else {
        b[u][0] = ldg_stream(reinterpret_cast<const float4*>(xf + base) + 2 * iu);
        b[u][1] = ldg_stream(reinterpret_cast<const float4*>(xf + base) + 2 * iu + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * stride;
      if (iu >= nv) break;
      float d[8] = {a[u][0].x, a[u][0].y, a[u][0].z, a[u][0].w, a[u][1].x, a[u][1].y, a[u][1].z, a[u][1].w};
      float xv[8], o[8];
      if constexpr (X16) {
        unpack_bf16x8(bq[u], xv);
      } else {
        xv[0] = b[u][0].x; xv[1] = b[u][0].y; xv[2] = b[u][0].z; xv[3] = b[u][0].w;
        xv[4] = b[u][1].x; xv[5] = b[u][1].y; xv[6] = b[u][1].z; xv[7] = b[u][1].w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xhat = fmaf(xv[j], is[j], -mis[j]);
        float r = sc[j] * (d[j] - m1[j] - xhat * m2[j]);
        if (relu_mask && !(xv[j] > 0.f)) r = 0.f;
        o[j] = r;
        osum[j] += r;
      }
      if constexpr (DX16) {
        reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(dx_) + base)[iu] = pack_bf16x8(o);
      } else {
        float* dxf = static_cast<float*>(dx_) + base;
        stg_stream(reinterpret_cast<float4*>(dxf) + 2 * iu, make_float4(o[0], o[1], o[2], o[3]));
        stg_stream(reinterpret_cast<float4*>(dxf) + 2 * iu + 1, make_float4(o[4], o[5], o[6], o[7]));
      }
    }
  }
  if (dx_sum) {      // per-channel sum of dx (the bias gradient of the layer that produced x): block reduce, one atomic per channel
    __syncthreads();
    const int c0 = (int)((i0 * 8) % C);
    constexpr int PERL = C / 8;                      // lanes l and l + PERL hold the same channels
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = osum[j];
      for (int o = 16; o >= PERL; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) < PERL) atomicAdd(&s_sum[c0 + j], v);
    }
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(dx_sum + threadIdx.x, s_sum[threadIdx.x]);
  }
}

// ---- fused BatchNorm backward of a BatchNorm -> ConvTranspose3d(k3, s1) junction (bnt5 -> convt5) ------------------
// Box sums of a one-channel gradient: out[g][t] = sum_{n in g} sum_{v in x-grid} dy(n, v + tap_t - pad), t = (a,b,c) in
// 3x3x3, plus out[g][27] = sum of all of dy.  One block per (image, d-plane); a voxel's contribution to the nine
// (b, c) taps is decided by its (h, w), to the three a's by its d.
__global__ void __launch_bounds__(256)
box_sums_kernel(const float* __restrict__ dy, int group_size, int yD, int yH, int yW, long long y_img, int xD, int xH, int xW,
                int pD, int pH, int pW, double* out) {
  // one block per image: a thread keeps the nine (b, c) sums of the plane it is walking and folds them into its 27
  // tap sums for the a's that plane belongs to; 28 atomics per block at the end
  const int n = blockIdx.x;
  const float* p0 = dy + (size_t)n * y_img;
  const int plane = yH * yW;
  float acc[27], tot = 0.f;
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  // each thread owns fixed (h, w) positions (e = tid + k * 256), so its tap masks do not depend on the plane
  for (int e = threadIdx.x; e < plane; e += blockDim.x) {
    const int h = e / yW, w = e - h * yW;
    bool okb[3], okc[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) { const int xh = h - b + pH; okb[b] = xh >= 0 && xh < xH; }
#pragma unroll
    for (int c = 0; c < 3; ++c) { const int xw = w - c + pW; okc[c] = xw >= 0 && xw < xW; }
    float sa[3] = {0.f, 0.f, 0.f};          // sum over the planes each a covers
    for (int d = 0; d < yD; ++d) {
      const float v = __ldg(p0 + (size_t)d * plane + e);
      tot += v;
#pragma unroll
      for (int a = 0; a < 3; ++a) { const int xd = d - a + pD; if (xd >= 0 && xd < xD) sa[a] += v; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (okb[b] && okc[c]) acc[a * 9 + b * 3 + c] += sa[a];
  }
  __shared__ float red[28];
  if (threadIdx.x < 28) red[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 27; ++i) {
    const float r = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], r);
  }
  {
    const float r = warp_sum(tot);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[27], r);
  }
  __syncthreads();
  if (threadIdx.x < 28) atomicAdd(out + (size_t)(n / group_size) * 28 + threadIdx.x, (double)red[threadIdx.x]);
}

// From the per-group raw products R[g][t][c] = sum x(v,c) dy(v + t) (x = the BatchNorm INPUT, un-normalised), the box
// sums D0[g][t] and the layer's weights W[c][t] (ConvTranspose3d with one output channel): everything the junction's
// backward needs that is not per-voxel.
//   weight gradient   dW[c][t] += sum_g scale[g,c] R[g,t,c] + shift[g,c] D0[g,t];   dbias += sum_g D0[g][27]
//   BatchNorm sums    s1[g,c] = sum_t W[c,t] D0[g,t];  s2[g,c] = sum_t W[c,t] (istd[g,c] R[g,t,c] - mistd[g,c] D0[g,t])
//   (s1 = sum of the data gradient, s2 = sum of data gradient x xhat: the identities hold because every tap of a
//    full transposed convolution with stride 1 is inside the output for every input voxel)
//   dgamma[c] += sum_g s2, dbeta[c] += sum_g s1
//   apply coefficients  dx = (x > 0) * (A dy + B x + C):  A = scale, B = -scale m2 istd, C = scale (mistd m2 - m1),
//   m1 = s1 / count, m2 = s2 / count  ->  coef[g][c][0..2]
__global__ void bn_fused_finalize_kernel(const float* __restrict__ raw, const double* __restrict__ box, const float* __restrict__ w,
                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ istd, const float* __restrict__ mistd, int groups, int c,
                                         double count, float* dw, float* dbias, float* dgamma, float* dbeta, float* coef) {
  // one warp per channel; lanes over taps
  const int ch = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ch >= c) return;
  double dg = 0.0, db = 0.0;
  float dwt = 0.f;
  for (int g = 0; g < groups; ++g) {
    const int gc = g * c + ch;
    double s1 = 0.0, s2 = 0.0;
    if (lane < 27) {
      const double r = (double)raw[((size_t)g * 27 + lane) * c + ch], d0 = box[(size_t)g * 28 + lane];
      // ConvTranspose3d weight (cin, 1, 3, 3, 3), ROUNDED to bf16 like the operand of the data-gradient MMAs: the
      // sums must describe the data gradient that is actually computed (a systematic 2^-9 mismatch in m1 would
      // show up, un-averaged, in every voxel — and in the producer's bias gradient)
      const double wv = (double)__bfloat162float(__float2bfloat16(w[(size_t)ch * 27 + lane]));
      s1 = wv * d0;
      s2 = wv * ((double)istd[gc] * r - (double)mistd[gc] * d0);
      dwt += (float)((double)scale[gc] * r + (double)shift[gc] * d0);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const double m1 = s1 / count, m2 = s2 / count;
    if (lane == 0) {
      coef[3 * gc] = scale[gc];
      coef[3 * gc + 1] = (float)(-(double)scale[gc] * m2 * (double)istd[gc]);
      coef[3 * gc + 2] = (float)((double)scale[gc] * ((double)mistd[gc] * m2 - m1));
    }
    dg += s2; db += s1;
  }
  if (lane < 27) dw[(size_t)ch * 27 + lane] += dwt;
  if (lane == 0) {
    if (dgamma) dgamma[ch] += (float)dg;
    if (dbeta) dbeta[ch] += (float)db;
    if (ch == 0 && dbias) {
      double t = 0.0;
      for (int g = 0; g < groups; ++g) t += box[(size_t)g * 28 + 27];
      dbias[0] += (float)t;
    }
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, int groups, int c, float* dgamma,
                                     float* dbeta) {
  const int ch = threadIdx.x;
  if (ch >= c) return;
  double a = 0, b = 0;
  for (int g = 0; g < groups; ++g) {
    b += sums[2 * (g * c + ch)];
    a += sums[2 * (g * c + ch) + 1];
  }
  if (dgamma) dgamma[ch] += (float)a;
  if (dbeta) dbeta[ch] += (float)b;
}

// (n, c, spatial) -> (n, spatial, c) through a 32x33 shared tile
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // per image: src is (rows, cols), dst is (cols, rows)
  __shared__ float tile[32][33];
  const size_t img = (size_t)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[img + (size_t)r * cols + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) dst[img + (size_t)c2 * rows + r2] = tile[threadIdx.x][j];
  }
}

static int transpose(const float* src, float* dst, int n, int rows, int cols, cudaStream_t st) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32), n), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(src, dst, rows, cols);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_bn_stats(const float* x, int n, int group_size, long long spatial, int c, double* stats,
                           void* stream) {
  VG_CHECK_ARG(x && stats && n > 0 && group_size > 0 && n % group_size == 0, "bad arguments");
  long long per_img = spatial * c;
  int bx = (int)((per_img / 4 + 255) / 256);
  int cap = (4 * vg_sm_count() + n - 1) / n;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  if (c == 1) bn_stats_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, spatial, group_size, stats);
  else if (c == 8) bn_stats_kernel<8><<<grid, 256, 0, as_stream(stream)>>>(x, spatial, group_size, stats);
  else if (c == 16) bn_stats_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(x, spatial, group_size, stats);
  else { set_error("vg_bn_stats: channels must be 1, 8 or 16 (got %d)", c); return VG_EINVAL; }
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_bn_finalize(const double* stats, const float* gamma, const float* beta, int groups, int c,
                              double count, float* scale, float* shift, float* istd, float* mistd,
                              void* stream) {
  VG_CHECK_ARG(stats && gamma && beta && scale && shift && groups > 0 && c > 0, "bad arguments");
  bn_finalize_kernel<<<cdiv(groups * c, 128), 128, 0, as_stream(stream)>>>(stats, gamma, beta, groups, c, count,
                                                                           scale, shift, istd, mistd);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_bn_bwd_apply(const float* dy, const void* x_, const double* sums, const float* scale,
                               const float* istd, const float* mistd, int n, int group_size,
                               long long spatial, int c, double count, int relu_mask, int bf16_mask, void* dx_,
                               float* dgamma, float* dbeta, float* dx_chan_sum, void* stream) {
  const float* x = static_cast<const float*>(x_);
  float* dx = static_cast<float*>(dx_);
  VG_CHECK_ARG(x && sums && n > 0 && group_size > 0 && n % group_size == 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (dx && (bf16_mask & (VG_BF16_X | VG_BF16_DX))) {
    VG_CHECK_ARG(dy && scale && istd && mistd, "null coefficient");
    VG_CHECK_ARG(c == 8 || c == 16, "bf16 storage needs 8 or 16 channels");
    VG_CHECK_ARG(!(bf16_mask & VG_BF16_DX) || (const void*)dy != dx_, "bf16 dx cannot alias the fp32 dy");
    const long long per_img = spatial * c;
    int bx = (int)((per_img / 8 + 255) / 256);
    int cap = 8 * vg_sm_count() / n;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n);
    const bool x16 = (bf16_mask & VG_BF16_X) != 0, d16 = (bf16_mask & VG_BF16_DX) != 0;
#define VG_BN8(C, X, D) bn_bwd_apply8_kernel<C, X, D><<<grid, 256, 0, st>>>(dy, x_, sums, scale, istd, mistd, group_size, spatial, count, relu_mask, dx_, dx_chan_sum)
    if (c == 8) { if (x16 && d16) VG_BN8(8, true, true); else if (x16) VG_BN8(8, true, false); else VG_BN8(8, false, true); }
    else { if (x16 && d16) VG_BN8(16, true, true); else if (x16) VG_BN8(16, true, false); else VG_BN8(16, false, true); }
#undef VG_BN8
    VG_LAUNCH_CHECK();
  } else if (dx) {
    VG_CHECK_ARG(dy && scale && istd && mistd, "null coefficient");
    VG_CHECK_ARG(dx_chan_sum == nullptr, "dx_chan_sum is produced by the bf16-storage variant only");
    long long per_img = spatial * c;
    int bx = (int)((per_img / 4 + 255) / 256);
    int cap = 8 * vg_sm_count() / n;          // whole grid resident at once (8 CTAs of 256 threads per SM): no tail wave
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n);
    if (c == 1) bn_bwd_apply_kernel<1><<<grid, 256, 0, st>>>(dy, x, sums, scale, istd, mistd, group_size, spatial, count, relu_mask, dx);
    else if (c == 8) bn_bwd_apply_kernel<8><<<grid, 256, 0, st>>>(dy, x, sums, scale, istd, mistd, group_size, spatial, count, relu_mask, dx);
    else if (c == 16) bn_bwd_apply_kernel<16><<<grid, 256, 0, st>>>(dy, x, sums, scale, istd, mistd, group_size, spatial, count, relu_mask, dx);
    else { set_error("vg_bn_bwd_apply: channels must be 1, 8 or 16 (got %d)", c); return VG_EINVAL; }
    VG_LAUNCH_CHECK();
  }
  if (dgamma || dbeta) {
    bn_param_grad_kernel<<<1, 32, 0, st>>>(sums, n / group_size, c, dgamma, dbeta);
    VG_LAUNCH_CHECK();
  }
  return VG_OK;
}

extern "C" int vg_nchw_to_nhwc(const float* src, float* dst, int n, int c, long long spatial, void* stream) {
  VG_CHECK_ARG(src && dst && n > 0 && c > 0 && spatial > 0, "bad arguments");
  return transpose(src, dst, n, c, (int)spatial, as_stream(stream));
}
extern "C" int vg_nhwc_to_nchw(const float* src, float* dst, int n, int c, long long spatial, void* stream) {
  VG_CHECK_ARG(src && dst && n > 0 && c > 0 && spatial > 0, "bad arguments");
  return transpose(src, dst, n, (int)spatial, c, as_stream(stream));
}

extern "C" int vg_box_sums(const float* dy, int n, int group_size, const int32_t* y_dims, long long y_img_stride,
                           const int32_t* x_dims, const int32_t* pad, double* out, void* stream) {
  VG_CHECK_ARG(dy && out && y_dims && x_dims && pad && n > 0 && group_size > 0 && n % group_size == 0, "bad arguments");
  const long long img = y_img_stride ? y_img_stride : (long long)y_dims[0] * y_dims[1] * y_dims[2];
  box_sums_kernel<<<n, 256, 0, as_stream(stream)>>>(dy, group_size, y_dims[0], y_dims[1], y_dims[2], img, x_dims[0], x_dims[1],
                                                        x_dims[2], pad[0], pad[1], pad[2], out);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_bn_fused_finalize(const float* raw, const double* box, const float* w, const float* scale, const float* shift,
                                    const float* istd, const float* mistd, int groups, int c, double count, float* dw,
                                    float* dbias, float* dgamma, float* dbeta, float* coef, void* stream) {
  VG_CHECK_ARG(raw && box && w && scale && shift && istd && mistd && dw && coef, "null argument");
  VG_CHECK_ARG(groups > 0 && c > 0 && c <= 32, "at most 32 channels");
  bn_fused_finalize_kernel<<<1, 32 * c, 0, as_stream(stream)>>>(raw, box, w, scale, shift, istd, mistd, groups, c, count, dw, dbias,
                                                                dgamma, dbeta, coef);
  VG_LAUNCH_CHECK();
  return VG_OK;
}
