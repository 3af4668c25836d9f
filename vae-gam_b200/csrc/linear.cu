// Fully-connected layers fc1..fc8 (vae_reg_GP.py:197-210; used at :244-251, :255-258) and
// their gradients: one strided fp32 tile GEMM, C(m,n) (+)= sum_k A(m,k)*B(k,n), with an
// optional ReLU mask on A (the saved layer output) and a bias/activation epilogue.
// These layers are <1.5 % of the step's FLOPs; the kernel is sized for
// skinny M (= minibatch) rather than for tensor cores.
#include "common.cuh"

namespace vg {

constexpr int TM = 64, TN = 64, TK = 32;   // deep K tile: these GEMMs are latency-bound, few sequential K steps matter most

struct GemmArgs {
  const float* A; long long as_m, as_k;
  const float* MA;                  // mask operand, indexed like A (A_eff = MA > 0 ? A : 0) or null
  const float* B; long long bs_k, bs_n;
  float* C; long long cs_m, cs_n;
  const float* bias;                // (n) or null
  int m, n, k;
  int act;
  int accumulate;                   // C += result (atomic when split-k > 1)
  int ksplit;                       // gridDim.z
  float* rowsum;                    // (m) += sum_k A_eff(m,k)   (bias gradient, fused into the dW GEMM) or null
};

// AK / BK: the K index is the contiguous one of A / B (compile-time, so each tile load is one code path)
template <bool AK, bool BK>
__global__ void __launch_bounds__(256, 2) gemm_kernel(const GemmArgs a) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kper = ((a.k + a.ksplit - 1) / a.ksplit + TK - 1) / TK * TK;
  const int kbeg = blockIdx.z * kper;
  const int kend = min(a.k, kbeg + kper);
  float acc[4][4];
  float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // Register double buffering: the global loads of tile k0 + TK are in flight while tile k0 is multiplied
  // (these GEMMs are a few CTAs deep in K, so the load latency of every K step used to be exposed).
  // Element e of a tile: the faster-varying thread index runs along whichever of (m, k) is contiguous.
  constexpr int NA = TM * TK / 256, NB = TN * TK / 256;
  float ra[NA], rb[NB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const int e = tid + r * 256;
      int mm, kk;
      if (AK) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < a.m && gk < kend) {
        const long long off = gm * a.as_m + gk * a.as_k;
        v = __ldg(a.A + off);
        if (a.MA && !(__ldg(a.MA + off) > 0.f)) v = 0.f;
      }
      ra[r] = v;
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      const int e = tid + r * 256;
      int nn, kk;
      if (BK) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
      const int gn = n0 + nn, gk = k0 + kk;
      rb[r] = (gn < a.n && gk < kend) ? __ldg(a.B + gk * a.bs_k + gn * a.bs_n) : 0.f;
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const int e = tid + r * 256;
      int mm, kk;
      if (AK) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      sA[kk][mm] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      const int e = tid + r * 256;
      int nn, kk;
      if (BK) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
      sB[kk][nn] = rb[r];
    }
    __syncthreads();
    if (k0 + TK < kend) fetch(k0 + TK);
    if (a.rowsum && blockIdx.x == 0 && tx == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = 0.f;
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) s += sA[kk][ty * 4 + i];
        rs[i] += s;
      }
    }
#pragma unroll 8
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (a.rowsum && blockIdx.x == 0 && tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ty * 4 + i;
      if (gm < a.m) atomicAdd(a.rowsum + gm, rs[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.n) continue;
      float v = acc[i][j];
      float* dst = a.C + gm * a.cs_m + gn * a.cs_n;
      if (a.ksplit > 1) {
        atomicAdd(dst, v);
      } else {
        if (a.bias) v += __ldg(a.bias + gn);
        if (a.act == VG_ACT_RELU) v = fmaxf(v, 0.f);
        else if (a.act == VG_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
        if (a.accumulate) v += *dst;
        *dst = v;
      }
    }
  }
}

// db[n] += sum_m (mask ? dy*(y>0) : dy)
__global__ void colsum_kernel(const float* __restrict__ dy, const float* __restrict__ y, int m, int n, float* db) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  float s = 0.f;
  for (int r = 0; r < m; ++r) {
    float v = __ldg(dy + (size_t)r * n + col);
    if (y && !(__ldg(y + (size_t)r * n + col) > 0.f)) v = 0.f;
    s += v;
  }
  db[col] += s;
}

__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, int m, int n, int act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * n) return;
  float v = y[i] + (bias ? __ldg(bias + i % n) : 0.f);
  if (act == VG_ACT_RELU) v = fmaxf(v, 0.f);
  else if (act == VG_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
  y[i] = v;
}

// split the reduction when the output grid alone cannot fill the machine
static int pick_ksplit(int m, int n, int k) {
  const int blocks = cdiv(n, TN) * cdiv(m, TM);
  if (blocks >= 64 || k < 512) return 1;
  int s = 2 * vg_sm_count() / blocks;      // two resident CTAs per SM keep the K chains short
  const int maxs = k / 128;
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

static int gemm(GemmArgs a, cudaStream_t st) {
  if (a.ksplit < 1) a.ksplit = 1;
  dim3 grid(cdiv(a.n, TN), cdiv(a.m, TM), a.ksplit);
  const bool ak = a.as_k == 1, bk = a.bs_k == 1;
  if (ak && bk) gemm_kernel<true, true><<<grid, 256, 0, st>>>(a);
  else if (ak) gemm_kernel<true, false><<<grid, 256, 0, st>>>(a);
  else if (bk) gemm_kernel<false, true><<<grid, 256, 0, st>>>(a);
  else gemm_kernel<false, false><<<grid, 256, 0, st>>>(a);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_linear_fwd(const float* x, const float* w, const float* bias, float* y, int m, int n, int k,
                             int act, void* stream) {
  VG_CHECK_ARG(x && w && y && m > 0 && n > 0 && k > 0, "bad arguments");
  GemmArgs a{};
  a.A = x; a.as_m = k; a.as_k = 1;
  a.B = w; a.bs_k = 1; a.bs_n = k;
  a.C = y; a.cs_m = n; a.cs_n = 1;
  a.m = m; a.n = n; a.k = k;
  a.ksplit = pick_ksplit(m, n, k);
  cudaStream_t st = as_stream(stream);
  if (a.ksplit == 1) {
    a.bias = bias; a.act = act;
    return gemm(a, st);
  }
  VG_CUDA(cudaMemsetAsync(y, 0, (size_t)m * n * sizeof(float), st));
  VG_TRY(gemm(a, st));
  bias_act_kernel<<<cdiv((long long)m * n, 256), 256, 0, st>>>(y, bias, m, n, act);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_linear_bwd(const float* dy, const float* relu_out, const float* x, const float* w, float* dx,
                             float* dw, float* db, int m, int n, int k, void* stream) {
  VG_CHECK_ARG(dy && x && w && m > 0 && n > 0 && k > 0, "bad arguments");
  cudaStream_t st = as_stream(stream);
  if (dx) {  // dx (m,k) = dYm (m,n) @ W (n,k)
    GemmArgs a{};
    a.A = dy; a.MA = relu_out; a.as_m = n; a.as_k = 1;
    a.B = w; a.bs_k = k; a.bs_n = 1;
    a.C = dx; a.cs_m = k; a.cs_n = 1;
    a.m = m; a.n = k; a.k = n;
    a.ksplit = pick_ksplit(m, k, n);
    if (a.ksplit > 1) VG_CUDA(cudaMemsetAsync(dx, 0, (size_t)m * k * sizeof(float), st));
    VG_TRY(gemm(a, st));
  }
  if (dw) {  // dw (n,k) += dYm^T (n,m) @ X (m,k)
    GemmArgs a{};
    a.A = dy; a.MA = relu_out; a.as_m = 1; a.as_k = n;
    a.B = x; a.bs_k = k; a.bs_n = 1;
    a.C = dw; a.cs_m = k; a.cs_n = 1;
    a.m = n; a.n = k; a.k = m; a.accumulate = 1; a.ksplit = 1;
    a.rowsum = db;                 // db (n) += sum_m dYm(m, n): row sums of this GEMM's A operand
    VG_TRY(gemm(a, st));
  } else if (db) {
    colsum_kernel<<<cdiv(n, 128), 128, 0, st>>>(dy, relu_out, m, n, db);
    VG_LAUNCH_CHECK();
  }
  return VG_OK;
}
