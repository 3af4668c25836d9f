// Fused reconstruction + Gaussian likelihood + GLM regulariser, forward and backward
// (vae_reg_GP.py:380 cons = g*diff, :388-389 cdist regulariser,
// :390 x_rec accumulation, :401-405 Normal(x_rec, exp(-eps)).log_prob(x).sum(1)).
//
// HBM-bound by construction (~1 FLOP/B): ONE pass over the 9 decoder maps and x.
//   forward : reads 10*B*V fp32 (+9*V of eps/GLM per ROWS-row group, L2-resident)
//             -> logp (B), norms (8,B)                         algorithmic 40*B*V bytes
//   backward: reads the same, writes 9*B*V of d(pre-sigmoid)   algorithmic 76*B*V bytes
//             -> dg (8,B), deps (V)
// Data layout: every map row is padded to VP = round_up(V,4) floats so that rows are 16-byte
// aligned (V = 70315 is odd); x is the caller's dense (B,V) tensor and is read with scalar
// (still fully coalesced) loads.  A thread owns one float4 column (4 voxels), keeps eps / the
// 8 GLM values of those voxels in registers, and walks ROWS batch rows with 9 independent
// 128-bit streaming loads in flight per row.  Row sums are block-reduced once per CTA and
// written as per-CTA partials; a second tiny kernel adds them in a fixed order, so the
// result is deterministic (no float atomics on the outputs; deps uses fp32 RED by design).
#include "common.cuh"

namespace vg {

constexpr int ROWS = 4;          // batch rows per CTA row-group
constexpr int NMAP = 9;          // base + 8 covariate maps
constexpr int KCOV = 8;
constexpr int RL_THREADS = 256;
constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

__device__ __forceinline__ float4 ld4(const float* p) { return ldg_stream(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_keep(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct ReconArgs {
  const float* maps;   // (9, b, vp)
  const float* g;      // (8, b)
  const float* x;      // (b, v) dense
  const float* eps;    // (vp)
  const float* glm;    // (8, vp)
  int b;
  int v, vp;
  int ncols;           // vp / 4
  int col_chunks;      // gridDim.x
};

// Row sums: every warp reduces a row's NV values with shuffles and lane 0 adds them to the warp's
// own shared-memory slot (no atomics, fixed order), so no per-row accumulator lives in registers
// across the loop — the kernels stay under 85 registers and three CTAs share an SM.
__device__ __forceinline__ void warp_accumulate(float v, float* slot) {
  const float r = warp_sum(v);
  if ((threadIdx.x & 31) == 0) *slot += r;
}

// partial layout: [rowgroup][chunk][ROWS][9]  (slot 0 = logp, 1..8 = squared norms)
template <bool MAPS>
__global__ void __launch_bounds__(RL_THREADS, 3)
recon_fwd_kernel(const ReconArgs a, float* __restrict__ partial, float* cons_out, float* xrec_out) {
  __shared__ float s_g[ROWS][KCOV];
  __shared__ float s_acc[RL_THREADS / 32][ROWS * 9];
  const int rg = blockIdx.y;
  const int row0 = rg * ROWS;
  const int warp = threadIdx.x >> 5;
  const size_t map_stride = (size_t)a.b * a.vp;
  if (threadIdx.x < ROWS * KCOV) {
    const int r = threadIdx.x / KCOV, i = threadIdx.x % KCOV;
    s_g[r][i] = (row0 + r < a.b) ? __ldg(a.g + i * a.b + row0 + r) : 0.f;
  }
  for (int i = threadIdx.x; i < (RL_THREADS / 32) * ROWS * 9; i += RL_THREADS) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();

  // the loop bound is warp-uniform (whole warps stay in the shuffles); out-of-range lanes add zeros
  const int ncols_pad = (a.ncols + 31) & ~31;
  for (int col = blockIdx.x * RL_THREADS + threadIdx.x; col < ncols_pad; col += a.col_chunks * RL_THREADS) {
    const bool col_ok = col < a.ncols;
    const int v0 = (col_ok ? col : 0) * 4;
    const float4 e4 = ld4_keep(a.eps + v0);
    const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
    float ww[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) ww[l] = __expf(2.f * ee[l]);
#pragma unroll 1
    for (int r = 0; r < ROWS; ++r) {
      const int row = row0 + r;
      if (row >= a.b) break;
      float4 m4[NMAP];
#pragma unroll
      for (int j = 0; j < NMAP; ++j) m4[j] = ld4(a.maps + j * map_stride + (size_t)row * a.vp + v0);
      float xs[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) xs[l] = (v0 + l < a.v) ? __ldg(a.x + (size_t)row * a.v + v0 + l) : 0.f;
      float xr[4] = {m4[0].x, m4[0].y, m4[0].z, m4[0].w};
#pragma unroll
      for (int i = 0; i < KCOV; ++i) {
        const float gi = s_g[r][i];
        const float4 G = ld4_keep(a.glm + (size_t)i * a.vp + v0);     // L1-resident after the first row
        const float c[4] = {gi * m4[i + 1].x, gi * m4[i + 1].y, gi * m4[i + 1].z, gi * m4[i + 1].w};
        const float Gs[4] = {G.x, G.y, G.z, G.w};
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          xr[l] += c[l];
          const float d = c[l] - Gs[l];
          if (v0 + l < a.v) s = fmaf(d, d, s);
        }
        warp_accumulate(col_ok ? s : 0.f, &s_acc[warp][r * 9 + 1 + i]);
        if (MAPS && cons_out && col_ok) {
#pragma unroll
          for (int l = 0; l < 4; ++l)
            if (v0 + l < a.v) cons_out[((size_t)i * a.b + row) * a.v + v0 + l] = c[l];
        }
      }
      float lp = 0.f;
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float rr = xs[l] - xr[l];
        if (v0 + l < a.v) lp += fmaf(-0.5f * rr * rr, ww[l], ee[l] - HALF_LOG_2PI);
      }
      warp_accumulate(col_ok ? lp : 0.f, &s_acc[warp][r * 9]);
      if (MAPS && xrec_out && col_ok) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (v0 + l < a.v) xrec_out[(size_t)row * a.v + v0 + l] = xr[l];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ROWS * 9; i += RL_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < RL_THREADS / 32; ++w) s += s_acc[w][i];
    partial[((size_t)rg * a.col_chunks + blockIdx.x) * ROWS * 9 + i] = s;
  }
}

__global__ void recon_fwd_finalize(const float* __restrict__ partial, int b, int col_chunks, float* logp,
                                   float* norms) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // over b*9
  if (t >= b * 9) return;
  const int row = t / 9, slot = t % 9;
  const int rg = row / ROWS, r = row % ROWS;
  double s = 0.0;
  for (int c = 0; c < col_chunks; ++c) s += partial[(((size_t)rg * col_chunks + c) * ROWS + r) * 9 + slot];
  if (slot == 0) logp[row] = (float)s;
  else norms[(slot - 1) * b + row] = (float)sqrt(s);
}

// partial layout: [rowgroup][chunk][ROWS][8]  (dg partial sums)
__global__ void __launch_bounds__(RL_THREADS, 3)
recon_bwd_kernel(const ReconArgs a, const float* __restrict__ norms, float lam, float* __restrict__ dpre,
                 float* __restrict__ partial, float* __restrict__ deps) {
  __shared__ float s_g[ROWS][KCOV], s_cf[ROWS][KCOV];     // gain and lam*B/norm
  __shared__ float s_acc[RL_THREADS / 32][ROWS * KCOV];
  const int rg = blockIdx.y;
  const int row0 = rg * ROWS;
  const int warp = threadIdx.x >> 5;
  const size_t map_stride = (size_t)a.b * a.vp;
  const float invB = 1.f / (float)a.b;
  if (threadIdx.x < ROWS * KCOV) {
    const int r = threadIdx.x / KCOV, i = threadIdx.x % KCOV;
    const bool ok = row0 + r < a.b;
    s_g[r][i] = ok ? __ldg(a.g + i * a.b + row0 + r) : 0.f;
    const float nn = ok ? __ldg(norms + i * a.b + row0 + r) : 0.f;
    s_cf[r][i] = nn > 0.f ? lam * (float)a.b / nn : 0.f;
  }
  for (int i = threadIdx.x; i < (RL_THREADS / 32) * ROWS * KCOV; i += RL_THREADS) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();

  const int ncols_pad = (a.ncols + 31) & ~31;
  for (int col = blockIdx.x * RL_THREADS + threadIdx.x; col < ncols_pad; col += a.col_chunks * RL_THREADS) {
    const bool col_ok = col < a.ncols;
    const int v0 = (col_ok ? col : 0) * 4;
    const float4 e4 = ld4_keep(a.eps + v0);
    const float ww[4] = {__expf(2.f * e4.x), __expf(2.f * e4.y), __expf(2.f * e4.z), __expf(2.f * e4.w)};
    float de[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int r = 0; r < ROWS; ++r) {
      const int row = row0 + r;
      if (row >= a.b) break;
      float4 m4[NMAP];
#pragma unroll
      for (int j = 0; j < NMAP; ++j) m4[j] = ld4(a.maps + j * map_stride + (size_t)row * a.vp + v0);
      float xs[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) xs[l] = (v0 + l < a.v) ? __ldg(a.x + (size_t)row * a.v + v0 + l) : 0.f;
      float mm[NMAP][4];
#pragma unroll
      for (int j = 0; j < NMAP; ++j) {
        mm[j][0] = m4[j].x; mm[j][1] = m4[j].y; mm[j][2] = m4[j].z; mm[j][3] = m4[j].w;
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (!(v0 + l < a.v)) mm[j][l] = 0.f;      // row padding may hold anything (even NaN)
      }
      float xr[4] = {mm[0][0], mm[0][1], mm[0][2], mm[0][3]};
#pragma unroll
      for (int i = 0; i < KCOV; ++i)
#pragma unroll
        for (int l = 0; l < 4; ++l) xr[l] = fmaf(s_g[r][i], mm[i + 1][l], xr[l]);
      float A[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float rr = xs[l] - xr[l];
        const bool ok = v0 + l < a.v;
        A[l] = ok ? -rr * ww[l] * invB : 0.f;
        if (ok) de[l] += (rr * rr * ww[l] - 1.f) * invB;
      }
      float o[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) o[l] = A[l] * mm[0][l] * (1.f - mm[0][l]);
      if (col_ok)
        stg_stream(reinterpret_cast<float4*>(dpre + (size_t)row * a.vp + v0), make_float4(o[0], o[1], o[2], o[3]));
#pragma unroll
      for (int i = 0; i < KCOV; ++i) {
        const float4 G = ld4_keep(a.glm + (size_t)i * a.vp + v0);     // L1-resident after the first row
        const float Gs[4] = {G.x, G.y, G.z, G.w};
        const float gi = s_g[r][i], cfi = s_cf[r][i];
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const float D = mm[i + 1][l];
          float dc = A[l] + cfi * (gi * D - Gs[l]);     // d tot / d cons_i
          if (!(v0 + l < a.v)) dc = 0.f;
          s = fmaf(D, dc, s);
          o[l] = gi * dc * D * (1.f - D);
        }
        warp_accumulate(col_ok ? s : 0.f, &s_acc[warp][r * KCOV + i]);
        if (col_ok)
          stg_stream(reinterpret_cast<float4*>(dpre + (i + 1) * map_stride + (size_t)row * a.vp + v0),
                     make_float4(o[0], o[1], o[2], o[3]));
      }
    }
    if (col_ok) {
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (v0 + l < a.v) atomicAdd(deps + v0 + l, de[l]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ROWS * KCOV; i += RL_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < RL_THREADS / 32; ++w) s += s_acc[w][i];
    partial[((size_t)rg * a.col_chunks + blockIdx.x) * ROWS * KCOV + i] = s;
  }
}

__global__ void recon_bwd_finalize(const float* __restrict__ partial, int b, int col_chunks, float* dg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // over b*8
  if (t >= b * KCOV) return;
  const int row = t / KCOV, i = t % KCOV;
  const int rg = row / ROWS, r = row % ROWS;
  double s = 0.0;
  for (int c = 0; c < col_chunks; ++c) s += partial[(((size_t)rg * col_chunks + c) * ROWS + r) * KCOV + i];
  dg[i * b + row] = (float)s;
}

static int pick_chunks(int ncols, int rowgroups) {
  // aim for >= 4 CTAs per SM in total, at least one column per thread
  const int max_chunks = cdiv(ncols, RL_THREADS);
  int want = cdiv(4LL * vg_sm_count(), rowgroups);
  if (want < 1) want = 1;
  return want < max_chunks ? want : max_chunks;
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_recon_workspace_bytes(int b, long long v) {
  const int vp = (int)((v + 3) / 4 * 4);
  const int rgs = (b + ROWS - 1) / ROWS;
  const size_t max_chunks = (size_t)cdiv(vp / 4, RL_THREADS);
  return rgs * max_chunks * ROWS * 9 * sizeof(float) + 256;
}

extern "C" int vg_recon_loss_fwd(const float* maps, const float* g, const float* x, const float* eps,
                                 const float* glm, int b, long long v, float* logp, float* norms, float* cons,
                                 float* x_rec, void* workspace, size_t workspace_bytes, void* stream) {
  VG_CHECK_ARG(maps && g && x && eps && glm && logp && norms && b > 0 && v > 0, "bad arguments");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_recon_workspace_bytes(b, v), "workspace too small");
  ReconArgs a{};
  a.maps = maps; a.g = g; a.x = x; a.eps = eps; a.glm = glm;
  a.b = b; a.v = (int)v; a.vp = (int)((v + 3) / 4 * 4); a.ncols = a.vp / 4;
  const int rgs = (b + ROWS - 1) / ROWS;
  a.col_chunks = pick_chunks(a.ncols, rgs);
  cudaStream_t st = as_stream(stream);
  if (cons || x_rec) recon_fwd_kernel<true><<<dim3(a.col_chunks, rgs), RL_THREADS, 0, st>>>(a, (float*)workspace, cons, x_rec);
  else recon_fwd_kernel<false><<<dim3(a.col_chunks, rgs), RL_THREADS, 0, st>>>(a, (float*)workspace, nullptr, nullptr);
  VG_LAUNCH_CHECK();
  recon_fwd_finalize<<<cdiv(b * 9, 128), 128, 0, st>>>((const float*)workspace, b, a.col_chunks, logp, norms);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_recon_loss_bwd(const float* maps, const float* g, const float* x, const float* eps,
                                 const float* glm, const float* norms, int b, long long v, float lam,
                                 float* dpre, float* dg, float* deps, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  VG_CHECK_ARG(maps && g && x && eps && glm && norms && dpre && dg && deps && b > 0 && v > 0, "bad arguments");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_recon_workspace_bytes(b, v), "workspace too small");
  ReconArgs a{};
  a.maps = maps; a.g = g; a.x = x; a.eps = eps; a.glm = glm;
  a.b = b; a.v = (int)v; a.vp = (int)((v + 3) / 4 * 4); a.ncols = a.vp / 4;
  const int rgs = (b + ROWS - 1) / ROWS;
  a.col_chunks = pick_chunks(a.ncols, rgs);
  cudaStream_t st = as_stream(stream);
  VG_CUDA(cudaMemsetAsync(deps, 0, (size_t)a.vp * sizeof(float), st));
  recon_bwd_kernel<<<dim3(a.col_chunks, rgs), RL_THREADS, 0, st>>>(a, norms, lam, dpre, (float*)workspace, deps);
  VG_LAUNCH_CHECK();
  recon_bwd_finalize<<<cdiv(b * KCOV, 128), 128, 0, st>>>((const float*)workspace, b, a.col_chunks, dg);
  VG_LAUNCH_CHECK();
  return VG_OK;
}
