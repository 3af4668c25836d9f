// Weight gradients of Conv3d / ConvTranspose3d (autograd of vae_reg_GP.py:238-242, 260-264),
// shared-memory tiled.
//
//   dW[t][ci][co] = sum_{n, q} X[n, q*s + k_t (mode 0) | q (mode 1)][ci] * Y[n, q (mode 0) | q*s + k_t - p (mode 1)][co]
//
// mode 0 = Conv3d (q over the OUTPUT grid, the input is read at the tap-shifted position),
// mode 1 = ConvTranspose3d (q over the INPUT grid, dy is read at the tap-shifted position; no
// parity phases are needed in this formulation).  A CTA stages one box of q-voxels: the
// un-shifted operand tile and the shifted operand's halo box (zero-filled outside the grid, BN
// fold applied to in-range input voxels while staging), then every warp lane owns one
// (tap, ci-group) pair and walks the voxels of the tile keeping a CIL x COUT outer-product
// accumulator in registers; the un-shifted operand is a warp-wide broadcast from shared memory.
// CTAs are persistent over tiles, accumulators are flushed once per CTA (shared-memory reduce
// over the voxel-phase warps, then one global atomic per weight).
#include "common.cuh"

namespace vg {

struct WgGeom {
  int N, group_size;
  int mode, s;
  int kD, kH, kW, pD, pH, pW;
  int bD, bH, bW;        // base grid (q)
  int xD, xH, xW;        // module input grid
  int yD, yH, yW;        // module output grid
  int tD, tH, tW;        // tile (q voxels)
  int nTd, nTh, nTw;
  long long x_img, y_img;
  int wst_t, wst_ci, wst_co;
  int ntaps;
};

template <int C>
__device__ __forceinline__ void lds_vec(const float* p, float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      const float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = p[i];
  }
}

// stage a box of voxels of a channels-last tensor into shared memory, zero-filled outside
template <int C>
__device__ __forceinline__ void stage_box(float* dst, const float* __restrict__ src, int oD, int oH, int oW,
                                          int bd, int bh, int bw, int gD, int gH, int gW,
                                          const float* sc, const float* sh) {
  const int nvox = bd * bh * bw;
  constexpr int W = (C % 4 == 0) ? 4 : 1;
  constexpr int PER = C / W;
  for (int e = threadIdx.x; e < nvox * PER; e += blockDim.x) {
    const int v = e / PER, part = e - v * PER;
    const int l = v % bw, r = v / bw;
    const int j = r % bh, i = r / bh;
    const int gd = oD + i, gh = oH + j, gw = oW + l;
    const bool ok = gd >= 0 && gd < gD && gh >= 0 && gh < gH && gw >= 0 && gw < gW;
    if constexpr (W == 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        t = __ldg(reinterpret_cast<const float4*>(src + (((size_t)gd * gH + gh) * gW + gw) * C) + part);
        if (sc) {
          const int c0 = part * 4;
          t.x = fmaf(t.x, sc[c0], sh[c0]); t.y = fmaf(t.y, sc[c0 + 1], sh[c0 + 1]);
          t.z = fmaf(t.z, sc[c0 + 2], sh[c0 + 2]); t.w = fmaf(t.w, sc[c0 + 3], sh[c0 + 3]);
        }
      }
      reinterpret_cast<float4*>(dst + (size_t)v * C)[part] = t;
    } else {
      float t = 0.f;
      if (ok) {
        t = __ldg(src + (((size_t)gd * gH + gh) * gW + gw) * C + part);
        if (sc) t = fmaf(t, sc[part], sh[part]);
      }
      dst[(size_t)v * C + part] = t;
    }
  }
}

template <int CIN, int COUT, int CIL, int MODE>
__global__ void __launch_bounds__(256)
wgrad_tiled_kernel(const __grid_constant__ WgGeom g, const float* __restrict__ x, const float* __restrict__ dy,
                   const float* __restrict__ in_scale, const float* __restrict__ in_shift, float* dw) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CG = CIN / CIL;                // ci groups per tap
  const int pairs = g.ntaps * CG;
  const int pgroups = (pairs + 31) / 32;       // warps needed to cover all (tap, ci-group) pairs
  const int nwarps = blockDim.x >> 5;
  const int phases = nwarps / pgroups;         // voxel phases
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pg = warp % pgroups, phase = warp / pgroups;
  const int pair = pg * 32 + lane;
  const bool live = pair < pairs && phase < phases;
  const int tap = live ? pair / CG : 0;
  const int cg = live ? pair - tap * CG : 0;
  const int ka = tap / (g.kH * g.kW), kb = (tap / g.kW) % g.kH, kc = tap % g.kW;

  // shifted-operand box
  const int sD = (g.tD - 1) * g.s + g.kD, sH = (g.tH - 1) * g.s + g.kH, sW = (g.tW - 1) * g.s + g.kW;
  const int tile_vox = g.tD * g.tH * g.tW, box_vox = sD * sH * sW;
  float* Xs = smem;                                                         // mode 0: box, mode 1: tile
  float* Ys = smem + (((size_t)(MODE == 0 ? box_vox : tile_vox) * CIN + 3) & ~(size_t)3);      // mode 0: tile, mode 1: box

  float acc[CIL][COUT];
#pragma unroll
  for (int i = 0; i < CIL; ++i)
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[i][j] = 0.f;

  const int tiles_per_img = g.nTd * g.nTh * g.nTw;
  const long long ntiles = (long long)tiles_per_img * g.N;
  const int shift_base = (ka * sH + kb) * sW + kc;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = (int)(tile / tiles_per_img);
    int tr = (int)(tile - (long long)n * tiles_per_img);
    const int tw = tr % g.nTw; tr /= g.nTw;
    const int th = tr % g.nTh;
    const int td = tr / g.nTh;
    const int q0d = td * g.tD, q0h = th * g.tH, q0w = tw * g.tW;
    const float* xn = x + (size_t)n * g.x_img;
    const float* yn = dy + (size_t)n * g.y_img;
    const float* sc = in_scale ? in_scale + (n / g.group_size) * CIN : nullptr;
    const float* sh = in_scale ? in_shift + (n / g.group_size) * CIN : nullptr;
    __syncthreads();   // previous tile fully consumed
    if (MODE == 0) {
      stage_box<CIN>(Xs, xn, q0d * g.s, q0h * g.s, q0w * g.s, sD, sH, sW, g.xD, g.xH, g.xW, sc, sh);
      stage_box<COUT>(Ys, yn, q0d, q0h, q0w, g.tD, g.tH, g.tW, g.yD, g.yH, g.yW, nullptr, nullptr);
    } else {
      stage_box<CIN>(Xs, xn, q0d, q0h, q0w, g.tD, g.tH, g.tW, g.xD, g.xH, g.xW, sc, sh);
      stage_box<COUT>(Ys, yn, q0d * g.s - g.pD, q0h * g.s - g.pH, q0w * g.s - g.pW, sD, sH, sW, g.yD, g.yH, g.yW,
                      nullptr, nullptr);
    }
    __syncthreads();
    if (!live) continue;
    const int rows = g.tD * g.tH;
    for (int r = phase; r < rows; r += phases) {
      const int i = r / g.tH, j = r - i * g.tH;
      int sh_idx = shift_base + g.s * ((i * sH + j) * sW);     // shifted operand, voxel (i, j, 0)
      int un_idx = r * g.tW;                                   // un-shifted operand
      for (int l = 0; l < g.tW; ++l, sh_idx += g.s, ++un_idx) {
        float xv[CIL], yv[COUT];
        if (MODE == 0) {
          lds_vec<CIL>(Xs + (size_t)sh_idx * CIN + cg * CIL, xv);
          lds_vec<COUT>(Ys + (size_t)un_idx * COUT, yv);
        } else {
          lds_vec<CIL>(Xs + (size_t)un_idx * CIN + cg * CIL, xv);
          lds_vec<COUT>(Ys + (size_t)sh_idx * COUT, yv);
        }
#pragma unroll
        for (int a = 0; a < CIL; ++a)
#pragma unroll
          for (int b = 0; b < COUT; ++b) acc[a][b] = fmaf(xv[a], yv[b], acc[a][b]);
      }
    }
  }
  // ---- flush: reduce the voxel-phase warps through shared memory, then one global atomic per weight
  __syncthreads();
  float* red = smem;                                  // [pairs][CIL*COUT]
  const int per_pair = CIL * COUT;
  for (int e = threadIdx.x; e < pairs * per_pair; e += blockDim.x) red[e] = 0.f;
  __syncthreads();
  if (live) {
#pragma unroll
    for (int a = 0; a < CIL; ++a)
#pragma unroll
      for (int b = 0; b < COUT; ++b) atomicAdd(&red[(size_t)pair * per_pair + a * COUT + b], acc[a][b]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < pairs * per_pair; e += blockDim.x) {
    const int p = e / per_pair, rem = e - p * per_pair;
    const int t = p / CG, c = (p - t * CG) * CIL + rem / COUT, co = rem % COUT;
    atomicAdd(dw + (size_t)t * g.wst_t + (size_t)c * g.wst_ci + (size_t)co * g.wst_co, red[e]);
  }
}

// dbias[c] += sum over all voxels of dy[.., c]
template <int C>
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ dy, long long spatial, long long img_stride, float* dbias) {
  __shared__ float sred[C];
  if (threadIdx.x < C) sred[threadIdx.x] = 0.f;
  __syncthreads();
  const float* p = dy + (size_t)blockIdx.y * img_stride;
  const long long total = spatial * C;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if constexpr (C % 4 == 0) {
    const long long nv = total / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = ldg_stream(reinterpret_cast<const float4*>(p) + i);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
    const int c0 = (int)((((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4) % C);
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&sred[c0 + j], s[j]);
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
      s[0] += __ldg(p + i);
    const float r = warp_sum(s[0]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sred[0], r);
  }
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(dbias + threadIdx.x, sred[threadIdx.x]);
}

static void pick_tile(const VgConvDesc* d, int cin, int cout, WgGeom& g) {
  // tile of q voxels: keep the staged floats under ~44 KB so two CTAs fit comfortably and no
  // opt-in is needed; W extent first (inner loop), then H, then D.
  const int budget = 11000;   // floats
  int best[3] = {1, 1, 1};
  long long best_score = -1;
  const int candD[] = {1, 2, 3, 4, 6, 8}, candH[] = {1, 2, 4, 6, 8, 10}, candW[] = {4, 6, 7, 8, 14, 16, 17, 33, 35};
  for (int a : candD)
    for (int b : candH)
      for (int c : candW) {
        if (a > g.bD || b > g.bH || c > g.bW) continue;
        const long long tile = (long long)a * b * c;
        const long long box = (long long)((a - 1) * g.s + g.kD) * ((b - 1) * g.s + g.kH) * ((c - 1) * g.s + g.kW);
        const long long fl = g.mode == 0 ? box * cin + tile * cout : tile * cin + box * cout;
        if (fl > budget) continue;
        // useful fraction of the padded tile grid, then prefer large tiles (halo amortisation)
        const long long nT = (long long)((g.bD + a - 1) / a) * ((g.bH + b - 1) / b) * ((g.bW + c - 1) / c);
        const double eff = (double)g.bD * g.bH * g.bW / (double)(nT * tile);
        const long long score = (long long)(eff * eff * (double)tile * 1000.0 / (double)fl * 1000.0);
        if (score > best_score) { best_score = score; best[0] = a; best[1] = b; best[2] = c; }
      }
  g.tD = best[0]; g.tH = best[1]; g.tW = best[2];
  g.nTd = (g.bD + g.tD - 1) / g.tD; g.nTh = (g.bH + g.tH - 1) / g.tH; g.nTw = (g.bW + g.tW - 1) / g.tW;
}

template <int CIN, int COUT, int CIL>
static int launch_wgrad_tiled(const WgGeom& g, const float* x, const float* dy, const float* sc, const float* sh,
                              float* dw, cudaStream_t st) {
  const int pairs = g.ntaps * (CIN / CIL);
  const int pgroups = (pairs + 31) / 32;
  int nwarps = 8;
  if (pgroups > 8) nwarps = pgroups;
  nwarps = nwarps / pgroups * pgroups;
  const long long tile = (long long)g.tD * g.tH * g.tW;
  const long long box = (long long)((g.tD - 1) * g.s + g.kD) * ((g.tH - 1) * g.s + g.kH) * ((g.tW - 1) * g.s + g.kW);
  size_t fl = g.mode == 0 ? box * CIN + tile * COUT : tile * CIN + box * COUT;
  const size_t red = (size_t)pairs * CIL * COUT;
  if (red > fl) fl = red;
  const size_t smem = fl * sizeof(float) + 32;
  const long long ntiles = (long long)g.nTd * g.nTh * g.nTw * g.N;
  long long blocks = 2LL * vg_sm_count();
  if (blocks > ntiles) blocks = ntiles;
  if (g.mode == 0) {
    if (smem > 48 * 1024)
      VG_CUDA(cudaFuncSetAttribute(wgrad_tiled_kernel<CIN, COUT, CIL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tiled_kernel<CIN, COUT, CIL, 0><<<(unsigned)blocks, nwarps * 32, smem, st>>>(g, x, dy, sc, sh, dw);
  } else {
    if (smem > 48 * 1024)
      VG_CUDA(cudaFuncSetAttribute(wgrad_tiled_kernel<CIN, COUT, CIL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tiled_kernel<CIN, COUT, CIL, 1><<<(unsigned)blocks, nwarps * 32, smem, st>>>(g, x, dy, sc, sh, dw);
  }
  VG_LAUNCH_CHECK();
  return VG_OK;
}

}  // namespace vg

using namespace vg;

// Declared in conv.cu's public entry point; kept separate so the two kernels build in parallel.
int vg_conv_wgrad_tiled(const VgConvDesc* d, const float* x, const float* dy, const float* in_scale,
                        const float* in_shift, float* dw, float* dbias, cudaStream_t st) {
  WgGeom g{};
  g.N = d->n; g.group_size = d->group_size;
  g.mode = d->transposed ? 1 : 0;
  g.s = d->stride;
  g.kD = d->k[0]; g.kH = d->k[1]; g.kW = d->k[2];
  g.pD = d->pad[0]; g.pH = d->pad[1]; g.pW = d->pad[2];
  const int* base = d->transposed ? d->in : d->out;
  g.bD = base[0]; g.bH = base[1]; g.bW = base[2];
  g.xD = d->in[0]; g.xH = d->in[1]; g.xW = d->in[2];
  g.yD = d->out[0]; g.yH = d->out[1]; g.yW = d->out[2];
  g.x_img = d->x_img_stride ? d->x_img_stride : (long long)d->in[0] * d->in[1] * d->in[2] * d->cin;
  g.y_img = d->y_img_stride ? d->y_img_stride : (long long)d->out[0] * d->out[1] * d->out[2] * d->cout;
  const int K = g.kD * g.kH * g.kW;
  g.ntaps = K;
  g.wst_t = 1;
  if (!d->transposed) { g.wst_ci = K; g.wst_co = d->cin * K; }        // w[co][ci][K]
  else                { g.wst_ci = d->cout * K; g.wst_co = K; }        // w[ci][co][K]
  pick_tile(d, d->cin, d->cout, g);
  int rc = VG_EINVAL;
  const int ci = d->cin, co = d->cout;
  if (!dw) rc = VG_OK;                      // weight gradient already produced by the tensor-core kernel
  else if (ci == 1 && co == 8) rc = launch_wgrad_tiled<1, 8, 1>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 8 && co == 1) rc = launch_wgrad_tiled<8, 1, 8>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 8 && co == 8) rc = launch_wgrad_tiled<8, 8, 4>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 8 && co == 16) rc = launch_wgrad_tiled<8, 16, 4>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 16 && co == 8) rc = launch_wgrad_tiled<16, 8, 8>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 16 && co == 16) rc = launch_wgrad_tiled<16, 16, 4>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 1 && co == 1) rc = launch_wgrad_tiled<1, 1, 1>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 1 && co == 16) rc = launch_wgrad_tiled<1, 16, 1>(g, x, dy, in_scale, in_shift, dw, st);
  else if (ci == 16 && co == 1) rc = launch_wgrad_tiled<16, 1, 8>(g, x, dy, in_scale, in_shift, dw, st);
  else set_error("unsupported channel pair (%d,%d): channels must be in {1,8,16}", ci, co);
  if (rc != VG_OK) return rc;
  if (dbias) {
    const long long spatial = (long long)d->out[0] * d->out[1] * d->out[2];
    int bx = (int)((spatial * co / 4 + 255) / 256);
    int cap = (4 * vg_sm_count() + d->n - 1) / d->n;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, d->n);
    if (co == 1) bias_grad_kernel<1><<<grid, 256, 0, st>>>(dy, spatial, g.y_img, dbias);
    else if (co == 8) bias_grad_kernel<8><<<grid, 256, 0, st>>>(dy, spatial, g.y_img, dbias);
    else bias_grad_kernel<16><<<grid, 256, 0, st>>>(dy, spatial, g.y_img, dbias);
    VG_LAUNCH_CHECK();
  }
  return VG_OK;
}
