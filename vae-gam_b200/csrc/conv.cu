// fp32 direct 3-D convolution family for sm_100a: one "tap-list gather" formulation serves
// Conv3d / ConvTranspose3d forward, both data gradients and both weight gradients.
//
//   out[n, q*sout + r, co] = epi( sum_t sum_ci pro(in[n, q*sin + off_t, ci]) * W[t][ci][co] )
//
// * Conv3d forward  (vae_reg_GP.py:238-242): sin = stride, taps = all kernel offsets.
// * ConvTranspose3d forward (vae_reg_GP.py:260-264): gather form of SURVEY Appendix B; a
//   stride-2 layer is split into its 8 output-parity phases, each a stride-1 gather over the
//   taps whose parity matches (sout = 2, r = phase).
// * data gradients are the same two forms with the channel roles of W swapped.
// * weight gradients correlate the (affine-folded) input with dy over the same geometry.
// Layout: activations channels-last (N,D,H,W,C) so one voxel's channels are one or more
// 128-bit loads and a warp's loads are contiguous; weights are read straight from the
// PyTorch layouts through (tap, ci, co) strides and staged once per CTA in shared memory.
// BatchNorm (batch statistics, per image group) never runs as its own pass: `pro` applies
// scale/shift to in-range taps only (zero padding pads the NORMALISED tensor) and `epi`
// accumulates the next layer's statistics, or the BatchNorm-backward sums.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_geom.cuh"

namespace vg {

template <int C>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = __ldg(p + i);
  }
}
template <int C>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i) p[i] = v[i];
  }
}

// ROW: unit input stride and the taps are a complete 3x3x3 cube sorted by (dd, dh, dw) — a thread then loads each
// (dd, dh) input row segment ONCE (VT + 2 voxels) and applies the three dw taps to it, instead of VT loads per tap.
template <int CIN, int COUT, int VT, bool ROW = false>
__global__ void __launch_bounds__(256)
gather_kernel(const __grid_constant__ Geom g, const GatherArgs a) {
  extern __shared__ float sw[];  // [ntaps][CIN][COUT]
  __shared__ double sred[2 * COUT];
  const int tid = threadIdx.x;
  {
    const int per_tap = CIN * COUT;
    const int nw = g.ntaps * per_tap;
    for (int i = tid; i < nw; i += blockDim.x) {
      int t = i / per_tap, rem = i - t * per_tap;
      int ci = rem / COUT, co = rem - ci * COUT;
      sw[i] = __ldg(a.w + (size_t)g.taps[t].widx * g.wst_t + (size_t)ci * g.wst_ci + (size_t)co * g.wst_co);
    }
    if (tid < 2 * COUT) sred[tid] = 0.0;
  }
  __syncthreads();

  const int n = blockIdx.y;
  const int grp = n / g.group_size;
  const int nWc = (g.qW + VT - 1) / VT;
  const int items = g.qD * g.qH * nWc;
  const int item = blockIdx.x * blockDim.x + tid;
  const bool active = item < items;
  const int qw0 = (item % nWc) * VT;
  const int qh = (item / nWc) % g.qH;
  const int qd = item / (nWc * g.qH);

  float acc[VT][COUT];
#pragma unroll
  for (int v = 0; v < VT; ++v)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[v][c] = 0.f;

  if (active) {
    float sc[CIN], sh[CIN];
    const bool affine = a.in_scale != nullptr;
    if (affine) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        sc[c] = __ldg(a.in_scale + grp * CIN + c);
        sh[c] = __ldg(a.in_shift + grp * CIN + c);
      }
    }
    const float* in_n = a.in + (size_t)n * g.in_img;
    if constexpr (ROW) {
      const int lo_d = g.taps[0].dd, lo_h = g.taps[0].dh, lo_w = g.taps[0].dw;       // tap t = (a * 3 + b) * 3 + c
      for (int ab = 0; ab < 9; ++ab) {
        const int id = qd + lo_d + ab / 3, ih = qh + lo_h + ab % 3;
        if (id < 0 || id >= g.inD || ih < 0 || ih >= g.inH) continue;
        const float* rowp = in_n + ((size_t)id * g.inH + ih) * g.inW * CIN;
        float xr[VT + 2][CIN];
#pragma unroll
        for (int v = 0; v < VT + 2; ++v) {
          const int iw = qw0 + v + lo_w;
          if (iw >= 0 && iw < g.inW) {
            load_vec<CIN>(rowp + (size_t)iw * CIN, xr[v]);
            if (affine) {
#pragma unroll
              for (int c = 0; c < CIN; ++c) xr[v][c] = fmaf(xr[v][c], sc[c], sh[c]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < CIN; ++c) xr[v][c] = 0.f;
          }
        }
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          const float* wt = sw + (ab * 3 + c3) * (CIN * COUT);
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            float wr[COUT];
            if constexpr (COUT % 4 == 0) {
#pragma unroll
              for (int i = 0; i < COUT / 4; ++i) {
                float4 q4 = reinterpret_cast<const float4*>(wt + ci * COUT)[i];
                wr[4 * i] = q4.x; wr[4 * i + 1] = q4.y; wr[4 * i + 2] = q4.z; wr[4 * i + 3] = q4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < COUT; ++i) wr[i] = wt[ci * COUT + i];
            }
#pragma unroll
            for (int v = 0; v < VT; ++v)
#pragma unroll
              for (int co = 0; co < COUT; ++co) acc[v][co] = fmaf(xr[v + c3][ci], wr[co], acc[v][co]);
          }
        }
      }
    } else
    for (int t = 0; t < g.ntaps; ++t) {
      const Tap tp = g.taps[t];
      const int id = qd * g.sin + tp.dd;
      const int ih = qh * g.sin + tp.dh;
      if (g.check && (id < 0 || id >= g.inD || ih < 0 || ih >= g.inH)) continue;
      const size_t row = ((size_t)id * g.inH + ih) * g.inW;
      float xv[VT][CIN];
#pragma unroll
      for (int v = 0; v < VT; ++v) {
        const int iw = (qw0 + v) * g.sin + tp.dw;
        const bool ok = (qw0 + v < g.qW) && (!g.check || (iw >= 0 && iw < g.inW));
        if (ok) {
          load_vec<CIN>(in_n + (row + iw) * CIN, xv[v]);
          if (affine) {
#pragma unroll
            for (int c = 0; c < CIN; ++c) xv[v][c] = fmaf(xv[v][c], sc[c], sh[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < CIN; ++c) xv[v][c] = 0.f;
        }
      }
      const float* wt = sw + t * (CIN * COUT);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float wr[COUT];
        if constexpr (COUT % 4 == 0) {
#pragma unroll
          for (int i = 0; i < COUT / 4; ++i) {
            float4 q4 = reinterpret_cast<const float4*>(wt + ci * COUT)[i];
            wr[4 * i] = q4.x; wr[4 * i + 1] = q4.y; wr[4 * i + 2] = q4.z; wr[4 * i + 3] = q4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < COUT; ++i) wr[i] = wt[ci * COUT + i];
        }
#pragma unroll
        for (int v = 0; v < VT; ++v)
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[v][co] = fmaf(xv[v][ci], wr[co], acc[v][co]);
      }
    }
  }

  // ---- epilogue
  const bool want_stats = a.stats != nullptr;
  const bool want_bn = a.aux_mode == 2;
  float s1[COUT], s2[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) s1[c] = s2[c] = 0.f;
  if (active) {
    float bias[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias[c] = a.bias ? __ldg(a.bias + c) : 0.f;
    float istd[COUT], mistd[COUT];
    if (want_bn) {
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        istd[c] = __ldg(a.aux_istd + grp * COUT + c);
        mistd[c] = __ldg(a.aux_mistd + grp * COUT + c);
      }
    }
    const int od = qd * g.sout + g.rD, oh = qh * g.sout + g.rH;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
      if (qw0 + v >= g.qW) continue;
      const int ow = (qw0 + v) * g.sout + g.rW;
      const size_t o = (size_t)n * g.out_img + (((size_t)od * g.outH + oh) * g.outW + ow) * COUT;
      float y[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        float t = acc[v][c] + bias[c];
        if (a.act == VG_ACT_RELU) t = fmaxf(t, 0.f);
        else if (a.act == VG_ACT_SIGMOID) t = 1.f / (1.f + __expf(-t));
        y[c] = t;
      }
      if (a.aux_mode != 0) {
        float ax[COUT];
        load_vec<COUT>(a.aux + o, ax);
        if (a.aux_mode == 1) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) y[c] = ax[c] > 0.f ? y[c] : 0.f;
        } else {
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            const float xh = fmaf(ax[c], istd[c], -mistd[c]);
            s1[c] += y[c];
            s2[c] = fmaf(y[c], xh, s2[c]);
          }
        }
      }
      if (want_stats) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          s1[c] += y[c];
          s2[c] = fmaf(y[c], y[c], s2[c]);
        }
      }
      if (a.out) store_vec<COUT>(a.out + o, y);
    }
  }
  if (want_stats || want_bn) {
    const int lane = tid & 31;
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      const float r1 = warp_sum(s1[c]);
      const float r2 = warp_sum(s2[c]);
      if (lane == 0) {
        atomicAdd(&sred[2 * c], (double)r1);
        atomicAdd(&sred[2 * c + 1], (double)r2);
      }
    }
    __syncthreads();
    if (tid < 2 * COUT) {
      double* dst = want_stats ? a.stats : a.aux_sums;
      atomicAdd(dst + (size_t)grp * COUT * 2 + tid, sred[tid]);
    }
  }
}

// --------------------------------------------------------------------------- geometry
static int check_desc(const VgConvDesc* d) {
  VG_CHECK_ARG(d != nullptr, "null descriptor");
  VG_CHECK_ARG(d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
  VG_CHECK_ARG(d->n > 0 && d->group_size > 0 && d->n % d->group_size == 0, "n % group_size != 0");
  for (int i = 0; i < 3; ++i) {
    VG_CHECK_ARG(d->k[i] >= 1 && d->k[i] <= 5, "kernel extent must be 1..5");
    int expect = d->transposed
                     ? (d->in[i] - 1) * d->stride - 2 * d->pad[i] + d->k[i] + d->opad[i]
                     : (d->in[i] - d->k[i]) / d->stride + 1;
    VG_CHECK_ARG(expect == d->out[i], "output size does not match the PyTorch formula");
    VG_CHECK_ARG(d->transposed || d->pad[i] == 0, "Conv3d padding is not supported (reference uses none)");
  }
  return VG_OK;
}

// kind 0 reads x and writes y; kind 1 reads dy and writes dx
static void set_img_strides(const VgConvDesc* d, int kind, Geom& g) {
  const long long xs = d->x_img_stride ? d->x_img_stride : (long long)d->in[0] * d->in[1] * d->in[2] * d->cin;
  const long long ys = d->y_img_stride ? d->y_img_stride : (long long)d->out[0] * d->out[1] * d->out[2] * d->cout;
  g.in_img = kind == 0 ? xs : ys;
  g.out_img = kind == 0 ? ys : xs;
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// kind 0: module forward; kind 1: data gradient.  Returns number of phase geometries.
static int build_geoms(const VgConvDesc* d, int kind, Geom* out /* up to 8 */) {
  const int K = d->k[0] * d->k[1] * d->k[2];
  const int s = d->stride;
  int ng = 0;
  const bool gather_plain = (kind == 0 && !d->transposed) || (kind == 1 && d->transposed);
  Geom base{};
  base.N = d->n;
  base.group_size = d->group_size;
  // weight strides for (tap, ci_eff, co_eff)
  if (!d->transposed) {  // w[co][ci][K]
    if (kind == 0) { base.wst_ci = K; base.wst_co = d->cin * K; }
    else           { base.wst_ci = d->cin * K; base.wst_co = K; }   // ci_eff = co, co_eff = ci
  } else {               // w[ci][co][K]
    if (kind == 0) { base.wst_ci = d->cout * K; base.wst_co = K; }
    else           { base.wst_ci = K; base.wst_co = d->cout * K; }  // ci_eff = co, co_eff = ci
  }
  base.wst_t = 1;
  if (gather_plain) {
    // q runs over the SMALL grid; input index = q*s + (k - p)
    Geom g = base;
    const int* small = d->transposed ? d->in : d->out;   // output of this gather
    const int* large = d->transposed ? d->out : d->in;   // tensor being read
    g.inD = large[0]; g.inH = large[1]; g.inW = large[2];
    g.outD = small[0]; g.outH = small[1]; g.outW = small[2];
    g.qD = small[0]; g.qH = small[1]; g.qW = small[2];
    g.sin = s; g.sout = 1; g.rD = g.rH = g.rW = 0;
    g.check = d->transposed ? 1 : 0;
    set_img_strides(d, kind, g);
    int t = 0;
    for (int a = 0; a < d->k[0]; ++a)
      for (int b = 0; b < d->k[1]; ++b)
        for (int c = 0; c < d->k[2]; ++c) {
          g.taps[t].dd = (int8_t)(a - d->pad[0]);
          g.taps[t].dh = (int8_t)(b - d->pad[1]);
          g.taps[t].dw = (int8_t)(c - d->pad[2]);
          g.taps[t].widx = (a * d->k[1] + b) * d->k[2] + c;
          ++t;
        }
    g.ntaps = t;
    out[ng++] = g;
  } else {
    // q runs over one parity phase of the LARGE grid; input index = q + (r + p - k)/s
    const int* small = d->transposed ? d->in : d->out;   // tensor being read
    const int* large = d->transposed ? d->out : d->in;   // output of this gather
    for (int r0 = 0; r0 < s; ++r0)
      for (int r1 = 0; r1 < s; ++r1)
        for (int r2 = 0; r2 < s; ++r2) {
          const int r[3] = {r0, r1, r2};
          Geom g = base;
          g.inD = small[0]; g.inH = small[1]; g.inW = small[2];
          g.outD = large[0]; g.outH = large[1]; g.outW = large[2];
          g.sin = 1; g.sout = s; g.rD = r0; g.rH = r1; g.rW = r2;
          g.check = 1;
          set_img_strides(d, kind, g);
          int q[3];
          bool empty = false;
          for (int i = 0; i < 3; ++i) {
            q[i] = large[i] > r[i] ? (large[i] - r[i] + s - 1) / s : 0;
            if (q[i] == 0) empty = true;
          }
          if (empty) continue;
          g.qD = q[0]; g.qH = q[1]; g.qW = q[2];
          int t = 0;
          for (int a = 0; a < d->k[0]; ++a) {
            if (((r0 + d->pad[0] - a) % s + s) % s != 0) continue;
            for (int b = 0; b < d->k[1]; ++b) {
              if (((r1 + d->pad[1] - b) % s + s) % s != 0) continue;
              for (int c = 0; c < d->k[2]; ++c) {
                if (((r2 + d->pad[2] - c) % s + s) % s != 0) continue;
                g.taps[t].dd = (int8_t)floordiv(r0 + d->pad[0] - a, s);
                g.taps[t].dh = (int8_t)floordiv(r1 + d->pad[1] - b, s);
                g.taps[t].dw = (int8_t)floordiv(r2 + d->pad[2] - c, s);
                g.taps[t].widx = (a * d->k[1] + b) * d->k[2] + c;
                ++t;
              }
            }
          }
          g.ntaps = t;
          out[ng++] = g;
        }
  }
  return ng;
}

template <int CIN, int COUT, int VT, bool ROW = false>
static int launch_gather_t(const Geom& g, const GatherArgs& a, cudaStream_t st) {
  const int nWc = (g.qW + VT - 1) / VT;
  const long long items = (long long)g.qD * g.qH * nWc;
  int threads = items >= 256 ? 256 : (int)((items + 31) / 32 * 32);
  if (threads < 32) threads = 32;
  dim3 grid(cdiv(items, threads), g.N);
  const size_t smem = (size_t)g.ntaps * CIN * COUT * sizeof(float);
  gather_kernel<CIN, COUT, VT, ROW><<<grid, threads, smem, st>>>(g, a);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

// unit input stride and a complete 3x3x3 tap cube: `sorted` = g with its taps ordered by (dd, dh, dw) (row-reuse kernel)
static bool full_cube(const Geom& g, Geom& sorted) {
  if (g.sin != 1 || g.ntaps != 27) return false;
  sorted = g;
  for (int i = 1; i < 27; ++i) {
    Tap t = sorted.taps[i];
    int j = i - 1;
    auto key = [](const Tap& x) { return (x.dd + 64) * 16384 + (x.dh + 64) * 128 + (x.dw + 64); };
    while (j >= 0 && key(sorted.taps[j]) > key(t)) { sorted.taps[j + 1] = sorted.taps[j]; --j; }
    sorted.taps[j + 1] = t;
  }
  const Tap c0 = sorted.taps[0];
  for (int t = 0; t < 27; ++t)
    if (sorted.taps[t].dd != c0.dd + t / 9 || sorted.taps[t].dh != c0.dh + (t / 3) % 3 || sorted.taps[t].dw != c0.dw + t % 3) return false;
  return true;
}

// Process default of the arithmetic mode (0 fp32 CUDA cores, 1 bf16 tcgen05, 2 mixed: bf16 tcgen05 with the
// encoder's forward convolutions in fp32).  Only consulted when a descriptor / step config says VG_ARITH_DEFAULT.
static int g_conv_mode = -1;
int conv_mode() {
  if (g_conv_mode < 0) {
    const char* e = getenv("VAEGAM_CONV_MODE");
    if (e && (e[0] == '0' || e[0] == 'f')) g_conv_mode = 0;        // "0" / "fp32" -> CUDA cores
    else if (e && (e[0] == '1' || e[0] == 'b')) g_conv_mode = 1;   // "1" / "bf16" -> tensor cores everywhere
    else g_conv_mode = 2;
  }
  return g_conv_mode;
}
// VG_ARITH_* of a descriptor / step config -> 0 (fp32), 1 (bf16 tensor cores) or 2 (mixed, step level only)
int resolve_arith(int arith) {
  if (arith == VG_ARITH_FP32) return 0;
  if (arith == VG_ARITH_BF16) return 1;
  if (arith == VG_ARITH_MIXED) return 2;
  return conv_mode();
}
static inline bool desc_tc(const VgConvDesc* d) { return resolve_arith(d->arith) != 0; }

// The plane-folded kernel is persistent with a per-CTA set-up (weight blocks, TMEM): worth it from a few
// hundred thousand output voxels per launch
static long long g_t2_min_voxels = -1;
static bool tc2_worthwhile(const Geom* gs, int ng) {
  long long vox = 0;
  for (int i = 0; i < ng; ++i) vox += (long long)gs[i].N * gs[i].qD * gs[i].qH * gs[i].qW;
  if (g_t2_min_voxels < 0) { const char* e = getenv("VAEGAM_T2_MIN_VOXELS"); g_t2_min_voxels = e ? atoll(e) : 400000; }   // process default
  return vox >= g_t2_min_voxels;
}

// bf16 storage is implemented by the plane-folded tensor-core kernel only
static bool wants_bf16(const GatherArgs& a) { return a.in_bf16 || a.out_bf16 || a.aux_bf16; }
static int check_bf16(int cin, int cout, const GatherArgs& a) {
  VG_CHECK_ARG(!a.in_bf16 || cin == 8, "bf16 input storage needs 8 input channels");
  VG_CHECK_ARG(!(a.out_bf16 || a.aux_bf16) || cout % 8 == 0, "bf16 output storage needs 8 or 16 output channels");
  return VG_OK;
}
static int no_bf16_path() {
  set_error("bf16 activation storage (VgConvDesc.bf16_mask) is only served by the plane-folded tensor-core kernel; "
            "this geometry / arithmetic falls outside it");
  return VG_EINVAL;
}

// strided 16-channel gathers (convt2's data gradient, conv4's forward) have no other tensor-core kernel: the
// plane-folded one serves them whatever their size
static bool tc2_only(int cin, const Geom& g) { return g.sin == 2 && cin == 16; }

static int launch_gather(bool tc, int cin, int cout, const Geom& g, const GatherArgs& a, cudaStream_t st) {
  if (tc && (tc2_worthwhile(&g, 1) || wants_bf16(a) || tc2_only(cin, g)) && tc2_supported(cin, cout, &g, 1))
    return launch_tc2_gather(cin, cout, &g, 1, a, st);
  if (wants_bf16(a)) return no_bf16_path();
  if (tc && tc_supported(cin, cout, g)) return launch_tc_gather(cin, cout, g, a, st);
  {
    Geom gs;                            // stride-1 3x3x3 layers with few input channels: each input row loaded once per thread
    if ((cin == 1 && cout == 8) || (cin == 8 && cout == 16)) {
      if (full_cube(g, gs)) return cin == 1 ? launch_gather_t<1, 8, 6, true>(gs, a, st) : launch_gather_t<8, 16, 4, true>(gs, a, st);
    }
  }
  if (cin == 1 && cout == 8) return launch_gather_t<1, 8, 4>(g, a, st);
  if (cin == 8 && cout == 1) return launch_gather_t<8, 1, 4>(g, a, st);
  if (cin == 8 && cout == 8) return launch_gather_t<8, 8, 4>(g, a, st);
  if (cin == 8 && cout == 16) return launch_gather_t<8, 16, 4>(g, a, st);
  if (cin == 16 && cout == 8) return launch_gather_t<16, 8, 4>(g, a, st);
  if (cin == 16 && cout == 16) return launch_gather_t<16, 16, 2>(g, a, st);
  if (cin == 1 && cout == 1) return launch_gather_t<1, 1, 4>(g, a, st);
  if (cin == 1 && cout == 16) return launch_gather_t<1, 16, 4>(g, a, st);
  if (cin == 16 && cout == 1) return launch_gather_t<16, 1, 4>(g, a, st);
  set_error("unsupported channel pair (%d,%d): channels must be in {1,8,16}", cin, cout);
  return VG_EINVAL;
}

// Helper streams for the output-parity phases of a small stride-2 layer: the phases write disjoint
// outputs (statistics are accumulated atomically), each is a latency-bound launch of a few CTAs, so they
// run side by side: fork from the caller's stream with one event, join back with one event per phase.
// Stream-capture safe (event record / wait only); VAEGAM_PHASE_STREAMS=0 serialises them.
constexpr int kPhaseStreams = 7;
struct PhaseStreams {
  cudaStream_t s[kPhaseStreams];
  cudaEvent_t fork, join[kPhaseStreams];
  bool ok = false;
};
// One set per host thread and device: entry points called from different threads or for different devices
// never share a stream or an event.
static PhaseStreams& phase_streams() {
  constexpr int kMaxDev = 16;
  static thread_local PhaseStreams per_dev[kMaxDev];
  static thread_local bool tried_dev[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { (void)cudaGetLastError(); dev = 0; }
  PhaseStreams& ps = per_dev[dev];
  bool& tried = tried_dev[dev];
  if (!tried) {
    tried = true;
    const char* e = getenv("VAEGAM_PHASE_STREAMS");
    if (!(e && e[0] == '0')) {
      ps.ok = cudaEventCreateWithFlags(&ps.fork, cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < kPhaseStreams; ++i)
        ps.ok = ps.ok && cudaStreamCreateWithFlags(&ps.s[i], cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&ps.join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ps.ok) (void)cudaGetLastError();
  }
  return ps;
}

// every gather of one layer pass: one fused multi-phase launch when the plane-folded kernel covers it
static int launch_all(bool tc, int cin, int cout, const Geom* gs, int ng, const GatherArgs& a, cudaStream_t st) {
  if (wants_bf16(a)) VG_TRY(check_bf16(cin, cout, a));
  if (ng > 1 && tc && (tc2_worthwhile(gs, ng) || wants_bf16(a)) && tc2_supported(cin, cout, gs, ng))
    return launch_tc2_gather(cin, cout, gs, ng, a, st);
  if (ng > 1 && wants_bf16(a)) return no_bf16_path();
  PhaseStreams& ps = phase_streams();
  if (ng < 2 || ng > kPhaseStreams + 1 || !ps.ok) {
    for (int i = 0; i < ng; ++i) VG_TRY(launch_gather(tc, cin, cout, gs[i], a, st));
    return VG_OK;
  }
  VG_CUDA(cudaEventRecord(ps.fork, st));
  int rc = launch_gather(tc, cin, cout, gs[0], a, st);
  for (int i = 1; i < ng; ++i) {
    cudaStream_t hs = ps.s[i - 1];
    VG_CUDA(cudaStreamWaitEvent(hs, ps.fork, 0));
    if (rc == VG_OK) rc = launch_gather(tc, cin, cout, gs[i], a, hs);
    VG_CUDA(cudaEventRecord(ps.join[i - 1], hs));          // always joined, also after a failed launch
    VG_CUDA(cudaStreamWaitEvent(st, ps.join[i - 1], 0));
  }
  return rc;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_set_conv_mode(int mode) {
  VG_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0 (fp32 CUDA cores), 1 (bf16 tcgen05) or 2 (mixed)");
  g_conv_mode = mode;
  return VG_OK;
}
extern "C" int vg_get_conv_mode(void) { return conv_mode(); }
extern "C" int vg_set_conv_tuning(const char* key, long long value) {
  VG_CHECK_ARG(key != nullptr, "null key");
  if (strcmp(key, "t2_min_voxels") == 0) { VG_CHECK_ARG(value >= 0, "negative threshold"); g_t2_min_voxels = value; return VG_OK; }
  set_error("unknown tuning key '%s'", key);
  return VG_EINVAL;
}

// Which kernel serves each gather of a layer (kind 0 forward, 1 data gradient) in the current
// convolution mode; one line per launch.  Host only, no device work.
extern "C" int vg_conv_describe(const VgConvDesc* d, int kind, char* buf, size_t cap) {
  VG_TRY(check_desc(d));
  VG_CHECK_ARG(buf && cap > 0 && (kind == 0 || kind == 1), "bad arguments");
  Geom gs[8];
  const int ng = build_geoms(d, kind, gs);
  const int cin = kind == 0 ? d->cin : d->cout, cout = kind == 0 ? d->cout : d->cin;
  const bool use_tc = desc_tc(d);
  const bool in16 = (d->bf16_mask & (kind == 0 ? VG_BF16_X : VG_BF16_Y)) != 0;      // the tensor this gather reads
  const bool any16 = d->bf16_mask != 0;
  size_t off = 0;
  buf[0] = 0;
  if (ng > 1 && use_tc && (tc2_worthwhile(gs, ng) || any16) && tc2_supported(cin, cout, gs, ng)) {
    const int n = tc2_describe(cin, cout, gs, ng, buf, cap, in16);
    if (n > 0 && (size_t)n + 2 < cap) { buf[n] = '\n'; buf[n + 1] = 0; }
    return 1;
  }
  for (int i = 0; i < ng && off + 8 < cap; ++i) {
    int n = 0;
    if (use_tc && (tc2_worthwhile(&gs[i], 1) || any16 || tc2_only(cin, gs[i]))) n = tc2_describe(cin, cout, &gs[i], 1, buf + off, cap - off, in16);
    if (n <= 0) {
      const bool tc = use_tc && tc_supported(cin, cout, gs[i]);
      n = snprintf(buf + off, cap - off, "%s cin=%d cout=%d q=(%d,%d,%d) taps=%d", tc ? "tc1" : "fp32", cin, cout,
                   gs[i].qD, gs[i].qH, gs[i].qW, gs[i].ntaps);
    }
    off += (size_t)n < cap - off ? (size_t)n : cap - off - 1;
    if (off + 2 < cap) { buf[off++] = '\n'; buf[off] = 0; }
  }
  return ng;
}

extern "C" int vg_conv_fwd(const VgConvDesc* d, const void* x, const float* w, const float* bias,
                           const float* in_scale, const float* in_shift, void* y, int act,
                           double* out_stats, void* stream) {
  VG_TRY(check_desc(d));
  VG_CHECK_ARG(x && w && y, "null tensor");
  Geom gs[8];
  const int ng = build_geoms(d, 0, gs);
  GatherArgs a{};
  a.in = static_cast<const float*>(x); a.w = w; a.bias = bias; a.in_scale = in_scale; a.in_shift = in_shift;
  a.out = static_cast<float*>(y); a.act = act; a.stats = out_stats;
  a.in_bf16 = (d->bf16_mask & VG_BF16_X) != 0; a.out_bf16 = (d->bf16_mask & VG_BF16_Y) != 0;
  return launch_all(desc_tc(d), d->cin, d->cout, gs, ng, a, as_stream(stream));
}

extern "C" int vg_conv_dgrad(const VgConvDesc* d, const void* dy_, const float* w, void* dx_,
                             const void* mask_act_, const void* bn_x_, const float* bn_istd,
                             const float* bn_mistd, double* bn_sums, void* stream) {
  const float* dy = static_cast<const float*>(dy_);
  float* dx = static_cast<float*>(dx_);
  const float* mask_act = static_cast<const float*>(mask_act_);
  const float* bn_x = static_cast<const float*>(bn_x_);
  VG_TRY(check_desc(d));
  VG_CHECK_ARG(dy && w, "null tensor");
  VG_CHECK_ARG(!(mask_act && bn_x), "mask_act and bn_x are exclusive");
  VG_CHECK_ARG(!bn_x || (bn_istd && bn_mistd && bn_sums), "bn_x needs istd/mistd/sums");
  Geom gs[8];
  const int ng = build_geoms(d, 1, gs);
  GatherArgs a{};
  a.in = dy; a.w = w; a.out = dx; a.act = VG_ACT_NONE;
  a.in_bf16 = (d->bf16_mask & VG_BF16_Y) != 0; a.out_bf16 = (d->bf16_mask & VG_BF16_DX) != 0;
  a.aux_bf16 = (d->bf16_mask & VG_BF16_X) != 0 && (mask_act || bn_x);
  if (mask_act) { a.aux = mask_act; a.aux_mode = 1; }
  if (bn_x) { a.aux = bn_x; a.aux_mode = 2; a.aux_istd = bn_istd; a.aux_mistd = bn_mistd; a.aux_sums = bn_sums; }
  // the gather reads dy (cout channels) and produces cin channels
  return launch_all(desc_tc(d), d->cout, d->cin, gs, ng, a, as_stream(stream));
}

namespace vg {
// wgrad_mma.cu: bf16 mma.sync weight gradient; VG_OK, a negative error, or 1 = channel pair not covered
int wgrad_mma(const VgConvDesc* d, const void* x, const void* dy, const float* in_scale, const float* in_shift,
              float* dw, float* dbias, cudaStream_t st, float* dw_group = nullptr);
}
int vg_conv_wgrad_tiled(const VgConvDesc* d, const float* x, const float* dy, const float* in_scale,
                        const float* in_shift, float* dw, float* dbias, cudaStream_t st);   // wgrad.cu

extern "C" int vg_conv_wgrad(const VgConvDesc* d, const void* x, const void* dy, const float* in_scale,
                             const float* in_shift, float* dw, float* dbias, void* stream) {
  VG_TRY(check_desc(d));
  VG_CHECK_ARG(x && dy && dw, "null tensor");
  if (desc_tc(d)) {
    const int rc = wgrad_mma(d, x, dy, in_scale, in_shift, dw, dbias, as_stream(stream));   // bias gradient fused
    if (rc <= 0) return rc;
  }
  if (d->bf16_mask & (VG_BF16_X | VG_BF16_Y)) return no_bf16_path();
  return vg_conv_wgrad_tiled(d, static_cast<const float*>(x), static_cast<const float*>(dy), in_scale, in_shift, dw, dbias,
                             as_stream(stream));
}

// ---- fused BatchNorm-backward junction (BatchNorm -> stride-1 ConvTranspose3d), see include/vaegam.h
extern "C" int vg_conv_wgrad_grouped(const VgConvDesc* d, const void* x, const void* dy, float* raw, void* stream) {
  VG_TRY(check_desc(d));
  VG_CHECK_ARG(x && dy && raw, "null tensor");
  VG_CHECK_ARG(desc_tc(d), "grouped weight gradients exist in the tensor-core arithmetic only");
  const int rc = wgrad_mma(d, x, dy, nullptr, nullptr, raw /* unused */, nullptr, as_stream(stream), raw);
  if (rc > 0) { set_error("vg_conv_wgrad_grouped: channel pair (%d,%d) is not covered", d->cin, d->cout); return VG_EINVAL; }
  return rc;
}

extern "C" int vg_conv_dgrad_bn_apply(const VgConvDesc* d, const void* dy_, const float* w, void* dx_, const void* bn_x_,
                                      const float* coef, float* dx_chan_sum, void* stream) {
  VG_TRY(check_desc(d));
  VG_CHECK_ARG(dy_ && w && dx_ && bn_x_ && coef, "null tensor");
  Geom gs[8];
  const int ng = build_geoms(d, 1, gs);
  GatherArgs a{};
  a.in = static_cast<const float*>(dy_); a.w = w; a.out = static_cast<float*>(dx_); a.act = VG_ACT_NONE;
  a.in_bf16 = (d->bf16_mask & VG_BF16_Y) != 0; a.out_bf16 = (d->bf16_mask & VG_BF16_DX) != 0;
  a.aux_bf16 = (d->bf16_mask & VG_BF16_X) != 0;
  a.aux = static_cast<const float*>(bn_x_); a.aux_mode = 3; a.aux_coef = coef; a.chan_sum = dx_chan_sum;
  // only the plane-folded tensor-core kernel implements this epilogue
  const int cin = d->cout, cout = d->cin;
  if (a.in_bf16 || a.out_bf16 || a.aux_bf16) VG_TRY(check_bf16(cin, cout, a));
  if (!desc_tc(d) || !tc2_supported(cin, cout, gs, ng)) {
    set_error("vg_conv_dgrad_bn_apply: the fused BatchNorm-backward epilogue needs the plane-folded tensor-core kernel");
    return VG_EINVAL;
  }
  return launch_tc2_gather(cin, cout, gs, ng, a, as_stream(stream));
}
