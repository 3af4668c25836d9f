// Tensor-core weight gradients of Conv3d / ConvTranspose3d (autograd of vae_reg_GP.py:238-242,
// 260-264) in bf16 with fp32 accumulation (convolution mode 1).
//
//   dW[t][ci][co] = sum_{n, q} X[n, q*s + k_t (mode 0) | q (mode 1)][ci] * dY[n, q (mode 0) | q*s + k_t - p (mode 1)][co]
//
// The result is a few thousand numbers reduced over tens of millions of voxels, so the GEMM is
// D[(tap, shifted channel)][unshifted channel] with K = voxels: M <= 360, N <= 16.  A tcgen05
// formulation would have to keep one TMEM accumulator per (dh,dw) shift (> 512 columns for the
// stride-2 layers) and re-read both operands from shared memory for every 16 voxels, so this
// kernel uses warp-level mma.sync.m16n8k16 with register accumulators instead: each warp owns a
// private copy of the whole (or half the) weight gradient, walks 16-voxel chunks of the CTA's
// tile, and gets its fragments with ldmatrix.trans straight from channels-last bf16 tiles — the
// tap shift and the stride are just per-lane row addresses.  CTAs are persistent over tiles;
// accumulators are flushed once per CTA (shared-memory reduce, then one global atomic per weight).
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace vg {

struct WmGeom {
  int N, group_size;
  int mode, s;
  int kD, kH, kW, pD, pH, pW;
  int bD, bH, bW;        // base grid (q)
  int xD, xH, xW;        // module input grid
  int yD, yH, yW;        // module output grid
  int tD, tH, tW;        // tile (q voxels)
  int nTd, nTh, nTw;
  int sD, sH, sW;        // shifted-operand box of a tile
  long long x_img, y_img;
  int wst_t, wst_ci, wst_co;
  int ntaps, tg, taps_per_group;
  uint32_t mul_tH, mul_tW, mul_sH, mul_sW;   // ceil(2^32 / d) for the staging index decomposition (0: d == 1)
  int x_bf16, y_bf16;    // x / dy stored as bf16 (8 or 16 channels): staged by a straight 16-byte copy when nothing is folded
  int un_tma, sh_tma;    // that copy is ONE cp.async.bulk.tensor box per tile (tensor maps in the kernel parameters)
  int nbuf;              // operand buffers in shared memory: 2 with TMA staging (the next tile's boxes travel during the MMAs)
};

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// exact e / d for e * d < 2^32 with mul = ceil(2^32 / d) (d >= 2); d == 1 is passed as mul = 0
__device__ __forceinline__ int fast_div(int e, uint32_t mul) { return mul ? (int)__umulhi((uint32_t)e, mul) : e; }

// Stage a box of voxels of a channels-last fp32 tensor as bf16 (BatchNorm fold on in-range voxels, zeros
// outside).  Flat element index (full lane utilisation for any box shape, divisions by multiply-high), four
// independent elements per thread in flight before the first is converted.  BIAS: also sum the raw values
// of the voxels this tile OWNS (own*: global coordinate ranges) per channel — a thread always handles the
// same 8-channel part (blockDim is a multiple of the parts per voxel), so the sums live in 8 registers.
template <int C, bool BIAS, int UV>
__device__ __forceinline__ void stage_box_bf16(__nv_bfloat16* dst, const float* __restrict__ src, int oD, int oH, int oW,
                                               int bh, int bw, int nvox, uint32_t mul_h, uint32_t mul_w, int gD, int gH,
                                               int gW, const float* sc, const float* sh, const int (&own)[6],
                                               float (&bs)[8], bool src16 = false, __nv_bfloat16* dst1 = nullptr) {
  constexpr int PER = C >= 8 ? C / 8 : 1;        // 16-byte bf16 chunks per voxel
  // elements in flight per thread (registers are shared with the accumulators)
  constexpr int U = UV;
  const int nel = nvox * PER;
  for (int e0 = threadIdx.x; e0 < nel; e0 += U * blockDim.x) {
    float4 va[U], vb[U];
    float v1[U];
    int state[U];                                // 0: beyond the box, 1: zero fill, 2: loaded, 3: loaded + owned
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * blockDim.x;
      state[u] = 0;
      if (e >= nel) continue;
      const int v = e / PER, part = e & (PER - 1);
      const int r = fast_div(v, mul_w), l = v - r * bw;
      const int i = fast_div(r, mul_h), j = r - i * bh;
      const int gd = oD + i, gh = oH + j, gw = oW + l;
      state[u] = 1;
      if (gd >= 0 && gd < gD && gh >= 0 && gh < gH && gw >= 0 && gw < gW) {
        state[u] = 2;
        if (BIAS && gd >= own[0] && gd < own[1] && gh >= own[2] && gh < own[3] && gw >= own[4] && gw < own[5]) state[u] = 3;
        const size_t off = (((size_t)gd * gH + gh) * gW + gw) * C;
        if constexpr (C >= 8) {
          if (src16) {
            const uint4 q = ldg_u4(reinterpret_cast<const __nv_bfloat16*>(src) + off + part * 8);
            va[u] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
          } else {
            const float4* p = reinterpret_cast<const float4*>(src + off) + part * 2;
            va[u] = __ldg(p);
            vb[u] = __ldg(p + 1);
          }
        } else {
          v1[u] = __ldg(src + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (state[u] == 0) continue;
      const int e = e0 + u * blockDim.x;
      if constexpr (C >= 8) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (state[u] >= 2 && src16 && !(BIAS && state[u] == 3) && !sc) {
          pk = make_uint4(__float_as_uint(va[u].x), __float_as_uint(va[u].y), __float_as_uint(va[u].z), __float_as_uint(va[u].w));
        } else if (state[u] >= 2) {
          float4 a = va[u], b = vb[u];
          if (src16) {
            float f[8];
            unpack_bf16x8(make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)), f);
            a = make_float4(f[0], f[1], f[2], f[3]);
            b = make_float4(f[4], f[5], f[6], f[7]);
          }
          if (BIAS && state[u] == 3) {
            bs[0] += a.x; bs[1] += a.y; bs[2] += a.z; bs[3] += a.w;
            bs[4] += b.x; bs[5] += b.y; bs[6] += b.z; bs[7] += b.w;
          }
          if (sc) {
            const int c0 = (e & (PER - 1)) * 8;
            a.x = fmaf(a.x, sc[c0], sh[c0]); a.y = fmaf(a.y, sc[c0 + 1], sh[c0 + 1]);
            a.z = fmaf(a.z, sc[c0 + 2], sh[c0 + 2]); a.w = fmaf(a.w, sc[c0 + 3], sh[c0 + 3]);
            b.x = fmaf(b.x, sc[c0 + 4], sh[c0 + 4]); b.y = fmaf(b.y, sc[c0 + 5], sh[c0 + 5]);
            b.z = fmaf(b.z, sc[c0 + 6], sh[c0 + 6]); b.w = fmaf(b.w, sc[c0 + 7], sh[c0 + 7]);
          }
          __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x, a.y), t1 = __floats2bfloat162_rn(a.z, a.w);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(b.x, b.y), t3 = __floats2bfloat162_rn(b.z, b.w);
          pk = make_uint4(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1),
                          *reinterpret_cast<uint32_t*>(&t2), *reinterpret_cast<uint32_t*>(&t3));
        }
        reinterpret_cast<uint4*>(dst)[e] = pk;
      } else {
        float t = 0.f;
        if (state[u] >= 2) {
          t = v1[u];
          if (BIAS && state[u] == 3) bs[0] += t;
          if (sc) t = fmaf(t, sc[0], sh[0]);
        }
        const __nv_bfloat16 tb = __float2bfloat16(t);
        dst[e] = tb;
        if (dst1 && e > 0) dst1[e - 1] = tb;       // second copy, one element ahead: odd pair addresses become aligned words
      }
    }
  }
}

// Same box, but the source is ALREADY bf16 and nothing has to be folded or summed: every 16-byte voxel part goes
// global -> shared with one cp.async (LDGSTS, zero-filled outside the tensor), no register round trip, and all
// copies of a tile are in flight together.  The caller waits with cp_async_wait_all() before its __syncthreads().
template <int C>
__device__ __forceinline__ void stage_box_async(__nv_bfloat16* dst, const __nv_bfloat16* __restrict__ src, int oD, int oH,
                                                int oW, int bh, int bw, int nvox, uint32_t mul_h, uint32_t mul_w, int gD,
                                                int gH, int gW) {
  static_assert(C % 8 == 0, "16-byte voxel parts");
  constexpr int PER = C / 8;
  const int nel = nvox * PER;
  const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(dst);
  for (int e = threadIdx.x; e < nel; e += blockDim.x) {
    const int v = e / PER, part = e & (PER - 1);
    const int r = fast_div(v, mul_w), l = v - r * bw;
    const int i = fast_div(r, mul_h), j = r - i * bh;
    const int gd = oD + i, gh = oH + j, gw = oW + l;
    const bool ok = gd >= 0 && gd < gD && gh >= 0 && gh < gH && gw >= 0 && gw < gW;
    const __nv_bfloat16* p = ok ? src + ((((size_t)gd * gH + gh) * gW + gw) * C + part * 8) : src;
    const uint32_t nbytes = ok ? 16u : 0u;            // 0 source bytes: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)e * 16u), "l"(p), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// CS: channels of the tap-shifted operand (1, 8, 16); CU: channels of the un-shifted operand (8, 16);
// MT: m16 tiles of (tap, shifted channel) rows a warp accumulates.
template <int CS, int CU, int MT, int MODE>
__global__ void __launch_bounds__(256, 2)
wgrad_mma_kernel(const __grid_constant__ WmGeom g, const float* __restrict__ x, const float* __restrict__ dy,
                 const float* __restrict__ in_scale, const float* __restrict__ in_shift, float* dw, float* dbias,
                 float* dw_group, const __grid_constant__ CUtensorMap tm_un, const __grid_constant__ CUtensorMap tm_sh) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_tma_bar[2];
  __shared__ int s_tapoff[48];
  __shared__ float s_bias[16];
  constexpr int NT = CU / 8;                       // n8 tiles
  constexpr int UV = MT * NT * 4 <= 16 ? 4 : 2;    // staged voxels in flight per thread: more when the accumulators are few
  constexpr int TPM = CS == 8 ? 2 : (CS == 16 ? 1 : 16);   // taps per m16 tile
  // one shifted channel: a box is ~14 elements per thread; the transposed layer (convt5) keeps all of them in flight at
  // once — each further round is one more exposed global-load latency per tile (the forward layer's variant would spill)
  constexpr int USH = CS >= 8 ? UV : (MODE == 1 ? 16 : 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvw = (blockDim.x >> 5) / g.tg;        // warps sharing a tap group split the voxel chunks
  const int tgi = warp % g.tg, vwi = warp / g.tg;
  const int tap_lo = tgi * g.taps_per_group;
  const int tap_hi = min(g.ntaps, tap_lo + g.taps_per_group);
  const int tile_vox = g.tD * g.tH * g.tW, box_vox = g.sD * g.sH * g.sW;
  const int nchunks = (tile_vox + 15) >> 4;
  const size_t sh_one = (((size_t)box_vox * CS * 2 + 64) + 127) & ~(size_t)127;           // + slack for clamped taps; TMA destinations are 128-byte aligned
  // One shifted channel: a second copy of the box, one element ahead, so that the (voxel, voxel + 1) pair of an A
  // fragment register is ONE aligned 32-bit word whatever the parity of the tap offset (pair_ok below).
  const size_t sh_bytes = CS == 1 ? 2 * sh_one : sh_one;
  const size_t un_bytes = ((size_t)nchunks * 16 * CU * 2 + 127) & ~(size_t)127;           // un-shifted operand: tile (+ pad chunk)
  // Operand buffers [shifted box | un-shifted tile] x nbuf: with TMA staging the boxes of tile i + 1 travel while
  // tile i is multiplied (two buffers, one mbarrier each).  Then the voxel-offset table and, with two buffers, the
  // reduction scratch of flush() (with one buffer it reuses the operand memory).
  const int nbuf = g.nbuf;
  const size_t buf_stride = sh_bytes + un_bytes;
  // box offset of every tile voxel (the tile geometry is the same for all tiles): no divisions in the chunk loop
  unsigned short* s_vo = reinterpret_cast<unsigned short*>(smem_raw + (size_t)nbuf * buf_stride);
  float* const red_scratch = nbuf == 2 ? reinterpret_cast<float*>(smem_raw + (size_t)nbuf * buf_stride + (((size_t)nchunks * 32 + 127) & ~(size_t)127))
                                       : reinterpret_cast<float*>(smem_raw);
  for (int v = threadIdx.x; v < nchunks * 16; v += blockDim.x) {
    // padded rows of the last chunk meet zero B rows; they keep the parity of their index (pair loads need even offsets)
    const int vv = v < tile_vox ? v : max(0, tile_vox - 2 + (v & 1) - (tile_vox & 1));
    const int l = vv % g.tW, rr = vv / g.tW;
    const int j = rr % g.tH, i = rr / g.tH;
    s_vo[v] = (unsigned short)(((i * g.s) * g.sH + j * g.s) * g.sW + l * g.s);
  }
  if (threadIdx.x < 16) s_bias[threadIdx.x] = 0.f;
  const bool any_tma = g.un_tma || g.sh_tma;
  if (any_tma && threadIdx.x == 0) {
    mbar_init(smem_u32(&s_tma_bar[0]), 1);
    mbar_init(smem_u32(&s_tma_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // zero the un-shifted rows of the last chunk's padding (tile_vox .. 16*nchunks) of every buffer: never overwritten
  for (int b = 0; b < nbuf; ++b)
    for (int e = threadIdx.x; e < (nchunks * 16 - tile_vox) * CU / 8; e += blockDim.x)
      reinterpret_cast<uint4*>(smem_raw + (size_t)b * buf_stride + sh_bytes + (size_t)tile_vox * CU * 2)[e] = make_uint4(0u, 0u, 0u, 0u);
  float bsum[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bsum[c] = 0.f;
  const uint32_t smem_addr0 = (uint32_t)__cvta_generic_to_shared(smem_raw);

  if (threadIdx.x < 48) {
    const int t = min((int)threadIdx.x, g.ntaps - 1);
    const int ka = t / (g.kH * g.kW), kb = (t / g.kW) % g.kH, kc = t % g.kW;
    s_tapoff[threadIdx.x] = (ka * g.sH + kb) * g.sW + kc;
  }

  // CS == 1 fast path (see the A fragments below): byte offset of a tap's pair word, per m16 tile and row half
  const bool pair_ok = CS == 1 && g.s == 1 && !(g.tW & 1) && !(g.sW & 1);
  uint32_t pair_off[CS == 1 ? MT : 1][2];
  if constexpr (CS == 1) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int t = min(tap_lo + mt * 16 + (lane >> 2) + hf * 8, max(tap_hi, 1) - 1);
        const int ka = t / (g.kH * g.kW), kb = (t / g.kW) % g.kH, kc = t % g.kW;
        const int off = (ka * g.sH + kb) * g.sW + kc;
        pair_off[mt][hf] = (off & 1) ? (uint32_t)sh_one + 2u * (uint32_t)(off - 1) : 2u * (uint32_t)off;
      }
  }

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  const int tiles_per_img = g.nTd * g.nTh * g.nTw;
  const long long ntiles = (long long)tiles_per_img * g.N;
  const int nred = g.ntaps * CS * CU;
  // reduce the warps' accumulators through shared memory (which holds no tile at that point), then one global
  // atomic per weight into `dst` laid out by (st_t, st_cs, st_cu)
  auto flush = [&](float* dst, int st_t, int st_cs, int st_cu) {
    float* red = red_scratch;                                   // [tap][shifted ch][un-shifted ch]
    for (int e = threadIdx.x; e < nred; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
    {
      const int gq = lane >> 2, q4 = lane & 3;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = mt * 16 + gq + (k >> 1) * 8;                // (tap, shifted channel) within the group
            const int colu = nt * 8 + 2 * q4 + (k & 1);                 // un-shifted channel
            int t, cs;
            if constexpr (CS == 1) { t = tap_lo + row; cs = 0; }
            else { t = tap_lo + row / CS; cs = row % CS; }
            if (t < tap_hi) atomicAdd(&red[((size_t)t * CS + cs) * CU + colu], acc[mt][nt][k]);
            acc[mt][nt][k] = 0.f;
          }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nred; e += blockDim.x) {
      const int colu = e % CU, rest = e / CU;
      const int cs = rest % CS, t = rest / CS;
      atomicAdd(dst + (size_t)t * st_t + (size_t)cs * st_cs + (size_t)colu * st_cu, red[e]);
    }
    __syncthreads();
  };
  // Grouped mode (dw_group != NULL): the products are accumulated PER BatchNorm statistics group, un-folded, into
  // dw_group[group][tap][cs][cu]; a CTA then walks a contiguous span of tiles (images in order), so it meets at
  // most a couple of groups and flushes when the group changes.
  const bool grouped = dw_group != nullptr;
  const long long span = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = grouped ? (long long)blockIdx.x * span : (long long)blockIdx.x;
  const long long t_end = grouped ? (t_begin + span < ntiles ? t_begin + span : ntiles) : ntiles;
  const long long t_step = grouped ? 1 : (long long)gridDim.x;
  int cur_grp = -1;
  // operands that are already bf16 with nothing to fold / sum go straight to shared memory: one TMA box per tile
  // (zero-filled outside the tensor) when the launch carries tensor maps, else one cp.async per voxel part
  const bool x16 = g.x_bf16 != 0, y16 = g.y_bf16 != 0;
  const bool x_async = x16 && !in_scale, y_async = y16 && !dbias;
  const bool skip_un = g.un_tma && (MODE == 0 ? y_async : x_async);                 // travels by TMA
  const bool skip_sh = CS >= 8 && g.sh_tma && (MODE == 0 ? x_async : y_async);
  const bool tma_wait = skip_un || skip_sh;
  auto tile_origin = [&](long long tile, int& n, int& td, int& th, int& tw) {
    n = (int)(tile / tiles_per_img);
    int tr = (int)(tile - (long long)n * tiles_per_img);
    tw = tr % g.nTw; tr /= g.nTw;
    th = tr % g.nTh;
    td = tr / g.nTh;
  };
  auto tma_issue = [&](long long tile, int buf) {          // thread 0 only
    int n, td, th, tw;
    tile_origin(tile, n, td, th, tw);
    const int q0d = td * g.tD, q0h = th * g.tH, q0w = tw * g.tW;
    const uint32_t bar = smem_u32(&s_tma_bar[buf]);
    const uint32_t sh_a = smem_addr0 + (uint32_t)((size_t)buf * buf_stride), un_a = sh_a + (uint32_t)sh_bytes;
    fence_async_smem();                        // earlier generic-proxy accesses of this buffer are ordered before the async writes
    mbar_arrive_expect_tx(bar, (skip_un ? (uint32_t)tile_vox * CU * 2u : 0u) + (skip_sh ? (uint32_t)box_vox * CS * 2u : 0u));
    // the maps' w coordinate counts 32-bit words: C / 2 per voxel
    if (skip_un) tma_load_4d(un_a, &tm_un, bar, q0w * (CU / 2), q0h, q0d, n);
    if (skip_sh) {
      if (MODE == 0) tma_load_4d(sh_a, &tm_sh, bar, q0w * g.s * (CS / 2), q0h * g.s, q0d * g.s, n);
      else tma_load_4d(sh_a, &tm_sh, bar, (q0w * g.s - g.pW) * (CS / 2), q0h * g.s - g.pH, q0d * g.s - g.pD, n);
    }
  };
  if (tma_wait && nbuf == 2 && threadIdx.x == 0 && t_begin < t_end) tma_issue(t_begin, 0);
  int tile_no = 0;
  for (long long tile = t_begin; tile < t_end; tile += t_step, ++tile_no) {
    int n, td, th, tw;
    tile_origin(tile, n, td, th, tw);
    const int q0d = td * g.tD, q0h = th * g.tH, q0w = tw * g.tW;
    const int buf = nbuf == 2 ? (tile_no & 1) : 0;
    __nv_bfloat16* const Ssh = reinterpret_cast<__nv_bfloat16*>(smem_raw + (size_t)buf * buf_stride);          // shifted operand: box
    __nv_bfloat16* const Ssh1 = CS == 1 ? reinterpret_cast<__nv_bfloat16*>(smem_raw + (size_t)buf * buf_stride + sh_one) : nullptr;
    __nv_bfloat16* const Sun = reinterpret_cast<__nv_bfloat16*>(smem_raw + (size_t)buf * buf_stride + sh_bytes);
    const uint32_t ssh_addr = smem_addr0 + (uint32_t)((size_t)buf * buf_stride), sun_addr = ssh_addr + (uint32_t)sh_bytes;
    // bf16 tensors have the same element offsets at half the element size
    const float* xn = x16 ? reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)n * g.x_img)
                          : x + (size_t)n * g.x_img;
    const float* yn = y16 ? reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)n * g.y_img)
                          : dy + (size_t)n * g.y_img;
    const float* sc = in_scale ? in_scale + (n / g.group_size) * (MODE == 0 ? CS : CU) : nullptr;
    const float* sh = in_scale ? in_shift + (n / g.group_size) * (MODE == 0 ? CS : CU) : nullptr;
    __syncthreads();   // previous tile fully consumed
    if (grouped) {
      const int grp = n / g.group_size;
      if (cur_grp >= 0 && grp != cur_grp) flush(dw_group + (size_t)cur_grp * nred, CS * CU, CU, 1);
      cur_grp = grp;
    }
    if (tma_wait && threadIdx.x == 0) {
      if (nbuf == 2) { if (tile + t_step < t_end) tma_issue(tile + t_step, buf ^ 1); }       // the NEXT tile, into the other buffer
      else tma_issue(tile, 0);
    }
    // un-shifted tile: clipped to the base grid (voxels of the tile beyond it contribute zeros)
    const int none[6] = {0, 0, 0, 0, 0, 0};
    if (MODE == 0) {
      // dY is the un-shifted tile: tiles partition the output grid
      const int own[6] = {q0d, q0d + g.tD, q0h, q0h + g.tH, q0w, q0w + g.tW};
      if constexpr (CS >= 8) {
        if (x_async && !skip_sh) stage_box_async<CS>(Ssh, reinterpret_cast<const __nv_bfloat16*>(xn), q0d * g.s, q0h * g.s, q0w * g.s, g.sH, g.sW,
                                         box_vox, g.mul_sH, g.mul_sW, g.xD, g.xH, g.xW);
      }
      if (!(CS >= 8 && x_async))
      stage_box_bf16<CS, false, USH>(Ssh, xn, q0d * g.s, q0h * g.s, q0w * g.s, g.sH, g.sW, box_vox, g.mul_sH, g.mul_sW, g.xD, g.xH,
                                g.xW, sc, sh, none, bsum, x16, Ssh1);
      if (skip_un) {
      } else if (y_async) stage_box_async<CU>(Sun, reinterpret_cast<const __nv_bfloat16*>(yn), q0d, q0h, q0w, g.tH, g.tW, tile_vox, g.mul_tH,
                                       g.mul_tW, g.yD, g.yH, g.yW);
      else if (dbias) stage_box_bf16<CU, true, UV>(Sun, yn, q0d, q0h, q0w, g.tH, g.tW, tile_vox, g.mul_tH, g.mul_tW, g.yD, g.yH, g.yW,
                                          nullptr, nullptr, own, bsum, y16);
      else stage_box_bf16<CU, false, UV>(Sun, yn, q0d, q0h, q0w, g.tH, g.tW, tile_vox, g.mul_tH, g.mul_tW, g.yD, g.yH, g.yW,
                                     nullptr, nullptr, none, bsum, y16);
    } else {
      // dY is the shifted box (boxes of neighbouring tiles overlap): a tile owns the outputs
      // [q0*s - p, (q0+t)*s - p), the first / last tile of a dimension also the grid's margins
      const int own[6] = {td == 0 ? 0 : q0d * g.s - g.pD, td == g.nTd - 1 ? g.yD : (q0d + g.tD) * g.s - g.pD,
                          th == 0 ? 0 : q0h * g.s - g.pH, th == g.nTh - 1 ? g.yH : (q0h + g.tH) * g.s - g.pH,
                          tw == 0 ? 0 : q0w * g.s - g.pW, tw == g.nTw - 1 ? g.yW : (q0w + g.tW) * g.s - g.pW};
      if (skip_un) {
      } else if (x_async) stage_box_async<CU>(Sun, reinterpret_cast<const __nv_bfloat16*>(xn), q0d, q0h, q0w, g.tH, g.tW, tile_vox, g.mul_tH,
                                       g.mul_tW, g.xD, g.xH, g.xW);
      else
      stage_box_bf16<CU, false, UV>(Sun, xn, q0d, q0h, q0w, g.tH, g.tW, tile_vox, g.mul_tH, g.mul_tW, g.xD, g.xH, g.xW, sc, sh,
                                none, bsum, x16);
      if constexpr (CS >= 8) {
        if (y_async && !skip_sh) stage_box_async<CS>(Ssh, reinterpret_cast<const __nv_bfloat16*>(yn), q0d * g.s - g.pD, q0h * g.s - g.pH,
                                         q0w * g.s - g.pW, g.sH, g.sW, box_vox, g.mul_sH, g.mul_sW, g.yD, g.yH, g.yW);
      }
      if (CS >= 8 && y_async) {
      } else if (dbias) stage_box_bf16<CS, true, USH>(Ssh, yn, q0d * g.s - g.pD, q0h * g.s - g.pH, q0w * g.s - g.pW, g.sH, g.sW, box_vox,
                                          g.mul_sH, g.mul_sW, g.yD, g.yH, g.yW, nullptr, nullptr, own, bsum, y16, Ssh1);
      else stage_box_bf16<CS, false, USH>(Ssh, yn, q0d * g.s - g.pD, q0h * g.s - g.pH, q0w * g.s - g.pW, g.sH, g.sW, box_vox,
                                     g.mul_sH, g.mul_sW, g.yD, g.yH, g.yW, nullptr, nullptr, none, bsum, y16, Ssh1);
    }
    cp_async_wait_all();
    if (tma_wait)                                // every thread observes the completion: the boxes are visible to it
      mbar_wait(smem_u32(&s_tma_bar[buf]), (uint32_t)((nbuf == 2 ? tile_no >> 1 : tile_no) & 1));
    __syncthreads();
    if (tap_lo >= tap_hi) continue;

    for (int c = vwi; c < nchunks; c += nvw) {
      // ---- B fragments: un-shifted operand, 16 voxels x CU channels
      uint32_t bf[NT][2];
      {
        const int mat = lane >> 3, r = lane & 7;
        if constexpr (CU == 8) {
          const int v = c * 16 + (mat & 1) * 8 + r;
          uint32_t t2[2];
          ldsm_x2_t(sun_addr + (uint32_t)v * 16u, t2);
          bf[0][0] = t2[0]; bf[0][1] = t2[1];
        } else {
          const int v = c * 16 + (mat & 1) * 8 + r;
          uint32_t t4[4];
          ldsm_x4_t(sun_addr + (uint32_t)v * 32u + (uint32_t)(mat >> 1) * 16u, t4);
          bf[0][0] = t4[0]; bf[0][1] = t4[1]; bf[1][0] = t4[2]; bf[1][1] = t4[3];
        }
      }
      // ---- A fragments: tap-shifted operand
      if constexpr (CS >= 8) {
        const int mat = lane >> 3, r = lane & 7;
        const int base = s_vo[c * 16 + (mat >> 1) * 8 + r];                 // box voxel of the un-tapped position
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int t0 = tap_lo + mt * TPM;
          if (t0 >= tap_hi) break;
          uint32_t af[4];
          if constexpr (CS == 8) {
            const int t = min(t0 + (mat & 1), tap_hi - 1);
            ldsm_x4_t(ssh_addr + (uint32_t)(base + s_tapoff[t]) * 16u, af);
          } else {
            ldsm_x4_t(ssh_addr + (uint32_t)(base + s_tapoff[t0]) * 32u + (uint32_t)(mat & 1) * 16u, af);
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[mt][nt], af, bf[nt][0], bf[nt][1]);
        }
      } else {
        // one shifted channel: rows of the m16 tile are taps, fragments are gathered element-wise
        const int gq = lane >> 2, q4 = lane & 3;
        if (pair_ok) {
          // even tile and box widths, unit stride: voxels (2 q4, 2 q4 + 1) are neighbours along w at an even box offset,
          // so a register is one 32-bit word of the copy the tap's parity selects
          const uint32_t a0 = ssh_addr + 2u * (uint32_t)s_vo[c * 16 + 2 * q4], a2 = ssh_addr + 2u * (uint32_t)s_vo[c * 16 + 2 * q4 + 8];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (tap_lo + mt * 16 >= tap_hi) break;
            uint32_t af[4];
            af[0] = lds_u32(a0 + pair_off[mt][0]); af[1] = lds_u32(a0 + pair_off[mt][1]);
            af[2] = lds_u32(a2 + pair_off[mt][0]); af[3] = lds_u32(a2 + pair_off[mt][1]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[mt][nt], af, bf[nt][0], bf[nt][1]);
          }
          continue;
        }
        int vo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) vo[k] = s_vo[c * 16 + 2 * q4 + (k & 1) + (k >> 1) * 8];
        const unsigned short* S16 = reinterpret_cast<const unsigned short*>(Ssh);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (tap_lo + mt * 16 >= tap_hi) break;
          const int ta = min(tap_lo + mt * 16 + gq, tap_hi - 1), tb = min(tap_lo + mt * 16 + gq + 8, tap_hi - 1);
          const int oa = s_tapoff[ta], ob = s_tapoff[tb];
          uint32_t af[4];
          af[0] = (uint32_t)S16[vo[0] + oa] | ((uint32_t)S16[vo[1] + oa] << 16);
          af[1] = (uint32_t)S16[vo[0] + ob] | ((uint32_t)S16[vo[1] + ob] << 16);
          af[2] = (uint32_t)S16[vo[2] + oa] | ((uint32_t)S16[vo[3] + oa] << 16);
          af[3] = (uint32_t)S16[vo[2] + ob] | ((uint32_t)S16[vo[3] + ob] << 16);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[mt][nt], af, bf[nt][0], bf[nt][1]);
        }
      }
    }
  }

  // ---- bias gradient: a lane's sums belong to the 8-channel part (lane mod parts) of dY
  if (dbias) {
    constexpr int CY = MODE == 0 ? CU : CS;
    constexpr int PERY = CY >= 8 ? CY / 8 : 1;
    const int part = lane & (PERY - 1);
#pragma unroll
    for (int c = 0; c < (CY >= 8 ? 8 : 1); ++c) {
      float v = bsum[c];
      // lanes with equal (lane mod PERY) hold the same channels
      for (int o = 16; o >= PERY; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < PERY) atomicAdd(&s_bias[part * 8 + c], v);
    }
  }
  // ---- flush: reduce the warps through shared memory, then one global atomic per weight
  __syncthreads();
  if (dbias && threadIdx.x < (MODE == 0 ? CU : CS)) atomicAdd(dbias + threadIdx.x, s_bias[threadIdx.x]);
  if (grouped) {
    if (cur_grp >= 0) flush(dw_group + (size_t)cur_grp * nred, CS * CU, CU, 1);
  } else {
    // (shifted, un-shifted) channel -> (ci, co) of the PyTorch weight layout
    flush(dw, g.wst_t, MODE == 0 ? g.wst_ci : g.wst_co, MODE == 0 ? g.wst_co : g.wst_ci);
  }
}

// shared-memory bytes of a launch (mirrors the kernel's layout)
static size_t wm_smem_bytes(int cs, int cu, long long tile, long long box, int ntaps, int nbuf) {
  const size_t sh_one = (((size_t)box * cs * 2 + 64) + 127) & ~(size_t)127;
  const size_t sh_bytes = cs == 1 ? 2 * sh_one : sh_one;
  const size_t nch = (size_t)((tile + 15) / 16);
  const size_t un_bytes = (nch * 16 * cu * 2 + 127) & ~(size_t)127;
  const size_t svo = (nch * 32 + 127) & ~(size_t)127;
  const size_t red = (size_t)ntaps * cs * cu * sizeof(float);
  size_t smem = (size_t)nbuf * (sh_bytes + un_bytes) + svo + (nbuf == 2 ? red : 0) + 64;
  return smem > red ? smem : red;
}

static void wm_pick_tile(int cs, int cu, WmGeom& g) {
  // q-voxel tile: shared memory under ~72 KB (~110 KB with two operand buffers) so that two CTAs share an SM; prefer
  // tiles whose staged box is small relative to the useful voxels (halo amortisation) and that tile the grid evenly
  const long long budget = (g.nbuf == 2 ? 110 : 72) * 1024;
  int best[3] = {1, 1, 1};
  double best_score = -1.0;
  const int candD[] = {1, 2, 3, 4, 6, 8}, candH[] = {1, 2, 4, 6, 8, 12}, candW[] = {4, 7, 8, 14, 16, 17, 33, 35};
  for (int a : candD)
    for (int b : candH)
      for (int c0 : candW) {
        int c = c0;
        if (cs == 1 && g.s == 1) {               // one shifted channel: even widths only (pair loads of the A fragments);
          if (c > g.bW) c = g.bW;                // the full row rounds UP to even, the extra voxel is outside the grid = zeros
          c += c & 1;
          if (a > g.bD || b > g.bH) continue;
        } else if (a > g.bD || b > g.bH || c > g.bW) continue;
        const long long tile = (long long)a * b * c;
        const long long box = (long long)((a - 1) * g.s + g.kD) * ((b - 1) * g.s + g.kH) * ((c - 1) * g.s + g.kW);
        const long long bytes = (long long)wm_smem_bytes(cs, cu, tile, box, g.ntaps, g.nbuf) + 192;
        if (box > 65535) continue;
        if (bytes > budget) continue;
        const long long nT = (long long)((g.bD + a - 1) / a) * ((g.bH + b - 1) / b) * ((g.bW + c - 1) / c);
        const double useful = (double)g.bD * g.bH * g.bW;
        // staged bytes per useful voxel (lower is better), chunk quantisation
        const double cost = (double)nT * (double)(box * cs + tile * cu) / useful;
        const double chunk_eff = (double)tile / (double)((tile + 15) / 16 * 16);
        const double score = chunk_eff / cost;
        if (score > best_score) { best_score = score; best[0] = a; best[1] = b; best[2] = c; }
      }
  g.tD = best[0]; g.tH = best[1]; g.tW = best[2];
  g.nTd = (g.bD + g.tD - 1) / g.tD; g.nTh = (g.bH + g.tH - 1) / g.tH; g.nTw = (g.bW + g.tW - 1) / g.tW;
  g.sD = (g.tD - 1) * g.s + g.kD; g.sH = (g.tH - 1) * g.s + g.kH; g.sW = (g.tW - 1) * g.s + g.kW;
  auto magic = [](int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ULL + (uint64_t)d - 1) / (uint64_t)d); };
  g.mul_tH = magic(g.tH); g.mul_tW = magic(g.tW); g.mul_sH = magic(g.sH); g.mul_sW = magic(g.sW);
}

template <int CS, int CU, int MT>
static int launch_wm(const WmGeom& g_in, const float* x, const float* dy, const float* sc, const float* sh, float* dw,
                     float* dbias, cudaStream_t st, float* dw_group = nullptr) {
  WmGeom g = g_in;
  const long long tile = (long long)g.tD * g.tH * g.tW, box = (long long)g.sD * g.sH * g.sW;
  const size_t smem = wm_smem_bytes(CS, CU, tile, box, g.ntaps, g.nbuf);
  // bf16 operands with nothing to fold or sum: one TMA box per tile (the kernel falls back to cp.async without a map)
  alignas(64) CUtensorMap tm_un, tm_sh;
  memset(&tm_un, 0, sizeof(tm_un));
  memset(&tm_sh, 0, sizeof(tm_sh));
  {
    const bool x_plain = g.x_bf16 && !sc, y_plain = g.y_bf16 && !dbias;
    const bool un_plain = g.mode == 0 ? y_plain : x_plain, sh_plain = CS >= 8 && (g.mode == 0 ? x_plain : y_plain);
    const void* un_base = g.mode == 0 ? (const void*)dy : (const void*)x;
    const void* sh_base = g.mode == 0 ? (const void*)x : (const void*)dy;
    const int* ug = g.mode == 0 ? &g.yD : &g.xD;          // (D, H, W) of the un-shifted / shifted operand's grid
    const int* sg = g.mode == 0 ? &g.xD : &g.yD;
    const long long un_img = g.mode == 0 ? g.y_img : g.x_img, sh_img = g.mode == 0 ? g.x_img : g.y_img;
    g.un_tma = un_plain && make_tmap_voxels(un_base, CU, ug[2], ug[1], ug[0], g.N, un_img, g.tW, g.tH, g.tD, &tm_un);
    g.sh_tma = sh_plain && make_tmap_voxels(sh_base, CS, sg[2], sg[1], sg[0], g.N, sh_img, g.sW, g.sH, g.sD, &tm_sh);
  }
  const long long ntiles = (long long)g.nTd * g.nTh * g.nTw * g.N;
  long long blocks = 2LL * vg_sm_count();
  if (blocks > ntiles) blocks = ntiles;
  static const bool wm_debug = getenv("VAEGAM_WM_DEBUG") != nullptr;
  if (wm_debug) {
    int occ = -1;
    if (g.mode == 0) { cudaFuncSetAttribute(wgrad_mma_kernel<CS, CU, MT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wgrad_mma_kernel<CS, CU, MT, 0>, 256, smem); }
    else { cudaFuncSetAttribute(wgrad_mma_kernel<CS, CU, MT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wgrad_mma_kernel<CS, CU, MT, 1>, 256, smem); }
    fprintf(stderr, "[wm] cs=%d cu=%d mode=%d tile=(%d,%d,%d) box=(%d,%d,%d) nT=(%d,%d,%d) smem=%zu nbuf=%d un_tma=%d sh_tma=%d occ=%d ntiles=%lld\n", CS, CU,
            g.mode, g.tD, g.tH, g.tW, g.sD, g.sH, g.sW, g.nTd, g.nTh, g.nTw, smem, g.nbuf, g.un_tma, g.sh_tma, occ, ntiles);
  }
  if (g.mode == 0) {
    VG_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<CS, CU, MT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_mma_kernel<CS, CU, MT, 0><<<(unsigned)blocks, 256, smem, st>>>(g, x, dy, sc, sh, dw, dbias, dw_group, tm_un, tm_sh);
  } else {
    VG_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<CS, CU, MT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_mma_kernel<CS, CU, MT, 1><<<(unsigned)blocks, 256, smem, st>>>(g, x, dy, sc, sh, dw, dbias, dw_group, tm_un, tm_sh);
  }
  VG_LAUNCH_CHECK();
  return VG_OK;
}

// returns VG_OK, or 1 when the channel pair is not covered (caller falls back to the fp32 kernel)
int wgrad_mma(const VgConvDesc* d, const void* x_, const void* dy_, const float* in_scale, const float* in_shift,
              float* dw, float* dbias, cudaStream_t st, float* dw_group) {
  const float* x = static_cast<const float*>(x_);
  const float* dy = static_cast<const float*>(dy_);
  WmGeom g{};
  g.x_bf16 = (d->bf16_mask & VG_BF16_X) != 0;
  g.y_bf16 = (d->bf16_mask & VG_BF16_Y) != 0;
  if ((g.x_bf16 && d->cin % 8) || (g.y_bf16 && d->cout % 8)) { set_error("bf16 storage needs 8 or 16 channels"); return VG_EINVAL; }
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
  if (K > 48) return 1;
  g.ntaps = K;
  g.wst_t = 1;
  if (!d->transposed) { g.wst_ci = K; g.wst_co = d->cin * K; }        // w[co][ci][K]
  else                { g.wst_ci = d->cout * K; g.wst_co = K; }        // w[ci][co][K]
  const int cs = d->transposed ? d->cout : d->cin, cu = d->transposed ? d->cin : d->cout;
  // 1-channel images need 16-byte aligned rows only through the vector path (C >= 8): always true for dense tensors
  if ((cs >= 8 && ((d->transposed ? g.y_img : g.x_img) % 4)) || ((d->transposed ? g.x_img : g.y_img) % 4)) return 1;
  auto groups = [&](int tpm, int mt) {       // tap groups so that a warp's m16 tiles fit MT
    const int tiles = (K + tpm - 1) / tpm;
    int tg = (tiles + mt - 1) / mt;
    while (8 % tg) ++tg;
    g.tg = tg;
    g.taps_per_group = ((tiles + tg - 1) / tg) * tpm;
  };
  {
    // plain bf16 operands travel by TMA, double-buffered; the tile is picked for that shared-memory budget
    const bool x_plain = g.x_bf16 && !in_scale, y_plain = g.y_bf16 && !dbias;
    const bool un_plain = g.mode == 0 ? y_plain : x_plain;
    // two buffers only where they do not shrink the tile: with an 8/16-channel shifted box the halo dominates the
    // bytes and a smaller tile costs more than the prefetch wins (convt4: 0.32 vs 0.24 ms)
    g.nbuf = (tma_available() && un_plain && cs == 1) ? 2 : 1;
  }
  wm_pick_tile(cs, cu, g);
  if (cs == 1 && cu == 8) { groups(16, 2); return launch_wm<1, 8, 2>(g, x, dy, in_scale, in_shift, dw, dbias, st, dw_group); }
  // MT keeps the accumulators at <= 56 registers so that two CTAs share an SM without spills
  if (cs == 8 && cu == 8) { groups(2, 12); return launch_wm<8, 8, 12>(g, x, dy, in_scale, in_shift, dw, dbias, st, dw_group); }
  if (cs == 8 && cu == 16) { groups(2, 7); return launch_wm<8, 16, 7>(g, x, dy, in_scale, in_shift, dw, dbias, st, dw_group); }
  if (cs == 16 && cu == 16) { groups(1, 7); return launch_wm<16, 16, 7>(g, x, dy, in_scale, in_shift, dw, dbias, st, dw_group); }
  if (cs == 16 && cu == 8) { groups(1, 14); return launch_wm<16, 8, 14>(g, x, dy, in_scale, in_shift, dw, dbias, st, dw_group); }
  return 1;
}

}  // namespace vg
