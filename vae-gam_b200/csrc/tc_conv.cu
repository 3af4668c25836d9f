// Implicit-GEMM 3-D convolution on the 5th-generation tensor cores (tcgen05 + TMEM), bf16
// operands, fp32 accumulation.  It serves every unit-input-stride tap-list gather of conv.cu
// (Conv3d / ConvTranspose3d forward incl. the stride-2 parity phases, and the matching data
// gradients) whose input has 8 or 16 channels:
//
//   D[m = voxel (h,w) of a 16x8 tile][n = cout] += A[m][k = (tap, ci)] * B[k][n]
//
// * A is never materialised.  A CTA keeps a ring of input d-planes in shared memory as
//   channels-last bf16 halo tiles [18][11][8ch] (16 B per voxel).  In that layout an 8-voxel run
//   along w IS a no-swizzle K-major core matrix (8 rows x 16 B), consecutive h rows are a
//   constant SBO apart, so "the im2col matrix of tap (dd,dh,dw)" is just a shared-memory
//   descriptor whose start address is shifted by the tap offset; two taps adjacent in w (or the
//   two 8-channel halves of a 16-channel voxel) form the K=16 of one tcgen05.mma through
//   LBO = 16 B (resp. the half-plane stride).  Core matrices of different MMAs overlap freely
//   (validated on hardware by tc_probe.cu).
// * The planes are staged by the CTA's threads (fp32 global -> BatchNorm fold -> bf16 ->
//   st.shared, zero outside the grid), made visible to the async proxy with
//   fence.proxy.async; one elected thread issues the MMAs of an output plane and commits to an
//   mbarrier; accumulators are double-buffered in TMEM so the epilogue of plane p-1
//   (tcgen05.ld -> bias / activation / statistics or backward epilogues -> global) overlaps
//   the MMAs of plane p and the staging of plane p+1.
// * B (weights) is converted once per CTA into per-MMA canonical 16x16 bf16 blocks.
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_geom.cuh"

namespace vg {

constexpr int TC_TH = 16, TC_TW = 8;       // output tile (h, w) = 128 GEMM rows
constexpr int TC_MAX_MMA = 32;

struct TcPlan {
  int nmma;
  int lo_d, lo_h, lo_w;        // smallest tap offsets
  int span_d, span_h, span_w;  // largest - smallest
  int PH, PW;                  // staged box (voxels)
  int NP;                      // ring slots
  int dchunk, nchunks;         // q planes per CTA, CTAs per column
  int8_t dd[TC_MAX_MMA], dh[TC_MAX_MMA], dw[TC_MAX_MMA];   // offsets relative to lo_*
  int8_t t0[TC_MAX_MMA], t1[TC_MAX_MMA];                   // tap ids of K-chunk 0 / 1 (-1 = zero weights)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (sm_100); no swizzle, K-major
  return d;
}

// kind::f16, A = B = bf16, D = f32, both K-major, M = 128, N = 16
__device__ __forceinline__ uint32_t umma_idesc_m128_n16() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  // bounded: a tensor-core operation that never completes must surface as an error, not a hang
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1 << 22)) asm volatile("trap;");
  }
}

template <int NCOL>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NCOL]) {
  uint32_t r[NCOL];
  if constexpr (NCOL == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  } else if constexpr (NCOL == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NCOL; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(128)
tc_gather_kernel(const __grid_constant__ Geom g, const GatherArgs a, const __grid_constant__ TcPlan pl) {
  constexpr int NG = CIN / 8;                  // 8-channel groups per voxel
  constexpr int NLD = COUT >= 16 ? 16 : (COUT == 8 ? 8 : 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double sred[2 * COUT];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group_bytes = pl.PH * pl.PW * 16;
  const int slot_bytes = NG * group_bytes;
  uint8_t* tiles = smem;                                   // NP slots
  uint8_t* wblk = smem + (size_t)pl.NP * slot_bytes;       // nmma blocks of 512 B
  wblk = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wblk) + 127) & ~uintptr_t(127));

  // ---- one-time setup: barriers, TMEM, weights
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 2 * COUT) sred[tid] = 0.0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int e = tid; e < pl.nmma * 256; e += blockDim.x) {  // 16 (n) x 16 (k) per block
    const int i = e >> 8, rem = e & 255, n = rem >> 4, k = rem & 15;
    const int chunk = k >> 3, cl = k & 7;
    const int tap = (NG == 2) ? pl.t0[i] : (chunk == 0 ? pl.t0[i] : pl.t1[i]);
    const int ci = (NG == 2) ? chunk * 8 + cl : cl;
    float w = 0.f;
    if (tap >= 0 && n < COUT)
      w = __ldg(a.w + (size_t)g.taps[tap].widx * g.wst_t + (size_t)ci * g.wst_ci + (size_t)n * g.wst_co);
    reinterpret_cast<__nv_bfloat16*>(wblk)[(i * 512 + (n >> 3) * 256 + chunk * 128 + (n & 7) * 16 + cl * 2) >> 1] =
        __float2bfloat16(w);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  // ---- this CTA's column: image n, tile (th, tw), all d planes of the q grid
  const int tilesW = (g.qW + TC_TW - 1) / TC_TW, tilesH = (g.qH + TC_TH - 1) / TC_TH;
  const int col = blockIdx.x / pl.nchunks;
  const int qd_beg = (blockIdx.x - col * pl.nchunks) * pl.dchunk;
  const int qd_end = min(g.qD, qd_beg + pl.dchunk);
  const int n = col / (tilesH * tilesW);
  const int trem = col - n * (tilesH * tilesW);
  const int h0 = (trem / tilesW) * TC_TH, w0 = (trem % tilesW) * TC_TW;
  const int grp = n / g.group_size;
  const float* in_n = a.in + (size_t)n * g.in_img;

  float sc[CIN], sh[CIN];
  const bool affine = a.in_scale != nullptr;
  // each thread stages at most two fixed voxels of every plane (PH*PW <= 198 <= 2*128): their
  // in-plane global offset (or -1 outside the grid / unused) and shared-memory offset are fixed
  // for the whole column, so only the plane base moves in the loop.
  int goff[2], soff[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int v = tid + k * 128;
    goff[k] = -1;
    soff[k] = v * 16;
    if (v < pl.PH * pl.PW) {
      const int i = v / pl.PW, j = v - i * pl.PW;
      const int gh = h0 + pl.lo_h + i, gw = w0 + pl.lo_w + j;
      if (gh >= 0 && gh < g.inH && gw >= 0 && gw < g.inW) goff[k] = (gh * g.inW + gw) * CIN;
      else goff[k] = -2;                         // inside the staged box but outside the grid: store zeros
    }
  }
  const size_t plane_floats = (size_t)g.inH * g.inW * CIN;
  float4 pre[2][2 * NG];                          // prefetch registers: loads of the NEXT plane stay in flight
  bool pre_ok[2];

  auto load_plane = [&](int rel) {                // rel = plane index relative to lo_d; issues the global loads
    const int gd = rel + pl.lo_d;                 // input plane (unit input stride)
    const bool d_ok = gd >= 0 && gd < g.inD;
    const float* base = in_n + (size_t)(d_ok ? gd : 0) * plane_floats;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      pre_ok[k] = d_ok && goff[k] >= 0;
      if (pre_ok[k]) {
        const float4* p = reinterpret_cast<const float4*>(base + goff[k]);
#pragma unroll
        for (int q = 0; q < 2 * NG; ++q) pre[k][q] = __ldg(p + q);
      }
    }
  };
  auto store_plane = [&](int rel) {               // fold + convert the prefetched voxels into the ring slot
    uint8_t* slot = tiles + (size_t)(rel % pl.NP) * slot_bytes;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (goff[k] == -1) continue;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (pre_ok[k]) {
          const float4 lo = pre[k][2 * gi], hi = pre[k][2 * gi + 1];
          float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          if (affine) {
#pragma unroll
            for (int c = 0; c < 8; ++c) f[c] = fmaf(f[c], sc[gi * 8 + c], sh[gi * 8 + c]);
          }
          pk = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
        *reinterpret_cast<uint4*>(slot + (size_t)gi * group_bytes + soff[k]) = pk;
      }
    }
  };

  if (affine) {
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      sc[c] = __ldg(a.in_scale + grp * CIN + c);
      sh[c] = __ldg(a.in_shift + grp * CIN + c);
    }
  }

  // epilogue state
  const int m = warp * 32 + lane;              // GEMM row = TMEM lane
  const int eh = h0 + (m >> 3), ew = w0 + (m & 7);
  const bool vox_ok = eh < g.qH && ew < g.qW;
  float bias[COUT], istd[COUT], mistd[COUT], s1[COUT], s2[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) {
    bias[c] = a.bias ? __ldg(a.bias + c) : 0.f;
    s1[c] = s2[c] = 0.f;
    istd[c] = mistd[c] = 0.f;
  }
  const bool want_stats = a.stats != nullptr;
  const bool want_bn = a.aux_mode == 2;
  if (want_bn) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      istd[c] = __ldg(a.aux_istd + grp * COUT + c);
      mistd[c] = __ldg(a.aux_mistd + grp * COUT + c);
    }
  }

  auto epilogue = [&](int qd, int it) {           // it = iteration index within this CTA
    const int b = it & 1;
    mbar_wait(smem_u32(&mbar[b]), (uint32_t)((it >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;");
    float acc[NLD];
    tmem_ld<NLD>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * 16), acc);
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (!vox_ok) return;
    const int od = qd * g.sout + g.rD, oh = eh * g.sout + g.rH, ow = ew * g.sout + g.rW;
    const size_t o = (size_t)n * g.out_img + (((size_t)od * g.outH + oh) * g.outW + ow) * COUT;
    float y[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      float t = acc[c] + bias[c];
      if (a.act == VG_ACT_RELU) t = fmaxf(t, 0.f);
      else if (a.act == VG_ACT_SIGMOID) t = 1.f / (1.f + __expf(-t));
      y[c] = t;
    }
    if (a.aux_mode != 0) {
      float ax[COUT];
      if constexpr (COUT % 4 == 0) {
#pragma unroll
        for (int i = 0; i < COUT / 4; ++i) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(a.aux + o) + i);
          ax[4 * i] = t.x; ax[4 * i + 1] = t.y; ax[4 * i + 2] = t.z; ax[4 * i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < COUT; ++i) ax[i] = __ldg(a.aux + o + i);
      }
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
    if (a.out) {
      if constexpr (COUT % 4 == 0) {
#pragma unroll
        for (int i = 0; i < COUT / 4; ++i)
          reinterpret_cast<float4*>(a.out + o)[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < COUT; ++i) a.out[o + i] = y[i];
      }
    }
  };

  // ---- main loop over output planes
  const uint32_t idesc = umma_idesc_m128_n16();
  const uint32_t tiles_addr = smem_u32(tiles), wblk_addr = smem_u32(wblk);
  const uint32_t sbo = (uint32_t)pl.PW * 16u;
  const uint32_t lbo = NG == 2 ? (uint32_t)group_bytes : 16u;
  // planes are addressed relative to lo_d: output plane qd reads ring planes qd .. qd + span_d
  for (int rel = qd_beg; rel < qd_beg + pl.span_d; ++rel) { load_plane(rel); store_plane(rel); }
  load_plane(qd_beg + pl.span_d);
  for (int qd = qd_beg; qd < qd_end; ++qd) {
    store_plane(qd + pl.span_d);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    const int it = qd - qd_beg;
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t d_tmem = tmem_base + (uint32_t)((it & 1) * 16);
      for (int i = 0; i < pl.nmma; ++i) {
        const uint32_t slot = (uint32_t)((qd + pl.dd[i]) % pl.NP);
        const uint32_t a_addr = tiles_addr + slot * (uint32_t)slot_bytes + (uint32_t)(pl.dh[i] * pl.PW + pl.dw[i]) * 16u;
        umma_bf16(d_tmem, umma_desc(a_addr, lbo, sbo), umma_desc(wblk_addr + (uint32_t)i * 512u, 128u, 256u), idesc,
                  i > 0 ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mbar[it & 1]))
                   : "memory");
    }
    if (qd + 1 < qd_end) load_plane(qd + 1 + pl.span_d);      // in flight during the epilogue below
    if (it > 0) epilogue(qd - 1, it - 1);
  }
  epilogue(qd_end - 1, qd_end - 1 - qd_beg);

  // ---- statistics flush + teardown
  if (want_stats || want_bn) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      const float r1 = warp_sum(s1[c]);
      const float r2 = warp_sum(s2[c]);
      if (lane == 0) {
        atomicAdd(&sred[2 * c], (double)r1);
        atomicAdd(&sred[2 * c + 1], (double)r2);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if ((want_stats || want_bn) && tid < 2 * COUT) {
    double* dst = want_stats ? a.stats : a.aux_sums;
    atomicAdd(dst + (size_t)grp * COUT * 2 + tid, sred[tid]);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base));
}

// Can this geometry run on the tensor-core path?
bool tc_supported(int cin, int cout, const Geom& g) {
  if (g.sin != 1) return false;
  if (cin != 8 && cin != 16) return false;
  if (cout != 1 && cout != 8 && cout != 16) return false;
  if (g.ntaps < 1) return false;
  int lo[3] = {127, 127, 127}, hi[3] = {-127, -127, -127};
  for (int t = 0; t < g.ntaps; ++t) {
    const int o[3] = {g.taps[t].dd, g.taps[t].dh, g.taps[t].dw};
    for (int i = 0; i < 3; ++i) { lo[i] = o[i] < lo[i] ? o[i] : lo[i]; hi[i] = o[i] > hi[i] ? o[i] : hi[i]; }
  }
  for (int i = 0; i < 3; ++i)
    if (hi[i] - lo[i] > 2) return false;
  return true;
}

static void build_plan(int cin, const Geom& g, TcPlan& pl) {
  int lo[3] = {127, 127, 127}, hi[3] = {-127, -127, -127};
  for (int t = 0; t < g.ntaps; ++t) {
    const int o[3] = {g.taps[t].dd, g.taps[t].dh, g.taps[t].dw};
    for (int i = 0; i < 3; ++i) { lo[i] = o[i] < lo[i] ? o[i] : lo[i]; hi[i] = o[i] > hi[i] ? o[i] : hi[i]; }
  }
  pl.lo_d = lo[0]; pl.lo_h = lo[1]; pl.lo_w = lo[2];
  pl.span_d = hi[0] - lo[0]; pl.span_h = hi[1] - lo[1]; pl.span_w = hi[2] - lo[2];
  pl.PH = TC_TH + pl.span_h;
  pl.PW = TC_TW + pl.span_w + 1;     // +1: the second K-chunk of an unpaired tap reads one voxel further
  pl.NP = pl.span_d + 2;
  pl.dchunk = g.qD;
  pl.nchunks = 1;
  pl.nmma = 0;
  bool used[kMaxTaps] = {false};
  auto find = [&](int dd, int dh, int dw) {
    for (int t = 0; t < g.ntaps; ++t)
      if (!used[t] && g.taps[t].dd == dd && g.taps[t].dh == dh && g.taps[t].dw == dw) return t;
    return -1;
  };
  for (int t = 0; t < g.ntaps; ++t) {
    if (used[t]) continue;
    used[t] = true;
    const int i = pl.nmma++;
    pl.dd[i] = (int8_t)(g.taps[t].dd - lo[0]);
    pl.dh[i] = (int8_t)(g.taps[t].dh - lo[1]);
    pl.dw[i] = (int8_t)(g.taps[t].dw - lo[2]);
    pl.t0[i] = (int8_t)t;
    pl.t1[i] = -1;
    if (cin == 8) {   // pair with the tap one voxel further along w, if it exists
      const int u = find(g.taps[t].dd, g.taps[t].dh, g.taps[t].dw + 1);
      if (u >= 0) { used[u] = true; pl.t1[i] = (int8_t)u; }
    }
  }
}

template <int CIN, int COUT>
static int launch_tc_t(const Geom& g, const GatherArgs& a, const TcPlan& pl, cudaStream_t st) {
  const int tilesW = (g.qW + TC_TW - 1) / TC_TW, tilesH = (g.qH + TC_TH - 1) / TC_TH;
  const long long cols = (long long)g.N * tilesH * tilesW * pl.nchunks;
  const size_t smem = (size_t)pl.NP * (CIN / 8) * pl.PH * pl.PW * 16 + (size_t)pl.nmma * 512 + 1024 + 128;
  VG_CUDA(cudaFuncSetAttribute(tc_gather_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_gather_kernel<CIN, COUT><<<(unsigned)cols, 128, smem, st>>>(g, a, pl);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

// sort taps so that w-neighbours are visited in increasing dw (pairing walks left to right)
int launch_tc_gather(int cin, int cout, const Geom& g_in, const GatherArgs& a, cudaStream_t st) {
  Geom g = g_in;
  for (int i = 1; i < g.ntaps; ++i) {          // insertion sort by (dd, dh, dw)
    Tap t = g.taps[i];
    int j = i - 1;
    auto key = [](const Tap& x) { return (x.dd + 64) * 16384 + (x.dh + 64) * 128 + (x.dw + 64); };
    while (j >= 0 && key(g.taps[j]) > key(t)) { g.taps[j + 1] = g.taps[j]; --j; }
    g.taps[j + 1] = t;
  }
  TcPlan pl{};
  build_plan(cin, g, pl);
  {  // split columns along d until the grid can fill the machine (>= 4 CTAs per SM), keeping >= 6 planes per CTA
    const int tilesW = (g.qW + TC_TW - 1) / TC_TW, tilesH = (g.qH + TC_TH - 1) / TC_TH;
    const long long cols = (long long)g.N * tilesH * tilesW;
    const long long want = 4LL * vg_sm_count();
    int nch = (int)((want + cols - 1) / cols);
    const int max_ch = g.qD / 6 > 0 ? g.qD / 6 : 1;
    if (nch > max_ch) nch = max_ch;
    if (nch < 1) nch = 1;
    pl.dchunk = (g.qD + nch - 1) / nch;
    pl.nchunks = (g.qD + pl.dchunk - 1) / pl.dchunk;
  }
  if (pl.nmma > TC_MAX_MMA) { set_error("tensor-core plan needs %d MMAs (max %d)", pl.nmma, TC_MAX_MMA); return VG_EINVAL; }
  if (cin == 8 && cout == 1) return launch_tc_t<8, 1>(g, a, pl, st);
  if (cin == 8 && cout == 8) return launch_tc_t<8, 8>(g, a, pl, st);
  if (cin == 8 && cout == 16) return launch_tc_t<8, 16>(g, a, pl, st);
  if (cin == 16 && cout == 1) return launch_tc_t<16, 1>(g, a, pl, st);
  if (cin == 16 && cout == 8) return launch_tc_t<16, 8>(g, a, pl, st);
  if (cin == 16 && cout == 16) return launch_tc_t<16, 16>(g, a, pl, st);
  set_error("tensor-core path: unsupported channel pair (%d,%d)", cin, cout);
  return VG_EINVAL;
}

}  // namespace vg
