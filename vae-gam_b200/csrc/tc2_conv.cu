// Plane-folded implicit-GEMM 3-D convolution on tcgen05 / TMEM (bf16 operands, fp32 accumulate).
//
// The layers of VAE-GAM have 1, 8 or 16 channels, so a plain implicit GEMM (N = cout) leaves the
// tensor cores waiting for shared memory: every MMA re-reads a 128-row A tile to produce 8 or 16
// columns.  This kernel folds OUTPUT d-PLANES into the N dimension instead (a block-Toeplitz
// weight matrix along d):
//
//   D[row = (h,w) voxel of the plane][n = (j, co)] += A[row][k = (i - 2p, ci)] * T[k][n]
//   A = input planes i = 2p, 2p+1 of the window (K = 16 = two 8-channel voxels, LBO = plane stride)
//   T[(i,ci)][(j,co)] = W[dd = i - j + lo_d, dh, dw][ci][co]   (0 where no such tap)
//
// so one read of A serves up to four output planes (N = 4*cout) and the (dh,dw) taps are
// shifted-window descriptors over the same staged plane: rows are the plane's voxels flattened
// with a padded pitch, so a tap is a constant row shift and a tile is 128 consecutive rows.
// Inputs with ONE channel use the 8 consecutive w-voxels of a row as the K-chunk (the dw taps sit
// inside K).  Stride-2 transposed layers run as their output-parity phases (sout = 2).
//
// Roles (416 threads, one persistent CTA per SM):
//   warps 0-3   epilogue: tcgen05.ld -> bias / activation / BatchNorm statistics / backward masks
//               -> global; re-zero the accumulator buffer (all MMAs accumulate)
//   warps 4-11  producers: fp32 global -> BatchNorm fold -> bf16 -> ring of plane pairs in smem
//   warp 12     one lane issues the tcgen05.mma list of a block of OB output planes
// Pipelines: full/empty mbarriers per ring slot (producers <-> MMA via tcgen05.commit), and
// full/empty per accumulator buffer (MMA <-> epilogue), accumulators double-buffered in TMEM.
#include "common.cuh"
#include "conv_geom.cuh"
#include "tc_common.cuh"

namespace vg {

constexpr int T2_EPI_WARPS = 4, T2_PROD_WARPS = 8;
constexpr int T2_PROD_THREADS = T2_PROD_WARPS * 32;
constexpr int T2_THREADS = (T2_EPI_WARPS + T2_PROD_WARPS + 1) * 32;
constexpr int T2_MAX_MMA = 96, T2_MAX_BLK = 96, T2_MAX_PAIR = 10, T2_MAX_RING = 16;
constexpr int T2_MAX_CHUNK = 3;                 // staged 16-byte chunks per producer thread and plane

struct T2Mma {
  uint16_t a_shift;    // row shift of the A window inside the staged plane
  uint16_t b_off16;    // weight block offset / 16 bytes
  uint8_t n8;          // N >> 3
  uint8_t dcol;        // first accumulator column
  uint8_t pad[2];
};
struct T2Blk {
  int8_t i0, j0, nj, dh, dw, pad[3];   // window plane of K-chunk 0; output planes [j0, j0+nj); tap offsets relative to lo_*
};
struct T2Plan {
  int PW, RTOT, TR, ntiles, SR, OB, NPAIR, R, nrb, ACCW, tmem_cols;
  int lo_d, lo_h, lo_w, span_d, span_h, span_w;
  int dchunk, ndchunks;
  int nblk, wbytes;
  int pair_begin[T2_MAX_PAIR + 1];
  T2Mma mma[T2_MAX_MMA];
  T2Blk blk[T2_MAX_BLK];
  uint16_t blk_off16[T2_MAX_BLK];
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(T2_THREADS, 1)
tc2_kernel(const __grid_constant__ Geom g, const GatherArgs a, const __grid_constant__ T2Plan pl) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[T2_MAX_RING], empty_bar[T2_MAX_RING], accf_bar[2], acce_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int lut[5 * 3 * 3];
  __shared__ float s_bias[16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SRB = pl.SR * 16;                 // bytes per staged plane
  const int PAIRB = 2 * SRB;
  uint8_t* ring = smem;
  uint8_t* wts = smem + (size_t)pl.R * PAIRB;
  const int H2 = pl.OB / 2;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < pl.R; ++s) {
      mbar_init(smem_u32(&full_bar[s]), T2_PROD_THREADS);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&accf_bar[b]), 1);
      mbar_init(smem_u32(&acce_bar[b]), T2_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 45) lut[tid] = -1;
  if (tid < 16) s_bias[tid] = (a.bias && tid < COUT) ? __ldg(a.bias + tid) : 0.f;
  if (warp == T2_EPI_WARPS + T2_PROD_WARPS) {
    const uint32_t dst = smem_u32(&tmem_base_s);
    switch (pl.tmem_cols) {
      case 32: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(dst)); break;
      case 64: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(dst)); break;
      case 128: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst)); break;
      case 256: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst)); break;
      default: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst)); break;
    }
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  if (tid < g.ntaps) {
    const Tap tp = g.taps[tid];
    lut[(tp.dd - pl.lo_d) * 9 + (tp.dh - pl.lo_h) * 3 + (tp.dw - pl.lo_w)] = tp.widx;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // Toeplitz weight blocks: element (n, k) of a block = W[tap(i - j, dh, dw)][ci][co], canonical K-major layout
  for (int b = 0; b < pl.nblk; ++b) {
    const T2Blk bk = pl.blk[b];
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(wts + (size_t)pl.blk_off16[b] * 16);
    const int nel = bk.nj * COUT * 16;
    for (int e = tid; e < nel; e += T2_THREADS) {
      const int n = e >> 4, k = e & 15;
      const int chunk = k >> 3, el = k & 7;
      const int j = bk.j0 + n / COUT, co = n % COUT;
      const int ddr = bk.i0 + chunk - j;
      int ci, dwr;
      if constexpr (CIN == 1) { ci = 0; dwr = el; } else { ci = el; dwr = bk.dw; }
      float w = 0.f;
      if (ddr >= 0 && ddr <= pl.span_d && dwr <= pl.span_w) {
        const int widx = lut[ddr * 9 + bk.dh * 3 + dwr];
        if (widx >= 0) w = __ldg(a.w + (size_t)widx * g.wst_t + (size_t)ci * g.wst_ci + (size_t)co * g.wst_co);
      }
      dst[(((n >> 3) * 256 + chunk * 128 + (n & 7) * 16) >> 1) + el] = __float2bfloat16(w);
    }
  }
  if (warp < T2_EPI_WARPS) {                    // all MMAs accumulate: start from zero
    for (int c = 0; c < pl.tmem_cols; c += 16) tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c);
    tmem_st_wait();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int ncols = g.N * pl.ntiles * pl.ndchunks;
  int pair_base = 0, block_base = 0;            // running counters, identical in every role

  if (warp < T2_EPI_WARPS) {
    // ================================================================ epilogue warps
    const bool want_stats = a.stats != nullptr, want_bn = a.aux_mode == 2;
    constexpr int JG = COUT == 1 ? 16 : (COUT == 8 ? 4 : 2);      // output planes per TMEM load group
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      const int dc = col % pl.ndchunks, t = (col / pl.ndchunks) % pl.ntiles, n = col / (pl.ndchunks * pl.ntiles);
      const int qd0 = dc * pl.dchunk, qd1 = min(g.qD, qd0 + pl.dchunk);
      const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
      const int grp = n / g.group_size;
      float istd[COUT], mistd[COUT], s1[COUT], s2[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        s1[c] = s2[c] = 0.f;
        istd[c] = want_bn ? __ldg(a.aux_istd + grp * COUT + c) : 0.f;
        mistd[c] = want_bn ? __ldg(a.aux_mistd + grp * COUT + c) : 0.f;
      }
      const size_t plane_out = (size_t)g.outH * g.outW * COUT;
      for (int b = 0; b < nblocks; ++b) {
        const int bg = block_base + b, buf = bg & 1;
        mbar_wait(smem_u32(&accf_bar[buf]), (uint32_t)((bg >> 1) & 1));
        tc_fence_after();
        for (int rb = 0; rb < pl.nrb; ++rb) {
          const int r = t * pl.TR + rb * 128 + tid;
          const int qh = r / pl.PW, qw = r - qh * pl.PW;
          const bool row_ok = r < pl.RTOT && qw < g.qW;
          const size_t o_row = (size_t)n * g.out_img +
                               ((size_t)(qh * g.sout + g.rH) * g.outW + (size_t)(qw * g.sout + g.rW)) * COUT;
          const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((buf * pl.nrb + rb) * pl.ACCW);
          for (int j0 = 0; j0 < pl.OB; j0 += JG) {
            float acc[JG][COUT];
            {
              if constexpr (COUT == 1) {
                uint32_t rr[16];
                tmem_ld16(tacc + (uint32_t)j0, rr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j][0] = __uint_as_float(rr[j]);
              } else if constexpr (COUT == 8) {
                uint32_t rr[JG][8];
#pragma unroll
                for (int j = 0; j < JG; ++j) tmem_ld8(tacc + (uint32_t)((j0 + j) * 8), rr[j]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < JG; ++j)
#pragma unroll
                  for (int c = 0; c < 8; ++c) acc[j][c] = __uint_as_float(rr[j][c]);
              } else {
                uint32_t rr[JG][16];
#pragma unroll
                for (int j = 0; j < JG; ++j) tmem_ld16(tacc + (uint32_t)((j0 + j) * 16), rr[j]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < JG; ++j)
#pragma unroll
                  for (int c = 0; c < 16; ++c) acc[j][c] = __uint_as_float(rr[j][c]);
              }
            }
            if (!row_ok) continue;
            // issue the auxiliary loads of the whole group first (memory-level parallelism)
            float ax[JG][COUT];
            bool pok[JG];
            size_t oo[JG];
#pragma unroll
            for (int j = 0; j < JG; ++j) {
              const int qd = qd0 + b * pl.OB + j0 + j;
              pok[j] = qd < qd1;
              oo[j] = o_row + (size_t)(qd * g.sout + g.rD) * plane_out;
              if (a.aux_mode != 0 && pok[j]) {
                if constexpr (COUT % 4 == 0) {
#pragma unroll
                  for (int i = 0; i < COUT / 4; ++i) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(a.aux + oo[j]) + i);
                    ax[j][4 * i] = v.x; ax[j][4 * i + 1] = v.y; ax[j][4 * i + 2] = v.z; ax[j][4 * i + 3] = v.w;
                  }
                } else {
#pragma unroll
                  for (int c = 0; c < COUT; ++c) ax[j][c] = __ldg(a.aux + oo[j] + c);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < JG; ++j) {
              if (!pok[j]) continue;
              float y[COUT];
#pragma unroll
              for (int c = 0; c < COUT; ++c) {
                float v = acc[j][c] + s_bias[c];
                if (a.act == VG_ACT_RELU) v = fmaxf(v, 0.f);
                else if (a.act == VG_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
                y[c] = v;
              }
              if (a.aux_mode == 1) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) y[c] = ax[j][c] > 0.f ? y[c] : 0.f;
              } else if (a.aux_mode == 2) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                  const float xh = fmaf(ax[j][c], istd[c], -mistd[c]);
                  s1[c] += y[c];
                  s2[c] = fmaf(y[c], xh, s2[c]);
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
                    reinterpret_cast<float4*>(a.out + oo[j])[i] =
                        make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                } else {
#pragma unroll
                  for (int c = 0; c < COUT; ++c) a.out[oo[j] + c] = y[c];
                }
              }
            }
          }
          // this row block of the buffer is drained: zero it for the block after next
          for (int c = 0; c < pl.ACCW; c += 16) tmem_zero16(tacc + (uint32_t)c);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&acce_bar[buf]));
      }
      if (want_stats || want_bn) {
        double* dst = (want_stats ? a.stats : a.aux_sums) + (size_t)grp * COUT * 2;
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float r1 = warp_sum(s1[c]);
          const float r2 = warp_sum(s2[c]);
          if (lane == 0) {
            atomicAdd(dst + 2 * c, (double)r1);
            atomicAdd(dst + 2 * c + 1, (double)r2);
          }
        }
      }
      pair_base += nblocks * H2 + (pl.NPAIR - H2);
      block_base += nblocks;
    }
  } else if (warp < T2_EPI_WARPS + T2_PROD_WARPS) {
    // ================================================================ producer warps
    const int ptid = tid - T2_EPI_WARPS * 32;
    const bool affine = a.in_scale != nullptr;
    const size_t plane_in = (size_t)g.inH * g.inW * CIN;
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      const int dc = col % pl.ndchunks, t = (col / pl.ndchunks) % pl.ntiles, n = col / (pl.ndchunks * pl.ntiles);
      const int qd0 = dc * pl.dchunk, qd1 = min(g.qD, qd0 + pl.dchunk);
      const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
      const int npairs = nblocks * H2 + (pl.NPAIR - H2);
      const int grp = n / g.group_size;
      const float* in_n = a.in + (size_t)n * g.in_img;
      float sc[CIN], sh[CIN];
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        sc[c] = affine ? __ldg(a.in_scale + grp * CIN + c) : 1.f;
        sh[c] = affine ? __ldg(a.in_shift + grp * CIN + c) : 0.f;
      }
      // the chunks this thread stages are the same for every plane of the column
      int goff[T2_MAX_CHUNK];      // CIN 8: float offset of the voxel inside its plane; -2 zero chunk; -1 none
      int wlim[T2_MAX_CHUNK];      // CIN 1: number of in-range w elements of the chunk
#pragma unroll
      for (int k = 0; k < T2_MAX_CHUNK; ++k) {
        const int s = ptid + k * T2_PROD_THREADS;
        goff[k] = -1; wlim[k] = 0;
        if (s < pl.SR) {
          const int rr = t * pl.TR + s;
          const int hh = rr / pl.PW, ww = rr - hh * pl.PW;
          const int ih = hh + pl.lo_h, iw = ww + pl.lo_w;
          goff[k] = -2;
          if constexpr (CIN == 1) {
            if (ih >= 0 && ih < g.inH && iw >= 0 && iw < g.inW) {
              goff[k] = ih * g.inW + iw;
              wlim[k] = min(pl.span_w + 1, g.inW - iw);
            }
          } else {
            if (ih >= 0 && ih < g.inH && iw >= 0 && iw < g.inW) goff[k] = (ih * g.inW + iw) * CIN;
          }
        }
      }
      for (int P = 0; P < npairs; ++P) {
        const int G = pair_base + P, slot = G % pl.R, use = G / pl.R;
        if (use > 0) mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)((use - 1) & 1));
        uint8_t* dst0 = ring + (size_t)slot * PAIRB;
        if constexpr (CIN == 8) {
          float4 v[2][T2_MAX_CHUNK][2];
          bool ok[2][T2_MAX_CHUNK];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int ip = qd0 + pl.lo_d + 2 * P + hf;
            const bool p_ok = ip >= 0 && ip < g.inD;
            const float* base = in_n + (size_t)(p_ok ? ip : 0) * plane_in;
#pragma unroll
            for (int k = 0; k < T2_MAX_CHUNK; ++k) {
              ok[hf][k] = p_ok && goff[k] >= 0;
              if (ok[hf][k]) {
                const float4* p = reinterpret_cast<const float4*>(base + goff[k]);
                v[hf][k][0] = __ldg(p);
                v[hf][k][1] = __ldg(p + 1);
              }
            }
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int k = 0; k < T2_MAX_CHUNK; ++k) {
              if (goff[k] == -1) continue;
              uint4 pk = make_uint4(0u, 0u, 0u, 0u);
              if (ok[hf][k]) {
                const float4 lo = v[hf][k][0], hi = v[hf][k][1];
                pk = make_uint4(pack_bf16(fmaf(lo.x, sc[0], sh[0]), fmaf(lo.y, sc[1], sh[1])),
                                pack_bf16(fmaf(lo.z, sc[2], sh[2]), fmaf(lo.w, sc[3], sh[3])),
                                pack_bf16(fmaf(hi.x, sc[4], sh[4]), fmaf(hi.y, sc[5], sh[5])),
                                pack_bf16(fmaf(hi.z, sc[6], sh[6]), fmaf(hi.w, sc[7], sh[7])));
              }
              *reinterpret_cast<uint4*>(dst0 + (size_t)hf * SRB + (size_t)(ptid + k * T2_PROD_THREADS) * 16) = pk;
            }
        } else {
          // one input channel: the K-chunk of a row is its own voxel and the next span_w voxels along w
          float v[2][T2_MAX_CHUNK][3];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int ip = qd0 + pl.lo_d + 2 * P + hf;
            const bool p_ok = ip >= 0 && ip < g.inD;
            const float* base = in_n + (size_t)(p_ok ? ip : 0) * plane_in;
#pragma unroll
            for (int k = 0; k < T2_MAX_CHUNK; ++k)
#pragma unroll
              for (int e = 0; e < 3; ++e) {
                v[hf][k][e] = 0.f;
                if (p_ok && goff[k] >= 0 && e < wlim[k]) v[hf][k][e] = fmaf(__ldg(base + goff[k] + e), sc[0], sh[0]);
              }
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int k = 0; k < T2_MAX_CHUNK; ++k) {
              if (goff[k] == -1) continue;
              const uint4 pk = make_uint4(pack_bf16(v[hf][k][0], v[hf][k][1]), pack_bf16(v[hf][k][2], 0.f), 0u, 0u);
              *reinterpret_cast<uint4*>(dst0 + (size_t)hf * SRB + (size_t)(ptid + k * T2_PROD_THREADS) * 16) = pk;
            }
        }
        fence_async_smem();
        mbar_arrive(smem_u32(&full_bar[slot]));
      }
      pair_base += npairs;
      block_base += nblocks;
    }
  } else if (lane == 0) {
    // ================================================================ MMA issuer (one thread)
    const uint32_t ring_addr = smem_u32(ring), w_addr = smem_u32(wts);
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      const int dc = col % pl.ndchunks;
      const int qd0 = dc * pl.dchunk, qd1 = min(g.qD, qd0 + pl.dchunk);
      const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
      for (int b = 0; b < nblocks; ++b) {
        const int bg = block_base + b, buf = bg & 1;
        mbar_wait(smem_u32(&acce_bar[buf]), (uint32_t)(((bg >> 1) & 1) ^ 1));
        tc_fence_after();
        for (int p = 0; p < pl.NPAIR; ++p) {
          const int G = pair_base + b * H2 + p, slot = G % pl.R, use = G / pl.R;
          mbar_wait(smem_u32(&full_bar[slot]), (uint32_t)(use & 1));
          tc_fence_after();
          const uint32_t a_slot = ring_addr + (uint32_t)slot * (uint32_t)PAIRB;
          for (int m = pl.pair_begin[p]; m < pl.pair_begin[p + 1]; ++m) {
            const T2Mma mm = pl.mma[m];
            const uint64_t bdesc = umma_desc(w_addr + (uint32_t)mm.b_off16 * 16u, 128u, 256u);
            const uint32_t idesc = umma_idesc_m128((uint32_t)mm.n8 << 3);
            for (int rb = 0; rb < pl.nrb; ++rb) {
              const uint64_t adesc = umma_desc(a_slot + (uint32_t)(rb * 128 + mm.a_shift) * 16u, (uint32_t)SRB, 128u);
              umma_bf16(tmem_base + (uint32_t)((buf * pl.nrb + rb) * pl.ACCW + mm.dcol), adesc, bdesc, idesc, 1u);
            }
          }
          if (p < H2 || b == nblocks - 1) umma_commit(smem_u32(&empty_bar[slot]));
        }
        umma_commit(smem_u32(&accf_bar[buf]));
      }
      pair_base += nblocks * H2 + (pl.NPAIR - H2);
      block_base += nblocks;
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == T2_EPI_WARPS + T2_PROD_WARPS) {
    tc_fence_after();
    switch (pl.tmem_cols) {
      case 32: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base)); break;
      case 64: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base)); break;
      case 128: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base)); break;
      case 256: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base)); break;
      default: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base)); break;
    }
  }
}

// ------------------------------------------------------------------------------- host planner
static constexpr int kT2SmemBudget = 214 * 1024;

static bool t2_build_plan(int cin, int cout, const Geom& g, T2Plan& pl) {
  if (g.sin != 1 || g.ntaps < 1) return false;
  if (cin != 1 && cin != 8) return false;
  if (cout != 1 && cout != 8 && cout != 16) return false;
  int lo[3] = {127, 127, 127}, hi[3] = {-127, -127, -127};
  for (int t = 0; t < g.ntaps; ++t) {
    const int o[3] = {g.taps[t].dd, g.taps[t].dh, g.taps[t].dw};
    for (int i = 0; i < 3; ++i) { lo[i] = o[i] < lo[i] ? o[i] : lo[i]; hi[i] = o[i] > hi[i] ? o[i] : hi[i]; }
  }
  pl.lo_d = lo[0]; pl.lo_h = lo[1]; pl.lo_w = lo[2];
  pl.span_d = hi[0] - lo[0]; pl.span_h = hi[1] - lo[1]; pl.span_w = hi[2] - lo[2];
  if (pl.span_d > 2 || pl.span_h > 2 || pl.span_w > 2) return false;
  int lut[45];
  for (int i = 0; i < 45; ++i) lut[i] = -1;
  for (int t = 0; t < g.ntaps; ++t)
    lut[(g.taps[t].dd - lo[0]) * 9 + (g.taps[t].dh - lo[1]) * 3 + (g.taps[t].dw - lo[2])] = g.taps[t].widx;

  pl.PW = cin == 1 ? g.qW : g.qW + pl.span_w;
  pl.RTOT = g.qH * pl.PW;
  pl.OB = cout == 1 ? 16 : 8;
  pl.NPAIR = (pl.OB + pl.span_d + 1) / 2;
  pl.ACCW = pl.OB * cout;
  if (pl.NPAIR > T2_MAX_PAIR) return false;

  // MMA list + weight blocks (deduplicated: interior pairs share one shift-invariant block)
  struct Key { int rel, nj, dh, dw, off16; };
  Key keys[T2_MAX_BLK];
  int nblk = 0, nmma = 0, woff = 0;
  for (int p = 0; p < pl.NPAIR; ++p) {
    pl.pair_begin[p] = nmma;
    const int j0 = cout == 1 ? 0 : (2 * p - 2 > 0 ? 2 * p - 2 : 0);
    const int j1 = cout == 1 ? pl.OB - 1 : (2 * p + 1 < pl.OB - 1 ? 2 * p + 1 : pl.OB - 1);
    if (j1 < j0) continue;
    const int nj = j1 - j0 + 1;
    for (int dh = 0; dh <= pl.span_h; ++dh)
      for (int dw = 0; dw <= (cin == 1 ? 0 : pl.span_w); ++dw) {
        bool any = false;                       // does the block hold any tap?
        for (int chunk = 0; chunk < 2 && !any; ++chunk)
          for (int j = j0; j <= j1 && !any; ++j) {
            const int ddr = 2 * p + chunk - j;
            if (ddr < 0 || ddr > pl.span_d) continue;
            for (int e = 0; e <= (cin == 1 ? pl.span_w : 0); ++e)
              if (lut[ddr * 9 + dh * 3 + (cin == 1 ? e : dw)] >= 0) any = true;
          }
        if (!any) continue;
        int found = -1;
        for (int k = 0; k < nblk; ++k)
          if (keys[k].rel == 2 * p - j0 && keys[k].nj == nj && keys[k].dh == dh && keys[k].dw == dw) found = k;
        if (found < 0) {
          if (nblk >= T2_MAX_BLK) return false;
          found = nblk++;
          keys[found] = Key{2 * p - j0, nj, dh, dw, woff >> 4};
          pl.blk[found].i0 = (int8_t)(2 * p); pl.blk[found].j0 = (int8_t)j0; pl.blk[found].nj = (int8_t)nj;
          pl.blk[found].dh = (int8_t)dh; pl.blk[found].dw = (int8_t)dw;
          pl.blk_off16[found] = (uint16_t)(woff >> 4);
          woff += nj * cout * 32;
        }
        if (nmma >= T2_MAX_MMA) return false;
        T2Mma& m = pl.mma[nmma++];
        m.a_shift = (uint16_t)(dh * pl.PW + (cin == 1 ? 0 : dw));
        m.b_off16 = (uint16_t)keys[found].off16;
        m.n8 = (uint8_t)((nj * cout) >> 3);
        m.dcol = (uint8_t)(j0 * cout);
      }
  }
  pl.pair_begin[pl.NPAIR] = nmma;
  pl.nblk = nblk;
  pl.wbytes = (woff + 1023) & ~1023;

  // rows per tile: as many 128-row blocks as TMEM (2 buffers) and shared memory allow
  int nrb = 512 / (2 * pl.ACCW);
  if (nrb > 4) nrb = 4;
  const int need = (pl.RTOT + 127) / 128;
  if (nrb > need) nrb = need;
  for (;; --nrb) {
    if (nrb < 1) return false;
    pl.nrb = nrb;
    pl.TR = 128 * nrb;
    pl.SR = pl.TR + pl.span_h * pl.PW + (cin == 1 ? 0 : pl.span_w);
    pl.SR = (pl.SR + 7) & ~7;
    pl.R = pl.NPAIR + 3;
    if (pl.R > T2_MAX_RING) pl.R = T2_MAX_RING;
    if (pl.SR > T2_MAX_CHUNK * T2_PROD_THREADS) continue;
    while (pl.R > pl.NPAIR + 1 && (size_t)pl.R * 2 * pl.SR * 16 + pl.wbytes > (size_t)kT2SmemBudget) --pl.R;
    if ((size_t)pl.R * 2 * pl.SR * 16 + pl.wbytes <= (size_t)kT2SmemBudget) break;
  }
  pl.ntiles = (pl.RTOT + pl.TR - 1) / pl.TR;
  int tc = 32;
  while (tc < 2 * pl.nrb * pl.ACCW) tc <<= 1;
  if (tc > 512) return false;
  pl.tmem_cols = tc;

  // split columns along d until the persistent grid has at least ~2 columns per SM
  const int nblocks_all = (g.qD + pl.OB - 1) / pl.OB;
  const long long cols = (long long)g.N * pl.ntiles;
  const long long want = 2LL * vg_sm_count();
  int nch = (int)((want + cols - 1) / cols);
  if (nch > nblocks_all) nch = nblocks_all;
  if (nch < 1) nch = 1;
  pl.dchunk = ((nblocks_all + nch - 1) / nch) * pl.OB;
  pl.ndchunks = (g.qD + pl.dchunk - 1) / pl.dchunk;
  return true;
}

bool tc2_supported(int cin, int cout, const Geom& g) {
  T2Plan pl{};
  return t2_build_plan(cin, cout, g, pl);
}

// human-readable plan (vg_conv_describe): tile shape, ring, MMA list size, shared memory
int tc2_describe(int cin, int cout, const Geom& g, char* buf, size_t cap) {
  T2Plan pl{};
  if (!t2_build_plan(cin, cout, g, pl)) return 0;
  const long long cols = (long long)g.N * pl.ntiles * pl.ndchunks;
  return snprintf(buf, cap,
                  "tc2 cin=%d cout=%d q=(%d,%d,%d) taps=%d PW=%d RTOT=%d TR=%d ntiles=%d SR=%d OB=%d NPAIR=%d R=%d ACCW=%d "
                  "tmem=%d nmma=%d nblk=%d wbytes=%d smem=%zu dchunk=%d ndchunks=%d cols=%lld",
                  cin, cout, g.qD, g.qH, g.qW, g.ntaps, pl.PW, pl.RTOT, pl.TR, pl.ntiles, pl.SR, pl.OB, pl.NPAIR, pl.R,
                  pl.ACCW, pl.tmem_cols, pl.pair_begin[pl.NPAIR], pl.nblk, pl.wbytes,
                  (size_t)pl.R * 2 * pl.SR * 16 + pl.wbytes, pl.dchunk, pl.ndchunks, cols);
}

template <int CIN, int COUT>
static int launch_tc2_t(const Geom& g, const GatherArgs& a, const T2Plan& pl, cudaStream_t st) {
  const size_t smem = (size_t)pl.R * 2 * pl.SR * 16 + pl.wbytes;
  VG_CUDA(cudaFuncSetAttribute(tc2_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long cols = (long long)g.N * pl.ntiles * pl.ndchunks;
  const int sms = vg_sm_count();
  const unsigned grid = (unsigned)(cols < sms ? cols : sms);
  tc2_kernel<CIN, COUT><<<grid, T2_THREADS, smem, st>>>(g, a, pl);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

int launch_tc2_gather(int cin, int cout, const Geom& g, const GatherArgs& a, cudaStream_t st) {
  T2Plan pl{};
  if (!t2_build_plan(cin, cout, g, pl)) { set_error("plane-folded tensor-core path: unsupported geometry"); return VG_EINVAL; }
  if (cin == 1 && cout == 8) return launch_tc2_t<1, 8>(g, a, pl, st);
  if (cin == 1 && cout == 16) return launch_tc2_t<1, 16>(g, a, pl, st);
  if (cin == 8 && cout == 1) return launch_tc2_t<8, 1>(g, a, pl, st);
  if (cin == 8 && cout == 8) return launch_tc2_t<8, 8>(g, a, pl, st);
  if (cin == 8 && cout == 16) return launch_tc2_t<8, 16>(g, a, pl, st);
  set_error("plane-folded tensor-core path: unsupported channel pair (%d,%d)", cin, cout);
  return VG_EINVAL;
}

}  // namespace vg
