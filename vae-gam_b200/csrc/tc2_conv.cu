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
// inside K); inputs with 16 channels put the two 8-channel halves of ONE plane into K.  A gather
// with input stride 2 (Conv3d stride 2, the data gradient of a stride-2 ConvTranspose3d) stages
// each input plane as its four (h,w)-parity sub-grids, so that every tap is again a unit-stride
// row shift inside one sub-grid.  A stride-2 transposed layer is ONE launch: its (up to 8) output-parity phases share
// the staged input planes and differ only in their MMA lists, weight blocks and output offsets.
//
// Roles (512 threads, one persistent CTA per SM), ES = 1 or 2 epilogue sets:
//   warps [0, 4*ES)      epilogue: tcgen05.ld -> bias / activation / BatchNorm statistics / backward
//                        masks -> global; re-zero the accumulator columns (all MMAs accumulate)
//   warps [4*ES, 12)     producers: fp32 global -> BatchNorm fold -> bf16 -> ring of plane pairs
//   warps 12..15         one per 128-row block of the tile: an elected lane issues the block's
//                        tcgen05.mma list of a (block, phase)
// Pipelines: full/empty mbarriers per ring slot (producers <-> MMA via tcgen05.commit), and
// full/empty per accumulator buffer (MMA <-> epilogue), accumulators double-buffered in TMEM.
#include <cuda.h>          // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint: no libcuda link)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "conv_geom.cuh"
#include "tc_common.cuh"

namespace vg {

constexpr int T2_THREADS = 16 * 32, T2_MMA_WARP = 12;   // warps 12..15: one MMA issuer per 128-row block
constexpr int T2_MAX_MMA = 224, T2_MAX_BLK = 224, T2_MAX_PAIR = 10, T2_MAX_RING = 16, T2_MAX_PH = 8;

// One tcgen05.mma of the list, pre-digested on the host so that the issuing lane only adds bases
// (the table sits in the kernel parameters = constant bank, read straight into uniform registers).
struct T2Mma {
  uint32_t a_shift;    // row shift of the A window inside the staged plane (16-byte units)
  uint32_t b_lo;       // low descriptor word of the weight block relative to the weight area: offset/16 | LBO field
  uint32_t idesc;      // instruction descriptor (M = 128, N = nj * cout)
  uint32_t dcol;       // first accumulator column
};
struct T2Blk {
  int8_t i0, j0, nj, dh, dw, ph, pad[2];   // window plane of K-chunk 0; output planes [j0, j0+nj); taps relative to lo_*
};
struct T2Phase {
  int16_t qD, qH, qW;
  int8_t rD, rH, rW;
  uint8_t pad;
  uint8_t pair_begin[T2_MAX_PAIR + 1];
  uint8_t pad2;
};
struct T2Plan {
  int PW, RTOT, TR, ntiles, SR, OB, NPAIR, R, nrb, ACCW, tmem_cols;
  int lo_d, lo_h, lo_w, span_d, span_h, span_w;
  int qDmax, dchunk, ndchunks;
  int nph, nmma, nblk, wbytes;
  int sd;                      // input stride of the gather (1 or 2)
  int pps;                     // input planes per ring slot: 2 (1 or 8 channels) or 1 (16 channels: K = the plane's two halves)
  int nsg;                     // (h,w)-parity sub-grids per staged plane: 1, or 4 when sd == 2
  int ppb;                     // ring slots consumed per block of OB output planes
  // TMA-direct staging (bf16 channels-last input with 8 channels, unit input stride): a tile is HB whole h-lines of
  // the row frame (rows = HB * PW), a ring slot is two TMA boxes [8 ch][PW][HB + halo lines] — exactly the staged
  // operand layout, out-of-range voxels zero-filled by the TMA unit.  hb == 0: staging by the producer warps.
  int hb, box_h;
  // saved-activation staging of the epilogue (bf16 aux with 8 channels whose rows are the plane's voxels in order:
  // 1-channel gathers, pitch == output width): the 128-row x 16-byte chunks an accumulator unit needs are fetched
  // one unit ahead with bulk asynchronous copies into a double buffer behind the weights.  0: loaded by the threads.
  int auxs;
  int groups;                  // BatchNorm groups of the launch when the fold moves into the weights (TMA mode): a CTA serves ONE group
  T2Phase ph[T2_MAX_PH];
  T2Mma mma[T2_MAX_MMA];
  T2Blk blk[T2_MAX_BLK];
  uint16_t blk_off16[T2_MAX_BLK];
};

// roles per (cin, cout): the side that moves more bytes gets more warps
__host__ __device__ constexpr int t2_epi_sets(int cin, int cout, int sd) { return (sd == 2 || (cin == 8 && cout == 1)) ? 1 : 2; }

constexpr int T2_MAX_CLS = 5;      // tap-validity classes per output dimension (k <= 5 would give more; TMA mode has k = 3)
// producer warps per (cin, epilogue sets): one input channel moves few bytes per row but waits for them — six producer
// warps (two MMA issuer warps then share the four row blocks) with shorter per-thread chunk lists.  (A second pair of
// look-ahead in the producers' registers measured slower: 0.38-0.41 vs 0.34 ms.)
__host__ __device__ constexpr int t2_prod_warps(int cin, int es) { return cin == 1 ? 6 : 12 - 4 * es; }
__host__ __device__ constexpr int t2_max_chunk(int cin, int es, int sd) {
  if (cin == 1) return 4;
  return sd == 2 ? 1 : (cin == 8 ? (es == 1 ? 2 : 4) : (cin == 16 ? (es == 1 ? 2 : 3) : (es == 1 ? 3 : 5)));
}

template <int CIN, int COUT, int SD, bool TMA>
__global__ void __launch_bounds__(T2_THREADS, 1)
tc2_kernel(const __grid_constant__ Geom g, const GatherArgs a, const __grid_constant__ T2Plan pl,
           const __grid_constant__ CUtensorMap tmap) {
  static_assert(!TMA || CIN == 8, "TMA-direct staging: 8 bf16 channels = one 16-byte word per voxel");
  // TMA mode has no producer warps: 12 epilogue warps (3 sets), warp 12 issues the TMA loads, warps 13..15 the MMAs
  constexpr int ES = TMA ? 3 : t2_epi_sets(CIN, COUT, SD);
  constexpr int NSG = SD == 2 ? 4 : 1;
  constexpr int EPI_WARPS = 4 * ES, PROD_WARPS = TMA ? 0 : t2_prod_warps(CIN, ES), PT = PROD_WARPS * 32;
  constexpr int MMA_WARP0 = TMA ? 13 : EPI_WARPS + PROD_WARPS, NMW = 16 - MMA_WARP0;      // MMA issuer warps
  constexpr int MAXC = t2_max_chunk(CIN, ES, SD);
  constexpr bool PIPE = SD == 1 && CIN != 16 && ((CIN == 1) || (ES == 1));   // register double-buffering of the staged pair
  constexpr int NJ = 16 / COUT;                        // output planes per epilogue item (16 TMEM columns)

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[T2_MAX_RING], empty_bar[T2_MAX_RING], accf_bar[2], acce_bar[2], auxf_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int lut[T2_MAX_PH][45];
  __shared__ float s_bias[16];
  // TMA mode: the BatchNorm shift cannot be added to the staged operand any more, so it becomes a bias that
  // depends on which taps fall inside the input — a (class_d, class_h, class_w) table built once per CTA
  __shared__ float s_btab[TMA ? T2_MAX_CLS * T2_MAX_CLS * T2_MAX_CLS * COUT : 1];
  __shared__ float s_tapsum[TMA ? 27 * COUT : 1];
  __shared__ uint8_t s_cls[3][64], s_cmask[3][8];
  __shared__ int s_ncls[3], s_midrange[2];     // [lo, hi): the output planes around the middle whose d-taps are all inside

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SRB = pl.SR * 16;                   // bytes per staged plane
  const int PAIRB = 2 * NSG * SRB;              // ring slot: [K half][sub-grid][rows]
  uint8_t* ring = smem;
  uint8_t* wts = smem + (size_t)pl.R * PAIRB;
  const int H2 = pl.ppb;                        // ring slots consumed per block

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < pl.R; ++s) {
      mbar_init(smem_u32(&full_bar[s]), TMA ? 1u : (uint32_t)PT);
      mbar_init(smem_u32(&empty_bar[s]), (uint32_t)min(pl.nrb, NMW));     // one commit per MMA issuer warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&accf_bar[b]), (uint32_t)min(pl.nrb, NMW));
      mbar_init(smem_u32(&acce_bar[b]), EPI_WARPS * 32);
      mbar_init(smem_u32(&auxf_bar[b]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = tid; i < T2_MAX_PH * 45; i += T2_THREADS) (&lut[0][0])[i] = -1;
  if (tid < 16) s_bias[tid] = (a.bias && tid < COUT) ? __ldg(a.bias + tid) : 0.f;
  if (warp == T2_MMA_WARP) {
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
    const Tap tp = g.taps[tid];                 // pad_ = phase of the tap
    lut[tp.pad_][(tp.dd - pl.lo_d) * 9 + (tp.dh - pl.lo_h) * 3 + (tp.dw - pl.lo_w)] = tp.widx;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // Column schedule.  Default: columns (image, tile, d-chunk) dealt round-robin to the persistent CTAs.  TMA mode
  // with a BatchNorm fold: the scale lives in the weight tiles, so a CTA is pinned to ONE statistics group
  // (group = blockIdx % groups) and walks that group's columns only.
  const int cols_per_img = pl.ntiles * pl.ndchunks;
  const bool grouped = TMA && pl.groups > 1;
  const int my_grp = grouped ? (int)(blockIdx.x % pl.groups) : 0;
  const int col_first = grouped ? (int)(blockIdx.x / pl.groups) : (int)blockIdx.x;
  const int col_step = grouped ? (int)((gridDim.x - my_grp + pl.groups - 1) / pl.groups) : (int)gridDim.x;
  const int col_count = grouped ? g.group_size * cols_per_img : g.N * cols_per_img;
  const int col_img0 = grouped ? my_grp * g.group_size : 0;
  const int col_imgs = grouped ? g.group_size : g.N;
  const int per_dc = col_imgs * pl.ntiles;                            // columns per d-chunk index
  // Columns inside a d-chunk are TILE major (all images of tile 0, then tile 1, ...): the last tile of the row frame may
  // hold fewer live 128-row blocks (t2_live_rb) and is then cheaper, so a CTA's round-robin share has to mix the tiles.
  auto live_rb = [&](int t) { return TMA ? pl.nrb : min(pl.nrb, (pl.RTOT - t * pl.TR + 127) >> 7); };
  const float* fold_scale = (TMA && a.in_scale) ? a.in_scale + (size_t)my_grp * CIN : nullptr;
  const float* fold_shift = (TMA && a.in_scale) ? a.in_shift + (size_t)my_grp * CIN : nullptr;

  // Toeplitz weight blocks: element (n, k) of a block = W[tap(i - j, dh, dw)][ci][co], canonical K-major layout.
  // The raw weights (a few KB, contiguous) are first copied into the still unused ring area, coalesced.
  {
    float* wraw = reinterpret_cast<float*>(ring);
    const int nraw = min(g.wst_ci, g.wst_co) * CIN * COUT;      // kernel volume * cin * cout
    for (int i = tid; i < nraw; i += T2_THREADS) wraw[i] = __ldg(a.w + i);
    __syncthreads();
    for (int b = 0; b < pl.nblk; ++b) {
      const T2Blk bk = pl.blk[b];
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(wts + (size_t)pl.blk_off16[b] * 16);
      const int nel = bk.nj * COUT * 16;
      for (int e = tid; e < nel; e += T2_THREADS) {
        const int n = e >> 4, k = e & 15;
        const int chunk = k >> 3, el = k & 7;
        const int j = bk.j0 + n / COUT, co = n % COUT;
        // bk.i0 = window plane of K half 0; the second half is the next plane (1 / 8 channels) or the same plane's
        // channels 8..15; an output plane j reads window plane SD*j + kd
        const int ddr = bk.i0 + (CIN == 16 ? 0 : chunk) - SD * j;
        int ci, dwr;
        if constexpr (CIN == 1) { ci = 0; dwr = el; } else if constexpr (CIN == 16) { ci = chunk * 8 + el; dwr = bk.dw; }
        else { ci = el; dwr = bk.dw; }
        float w = 0.f;
        if (ddr >= 0 && ddr <= pl.span_d && dwr <= pl.span_w) {
          const int widx = lut[bk.ph][ddr * 9 + bk.dh * 3 + dwr];
          if (widx >= 0) w = wraw[widx * g.wst_t + ci * g.wst_ci + co * g.wst_co];
          if constexpr (TMA) { if (fold_scale) w *= __ldg(fold_scale + ci); }     // BatchNorm scale folded into the weights
        }
        dst[(((n >> 3) * 256 + chunk * 128 + (n & 7) * 16) >> 1) + el] = __float2bfloat16(w);
      }
    }
  }
  if constexpr (TMA) {
    // ---- per-dimension tap-validity classes and the bias table (single phase, unit strides: the only layers folded)
    const float* wraw = reinterpret_cast<const float*>(ring);
    if (tid < 3) {
      const int In = tid == 0 ? g.inD : (tid == 1 ? g.inH : g.inW), Out = tid == 0 ? g.outD : (tid == 1 ? g.outH : g.outW);
      const int lo = tid == 0 ? pl.lo_d : (tid == 1 ? pl.lo_h : pl.lo_w), span = tid == 0 ? pl.span_d : (tid == 1 ? pl.span_h : pl.span_w);
      int ncls = 0;
      for (int o = 0; o < Out && o < 64; ++o) {
        int m = 0;
        if (fold_shift)
          for (int k = 0; k <= span; ++k) { const int i = o + lo + k; if (i >= 0 && i < In) m |= 1 << k; }
        int c = -1;
        for (int q = 0; q < ncls; ++q) if (s_cmask[tid][q] == m) c = q;
        if (c < 0 && ncls < T2_MAX_CLS) { c = ncls; s_cmask[tid][ncls++] = (uint8_t)m; }
        s_cls[tid][o] = (uint8_t)(c < 0 ? 0 : c);
      }
      s_ncls[tid] = ncls < 1 ? 1 : ncls;
      if (tid == 0) {
        const int lim = Out < 64 ? Out : 64, mid = lim / 2;
        int a0 = mid, a1 = mid + 1;
        while (a0 > 0 && s_cls[0][a0 - 1] == s_cls[0][mid]) --a0;
        while (a1 < lim && s_cls[0][a1] == s_cls[0][mid]) ++a1;
        s_midrange[0] = a0; s_midrange[1] = a1;
      }
    }
    for (int e = tid; e < 27 * COUT; e += T2_THREADS) {       // per-tap sum over input channels of shift * W
      const int t = e / COUT, co = e - t * COUT;
      float sacc = 0.f;
      const int widx = (t < 45 && pl.nph == 1) ? lut[0][t] : -1;   // t = kd * 9 + kh * 3 + kw (relative to lo_*)
      if (fold_shift && widx >= 0)
        for (int ci = 0; ci < CIN; ++ci) sacc = fmaf(__ldg(fold_shift + ci), wraw[widx * g.wst_t + ci * g.wst_ci + co * g.wst_co], sacc);
      s_tapsum[e] = sacc;
    }
    __syncthreads();
    const int nd_ = s_ncls[0], nh_ = s_ncls[1], nw_ = s_ncls[2];
    for (int e = tid; e < nd_ * nh_ * nw_ * COUT; e += T2_THREADS) {
      const int co = e % COUT, cw = (e / COUT) % nw_, ch = (e / (COUT * nw_)) % nh_, cd = e / (COUT * nw_ * nh_);
      float bsum = s_bias[co];
      const int md = s_cmask[0][cd], mh = s_cmask[1][ch], mw = s_cmask[2][cw];
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw)
            if (((md >> kd) & 1) && ((mh >> kh) & 1) && ((mw >> kw) & 1)) bsum += s_tapsum[(kd * 9 + kh * 3 + kw) * COUT + co];
      s_btab[e] = bsum;
    }
    __syncthreads();
    // the raw weights are no longer needed: clear the ring (rows beyond a TMA box are read by the last row block's
    // MMAs and must at least be finite)
    for (int i = tid; i < (pl.R * PAIRB) / 16; i += T2_THREADS) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp < 4) {                               // all MMAs accumulate: start from zero
    for (int c = 0; c < pl.tmem_cols; c += 16) tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c);
    tmem_st_wait();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  int pair_base = 0, acc_base = 0;              // running counters, identical in every role

  // Register reallocation between the warpgroups (setmaxnreg): the epilogue warps are the ones that need registers
  // (accumulator slice, look-ahead buffers of the saved activation, statistics), the TMA / MMA issuers need few.
  // Budget: 64 K registers per SM = sum over the four warpgroups of 128 threads x their limit.
  // (producer warps keep a register-staged plane pair: 8- and 16-channel inputs need ~100 registers there, the
  // 1-channel gathers less — and their epilogue, convt5's data gradient, the most)
  constexpr int REG_OTHER = TMA ? 56 : (ES == 1 ? 112 : (CIN == 1 ? 88 : 104));
  constexpr int REG_EPI = TMA ? 152 : (ES == 1 ? 176 : (CIN == 1 ? 168 : 152));
  static_assert(ES * REG_EPI + (4 - ES) * REG_OTHER <= 512, "register budget");
  if (warp < EPI_WARPS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_EPI));
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_OTHER));

  if (warp < EPI_WARPS) {
    // ================================================================ epilogue warps
    const bool want_stats = a.stats != nullptr, want_bn = a.aux_mode == 2, bn_apply = a.aux_mode == 3;
    const int eset = warp >> 2, etid = tid & 127;            // TMEM lane = etid
    const int ipr = pl.ACCW >> 4;                            // 16-column items per row block
    const int ipr_shift = 31 - __clz(ipr);                   // ACCW is 16, 32, 64 or 128: ipr is a power of two
    const uint32_t plane_out = (uint32_t)(g.outH * g.outW * COUT);   // offsets inside ONE image fit 32 bits
    const uint32_t plane_step = (uint32_t)g.sout * plane_out;
    const bool relu = a.act == VG_ACT_RELU, sigm = a.act == VG_ACT_SIGMOID;
    // the epilogue's switches as one key (see `process`); combinations a channel count cannot have map to the run-time path
    const int epi_key = (relu ? 1 : (sigm ? 2 : 0)) | ((a.aux_mode & 3) << 2) | (want_stats ? 0x10 : 0) |
                        (a.out_bf16 ? 0x20 : 0) | (a.aux_bf16 ? 0x40 : 0);
    // Item = 16 accumulator columns (NJ output planes x COUT) of one 128-row block.  Everything that depends on
    // the ROW only (division by the pitch, validity, offset inside the plane, bias class) is computed once per
    // row block and reused by the items that share it.
    struct Item { int rb, k, qd, bofs; bool row_ok; uint32_t off; };
    const int cls_hw = TMA ? s_ncls[1] * s_ncls[2] * COUT : 0;
    // bias of a voxel whose taps are all inside the input (the bulk): kept in registers; only border voxels (or border
    // planes) read the class table.  Interior class = the class of the middle coordinate.
    int cls_mid[3] = {0, 0, 0};
    constexpr bool BREG = COUT < 8;            // few channels: the bulk bias lives in registers, else it is a broadcast LDS
    float b_mid[BREG ? COUT : 1];
    const float* b_mid_s = s_bias;
    bool one_cls = true;
    if constexpr (TMA) {
      cls_mid[0] = s_cls[0][min(g.outD / 2, 63)]; cls_mid[1] = s_cls[1][min(g.outH / 2, 63)]; cls_mid[2] = s_cls[2][min(g.outW / 2, 63)];
      one_cls = s_ncls[0] * s_ncls[1] * s_ncls[2] == 1;
      b_mid_s = s_btab + cls_mid[0] * cls_hw + (cls_mid[1] * s_ncls[2] + cls_mid[2]) * COUT;
    }
    if constexpr (BREG) {
#pragma unroll
      for (int c = 0; c < COUT; ++c) b_mid[c] = b_mid_s[c];
    }
    const int mid_bofs = (cls_mid[1] * (TMA ? s_ncls[2] : 0) + cls_mid[2]) * COUT;
    const int mid_d0 = TMA ? s_midrange[0] : 0, mid_d1 = TMA ? s_midrange[1] : 0;
    const uint32_t pw_mul = (uint32_t)((0x100000000ULL + (uint64_t)pl.PW - 1) / (uint64_t)pl.PW);   // exact r / PW for r * PW < 2^32
    // ---- saved-activation staging (pl.auxs): thread 0 fetches the aux chunks of the NEXT accumulator unit with bulk
    // asynchronous copies while the current unit is processed; chunk (row block, plane) = 128 rows x 16 bytes, contiguous
    // in global memory because the rows are the plane's voxels in order.
    constexpr bool AUXS_OK = CIN == 1 && COUT == 8 && !TMA;
    const bool auxs = AUXS_OK && pl.auxs != 0;
    const uint32_t aux_s0 = smem_u32(wts) + (uint32_t)pl.wbytes;
    const uint8_t* const aux_sp = wts + pl.wbytes;
    auto aux_issue = [&](int n_, int t_, int qdf, int qde, int bufi) {
      const int plane_vox = g.outH * g.outW;
      const __nv_bfloat16* src_n = reinterpret_cast<const __nv_bfloat16*>(a.aux) + (size_t)n_ * g.out_img;
      uint32_t bytes = 0;
      for (int rbb = 0; rbb < pl.nrb; ++rbb) {
        const int r0 = t_ * pl.TR + rbb * 128;
        const int rows = min(128, pl.RTOT - r0);
        if (rows <= 0) continue;
        for (int j = 0; j < pl.OB; ++j) if (qdf + j < qde) bytes += (uint32_t)rows * 16u;
      }
      const uint32_t bar = smem_u32(&auxf_bar[bufi]);
      mbar_arrive_expect_tx(bar, bytes);
      for (int rbb = 0; rbb < pl.nrb; ++rbb) {
        const int r0 = t_ * pl.TR + rbb * 128;
        const int rows = min(128, pl.RTOT - r0);
        if (rows <= 0) continue;
        for (int j = 0; j < pl.OB; ++j)
          if (qdf + j < qde)
            bulk_g2s(aux_s0 + (uint32_t)(((bufi * pl.nrb + rbb) * pl.OB + j) * 2048), src_n + ((size_t)(qdf + j) * plane_vox + r0) * 8,
                     (uint32_t)rows * 16u, bar);
      }
    };
    if (auxs && tid == 0 && col_first < col_count) {       // first unit of this CTA
      const int dc = col_first / per_dc, t_ = (col_first - dc * per_dc) / col_imgs, n_ = col_img0 + (col_first - dc * per_dc) % col_imgs;
      const int qd0_ = dc * pl.dchunk, qd1_ = min(pl.qDmax, qd0_ + pl.dchunk);
      aux_issue(n_, t_, qd0_, min((int)pl.ph[0].qD, qd1_), 0);
    }
    for (int col = col_first; col < col_count; col += col_step) {
      // d-chunk major: the chunks of an image differ in length (the last one is short), so all the long columns come
      // first and a CTA's round-robin share mixes long and short ones whatever the parity of the grid
      const int dc = col / per_dc, t = (col - dc * per_dc) / col_imgs, n = col_img0 + (col - dc * per_dc) % col_imgs;
      const int qd0 = dc * pl.dchunk, qd1 = min(pl.qDmax, qd0 + pl.dchunk);
      const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
      const int grp = n / g.group_size;
      const int nitems = live_rb(t) * ipr;                   // items per accumulator: dead row blocks are never accumulated, never drained
      // image bases (64-bit once per column); bf16 tensors have the same element offsets at half the size
      float* const out_n = a.out ? (a.out_bf16 ? reinterpret_cast<float*>(reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t)n * g.out_img)
                                               : a.out + (size_t)n * g.out_img) : nullptr;
      const float* const aux_n = a.aux ? (a.aux_bf16 ? reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(a.aux) + (size_t)n * g.out_img)
                                                     : a.aux + (size_t)n * g.out_img) : nullptr;
      float istd[COUT], mistd[COUT], s1[COUT], s2[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        s1[c] = s2[c] = 0.f;
        istd[c] = want_bn ? __ldg(a.aux_istd + grp * COUT + c) : 0.f;
        mistd[c] = want_bn ? __ldg(a.aux_mistd + grp * COUT + c) : 0.f;
        if (bn_apply) {      // fused BatchNorm backward: the three per-channel coefficients reuse these registers
          istd[c] = __ldg(a.aux_coef + (grp * COUT + c) * 3);
          mistd[c] = __ldg(a.aux_coef + (grp * COUT + c) * 3 + 1);
          s2[c] = __ldg(a.aux_coef + (grp * COUT + c) * 3 + 2);
        }
      }
      for (int b = 0; b < nblocks; ++b)
        for (int ph = 0; ph < pl.nph; ++ph) {
          const T2Phase& P = pl.ph[ph];
          const int au = acc_base + b * pl.nph + ph, buf = au & 1;
          const int qd_end = min((int)P.qD, qd1);
          const uint32_t ph_off = (uint32_t)P.rD * plane_out + ((uint32_t)P.rH * (uint32_t)g.outW + (uint32_t)P.rW) * COUT;
          if (auxs && tid == 0) {                 // prefetch the aux chunks of the next unit (same column or the next one)
            int n_ = n, t_ = t, qdf = qd0 + (b + 1) * pl.OB, qde = qd_end;
            bool have = b + 1 < nblocks;
            if (!have && col + col_step < col_count) {
              const int c2 = col + col_step, dc2 = c2 / per_dc;
              t_ = (c2 - dc2 * per_dc) / col_imgs; n_ = col_img0 + (c2 - dc2 * per_dc) % col_imgs;
              qdf = dc2 * pl.dchunk;
              qde = min((int)P.qD, min(pl.qDmax, qdf + pl.dchunk));
              have = true;
            }
            if (have) {
              const int au1 = au + 1;
              mbar_wait(smem_u32(&acce_bar[au1 & 1]), (uint32_t)(((au1 >> 1) & 1) ^ 1));      // unit au - 1 drained that buffer
              aux_issue(n_, t_, qdf, qde, au1 & 1);
            }
          }
          // row cache: valid for the row block `c_rb` of this (column, phase)
          int c_rb = -1, c_bofs = mid_bofs;
          bool c_ok = false;
          uint32_t c_rowoff = 0;
          auto setup = [&](int it, Item& I) {
            I.rb = it >> ipr_shift; I.k = it & (ipr - 1);
            if (I.rb != c_rb) {
              c_rb = I.rb;
              int qh, qw;
              if constexpr (TMA) {                  // tile = hb whole lines of the row frame
                const int rl = I.rb * 128 + etid, lh = (int)__umulhi((uint32_t)rl, pw_mul);
                qh = t * pl.hb + lh; qw = rl - lh * pl.PW;
                c_ok = lh < pl.hb && qh < P.qH && qw < P.qW;
                c_bofs = (c_ok && !one_cls) ? (s_cls[1][min(qh, 63)] * s_ncls[2] + s_cls[2][min(qw, 63)]) * COUT : mid_bofs;
              } else {
                const int r = t * pl.TR + I.rb * 128 + etid;
                qh = (int)__umulhi((uint32_t)r, pw_mul); qw = r - qh * pl.PW;
                c_ok = qh < P.qH && qw < P.qW;
              }
              c_rowoff = ph_off + ((uint32_t)(qh * g.sout) * (uint32_t)g.outW + (uint32_t)(qw * g.sout)) * COUT;
            }
            I.row_ok = c_ok; I.bofs = c_bofs;
            I.qd = qd0 + b * pl.OB + I.k * NJ;
            I.off = (uint32_t)I.qd * plane_step + c_rowoff;
          };
          auto load_aux = [&](const Item& I, float (&ax)[NJ][COUT]) {
            if (a.aux_mode == 0 || !I.row_ok || auxs) return;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              if (I.qd + j >= qd_end) continue;
              const uint32_t off = I.off + (uint32_t)j * plane_step;
              if constexpr (COUT % 8 == 0) {
                if (a.aux_bf16) {                 // saved activation stored as bf16: one 16-byte word per 8 channels
                  const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(aux_n) + off;
#pragma unroll
                  for (int i = 0; i < COUT / 8; ++i) {      // raw words only: unpacked at use, so the load stays in flight
                    const uint4 q = ldg_u4(pb + 8 * i);
                    ax[j][4 * i] = __uint_as_float(q.x); ax[j][4 * i + 1] = __uint_as_float(q.y);
                    ax[j][4 * i + 2] = __uint_as_float(q.z); ax[j][4 * i + 3] = __uint_as_float(q.w);
                  }
                  continue;
                }
              }
              const float* p = aux_n + off;
              if constexpr (COUT % 4 == 0) {
#pragma unroll
                for (int i = 0; i < COUT / 4; ++i) {
                  const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
                  ax[j][4 * i] = v.x; ax[j][4 * i + 1] = v.y; ax[j][4 * i + 2] = v.z; ax[j][4 * i + 3] = v.w;
                }
              } else {
#pragma unroll
                for (int c = 0; c < COUT; ++c) ax[j][c] = __ldg(p + c);
              }
            }
          };
          // q8: optional raw bf16 words of the saved activation (deep-prefetch path, COUT == 8: one 16-byte word per plane)
          // KEY >= 0 (a literal at the inlined call sites of run_items below) fixes the epilogue's switches at compile
          // time: bits 0-1 activation, 2-3 auxiliary mode, 4 statistics, 5 bf16 output, 6 bf16 saved activation — the hot
          // variants; KEY < 0 reads them at run time
          auto process = [&](const int KEY, const Item& I, const float (&ax)[NJ][COUT], const uint4* q8 = nullptr) {
            const bool relu_ = KEY < 0 ? relu : (KEY & 3) == 1, sigm_ = KEY < 0 ? sigm : (KEY & 3) == 2;
            const int auxm_ = KEY < 0 ? a.aux_mode : (KEY >> 2) & 3;
            const bool stats_ = KEY < 0 ? want_stats : ((KEY >> 4) & 1) != 0;
            const bool out16_ = KEY < 0 ? a.out_bf16 != 0 : ((KEY >> 5) & 1) != 0;
            const bool aux16_ = KEY < 0 ? a.aux_bf16 != 0 : ((KEY >> 6) & 1) != 0;
            const uint32_t tcol = tmem_base + ((uint32_t)((etid >> 5) * 32) << 16) +
                                  (uint32_t)((buf * pl.nrb + I.rb) * pl.ACCW + I.k * 16);
            uint32_t rr[16];
            tmem_ld16(tcol, rr);
            tmem_ld_wait();
            tmem_zero16(tcol);                    // drained: ready for the (block, phase) after next
            if (!I.row_ok) return;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              if (I.qd + j >= qd_end) continue;
              const uint32_t off = I.off + (uint32_t)j * plane_step;
              float y[COUT];
              const float* bsrc = BREG ? nullptr : b_mid_s;        // border voxel / border plane: bias from the class table
              if constexpr (TMA) {
                const int od = I.qd + j;
                if (!one_cls && !(od >= mid_d0 && od < mid_d1 && I.bofs == mid_bofs))
                  bsrc = s_btab + s_cls[0][min(od, 63)] * cls_hw + I.bofs;
              }
#pragma unroll
              for (int c = 0; c < COUT; ++c) {
                float bc;
                if constexpr (BREG) bc = bsrc ? bsrc[c] : b_mid[c];
                else bc = bsrc[c];
                float v = __uint_as_float(rr[j * COUT + c]) + bc;
                if (relu_) v = fmaxf(v, 0.f);
                else if (sigm_) v = 1.f / (1.f + __expf(-v));
                y[c] = v;
              }
              if (auxm_ != 0) {
                float av[COUT];
#pragma unroll
                for (int c = 0; c < COUT; ++c) av[c] = ax[j][c];
                if constexpr (COUT % 8 == 0) {
                  if (aux16_) {
#pragma unroll
                    for (int c = 0; c < COUT / 2; ++c) {
                      const uint32_t wv = __float_as_uint(ax[j][c]);
                      av[2 * c] = bf16_lo(wv); av[2 * c + 1] = bf16_hi(wv);
                    }
                  }
                }
                if constexpr (COUT == 8) {
                  if (q8) unpack_bf16x8(q8[j], av);
                }
                if constexpr (AUXS_OK) {
                  if (auxs) {                     // staged chunk (buffer, row block, plane of the block): this thread's row
                    const uint4 q = *reinterpret_cast<const uint4*>(aux_sp + (size_t)(((buf * pl.nrb + I.rb) * pl.OB + I.k * NJ + j) * 2048) +
                                                                     (size_t)etid * 16);
                    unpack_bf16x8(q, av);
                  }
                }
                if (auxm_ == 1) {
#pragma unroll
                  for (int c = 0; c < COUT; ++c) y[c] = av[c] > 0.f ? y[c] : 0.f;
                } else if (auxm_ == 2) {
#pragma unroll
                  for (int c = 0; c < COUT; ++c) {
                    const float xh = fmaf(av[c], istd[c], -mistd[c]);
                    s1[c] += y[c];
                    s2[c] = fmaf(y[c], xh, s2[c]);
                  }
                } else {                          // out = (x > 0) * (A dy + B x + C); its sum = the producer's bias gradient
#pragma unroll
                  for (int c = 0; c < COUT; ++c) {
                    y[c] = av[c] > 0.f ? fmaf(istd[c], y[c], fmaf(mistd[c], av[c], s2[c])) : 0.f;
                    s1[c] += y[c];
                  }
                }
              }
              if (stats_) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                  s1[c] += y[c];
                  s2[c] = fmaf(y[c], y[c], s2[c]);
                }
              }
              if (out_n) {
                if constexpr (COUT % 8 == 0) {
                  if (out16_) {
                    __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(out_n) + off;
#pragma unroll
                    for (int i = 0; i < COUT / 8; ++i) {
                      float f[8];
#pragma unroll
                      for (int e = 0; e < 8; ++e) f[e] = y[8 * i + e];
                      reinterpret_cast<uint4*>(pb)[i] = pack_bf16x8(f);
                    }
                    continue;
                  }
                }
                float* p = out_n + off;
                if constexpr (COUT % 4 == 0) {
#pragma unroll
                  for (int i = 0; i < COUT / 4; ++i)
                    reinterpret_cast<float4*>(p)[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                } else {
#pragma unroll
                  for (int c = 0; c < COUT; ++c) p[c] = y[c];
                }
              }
            }
          };
          // items of this set, the auxiliary loads of the next item in flight while the current one is processed
          Item I0, I1;
          float ax0[NJ][COUT], ax1[NJ][COUT];
          int it = eset;
          bool deep_done = false;
          if constexpr (CIN == 1 && COUT == 8 && !TMA) {      // convt5's data gradient: the one launch this matters for
            // bf16 saved activation with 8 channels = ONE 16-byte word per plane: the registers of the two fp32
            // look-ahead buffers hold THREE items of raw words instead, i.e. the loads of the next two items are in
            // flight while one is processed (the epilogue was waiting on these loads, profiles/r2_tc2_source_hotspots.txt)
            if (a.aux_bf16 && a.aux_mode != 0 && !auxs) {
              deep_done = true;
              Item J0, J1, J2;
              uint4 q0[NJ], q1[NJ], q2[NJ];
              const __nv_bfloat16* const aux16 = reinterpret_cast<const __nv_bfloat16*>(aux_n);
              auto fetch = [&](int it_, Item& J, uint4 (&q)[NJ]) {
                setup(it_, J);
                if (!J.row_ok) return;
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                  if (J.qd + j < qd_end) q[j] = ldg_u4(aux16 + J.off + (uint32_t)j * plane_step);
              };
              if (it < nitems) fetch(it, J0, q0);
              if (it + ES < nitems) fetch(it + ES, J1, q1);
              mbar_wait(smem_u32(&accf_bar[buf]), (uint32_t)((au >> 1) & 1));
              tc_fence_after();
              // The fused BatchNorm-backward apply writing bf16 (no bias, no activation, no statistics) is THE hot
              // instance: a body without the runtime switches of the general epilogue, ~1/3 of its instructions
              const bool lean = bn_apply && a.out_bf16 && out_n && !a.bias && !relu && !sigm && !want_stats;
              auto process3 = [&](const Item& I, const uint4 (&q)[NJ]) {
                const uint32_t tcol = tmem_base + ((uint32_t)((etid >> 5) * 32) << 16) +
                                      (uint32_t)((buf * pl.nrb + I.rb) * pl.ACCW + I.k * 16);
                uint32_t rr[16];
                tmem_ld16(tcol, rr);
                tmem_ld_wait();
                tmem_zero16(tcol);
                if (!I.row_ok) return;
                __nv_bfloat16* const ob = reinterpret_cast<__nv_bfloat16*>(out_n);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                  if (I.qd + j >= qd_end) continue;
                  float av[8], y[8];
                  unpack_bf16x8(q[j], av);
#pragma unroll
                  for (int c = 0; c < 8; ++c) {     // dx = (x > 0) * (A dy + B x + C); its sum = the producer's bias gradient
                    const float v = fmaf(istd[c], __uint_as_float(rr[j * 8 + c]), fmaf(mistd[c], av[c], s2[c]));
                    y[c] = av[c] > 0.f ? v : 0.f;
                    s1[c] += y[c];
                  }
                  *reinterpret_cast<uint4*>(ob + I.off + (uint32_t)j * plane_step) = pack_bf16x8(y);
                }
              };
              auto run = [&](const Item& I, const uint4 (&q)[NJ]) {
                if (lean) process3(I, q);
                else process(-1, I, ax0, q);
              };
              while (it < nitems) {
                if (it + 2 * ES < nitems) fetch(it + 2 * ES, J2, q2);
                run(J0, q0);
                it += ES;
                if (it >= nitems) break;
                if (it + 2 * ES < nitems) fetch(it + 2 * ES, J0, q0);
                run(J1, q1);
                it += ES;
                if (it >= nitems) break;
                if (it + 2 * ES < nitems) fetch(it + 2 * ES, J1, q1);
                run(J2, q2);
                it += ES;
              }
            }
          }
          if (!deep_done) {
          if (it < nitems) { setup(it, I0); load_aux(I0, ax0); }
          mbar_wait(smem_u32(&accf_bar[buf]), (uint32_t)((au >> 1) & 1));
          tc_fence_after();
          if (auxs) mbar_wait(smem_u32(&auxf_bar[buf]), (uint32_t)((au >> 1) & 1));       // this unit's aux chunks have landed
          }
          // the item loop, instantiated once per hot combination of the epilogue switches and once with run-time switches
          auto run_items = [&](const int tag) {
            while (it < nitems) {
              if (it + ES < nitems) { setup(it + ES, I1); load_aux(I1, ax1); }
              process(tag, I0, ax0);
              it += ES;
              if (it >= nitems) break;
              if (it + ES < nitems) { setup(it + ES, I0); load_aux(I0, ax0); }
              process(tag, I1, ax1);
              it += ES;
            }
          };
          if (!deep_done) {
            switch (epi_key) {
              case 0x31: run_items(0x31); break;      // ReLU + next-layer statistics, bf16 out
              case 0x21: run_items(0x21); break;      // ReLU, bf16 out
              case 0x11: run_items(0x11); break;      // ReLU + next-layer statistics, fp32 out (convt2)
              case 0x00: run_items(0x00); break;      // plain fp32 out
              case 0x02: run_items(0x02); break;      // sigmoid, fp32 out
              case 0x64: run_items(0x64); break;      // ReLU mask from a bf16 activation, bf16 out
              case 0x04: run_items(0x04); break;      // ReLU mask, fp32
              case 0x08: run_items(0x08); break;      // BatchNorm-backward sums, fp32
              case 0x48: run_items(0x48); break;      // BatchNorm-backward sums from a bf16 activation
              default: run_items(-1); break;
            }
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&acce_bar[buf]));
        }
      if (bn_apply && a.chan_sum) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float r1 = warp_sum(s1[c]);
          if (lane == 0) atomicAdd(a.chan_sum + c, r1);
        }
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
      acc_base += nblocks * pl.nph;
    }
  } else if (TMA && warp < MMA_WARP0) {
    // ================================================================ TMA producer: one elected lane of one warp
    if constexpr (TMA) {
      if (warp == EPI_WARPS) {
        const uint32_t ring_a = smem_u32(ring);
        const uint32_t slot_tx = (uint32_t)(2 * NSG * pl.box_h * pl.PW * 16);     // 2 * NSG boxes of [8 ch][PW][box_h] bf16
        for (int col = col_first; col < col_count; col += col_step) {
          // d-chunk major: the chunks of an image differ in length (the last one is short), so all the long columns come
      // first and a CTA's round-robin share mixes long and short ones whatever the parity of the grid
      const int dc = col / per_dc, t = (col - dc * per_dc) / col_imgs, n = col_img0 + (col - dc * per_dc) % col_imgs;
          const int qd0 = dc * pl.dchunk, qd1 = min(pl.qDmax, qd0 + pl.dchunk);
          const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
          const int npairs = nblocks * H2 + (pl.NPAIR - H2);
          for (int P = 0; P < npairs; ++P) {
            const int G = pair_base + P, slot = G % pl.R, use = G / pl.R;
            if (use > 0) mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)((use - 1) & 1));
            if (elect_one()) {
              const uint32_t bar = smem_u32(&full_bar[slot]);
              mbar_arrive_expect_tx(bar, slot_tx);
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {    // plane ip may lie outside the input: the whole box is then zero-filled
                if constexpr (SD == 1) {
                  tma_load_4d(ring_a + (uint32_t)slot * (uint32_t)PAIRB + (uint32_t)hf * (uint32_t)SRB, &tmap, bar, 4 * pl.lo_w,
                              t * pl.hb + pl.lo_h, qd0 + pl.lo_d + 2 * P + hf, n);
                } else {                          // input stride 2: one strided box per (h, w)-parity sub-grid
#pragma unroll
                  for (int sg = 0; sg < NSG; ++sg)
                    tma_load_5d(ring_a + (uint32_t)slot * (uint32_t)PAIRB + (uint32_t)(hf * NSG + sg) * (uint32_t)SRB, &tmap, bar, 0,
                                pl.lo_w + (sg & 1), SD * t * pl.hb + pl.lo_h + (sg >> 1), SD * qd0 + pl.lo_d + 2 * P + hf, n);
                }
              }
            }
            __syncwarp();
          }
          pair_base += npairs;
          acc_base += nblocks * pl.nph;
        }
      }
    }
  } else if (!TMA && warp < MMA_WARP0) {
    // ================================================================ producer warps
    const int ptid = tid - EPI_WARPS * 32;
    const bool affine = a.in_scale != nullptr;
    const size_t plane_in = (size_t)g.inH * g.inW * CIN;
    for (int col = col_first; col < col_count; col += col_step) {
      // d-chunk major: the chunks of an image differ in length (the last one is short), so all the long columns come
      // first and a CTA's round-robin share mixes long and short ones whatever the parity of the grid
      const int dc = col / per_dc, t = (col - dc * per_dc) / col_imgs, n = col_img0 + (col - dc * per_dc) % col_imgs;
      const int qd0 = dc * pl.dchunk, qd1 = min(pl.qDmax, qd0 + pl.dchunk);
      const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
      const int npairs = nblocks * H2 + (pl.NPAIR - H2);
      const int grp = n / g.group_size;
      const float* in_n = a.in + (size_t)n * g.in_img;
      // the same tensor when it is stored as bf16 (CIN == 8: a voxel is ONE 16-byte word, already the staged format)
      const __nv_bfloat16* in_nb = reinterpret_cast<const __nv_bfloat16*>(a.in) + (size_t)n * g.in_img;
      const bool in16 = CIN == 8 && a.in_bf16 != 0;
      float sc[CIN], sh[CIN];
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        sc[c] = affine ? __ldg(a.in_scale + grp * CIN + c) : 1.f;
        sh[c] = affine ? __ldg(a.in_shift + grp * CIN + c) : 0.f;
      }
      // the chunks this thread stages are the same for every plane of the column
      int goff[NSG][MAXC]; // float offset of the chunk's first voxel inside its plane; -2 zero chunk; -1 none
      int wlim[MAXC];      // CIN 1: number of in-range w elements of the chunk
#pragma unroll
      for (int sg = 0; sg < NSG; ++sg)
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
          const int s = ptid + k * PT;
          goff[sg][k] = -1;
          if (sg == 0) wlim[k] = 0;
          if (s < pl.SR - (pl.nrb - live_rb(t)) * 128) {      // rows the live row blocks' shifted windows reach
            const int rr = t * pl.TR + s;
            const int hh = rr / pl.PW, ww = rr - hh * pl.PW;
            const int ih = SD * hh + (sg >> 1) + pl.lo_h, iw = SD * ww + (sg & 1) + pl.lo_w;
            goff[sg][k] = -2;
            if (ih >= 0 && ih < g.inH && iw >= 0 && iw < g.inW) {
              goff[sg][k] = (ih * g.inW + iw) * CIN;
              if (sg == 0) wlim[k] = min(pl.span_w + 1, g.inW - iw);
            }
          }
        }
      if constexpr (PIPE) {
        using Buf = typename std::conditional<CIN == 8, float4[2][MAXC][2], float[2][MAXC][3]>::type;
        auto load_pair = [&](int P, Buf& v) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int ip = qd0 + pl.lo_d + 2 * P + hf;
            const bool p_ok = ip >= 0 && ip < g.inD;
            const float* base = in_n + (size_t)(p_ok ? ip : 0) * plane_in;
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
              if constexpr (CIN == 8) {
                if (p_ok && goff[0][k] >= 0) {
                  if (in16) {
                    const uint4 q = ldg_u4(in_nb + (size_t)ip * plane_in + goff[0][k]);
                    v[hf][k][0] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
                  } else {
                    const float4* p = reinterpret_cast<const float4*>(base + goff[0][k]);
                    v[hf][k][0] = __ldg(p);
                    v[hf][k][1] = __ldg(p + 1);
                  }
                }
              } else {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                  v[hf][k][e] = 0.f;
                  if (p_ok && goff[0][k] >= 0 && e < wlim[k]) v[hf][k][e] = __ldg(base + goff[0][k] + e);
                }
              }
            }
          }
        };
        auto store_pair = [&](int P, const Buf& v) {
          const int G = pair_base + P, slot = G % pl.R, use = G / pl.R;
          if (use > 0) mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)((use - 1) & 1));
          uint8_t* dst0 = ring + (size_t)slot * PAIRB;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int ip = qd0 + pl.lo_d + 2 * P + hf;
            const bool p_ok = ip >= 0 && ip < g.inD;
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
              if (goff[0][k] == -1) continue;
              uint4 pk = make_uint4(0u, 0u, 0u, 0u);
              if (p_ok && goff[0][k] >= 0) {
                if constexpr (CIN == 8) {
                  if (in16) {
                    const float4 raw = v[hf][k][0];
                    pk = make_uint4(__float_as_uint(raw.x), __float_as_uint(raw.y), __float_as_uint(raw.z), __float_as_uint(raw.w));
                    if (affine) {
                      float f[8];
                      unpack_bf16x8(pk, f);
#pragma unroll
                      for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
                      pk = pack_bf16x8(f);
                    }
                  } else {
                  const float4 lo = v[hf][k][0], hi = v[hf][k][1];
                  pk = make_uint4(pack_bf16(fmaf(lo.x, sc[0], sh[0]), fmaf(lo.y, sc[1], sh[1])),
                                  pack_bf16(fmaf(lo.z, sc[2], sh[2]), fmaf(lo.w, sc[3], sh[3])),
                                  pack_bf16(fmaf(hi.x, sc[4], sh[4]), fmaf(hi.y, sc[5], sh[5])),
                                  pack_bf16(fmaf(hi.z, sc[6], sh[6]), fmaf(hi.w, sc[7], sh[7])));
                  }
                } else {
                  // one input channel: the K-chunk of a row is its own voxel and the next span_w voxels along w
                  float f[3];
#pragma unroll
                  for (int e = 0; e < 3; ++e) f[e] = e < wlim[k] ? fmaf(v[hf][k][e], sc[0], sh[0]) : 0.f;
                  pk = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], 0.f), 0u, 0u);
                }
              }
              *reinterpret_cast<uint4*>(dst0 + (size_t)hf * SRB + (size_t)(ptid + k * PT) * 16) = pk;
            }
          }
          fence_async_smem();
          mbar_arrive(smem_u32(&full_bar[slot]));
        };
        {
        Buf v0, v1;
        load_pair(0, v0);
        for (int P = 0; P < npairs; P += 2) {
          if (P + 1 < npairs) load_pair(P + 1, v1);
          store_pair(P, v0);
          if (P + 1 < npairs) {
            if (P + 2 < npairs) load_pair(P + 2, v0);
            store_pair(P + 1, v1);
          }
        }
        }
      } else {
        // generic path: one [K half][sub-grid] array at a time, its chunks in flight together
        for (int P = 0; P < npairs; ++P) {
          const int G = pair_base + P, slot = G % pl.R, use = G / pl.R;
          if (use > 0) mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)((use - 1) & 1));
          uint8_t* dst0 = ring + (size_t)slot * PAIRB;
          if constexpr (CIN == 16) {
            const int ip = SD * qd0 + pl.lo_d + P;
            const bool p_ok = ip >= 0 && ip < g.inD;
            const float* base = in_n + (size_t)(p_ok ? ip : 0) * plane_in;
#pragma unroll
            for (int sg = 0; sg < NSG; ++sg) {
              float4 v[MAXC][4];
#pragma unroll
              for (int k = 0; k < MAXC; ++k)
                if (p_ok && goff[sg][k] >= 0) {
                  const float4* p = reinterpret_cast<const float4*>(base + goff[sg][k]);
#pragma unroll
                  for (int q = 0; q < 4; ++q) v[k][q] = __ldg(p + q);
                }
#pragma unroll
              for (int k = 0; k < MAXC; ++k) {
                if (goff[sg][k] == -1) continue;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                  if (p_ok && goff[sg][k] >= 0) {
                    const float4 lo = v[k][2 * hf], hi = v[k][2 * hf + 1];
                    const int c0 = 8 * hf;
                    pk = make_uint4(pack_bf16(fmaf(lo.x, sc[c0], sh[c0]), fmaf(lo.y, sc[c0 + 1], sh[c0 + 1])),
                                    pack_bf16(fmaf(lo.z, sc[c0 + 2], sh[c0 + 2]), fmaf(lo.w, sc[c0 + 3], sh[c0 + 3])),
                                    pack_bf16(fmaf(hi.x, sc[c0 + 4], sh[c0 + 4]), fmaf(hi.y, sc[c0 + 5], sh[c0 + 5])),
                                    pack_bf16(fmaf(hi.z, sc[c0 + 6], sh[c0 + 6]), fmaf(hi.w, sc[c0 + 7], sh[c0 + 7])));
                  }
                  *reinterpret_cast<uint4*>(dst0 + (size_t)(hf * NSG + sg) * SRB + (size_t)(ptid + k * PT) * 16) = pk;
                }
              }
            }
          } else {
            float4 v[2][NSG][MAXC][CIN == 8 ? 2 : 1];
            bool pok[2];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int ip = SD * qd0 + pl.lo_d + 2 * P + hf;
              pok[hf] = ip >= 0 && ip < g.inD;
              const float* base = in_n + (size_t)(pok[hf] ? ip : 0) * plane_in;
#pragma unroll
              for (int sg = 0; sg < NSG; ++sg)
#pragma unroll
                for (int k = 0; k < MAXC; ++k)
                  if (pok[hf] && goff[sg][k] >= 0) {
                    if constexpr (CIN == 8) {
                      if (in16) {
                        const uint4 q = ldg_u4(in_nb + (size_t)ip * plane_in + goff[sg][k]);
                        v[hf][sg][k][0] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
                      } else {
                      const float4* p = reinterpret_cast<const float4*>(base + goff[sg][k]);
                      v[hf][sg][k][0] = __ldg(p);
                      v[hf][sg][k][1] = __ldg(p + 1);
                      }
                    } else {
                      v[hf][sg][k][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                      if (0 < wlim[k]) v[hf][sg][k][0].x = __ldg(base + goff[sg][k]);
                      if (1 < wlim[k]) v[hf][sg][k][0].y = __ldg(base + goff[sg][k] + 1);
                      if (2 < wlim[k]) v[hf][sg][k][0].z = __ldg(base + goff[sg][k] + 2);
                    }
                  }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
              for (int sg = 0; sg < NSG; ++sg)
#pragma unroll
                for (int k = 0; k < MAXC; ++k) {
                  if (goff[sg][k] == -1) continue;
                  uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                  if (pok[hf] && goff[sg][k] >= 0) {
                    if constexpr (CIN == 8) {
                      if (in16) {
                        const float4 raw = v[hf][sg][k][0];
                        pk = make_uint4(__float_as_uint(raw.x), __float_as_uint(raw.y), __float_as_uint(raw.z), __float_as_uint(raw.w));
                        if (affine) {
                          float f[8];
                          unpack_bf16x8(pk, f);
#pragma unroll
                          for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
                          pk = pack_bf16x8(f);
                        }
                      } else {
                      const float4 lo = v[hf][sg][k][0], hi = v[hf][sg][k][1];
                      pk = make_uint4(pack_bf16(fmaf(lo.x, sc[0], sh[0]), fmaf(lo.y, sc[1], sh[1])),
                                      pack_bf16(fmaf(lo.z, sc[2], sh[2]), fmaf(lo.w, sc[3], sh[3])),
                                      pack_bf16(fmaf(hi.x, sc[4], sh[4]), fmaf(hi.y, sc[5], sh[5])),
                                      pack_bf16(fmaf(hi.z, sc[6], sh[6]), fmaf(hi.w, sc[7], sh[7])));
                      }
                    } else {
                      const float4 t = v[hf][sg][k][0];
                      const float f0 = 0 < wlim[k] ? fmaf(t.x, sc[0], sh[0]) : 0.f;
                      const float f1 = 1 < wlim[k] ? fmaf(t.y, sc[0], sh[0]) : 0.f;
                      const float f2 = 2 < wlim[k] ? fmaf(t.z, sc[0], sh[0]) : 0.f;
                      pk = make_uint4(pack_bf16(f0, f1), pack_bf16(f2, 0.f), 0u, 0u);
                    }
                  }
                  *reinterpret_cast<uint4*>(dst0 + (size_t)(hf * NSG + sg) * SRB + (size_t)(ptid + k * PT) * 16) = pk;
                }
          }
          fence_async_smem();
          mbar_arrive(smem_u32(&full_bar[slot]));
        }
      }
      pair_base += npairs;
      acc_base += nblocks * pl.nph;
    }
  } else if (warp >= MMA_WARP0 && warp - MMA_WARP0 < pl.nrb) {
    // ================================================================ MMA warps (one per 128-row block; with fewer
    // issuer warps than row blocks a warp also takes the row blocks NMW further on)
    // The whole warp walks the loops (uniform control flow, waits included); one elected lane
    // issues the tcgen05.mma / tcgen05.commit instructions, so operands stay in uniform registers.
    const int rb0 = warp - MMA_WARP0;
    const uint32_t ring16 = (smem_u32(ring) >> 4), w16 = smem_u32(wts) >> 4;
    const uint32_t lbo_field = ((uint32_t)(NSG * SRB) >> 4) << 16;
    const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = (256u >> 4) | (1u << 14);
    for (int col = col_first; col < col_count; col += col_step) {
      const int dc = col / per_dc;
      const int qd0 = dc * pl.dchunk, qd1 = min(pl.qDmax, qd0 + pl.dchunk);
      const int nblocks = (qd1 - qd0 + pl.OB - 1) / pl.OB;
      const int nrb_t = live_rb((col - dc * per_dc) / col_imgs);     // a warp without a live row block still waits and commits
      for (int b = 0; b < nblocks; ++b)
        for (int ph = 0; ph < pl.nph; ++ph) {
          const int au = acc_base + b * pl.nph + ph, buf = au & 1;
          mbar_wait(smem_u32(&acce_bar[buf]), (uint32_t)(((au >> 1) & 1) ^ 1));
          tc_fence_after();
          const bool last_ph = ph == pl.nph - 1;
          for (int p = 0; p < pl.NPAIR; ++p) {
            const int G = pair_base + b * H2 + p, slot = G % pl.R, use = G / pl.R;
            if (ph == 0) {
              mbar_wait(smem_u32(&full_bar[slot]), (uint32_t)(use & 1));
              tc_fence_after();
            }
            if (elect_one()) {
              const int m1 = pl.ph[ph].pair_begin[p + 1];
              for (int rb = rb0; rb < nrb_t; rb += NMW) {
                const uint32_t d_buf = tmem_base + (uint32_t)((buf * pl.nrb + rb) * pl.ACCW);
                const uint32_t a_lo0 = ((ring16 + (uint32_t)rb * 128u + (uint32_t)slot * ((uint32_t)PAIRB >> 4)) & 0x3FFFu) | lbo_field;
#pragma unroll 4
                for (int m = pl.ph[ph].pair_begin[p]; m < m1; ++m) {
                  const T2Mma e = pl.mma[m];
                  umma_bf16(d_buf + e.dcol, ((uint64_t)a_hi << 32) | (uint64_t)(a_lo0 + e.a_shift),
                            ((uint64_t)b_hi << 32) | (uint64_t)(e.b_lo + w16), e.idesc, 1u);
                }
              }
              if (last_ph && (p < H2 || b == nblocks - 1)) umma_commit(smem_u32(&empty_bar[slot]));
              if (p == pl.NPAIR - 1) umma_commit(smem_u32(&accf_bar[buf]));
            }
            __syncwarp();
          }
        }
      pair_base += nblocks * H2 + (pl.NPAIR - H2);
      acc_base += nblocks * pl.nph;
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == T2_MMA_WARP) {
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
static constexpr int kT2SmemBudget = 212 * 1024;

// gs[0..ng): the gathers of one layer pass that share their input (ng > 1: output-parity phases).
// merged: gs[0] with every phase's taps (Tap::pad_ = phase).
static inline int t2_ceil_div(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

// tma: TMA-direct staging requested (bf16 input; see T2Plan::hb) — falls back to producer-warp staging (returns
// true with pl.hb == 0) when the geometry is outside what the TMA path covers.
// VAEGAM_T2_DEBUG=1: which check of the planner rejected a geometry (stderr)
#define T2_FAIL(id) do { static const bool dbg = getenv("VAEGAM_T2_DEBUG") != nullptr; \
    if (dbg) fprintf(stderr, "[t2 plan] cin=%d cout=%d phases=%d rejected at check %d (line %d)\n", cin, cout, ng, id, __LINE__); } while (0)
static bool t2_build_plan(int cin, int cout, const Geom* gs, int ng, Geom& merged, T2Plan& pl, bool tma = false,
                          bool affine = false, bool auxs = false) {
  pl.hb = 0; pl.box_h = 0; pl.groups = 1; pl.auxs = 0;
  if (ng < 1 || ng > T2_MAX_PH) { T2_FAIL(1); return false; }
  if (cin != 1 && cin != 8 && cin != 16) { T2_FAIL(2); return false; }
  if (cout != 1 && cout != 8 && cout != 16) { T2_FAIL(3); return false; }
  const int sd = gs[0].sin;
  if (sd != 1 && sd != 2) { T2_FAIL(4); return false; }
  if (sd == 2 && (ng != 1 || cin == 1 || cout == 1)) { T2_FAIL(5); return false; }     // strided gathers: 8 or 16 input channels, one phase
  if (cin == 16 && cout == 1) { T2_FAIL(6); return false; }
  merged = gs[0];
  int ntaps = 0;
  int lo[3] = {127, 127, 127}, hi[3] = {-127, -127, -127};
  int qmax[3] = {0, 0, 0};
  for (int p = 0; p < ng; ++p) {
    const Geom& g = gs[p];
    if (g.sin != sd || g.ntaps < 1 || g.sout != gs[0].sout) { T2_FAIL(7); return false; }
    if (g.inD != gs[0].inD || g.inH != gs[0].inH || g.inW != gs[0].inW) { T2_FAIL(8); return false; }
    if (g.qD > 32767 || g.qH > 32767 || g.qW > 32767) { T2_FAIL(9); return false; }
    for (int t = 0; t < g.ntaps; ++t) {
      if (ntaps >= kMaxTaps) { T2_FAIL(10); return false; }
      merged.taps[ntaps] = g.taps[t];
      merged.taps[ntaps].pad_ = (int8_t)p;
      ++ntaps;
      const int o[3] = {g.taps[t].dd, g.taps[t].dh, g.taps[t].dw};
      for (int i = 0; i < 3; ++i) { lo[i] = o[i] < lo[i] ? o[i] : lo[i]; hi[i] = o[i] > hi[i] ? o[i] : hi[i]; }
    }
    qmax[0] = g.qD > qmax[0] ? g.qD : qmax[0];
    qmax[1] = g.qH > qmax[1] ? g.qH : qmax[1];
    qmax[2] = g.qW > qmax[2] ? g.qW : qmax[2];
    pl.ph[p].qD = (int16_t)g.qD; pl.ph[p].qH = (int16_t)g.qH; pl.ph[p].qW = (int16_t)g.qW;
    pl.ph[p].rD = (int8_t)g.rD; pl.ph[p].rH = (int8_t)g.rH; pl.ph[p].rW = (int8_t)g.rW;
  }
  merged.ntaps = ntaps;
  pl.nph = ng;
  pl.sd = sd;
  pl.pps = cin == 16 ? 1 : 2;
  pl.nsg = sd == 2 ? 4 : 1;
  pl.lo_d = lo[0]; pl.lo_h = lo[1]; pl.lo_w = lo[2];
  pl.span_d = hi[0] - lo[0]; pl.span_h = hi[1] - lo[1]; pl.span_w = hi[2] - lo[2];
  if (pl.span_d > 4 || pl.span_h > 2 || pl.span_w > 2) { T2_FAIL(11); return false; }
  int lut[T2_MAX_PH][45];
  for (int p = 0; p < ng; ++p)
    for (int i = 0; i < 45; ++i) lut[p][i] = -1;
  for (int t = 0; t < ntaps; ++t) {
    const Tap& tp = merged.taps[t];
    lut[tp.pad_][(tp.dd - lo[0]) * 9 + (tp.dh - lo[1]) * 3 + (tp.dw - lo[2])] = tp.widx;
  }

  // row frame: a tap (kh, kw) = (sd*mh + parity_h, sd*mw + parity_w) is the row shift mh*PW + mw in its sub-grid
  const int sm_h = pl.span_h / sd, sm_w = pl.span_w / sd;
  pl.PW = cin == 1 ? qmax[2] : qmax[2] + sm_w;
  pl.RTOT = qmax[1] * pl.PW;
  pl.qDmax = qmax[0];
  pl.OB = cout == 1 ? 16 : (sd == 2 ? (cin == 16 ? 2 : 4) : 8);       // 16-channel strided gathers: one plane per slot, keep the window short
  const int wpl = sd * (pl.OB - 1) + pl.span_d + 1;      // input planes one block reads
  pl.NPAIR = (wpl + pl.pps - 1) / pl.pps;
  pl.ppb = sd * pl.OB / pl.pps;
  pl.ACCW = pl.OB * cout;
  if (pl.NPAIR > T2_MAX_PAIR || pl.ppb > pl.NPAIR) { T2_FAIL(12); return false; }

  // MMA list + weight blocks (deduplicated: interior slots share one shift-invariant block)
  struct Key { int ph, rel, nj, dh, dw, off16; };
  Key keys[T2_MAX_BLK];
  int m_sg[T2_MAX_MMA], m_mh[T2_MAX_MMA], m_mw[T2_MAX_MMA];
  int nblk = 0, nmma = 0, woff = 0;
  for (int ph = 0; ph < ng; ++ph)
    for (int p = 0; p <= pl.NPAIR; ++p) {
      if (nmma > 255) { T2_FAIL(13); return false; }
      pl.ph[ph].pair_begin[p] = (uint8_t)nmma;
      if (p == pl.NPAIR) break;
      const int i_lo = pl.pps * p, i_hi = i_lo + pl.pps - 1;
      int j0 = t2_ceil_div(i_lo - pl.span_d, sd), j1 = i_hi / sd;
      if (j0 < 0) j0 = 0;
      if (j1 > pl.OB - 1) j1 = pl.OB - 1;
      if (cout == 1) { j0 = 0; j1 = pl.OB - 1; }
      if (j1 < j0) continue;
      if (cout == 8) {                           // N = nj * 8 must be a multiple of 16; keep the column offset 16-aligned
        j0 &= ~1;
        if (((j1 - j0 + 1) & 1) != 0) ++j1;      // OB is even, so an even j1 is never the last plane
      }
      const int nj = j1 - j0 + 1;
      for (int sg = 0; sg < pl.nsg; ++sg)
        for (int mh = 0; sd * mh + (sg >> 1) <= pl.span_h; ++mh)
          for (int mw = 0; mw <= (cin == 1 ? 0 : sm_w) && sd * mw + (sg & 1) <= pl.span_w; ++mw) {
            const int kh = sd * mh + (sg >> 1), kw = sd * mw + (sg & 1);
            bool any = false;                     // does the block hold any tap?
            for (int chunk = 0; chunk < 2 && !any; ++chunk)
              for (int j = j0; j <= j1 && !any; ++j) {
                const int kd = i_lo + (cin == 16 ? 0 : chunk) - sd * j;
                if (kd < 0 || kd > pl.span_d) continue;
                for (int e = 0; e <= (cin == 1 ? pl.span_w : 0); ++e)
                  if (lut[ph][kd * 9 + kh * 3 + (cin == 1 ? e : kw)] >= 0) any = true;
              }
            if (!any) continue;
            int found = -1;
            for (int k = 0; k < nblk; ++k)
              if (keys[k].ph == ph && keys[k].rel == i_lo - sd * j0 && keys[k].nj == nj && keys[k].dh == kh && keys[k].dw == kw)
                found = k;
            if (found < 0) {
              if (nblk >= T2_MAX_BLK) { T2_FAIL(14); return false; }
              found = nblk++;
              keys[found] = Key{ph, i_lo - sd * j0, nj, kh, kw, woff >> 4};
              pl.blk[found].i0 = (int8_t)i_lo; pl.blk[found].j0 = (int8_t)j0; pl.blk[found].nj = (int8_t)nj;
              pl.blk[found].dh = (int8_t)kh; pl.blk[found].dw = (int8_t)kw; pl.blk[found].ph = (int8_t)ph;
              pl.blk_off16[found] = (uint16_t)(woff >> 4);
              woff += nj * cout * 32;
            }
            if (nmma >= T2_MAX_MMA) { T2_FAIL(15); return false; }
            T2Mma& m = pl.mma[nmma];
            m_sg[nmma] = sg; m_mh[nmma] = mh; m_mw[nmma] = cin == 1 ? 0 : mw;
            ++nmma;
            m.b_lo = (uint32_t)keys[found].off16 | (8u << 16);          // LBO = 128 bytes between the two K chunks
            m.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((nj * cout) >> 3) << 17) | ((128u >> 4) << 24);
            m.dcol = (uint32_t)(j0 * cout);
          }
    }
  pl.nmma = nmma;
  pl.nblk = nblk;
  pl.wbytes = (woff + 1023) & ~1023;
  if (pl.wbytes > 96 * 1024) { T2_FAIL(16); return false; }

  // (16 output channels stay on the producer-warp path: that instantiation spills in the 12-warp epilogue and is slower)
  if (tma && cin == 8 && cout != 16 && (sd == 1 || (ng == 1 && !affine && cout != 1)) && (ng == 1 || !affine) && (pl.span_d <= 2 || !affine) &&
      merged.outD <= 64 && merged.outH <= 64 && merged.outW <= 64 && pl.PW <= 64) {
    // ---- TMA-direct staging: tile = hb whole lines of the row frame, a slot = two boxes [8][PW][hb + halo]
    int nrb_max = 512 / (2 * pl.ACCW);
    if (nrb_max > 4) nrb_max = 4;
    // whole-line tiles quantise: pick the row-block count whose tiles waste the fewest MMA rows (ties: the larger tile)
    int order[4], nord = 0;
    {
      double effs[5] = {0, 0, 0, 0, 0};
      for (int q = 1; q <= nrb_max; ++q) {
        int hb = (128 * q) / pl.PW;
        if (hb > qmax[1]) hb = qmax[1];
        if (hb < 1) continue;
        const int nt = (qmax[1] + hb - 1) / hb;
        hb = (qmax[1] + nt - 1) / nt;
        effs[q] = (double)(qmax[1] * pl.PW) / ((double)nt * 128.0 * ((hb * pl.PW + 127) / 128)) + 1e-3 * q;
      }
      bool used[5] = {false, false, false, false, false};
      for (int k = 0; k < nrb_max; ++k) {
        int best_q = 0;
        for (int q = 1; q <= nrb_max; ++q) if (!used[q] && effs[q] > 0 && (best_q == 0 || effs[q] > effs[best_q])) best_q = q;
        if (!best_q) break;
        used[best_q] = true;
        order[nord++] = best_q;
      }
    }
    for (int oi = 0; oi < nord; ++oi) {                      // best tile whose ring still fits next to the weights
      const int nrb_try = order[oi];
      int hb = (128 * nrb_try) / pl.PW;
      if (hb > qmax[1]) hb = qmax[1];
      if (hb < 1) continue;
      const int nt = (qmax[1] + hb - 1) / hb;
      hb = (qmax[1] + nt - 1) / nt;                        // balanced tiles
      const int nrb = (hb * pl.PW + 127) / 128;
      const int box_h = hb + sm_h;
      int sr = box_h * pl.PW;
      const int reach = 128 * nrb + sm_h * pl.PW + sm_w;   // rows the last row block's shifted windows touch
      if (sr < reach) sr = reach;
      sr = (sr + 7) & ~7;
      const size_t slot = (size_t)2 * pl.nsg * sr * 16;
      int r = pl.NPAIR + 3;
      if (r > T2_MAX_RING) r = T2_MAX_RING;
      while (r > pl.NPAIR + 1 && (size_t)r * slot + pl.wbytes > (size_t)kT2SmemBudget) --r;
      if ((size_t)r * slot + pl.wbytes > (size_t)kT2SmemBudget || sd * box_h > 256) continue;
      pl.hb = hb; pl.box_h = box_h; pl.nrb = nrb; pl.TR = 128 * nrb; pl.SR = sr; pl.R = r;
      pl.ntiles = nt;
      pl.groups = affine ? merged.N / merged.group_size : 1;
      for (int m = 0; m < nmma; ++m) pl.mma[m].a_shift = (uint32_t)(m_sg[m] * pl.SR + m_mh[m] * pl.PW + m_mw[m]);
      int tc = 32;
      while (tc < 2 * pl.nrb * pl.ACCW) tc <<= 1;
      if (tc > 512) { T2_FAIL(17); return false; }
      pl.tmem_cols = tc;
      const int nblocks_all = (pl.qDmax + pl.OB - 1) / pl.OB;
      const long long cols = (long long)merged.N * pl.ntiles;
      const long long want = 2LL * vg_sm_count();
      int nch = (int)((want + cols - 1) / cols);
      if (nch > nblocks_all) nch = nblocks_all;
      if (nch < 1) nch = 1;
      pl.dchunk = ((nblocks_all + nch - 1) / nch) * pl.OB;
      pl.ndchunks = (pl.qDmax + pl.dchunk - 1) / pl.dchunk;
      return true;
    }
  }
  // rows per tile: as many 128-row blocks as TMEM (2 buffers), the producers' reach and shared memory allow
  const int es = t2_epi_sets(cin, cout, sd);
  const int max_sr = t2_max_chunk(cin, es, sd) * t2_prod_warps(cin, es) * 32;
  int nrb_max = 512 / (2 * pl.ACCW);
  if (nrb_max > 4) nrb_max = 4;
  const int need = (pl.RTOT + 127) / 128;
  if (nrb_max > need) nrb_max = need;
  // epilogue aux staging: rows must be the output plane's voxels in order (pitch == width, one phase, unit strides)
  const bool use_auxs = auxs && cin == 1 && cout == 8 && ng == 1 && sd == 1 && merged.sout == 1 && pl.PW == merged.outW &&
                        merged.qH == merged.outH && merged.qW == merged.outW;
  if (use_auxs && nrb_max > 2) nrb_max = 2;                 // two units of nrb x OB chunks of 2 KB must fit beside the ring
  int best = 0, best_r = 0;
  double best_eff = 0.0;
  for (int nrb = nrb_max; nrb >= 1; --nrb) {
    const int tr = 128 * nrb;
    const int sr = (tr + sm_h * pl.PW + (cin == 1 ? 0 : sm_w) + 7) & ~7;
    if (sr > max_sr) continue;
    const size_t slot = (size_t)2 * pl.nsg * sr * 16;
    const size_t aux_bytes = use_auxs ? (size_t)2 * nrb * pl.OB * 2048 : 0;
    int r = pl.NPAIR + 3;
    if (r > T2_MAX_RING) r = T2_MAX_RING;
    while (r > pl.NPAIR + 1 && (size_t)r * slot + pl.wbytes + aux_bytes > (size_t)kT2SmemBudget) --r;
    if ((size_t)r * slot + pl.wbytes + aux_bytes > (size_t)kT2SmemBudget) continue;
    const int nt = (pl.RTOT + tr - 1) / tr;
    // useful rows per staged row: tile quantisation and the h-halo that every tile re-stages
    const double eff = (double)pl.RTOT / ((double)nt * sr);
    if (eff > best_eff * 1.15) { best_eff = eff; best = nrb; best_r = r; }
  }
  if (best < 1) { T2_FAIL(18); return false; }
  pl.nrb = best;
  pl.TR = 128 * best;
  pl.SR = (pl.TR + sm_h * pl.PW + (cin == 1 ? 0 : sm_w) + 7) & ~7;
  pl.R = best_r;
  pl.auxs = use_auxs ? 1 : 0;
  for (int m = 0; m < nmma; ++m) pl.mma[m].a_shift = (uint32_t)(m_sg[m] * pl.SR + m_mh[m] * pl.PW + m_mw[m]);
  pl.ntiles = (pl.RTOT + pl.TR - 1) / pl.TR;
  int tc = 32;
  while (tc < 2 * pl.nrb * pl.ACCW) tc <<= 1;
  if (tc > 512) { T2_FAIL(19); return false; }
  pl.tmem_cols = tc;

  // split columns along d until the persistent grid has at least ~2 columns per SM
  const int nblocks_all = (pl.qDmax + pl.OB - 1) / pl.OB;
  const long long cols = (long long)merged.N * pl.ntiles;
  const long long want = 2LL * vg_sm_count();
  int nch = (int)((want + cols - 1) / cols);
  if (nch > nblocks_all) nch = nblocks_all;
  if (nch < 1) nch = 1;
  pl.dchunk = ((nblocks_all + nch - 1) / nch) * pl.OB;
  pl.ndchunks = (pl.qDmax + pl.dchunk - 1) / pl.dchunk;
  return true;
}

bool tc2_supported(int cin, int cout, const Geom* gs, int ng) {
  T2Plan pl;
  Geom merged;
  return t2_build_plan(cin, cout, gs, ng, merged, pl);
}

// human-readable plan (vg_conv_describe): tile shape, ring, MMA list size, shared memory
static bool tma_wanted() {      // VAEGAM_TMA=0 switches the TMA-direct staging off (producer-warp staging of the same bf16 tensors)
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAEGAM_TMA"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
int tc2_describe(int cin, int cout, const Geom* gs, int ng, char* buf, size_t cap, bool in_bf16) {
  T2Plan pl;
  Geom merged;
  if (!t2_build_plan(cin, cout, gs, ng, merged, pl, in_bf16 && tma_wanted(), false)) return 0;
  const long long cols = (long long)merged.N * pl.ntiles * pl.ndchunks;
  return snprintf(buf, cap,
                  "tc2 cin=%d cout=%d sin=%d phases=%d q=(%d,%d,%d) taps=%d PW=%d RTOT=%d TR=%d ntiles=%d SR=%d OB=%d NPAIR=%d R=%d "
                  "ACCW=%d tmem=%d nmma=%d nblk=%d wbytes=%d smem=%zu dchunk=%d ndchunks=%d cols=%lld tma_hb=%d",
                  cin, cout, pl.sd, ng, pl.qDmax, pl.RTOT / pl.PW, merged.qW, merged.ntaps, pl.PW, pl.RTOT, pl.TR, pl.ntiles,
                  pl.SR, pl.OB, pl.NPAIR, pl.R, pl.ACCW, pl.tmem_cols, pl.nmma, pl.nblk, pl.wbytes,
                  (size_t)pl.R * 2 * pl.nsg * pl.SR * 16 + pl.wbytes, pl.dchunk, pl.ndchunks, cols, pl.hb);
}

// ---- TMA tensor map of a channels-last bf16 tensor (N, D, H, W, 8).  A voxel is one 16-byte word and the voxels of
// an h-line are contiguous, so (channel, w) is ONE dimension of 4 * W 32-bit words: dims fastest-first (line words,
// h, d, n), a box is [4 * PW][box_h][1][1] — every box row is a whole padded line (a few hundred contiguous bytes,
// not 16).  Out-of-range coordinates (the halo of a transposed convolution, planes beyond the volume) are
// zero-filled by the hardware.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}
static bool tma_enabled() { return tma_wanted() && tma_encoder() != nullptr; }
static int make_tmap_strided(const Geom& g, const T2Plan& pl, const void* base, CUtensorMap& tm) {
  // input stride 2: 5-D (channel, w, h, d, n) with element strides 2 along w and h; a box of 2*PW x 2*box_h traversed
  // positions delivers the PW x box_h voxels of one parity sub-grid, densely packed
  const cuuint64_t dims[5] = {8, (cuuint64_t)g.inW, (cuuint64_t)g.inH, (cuuint64_t)g.inD, (cuuint64_t)g.N};
  const cuuint64_t strides[4] = {16, (cuuint64_t)g.inW * 16, (cuuint64_t)g.inW * g.inH * 16, (cuuint64_t)g.in_img * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)pl.PW * 2, (cuuint32_t)pl.box_h * 2, 1, 1};
  const cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  if (((uintptr_t)base & 15) || (strides[3] & 15)) { set_error("TMA staging: tensor not 16-byte aligned"); return VG_EINVAL; }
  const CUresult rc = tma_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (strided) failed (%d)", (int)rc); return VG_ECUDA; }
  return VG_OK;
}
static int make_tmap(const Geom& g, const T2Plan& pl, const void* base, CUtensorMap& tm) {
  if (pl.sd == 2) return make_tmap_strided(g, pl, base, tm);
  const cuuint64_t dims[4] = {(cuuint64_t)g.inW * 4, (cuuint64_t)g.inH, (cuuint64_t)g.inD, (cuuint64_t)g.N};
  const cuuint64_t strides[3] = {(cuuint64_t)g.inW * 16, (cuuint64_t)g.inW * g.inH * 16, (cuuint64_t)g.in_img * 2};
  const cuuint32_t box[4] = {(cuuint32_t)pl.PW * 4, (cuuint32_t)pl.box_h, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (((uintptr_t)base & 15) || (strides[2] & 15)) { set_error("TMA staging: tensor not 16-byte aligned"); return VG_EINVAL; }
  const CUresult rc = tma_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc); return VG_ECUDA; }
  return VG_OK;
}

bool tma_available() { return tma_enabled(); }
bool make_tmap_voxels(const void* base, int C, int W, int H, int D, int N, long long img_stride_elems, int bw, int bh, int bd,
                      CUtensorMap* tm) {
  // (channel, w) is ONE dimension of 32-bit words (C / 2 per voxel), so a box row is bw whole voxels = a contiguous
  // run of bw * C * 2 bytes (16-byte box rows are what makes a tensor copy slow)
  const int wpv = C / 2;
  if (!tma_enabled() || (C != 8 && C != 16) || bw < 1 || bh < 1 || bd < 1 || bw * wpv > 256 || bh > 256 || bd > 256) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W * wpv, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)W * C * 2, (cuuint64_t)W * H * C * 2, (cuuint64_t)img_stride_elems * 2};
  const cuuint32_t box[4] = {(cuuint32_t)(bw * wpv), (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (((uintptr_t)base & 15) || (strides[2] & 15)) return false;
  memset(tm, 0, sizeof(*tm));
  const CUresult rc = tma_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

template <int CIN, int COUT, int SD, bool TMA = false>
static int launch_tc2_t(const Geom& g, const GatherArgs& a, const T2Plan& pl, cudaStream_t st) {
  const size_t smem = (size_t)pl.R * 2 * pl.nsg * pl.SR * 16 + pl.wbytes + (pl.auxs ? (size_t)2 * pl.nrb * pl.OB * 2048 : 0);
  VG_CUDA(cudaFuncSetAttribute(tc2_kernel<CIN, COUT, SD, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long cols = (long long)g.N * pl.ntiles * pl.ndchunks;
  const int sms = vg_sm_count();
  unsigned grid = (unsigned)(cols < sms ? cols : sms);
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if constexpr (TMA) {
    VG_TRY(make_tmap(g, pl, a.in, tm));
    if (pl.groups > 1) {                          // a CTA serves one statistics group: equal CTA counts per group
      const long long per_group = (long long)g.group_size * pl.ntiles * pl.ndchunks;
      long long per = sms / pl.groups;
      if (per < 1) per = 1;
      if (per > per_group) per = per_group;
      grid = (unsigned)(per * pl.groups);
    }
  }
  tc2_kernel<CIN, COUT, SD, TMA><<<grid, T2_THREADS, smem, st>>>(g, a, pl, tm);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

int launch_tc2_gather(int cin, int cout, const Geom* gs, int ng, const GatherArgs& a, cudaStream_t st) {
  T2Plan pl;
  Geom merged;
  const bool want_tma = a.in_bf16 && tma_enabled();
  // Opt-in (VAEGAM_AUX_STAGING=1): measured on B200 the smaller tile it needs (two row blocks, so that two units of
  // chunks fit beside the ring) costs more than the prefetch wins — convt5's data gradient 0.65 ms vs 0.50 ms.
  static const bool aux_staging = [] { const char* e = getenv("VAEGAM_AUX_STAGING"); return e && e[0] == '1'; }();
  const bool want_auxs = aux_staging && a.aux_bf16 && a.aux_mode != 0 && a.aux != nullptr;
  if (!t2_build_plan(cin, cout, gs, ng, merged, pl, want_tma, a.in_scale != nullptr, want_auxs)) { set_error("plane-folded tensor-core path: unsupported geometry"); return VG_EINVAL; }
  if (pl.hb > 0) {                                // TMA-direct staging of a bf16 input
    if (pl.sd == 2 && cin == 8 && cout == 8) return launch_tc2_t<8, 8, 2, true>(merged, a, pl, st);
    if (pl.sd == 2 && cin == 8 && cout == 16) return launch_tc2_t<8, 16, 2, true>(merged, a, pl, st);
    if (cin == 8 && cout == 1) return launch_tc2_t<8, 1, 1, true>(merged, a, pl, st);
    if (cin == 8 && cout == 8) return launch_tc2_t<8, 8, 1, true>(merged, a, pl, st);
    if (cin == 8 && cout == 16) return launch_tc2_t<8, 16, 1, true>(merged, a, pl, st);
    set_error("TMA staging: unsupported channel pair (%d,%d)", cin, cout);
    return VG_EINVAL;
  }
  if (pl.sd == 2) {
    if (cin == 8 && cout == 8) return launch_tc2_t<8, 8, 2>(merged, a, pl, st);
    if (cin == 8 && cout == 16) return launch_tc2_t<8, 16, 2>(merged, a, pl, st);
    if (cin == 16 && cout == 16) return launch_tc2_t<16, 16, 2>(merged, a, pl, st);
  } else {
    if (cin == 1 && cout == 8) return launch_tc2_t<1, 8, 1>(merged, a, pl, st);
    if (cin == 1 && cout == 16) return launch_tc2_t<1, 16, 1>(merged, a, pl, st);
    if (cin == 8 && cout == 1) return launch_tc2_t<8, 1, 1>(merged, a, pl, st);
    if (cin == 8 && cout == 8) return launch_tc2_t<8, 8, 1>(merged, a, pl, st);
    if (cin == 8 && cout == 16) return launch_tc2_t<8, 16, 1>(merged, a, pl, st);
    if (cin == 16 && cout == 8) return launch_tc2_t<16, 8, 1>(merged, a, pl, st);
    if (cin == 16 && cout == 16) return launch_tc2_t<16, 16, 1>(merged, a, pl, st);
  }
  set_error("plane-folded tensor-core path: unsupported channel pair (%d,%d) at input stride %d", cin, cout, pl.sd);
  return VG_EINVAL;
}

}  // namespace vg
