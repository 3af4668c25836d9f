// Fused reconstruction + Gaussian likelihood + GLM regulariser, forward and backward
// (vae_reg_GP.py:380 cons = g*diff, :388-389 cdist regulariser,
// :390 x_rec accumulation, :401-405 Normal(x_rec, exp(-eps)).log_prob(x).sum(1)).
//
// HBM-bound by construction (~1 FLOP/B): ONE pass over the 9 decoder maps and x.
//   forward : reads 10*B*V fp32 (+9*V of eps/GLM, L1/L2-resident)
//             -> logp (B), norms (8,B)                         algorithmic 40*B*V bytes
//   backward: reads the same, writes 9*B*V of d(pre-sigmoid)   algorithmic 76*B*V bytes
//             -> dg (8,B), deps (V)
// Data layout: every map row is padded to VP = round_up(V,4) floats so that rows are 16-byte
// aligned (V = 70315 is odd); x is the caller's dense (B,V) tensor (rows not 16-byte aligned).
//
// Work decomposition (balanced, persistent): the (row-group, column) space is cut into WARP items of
// 4 batch rows x 8 float4 columns (a lane owns one row and 4 voxels; the 4 lanes sharing a column
// read eps / the 8 GLM values as one broadcast).  The items, ordered row-group-major, are dealt to
// G ~ (CTAs per SM) x SMs CTAs in equal contiguous spans, the warps of a CTA interleaved inside the
// span.  A wave is therefore always full and every CTA finishes at the same time, whatever B is (the
// former (chunk, row-group) grid ran 1.37 waves at B = 128).
//
// Memory pipeline: each lane streams ITS OWN 9 x 16 B of maps through a private
// shared-memory ring with cp.async (LDGSTS, L2-only), STAGES-1 items ahead of the arithmetic.  The
// bytes in flight are therefore constant (threads x (STAGES-1) x 144 B per SM) instead of dropping to
// zero while a warp computes, no register holds a load in flight, and — a lane only ever reads what it
// copied itself — the ring needs no barrier, only cp.async.wait_group.  x (rows only 4-byte aligned)
// comes through 4 registers requested one item ahead.
//
// Row sums stay in per-lane registers for the whole span and are reduced over the 8 lanes of a row
// once per (span, row-group) segment; a CTA's span crosses at most one row-group boundary, so it writes
// two segments of partials.  A second tiny kernel adds the partials in a fixed order, so the result
// is deterministic (no float atomics on the outputs; deps uses vector fp32 RED by design).
#include <type_traits>

#include "common.cuh"

namespace vg {

constexpr int ROWS = 4;          // batch rows per warp item
constexpr int WCOLS = 8;         // float4 columns per warp item
constexpr int NMAP = 9;          // base + 8 covariate maps
constexpr int KCOV = 8;
constexpr int NSLOT = 9;         // partial slots per row: fwd logp + 8 squared norms, bwd 8 dg (slot 8 unused)
constexpr int GV = 1 + KCOV;          // per-column constants: eps + 8 GLM values
constexpr int RING_VECS = NMAP;       // float4 per lane per stage: the 9 maps
constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

__device__ __forceinline__ float4 ld4_keep(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ReconArgs {
  const float* maps;   // (9, b, vp)
  const float* g;      // (8, b)
  const float* x;      // (b, v) dense
  const float* eps;    // (vp)
  const float* glm;    // (8, vp)
  int b;
  int v, vp;
  int ncols;           // vp / 4
  int n_c8;            // warp items per row-group = ceil(ncols / 8)
  int n_items;         // row-groups * n_c8
  int span;            // warp items per CTA (multiple of the CTA's warps, <= n_c8)
};

// Launch shape: warps per CTA, CTAs per SM, ring depth (shared memory: warps*32*stages*144 B per CTA).
struct Shape { int warps, ctas, stages; };
constexpr int NVARIANT = 4;
// vg_recon_tune(variant): 0 = 4 warps x 5 CTAs/SM x 2 stages, 1 = 4 x 3 x 3, 2 = 4 x 4 x 2, 3 = 8 x 2 x 3 (default; anything else selects it).
static int g_variant = 3;
static inline Shape shape_of(int variant) {
  return variant == 1 ? Shape{4, 3, 3} : variant == 2 ? Shape{4, 4, 2} : variant == 3 ? Shape{8, 2, 3} : Shape{4, 5, 2};
}
static inline size_t ring_bytes(const Shape& sh) { return (size_t)sh.stages * RING_VECS * sh.warps * 32 * sizeof(float4); }

struct ReconPlan { int n_rg, n_c8, n_items, span, grid; };

static ReconPlan make_plan(int b, long long v, const Shape& sh) {
  ReconPlan p;
  const int ncols = (int)((v + 3) / 4);
  p.n_rg = cdiv(b, ROWS);
  p.n_c8 = cdiv(ncols, WCOLS);
  p.n_items = p.n_rg * p.n_c8;
  const int target = sh.ctas * vg_sm_count();
  int span = cdiv(cdiv(p.n_items, target), sh.warps) * sh.warps;
  const int cap = p.n_c8 >= sh.warps ? p.n_c8 / sh.warps * sh.warps : sh.warps;   // <= 1 boundary per span
  if (span > cap) span = cap;
  p.span = span;
  p.grid = cdiv(p.n_items, span);
  return p;
}

// partial layout: [cta][segment 0..1][warp][row 0..3][NSLOT]
__device__ __forceinline__ size_t partial_index(int cta, int seg, int warp, int r, int nwarps) {
  return ((((size_t)cta * 2 + seg) * nwarps + warp) * ROWS + r) * NSLOT;
}

// sum over the 8 lanes that share a batch row (lanes r*8 .. r*8+7)
__device__ __forceinline__ float row_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

extern __shared__ float4 s_ring[];   // [stage][vec 0..9][thread]: consecutive lanes, consecutive 16-byte words

// A lane's input stream: its warp's items inside the CTA's span, through the lane-private ring.
template <int NW, int STAGES>
struct Stream {
  const ReconArgs& a;
  float4* ring;          // &s_ring[threadIdx.x]
  int r, cl;             // row inside the item, column inside the item
  int last;              // end of the CTA's span
  int next_item;         // next item to issue
  int issue_stage;

  __device__ Stream(const ReconArgs& args, int first, int last_)
      : a(args), ring(s_ring + threadIdx.x), r((threadIdx.x & 31) >> 3), cl(threadIdx.x & 7), last(last_),
        next_item(first + (threadIdx.x >> 5)), issue_stage(0) {}

  __device__ __forceinline__ float4* slot(int stage, int vec) const {
    return ring + (stage * RING_VECS + vec) * (NW * 32);
  }

  // copy this lane's share of `next_item` (if any) into the next ring stage; always commits one group
  __device__ __forceinline__ void issue() {
    if (next_item < last) {
      const int rg = next_item / a.n_c8;
      const int col = (next_item - rg * a.n_c8) * WCOLS + cl;
      const int row = rg * ROWS + r;
      if (row < a.b && col < a.ncols) {
        const int v0 = col * 4;
        const float* src = a.maps + (size_t)row * a.vp + v0;
        const size_t map_stride = (size_t)a.b * a.vp;
#pragma unroll
        for (int j = 0; j < NMAP; ++j) cp_async16(slot(issue_stage, j), src + j * map_stride);
      }
    }
    cp_async_commit();
    next_item += NW;
    issue_stage = issue_stage + 1 == STAGES ? 0 : issue_stage + 1;
  }

  // x rows are only 4-byte aligned (V is odd) and eps / GLM are shared by the 4 rows of an item (one
  // broadcast request per warp, L1/L2 hits): this lane's 4 voxels of x and its column's 9 constants come
  // through registers, requested one item ahead of their use so that the arithmetic never waits on L2.
  template <bool CONSTS>
  __device__ __forceinline__ void load_x(int item, float out[4], float4 cst[GV]) const {
#pragma unroll
    for (int l = 0; l < 4; ++l) out[l] = 0.f;
#pragma unroll
    for (int j = 0; j < GV; ++j) cst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (item < last) {
      const int rg = item / a.n_c8;
      const int col = (item - rg * a.n_c8) * WCOLS + cl;
      const int row = rg * ROWS + r;
      if (row < a.b && col < a.ncols) {
        const int v0 = col * 4;
        const float* xsrc = a.x + (size_t)row * a.v + v0;
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (v0 + l < a.v) out[l] = __ldg(xsrc + l);
        if (CONSTS) {
          cst[0] = ld4_keep(a.eps + v0);
#pragma unroll
          for (int i = 0; i < KCOV; ++i) cst[1 + i] = ld4_keep(a.glm + (size_t)i * a.vp + v0);
        }
      }
    }
  }
};

template <int NW, int CPS, int STAGES, bool MAPS>
__global__ void __launch_bounds__(NW * 32, CPS)
recon_fwd_kernel(const ReconArgs a, float* __restrict__ partial, float* cons_out, float* xrec_out) {
  __shared__ float s_g[2][ROWS][KCOV];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 3, cl = lane & 7;
  const int first = blockIdx.x * a.span;
  const int last = min(first + a.span, a.n_items);
  const int rg0 = first / a.n_c8;
  Stream<NW, STAGES> in(a, first, last);
#pragma unroll
  for (int s = 0; s < STAGES; ++s) in.issue();
  for (int t = threadIdx.x; t < 2 * ROWS * KCOV; t += NW * 32) {
    const int sg = t / (ROWS * KCOV), rr = (t / KCOV) % ROWS, i = t % KCOV;
    const int row = (rg0 + sg) * ROWS + rr;
    s_g[sg][rr][i] = row < a.b ? __ldg(a.g + i * a.b + row) : 0.f;
  }
  __syncthreads();

  int it = first + warp;
  int stage = 0;
  float xn[4];
  float4 cn[GV];
  in.template load_x<true>(it, xn, cn);
#pragma unroll 1
  for (int seg = 0; seg < 2; ++seg) {
    const int rg = rg0 + seg;
    const int lim = min(last, (rg + 1) * a.n_c8);
    const int row = rg * ROWS + r;
    const bool row_ok = row < a.b;
    float acc[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) acc[j] = 0.f;
    float gi[KCOV];
#pragma unroll
    for (int i = 0; i < KCOV; ++i) gi[i] = s_g[seg][r][i];
#pragma unroll 1
    for (; it < lim; it += NW) {
      cp_async_wait<STAGES - 1>();       // this lane's copy of item `it` has landed
      const float xs[4] = {xn[0], xn[1], xn[2], xn[3]};
      float4 cs[GV];
#pragma unroll
      for (int j = 0; j < GV; ++j) cs[j] = cn[j];
      in.template load_x<true>(it + NW, xn, cn);
      const int col = (it - rg * a.n_c8) * WCOLS + cl;
      const int v0 = col * 4;
      // TAIL: the last column holds fewer than 4 voxels (row padding may hold anything, even NaN)
      auto body = [&](auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        bool ok[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) ok[l] = TAIL ? v0 + l < a.v : true;
        const float4 e4 = cs[0];
        const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
        const float4 b4 = *in.slot(stage, 0);
        float xr[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < KCOV; ++i) {
          const float4 G = cs[1 + i];
          const float4 D = *in.slot(stage, i + 1);
          const float c[4] = {gi[i] * D.x, gi[i] * D.y, gi[i] * D.z, gi[i] * D.w};
          const float Gs[4] = {G.x, G.y, G.z, G.w};
          float s = 0.f;
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            xr[l] += c[l];
            const float d = c[l] - Gs[l];
            if (ok[l]) s = fmaf(d, d, s);
          }
          acc[1 + i] += s;
          if (MAPS && cons_out) {
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (ok[l]) cons_out[((size_t)i * a.b + row) * a.v + v0 + l] = c[l];
          }
        }
        float lp = 0.f;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const float rr = xs[l] - xr[l];
          if (ok[l]) lp += fmaf(-0.5f * rr * rr, __expf(2.f * ee[l]), ee[l] - HALF_LOG_2PI);
        }
        acc[0] += lp;
        if (MAPS && xrec_out) {
#pragma unroll
          for (int l = 0; l < 4; ++l)
            if (ok[l]) xrec_out[(size_t)row * a.v + v0 + l] = xr[l];
        }
      };
      if (row_ok && col < a.ncols) {
        if (v0 + 4 <= a.v) body(std::false_type{});
        else body(std::true_type{});
      }
      in.issue();                        // refill the stage just consumed
      stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
    __syncwarp();
    float mine = 0.f, extra = 0.f;
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      const float s = row_sum(acc[j]);
      if (j < 8) { if (cl == j) mine = s; }
      else extra = s;
    }
    float* out = partial + partial_index(blockIdx.x, seg, warp, r, NW);
    out[cl] = mine;
    if (cl == 0) out[8] = extra;
  }
  cp_async_wait<0>();
}

// One warp per (row, slot): the partials of the CTAs whose span meets the row's row-group, `nwarps` each,
// are added in double in a fixed order.
__global__ void recon_finalize(const float* __restrict__ partial, int b, int n_c8, int span, int n_items,
                               int nwarps, int nslot_used, int is_fwd, float* out0, float* out1) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= b * nslot_used) return;
  const int row = gw / nslot_used, slot = gw % nslot_used;
  const int rg = row / ROWS, r = row % ROWS;
  const int c_lo = (rg * n_c8) / span;
  const int c_hi = (min((rg + 1) * n_c8, n_items) - 1) / span;
  const int n = (c_hi - c_lo + 1) * nwarps;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) {
    const int c = c_lo + k / nwarps, w = k % nwarps;
    const int seg = rg - (c * span) / n_c8;
    s += (double)partial[partial_index(c, seg, w, r, nwarps) + slot];
  }
  s = warp_sum(s);
  if (lane != 0) return;
  if (is_fwd) {
    if (slot == 0) out0[row] = (float)s;                      // logp
    else out1[(slot - 1) * b + row] = (float)sqrt(s);         // norms (8, b)
  } else {
    out0[slot * b + row] = (float)s;                          // dg (8, b)
  }
}

template <int NW, int CPS, int STAGES>
__global__ void __launch_bounds__(NW * 32, CPS)
recon_bwd_kernel(const ReconArgs a, const float* __restrict__ norms, float lam, float* __restrict__ dpre,
                 float* __restrict__ partial, float* __restrict__ deps) {
  __shared__ float s_g[2][ROWS][KCOV], s_cf[2][ROWS][KCOV];     // gain and lam*B/norm
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 3, cl = lane & 7;
  const int first = blockIdx.x * a.span;
  const int last = min(first + a.span, a.n_items);
  const int rg0 = first / a.n_c8;
  const size_t map_stride = (size_t)a.b * a.vp;
  const float invB = 1.f / (float)a.b;
  Stream<NW, STAGES> in(a, first, last);
#pragma unroll
  for (int s = 0; s < STAGES; ++s) in.issue();
  for (int t = threadIdx.x; t < 2 * ROWS * KCOV; t += NW * 32) {
    const int sg = t / (ROWS * KCOV), rr = (t / KCOV) % ROWS, i = t % KCOV;
    const int row = (rg0 + sg) * ROWS + rr;
    const bool ok = row < a.b;
    s_g[sg][rr][i] = ok ? __ldg(a.g + i * a.b + row) : 0.f;
    const float nn = ok ? __ldg(norms + i * a.b + row) : 0.f;
    s_cf[sg][rr][i] = nn > 0.f ? lam * (float)a.b / nn : 0.f;
  }
  __syncthreads();

  int it = first + warp;
  int stage = 0;
  float xn[4];
  float4 cn[GV];
  in.template load_x<false>(it, xn, cn);
#pragma unroll 1
  for (int seg = 0; seg < 2; ++seg) {
    const int rg = rg0 + seg;
    const int lim = min(last, (rg + 1) * a.n_c8);
    const int row = rg * ROWS + r;
    const bool row_ok = row < a.b;
    float acc[KCOV];
#pragma unroll
    for (int i = 0; i < KCOV; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (; it < lim; it += NW) {       // warp-uniform bound: whole warps reach the shuffles below
      cp_async_wait<STAGES - 1>();     // this lane's copy of item `it` has landed
      const float xs[4] = {xn[0], xn[1], xn[2], xn[3]};
      in.template load_x<false>(it + NW, xn, cn);
      const int col = (it - rg * a.n_c8) * WCOLS + cl;
      const int v0 = col * 4;
      float de[4] = {0.f, 0.f, 0.f, 0.f};
      // TAIL: the last column holds fewer than 4 voxels (row padding may hold anything, even NaN)
      auto body = [&](auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        bool ok[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) ok[l] = TAIL ? v0 + l < a.v : true;
        const float4 e4 = ld4_keep(a.eps + v0);
        const float ww[4] = {__expf(2.f * e4.x), __expf(2.f * e4.y), __expf(2.f * e4.z), __expf(2.f * e4.w)};
        const float4 b4 = *in.slot(stage, 0);
        const float D0[4] = {ok[0] ? b4.x : 0.f, ok[1] ? b4.y : 0.f, ok[2] ? b4.z : 0.f, ok[3] ? b4.w : 0.f};
        float xr[4] = {D0[0], D0[1], D0[2], D0[3]};
#pragma unroll
        for (int i = 0; i < KCOV; ++i) {
          const float4 D = *in.slot(stage, i + 1);
          const float gi = s_g[seg][r][i];
          xr[0] = fmaf(gi, D.x, xr[0]); xr[1] = fmaf(gi, D.y, xr[1]);
          xr[2] = fmaf(gi, D.z, xr[2]); xr[3] = fmaf(gi, D.w, xr[3]);
        }
        float A[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const float rr = xs[l] - xr[l];
          A[l] = ok[l] ? -rr * ww[l] * invB : 0.f;
          de[l] = ok[l] ? (rr * rr * ww[l] - 1.f) * invB : 0.f;
        }
        float* out = dpre + (size_t)row * a.vp + v0;
        stg_stream(reinterpret_cast<float4*>(out),
                   make_float4(A[0] * D0[0] * (1.f - D0[0]), A[1] * D0[1] * (1.f - D0[1]),
                               A[2] * D0[2] * (1.f - D0[2]), A[3] * D0[3] * (1.f - D0[3])));
#pragma unroll
        for (int i = 0; i < KCOV; ++i) {
          const float4 G = ld4_keep(a.glm + (size_t)i * a.vp + v0);     // one broadcast request per warp
          const float4 D4 = *in.slot(stage, i + 1);
          const float Gs[4] = {G.x, G.y, G.z, G.w};
          const float Ds[4] = {D4.x, D4.y, D4.z, D4.w};
          const float gi = s_g[seg][r][i], cfi = s_cf[seg][r][i];
          float s = 0.f, o[4];
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            const float D = ok[l] ? Ds[l] : 0.f;
            const float dc = ok[l] ? A[l] + cfi * (gi * D - Gs[l]) : 0.f;     // d tot / d cons_i
            s = fmaf(D, dc, s);
            o[l] = gi * dc * D * (1.f - D);
          }
          acc[i] += s;
          stg_stream(reinterpret_cast<float4*>(out + (i + 1) * map_stride), make_float4(o[0], o[1], o[2], o[3]));
        }
      };
      if (row_ok && col < a.ncols) {
        if (v0 + 4 <= a.v) body(std::false_type{});
        else body(std::true_type{});
      }
      in.issue();                      // refill the stage just consumed
      stage = stage + 1 == STAGES ? 0 : stage + 1;
      // d eps: add the 4 rows of the item (lanes cl, cl+8, cl+16, cl+24), one vector RED per column
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        de[l] += __shfl_xor_sync(0xffffffffu, de[l], 8);
        de[l] += __shfl_xor_sync(0xffffffffu, de[l], 16);
      }
      if (r == 0 && col < a.ncols) atomicAdd(reinterpret_cast<float4*>(deps + v0), make_float4(de[0], de[1], de[2], de[3]));
    }
    __syncwarp();
    float mine = 0.f;
#pragma unroll
    for (int i = 0; i < KCOV; ++i) {
      const float s = row_sum(acc[i]);
      if (cl == i) mine = s;
    }
    partial[partial_index(blockIdx.x, seg, warp, r, NW) + cl] = mine;
  }
  cp_async_wait<0>();
}

static void fill_args(ReconArgs& a, const ReconPlan& p, const float* maps, const float* g, const float* x,
                      const float* eps, const float* glm, int b, long long v) {
  a.maps = maps; a.g = g; a.x = x; a.eps = eps; a.glm = glm;
  a.b = b; a.v = (int)v; a.vp = (int)((v + 3) / 4 * 4); a.ncols = a.vp / 4;
  a.n_c8 = p.n_c8; a.n_items = p.n_items; a.span = p.span;
}

// the rings are dynamic shared memory above the 48 KB static limit: opt in once per kernel
template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int NW, int CPS, int STAGES>
static int launch_fwd(const ReconArgs& a, int grid, float* ws, float* cons, float* x_rec, cudaStream_t st) {
  const size_t smem = ring_bytes(Shape{NW, CPS, STAGES});
  static bool ready = false;
  if (!ready) {
    VG_CUDA(allow_smem(recon_fwd_kernel<NW, CPS, STAGES, true>, smem));
    VG_CUDA(allow_smem(recon_fwd_kernel<NW, CPS, STAGES, false>, smem));
    ready = true;
  }
  if (cons || x_rec) recon_fwd_kernel<NW, CPS, STAGES, true><<<grid, NW * 32, smem, st>>>(a, ws, cons, x_rec);
  else recon_fwd_kernel<NW, CPS, STAGES, false><<<grid, NW * 32, smem, st>>>(a, ws, nullptr, nullptr);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

template <int NW, int CPS, int STAGES>
static int launch_bwd(const ReconArgs& a, int grid, const float* norms, float lam, float* dpre, float* ws,
                      float* deps, cudaStream_t st) {
  const size_t smem = ring_bytes(Shape{NW, CPS, STAGES});
  static bool ready = false;
  if (!ready) {
    VG_CUDA(allow_smem(recon_bwd_kernel<NW, CPS, STAGES>, smem));
    ready = true;
  }
  recon_bwd_kernel<NW, CPS, STAGES><<<grid, NW * 32, smem, st>>>(a, norms, lam, dpre, ws, deps);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

}  // namespace vg

using namespace vg;

extern "C" void vg_recon_tune(int variant) { g_variant = (variant >= 0 && variant < NVARIANT) ? variant : 3; }

// the work decomposition of a launch, for host-side inspection and tests: out = {row groups, warp items per
// row group, items, items per CTA (span), CTAs, warps per CTA}
extern "C" int vg_recon_plan(int b, long long v, int variant, int* out) {
  VG_CHECK_ARG(out && b > 0 && v >= 256, "bad arguments (v >= 256)");
  const Shape sh = shape_of((variant >= 0 && variant < NVARIANT) ? variant : 3);
  const ReconPlan p = make_plan(b, v, sh);
  out[0] = p.n_rg; out[1] = p.n_c8; out[2] = p.n_items; out[3] = p.span; out[4] = p.grid; out[5] = sh.warps;
  return VG_OK;
}

extern "C" size_t vg_recon_workspace_bytes(int b, long long v) {
  if (b <= 0 || v <= 0) return 256;
  size_t most = 0;
  for (int variant = 0; variant < NVARIANT; ++variant) {
    const Shape sh = shape_of(variant);
    const ReconPlan p = make_plan(b, v, sh);
    const size_t n = (size_t)p.grid * 2 * sh.warps * ROWS * NSLOT * sizeof(float);
    most = n > most ? n : most;
  }
  return most + 256;
}

extern "C" int vg_recon_loss_fwd(const float* maps, const float* g, const float* x, const float* eps,
                                 const float* glm, int b, long long v, float* logp, float* norms, float* cons,
                                 float* x_rec, void* workspace, size_t workspace_bytes, void* stream) {
  VG_CHECK_ARG(maps && g && x && eps && glm && logp && norms && b > 0 && v >= 256, "bad arguments (v >= 256)");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_recon_workspace_bytes(b, v), "workspace too small");
  const int variant = g_variant;
  const Shape sh = shape_of(variant);
  const ReconPlan p = make_plan(b, v, sh);
  ReconArgs a{};
  fill_args(a, p, maps, g, x, eps, glm, b, v);
  cudaStream_t st = as_stream(stream);
  float* ws = (float*)workspace;
  if (variant == 1) VG_TRY((launch_fwd<4, 3, 3>(a, p.grid, ws, cons, x_rec, st)));
  else if (variant == 2) VG_TRY((launch_fwd<4, 4, 2>(a, p.grid, ws, cons, x_rec, st)));
  else if (variant == 3) VG_TRY((launch_fwd<8, 2, 3>(a, p.grid, ws, cons, x_rec, st)));
  else VG_TRY((launch_fwd<4, 5, 2>(a, p.grid, ws, cons, x_rec, st)));
  recon_finalize<<<cdiv(b * 9 * 32, 256), 256, 0, st>>>(ws, b, p.n_c8, p.span, p.n_items, sh.warps, 9, 1, logp, norms);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

extern "C" int vg_recon_loss_bwd(const float* maps, const float* g, const float* x, const float* eps,
                                 const float* glm, const float* norms, int b, long long v, float lam,
                                 float* dpre, float* dg, float* deps, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  VG_CHECK_ARG(maps && g && x && eps && glm && norms && dpre && dg && deps && b > 0 && v >= 256, "bad arguments (v >= 256)");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_recon_workspace_bytes(b, v), "workspace too small");
  const int variant = g_variant;
  const Shape sh = shape_of(variant);
  const ReconPlan p = make_plan(b, v, sh);
  ReconArgs a{};
  fill_args(a, p, maps, g, x, eps, glm, b, v);
  cudaStream_t st = as_stream(stream);
  float* ws = (float*)workspace;
  VG_CUDA(cudaMemsetAsync(deps, 0, (size_t)a.vp * sizeof(float), st));
  if (variant == 1) VG_TRY((launch_bwd<4, 3, 3>(a, p.grid, norms, lam, dpre, ws, deps, st)));
  else if (variant == 2) VG_TRY((launch_bwd<4, 4, 2>(a, p.grid, norms, lam, dpre, ws, deps, st)));
  else if (variant == 3) VG_TRY((launch_bwd<8, 2, 3>(a, p.grid, norms, lam, dpre, ws, deps, st)));
  else VG_TRY((launch_bwd<4, 5, 2>(a, p.grid, norms, lam, dpre, ws, deps, st)));
  recon_finalize<<<cdiv(b * KCOV * 32, 256), 256, 0, st>>>(ws, b, p.n_c8, p.span, p.n_items, sh.warps, KCOV, 0, dg,
                                                           nullptr);
  VG_LAUNCH_CHECK();
  return VG_OK;
}
