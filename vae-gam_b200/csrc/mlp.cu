// Fused chains of the small fully-connected layers: the encoder heads fc2 -> fc31/32/33 -> fc41/42/43
// (vae_reg_GP.py:198-204, used at :245-251) and the decoder stem fc5 -> fc6 -> fc7 (:207-209, :255-257),
// forward and backward, ONE launch per chain and direction.
//
// These layers hold < 0.1 % of the step's FLOPs but were 23 dependent ~20 us GEMM launches (each a
// handful of CTAs waiting on global-load latency).  Every batch row is independent through a chain,
// so a CTA owns R (1..8) rows and walks the whole layer list with the activations (and, backward, their
// gradients) resident in shared memory.  All weights and biases of the chain (108 KB / 159 KB) are
// copied into shared memory up front with one burst of cp.async, so no layer waits on L2 again.
//   forward : y[r][n] = act(b[n] + sum_k x[r][k] W[n][k])   warp per 4 neurons, lanes along k
//             (conflict-free weight rows)
//   backward: dym = dy * (y > 0);  dx[r][k] += sum_n dym[r][n] W[n][k]   thread per k, n split over
//             thread groups (coalesced weight rows; fan-in of branches adds up in shared memory);
//             dW[n][k] += sum_r dym[r][n] x[r][k], db[n] += sum_r dym[r][n]   fp32 RED to global
// The big layers fc1 / fc8 (2.4 / 3.1 MB of weights) stay on the tiled GEMM in linear.cu.
#include "common.cuh"

namespace vg {

constexpr int MLP_THREADS = 256;
constexpr int MLP_KSTEPS = 7;        // lanes along k: K <= 32 * 7 = 224
constexpr int MLP_NB = 4;            // neurons per warp iteration

struct MlpOffsets {
  int off[VG_MLP_MAX_BUFS + 1];        // activation columns before buffer i (x rows_per_cta floats)
  int woff[VG_MLP_MAX_LAYERS + 1];     // weight floats before layer l
  int boff[VG_MLP_MAX_LAYERS + 1];     // bias floats before layer l
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// weights (parameter tensors are only 4-byte aligned inside the flat buffer) and biases -> shared memory
__device__ __forceinline__ void stage_weights(const VgMlp& m, const MlpOffsets& o, float* sw, float* sb) {
  for (int l = 0; l < m.nlayers; ++l) {
    const VgMlpLayer& L = m.layer[l];
    const int nw = L.n * L.k;
    for (int i = threadIdx.x; i < nw; i += MLP_THREADS) cp_async4(sw + o.woff[l] + i, L.w + i);
    for (int i = threadIdx.x; i < L.n; i += MLP_THREADS) {
      if (L.b) cp_async4(sb + o.boff[l] + i, L.b + i);
      else sb[o.boff[l] + i] = 0.f;
    }
  }
}

__device__ __forceinline__ float apply_act(float v, int act) { return act == VG_ACT_RELU ? fmaxf(v, 0.f) : v; }

template <int R>
__device__ __forceinline__ void dense_fwd(const float* __restrict__ sx, int K, const float* __restrict__ W,
                                          const float* __restrict__ b, int N, float* __restrict__ sy, int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n0 = warp * MLP_NB; n0 < N; n0 += (MLP_THREADS / 32) * MLP_NB) {
    float w[MLP_KSTEPS][MLP_NB];
#pragma unroll
    for (int s = 0; s < MLP_KSTEPS; ++s)
#pragma unroll
      for (int j = 0; j < MLP_NB; ++j) {
        const int k = lane + 32 * s;
        w[s][j] = (k < K && n0 + j < N) ? W[(n0 + j) * K + k] : 0.f;
      }
    float acc[MLP_NB][R];
#pragma unroll
    for (int j = 0; j < MLP_NB; ++j)
#pragma unroll
      for (int r = 0; r < R; ++r) acc[j][r] = 0.f;
#pragma unroll
    for (int s = 0; s < MLP_KSTEPS; ++s) {
      const int k = lane + 32 * s;
      if (k < K) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float xv = sx[r * K + k];
#pragma unroll
          for (int j = 0; j < MLP_NB; ++j) acc[j][r] = fmaf(w[s][j], xv, acc[j][r]);
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < MLP_NB; ++j)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float s = warp_sum(acc[j][r]);
        if (lane == j * R + r) mine = s;
      }
    if (lane < MLP_NB * R) {
      const int j = lane / R, r = lane % R;
      if (n0 + j < N) sy[r * N + n0 + j] = apply_act(mine + b[n0 + j], act);
    }
  }
}

// sdx[r][k] += sum_n sdym[r][n] * W[n][k]
template <int R>
__device__ __forceinline__ void dense_dx(const float* __restrict__ sdym, int N, const float* __restrict__ W, int K,
                                         float* sdx) {
  const int kpad = (K + 31) & ~31;
  const int groups = MLP_THREADS / kpad > 0 ? MLP_THREADS / kpad : 1;
  for (int t = threadIdx.x; t < kpad * groups; t += MLP_THREADS) {
    const int grp = t / kpad, k = t - grp * kpad;
    if (k >= K) continue;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    constexpr int U = 8;
    for (int n0 = grp * U; n0 < N; n0 += groups * U) {
      float w[U];
#pragma unroll
      for (int j = 0; j < U; ++j) w[j] = n0 + j < N ? W[(n0 + j) * K + k] : 0.f;
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (n0 + j < N) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmaf(w[j], sdym[r * N + n0 + j], acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) atomicAdd(sdx + r * K + k, acc[r]);
  }
}

// dW[n][k] += sum_r sdym[r][n] * sx[r][k];  db[n] += sum_r sdym[r][n]
template <int R>
__device__ __forceinline__ void dense_dw(const float* __restrict__ sdym, int N, const float* __restrict__ sx, int K,
                                         float* dW, float* db) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < N; n += MLP_THREADS / 32) {
    float d[R];
#pragma unroll
    for (int r = 0; r < R; ++r) d[r] = sdym[r * N + n];
    if (dW) {
      for (int k = lane; k < K; k += 32) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) s = fmaf(d[r], sx[r * K + k], s);
        atomicAdd(dW + (size_t)n * K + k, s);
      }
    }
    if (db && lane == 0) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) s += d[r];
      atomicAdd(db + n, s);
    }
  }
}

// rows [row0, row0 + nrows) of a (rows, width) buffer -> shared memory (asynchronously), zero beyond nrows
__device__ __forceinline__ void load_rows(float* dst, const float* src, int width, int row0, int nrows, int R) {
  const int live = src ? nrows * width : 0;
  for (int i = threadIdx.x; i < R * width; i += MLP_THREADS) {
    if (i < live) cp_async4(dst + i, src + (size_t)row0 * width + i);
    else dst[i] = 0.f;
  }
}
__device__ __forceinline__ void store_rows(float* dst, const float* src, int width, int row0, int nrows) {
  for (int i = threadIdx.x; i < nrows * width; i += MLP_THREADS) dst[(size_t)row0 * width + i] = src[i];
}

template <int R>
__global__ void __launch_bounds__(MLP_THREADS) mlp_fwd_kernel(const VgMlp m, const MlpOffsets o) {
  extern __shared__ float sm[];
  float* sw = sm + R * o.off[m.nbufs];
  float* sb = sw + o.woff[m.nlayers];
  const int row0 = blockIdx.x * R, nrows = min(R, m.rows - row0);
  stage_weights(m, o, sw, sb);
  for (int i = 0; i < m.nbufs; ++i)
    if (m.buf[i].role & VG_MLP_INPUT) load_rows(sm + R * o.off[i], m.buf[i].act, m.buf[i].width, row0, nrows, R);
  cp_async_commit_wait_all();
  __syncthreads();
  for (int l = 0; l < m.nlayers; ++l) {
    const VgMlpLayer& L = m.layer[l];
    float* sy = sm + R * o.off[L.out];
    dense_fwd<R>(sm + R * o.off[L.in], L.k, sw + o.woff[l], sb + o.boff[l], L.n, sy, L.act);
    __syncthreads();
    store_rows(m.buf[L.out].act, sy, L.n, row0, nrows);
  }
}

template <int R>
__global__ void __launch_bounds__(MLP_THREADS) mlp_bwd_kernel(const VgMlp m, const MlpOffsets o) {
  extern __shared__ float sm[];
  float* sact = sm;
  float* sgrad = sm + R * o.off[m.nbufs];
  float* sw = sgrad + R * o.off[m.nbufs];
  float* sb = sw + o.woff[m.nlayers];
  const int row0 = blockIdx.x * R, nrows = min(R, m.rows - row0);
  stage_weights(m, o, sw, sb);
  for (int i = 0; i < m.nbufs; ++i) {
    load_rows(sact + R * o.off[i], m.buf[i].act, m.buf[i].width, row0, nrows, R);
    load_rows(sgrad + R * o.off[i], (m.buf[i].role & VG_MLP_GRAD_IN) ? m.buf[i].grad : nullptr, m.buf[i].width, row0,
              nrows, R);
  }
  cp_async_commit_wait_all();
  __syncthreads();
  for (int l = m.nlayers - 1; l >= 0; --l) {
    const VgMlpLayer& L = m.layer[l];
    float* dy = sgrad + R * o.off[L.out];
    const float* y = sact + R * o.off[L.out];
    if (L.act == VG_ACT_RELU) {
      for (int i = threadIdx.x; i < R * L.n; i += MLP_THREADS)
        if (!(y[i] > 0.f)) dy[i] = 0.f;
      __syncthreads();
    }
    bool need_dx = (m.buf[L.in].role & VG_MLP_GRAD_OUT) != 0;
    for (int q = 0; q < l; ++q) need_dx = need_dx || m.layer[q].out == L.in;
    if (need_dx) dense_dx<R>(dy, L.n, sw + o.woff[l], L.k, sgrad + R * o.off[L.in]);
    dense_dw<R>(dy, L.n, sact + R * o.off[L.in], L.k, L.dw, L.db);
    __syncthreads();
  }
  for (int i = 0; i < m.nbufs; ++i)
    if (m.buf[i].role & VG_MLP_GRAD_OUT) store_rows(m.buf[i].grad, sgrad + R * o.off[i], m.buf[i].width, row0, nrows);
}

static int check_mlp(const VgMlp* m, MlpOffsets& o, bool backward) {
  VG_CHECK_ARG(m && m->nlayers >= 1 && m->nlayers <= VG_MLP_MAX_LAYERS && m->nbufs >= 2 && m->nbufs <= VG_MLP_MAX_BUFS,
               "layer / buffer count");
  VG_CHECK_ARG(m->rows > 0 && (m->rows_per_cta == 1 || m->rows_per_cta == 2 || m->rows_per_cta == 4 || m->rows_per_cta == 8),
               "rows, rows_per_cta (1, 2, 4 or 8)");
  o.off[0] = 0;
  for (int i = 0; i < m->nbufs; ++i) {
    VG_CHECK_ARG(m->buf[i].width > 0 && m->buf[i].act, "buffer width / activation pointer");
    if (backward && (m->buf[i].role & (VG_MLP_GRAD_IN | VG_MLP_GRAD_OUT))) VG_CHECK_ARG(m->buf[i].grad, "gradient pointer");
    o.off[i + 1] = o.off[i] + m->buf[i].width;
  }
  o.woff[0] = o.boff[0] = 0;
  for (int l = 0; l < m->nlayers; ++l) {
    const VgMlpLayer& L = m->layer[l];
    o.woff[l + 1] = o.woff[l] + L.n * L.k;
    o.boff[l + 1] = o.boff[l] + L.n;
    VG_CHECK_ARG(L.w && L.in >= 0 && L.in < m->nbufs && L.out >= 0 && L.out < m->nbufs && L.in != L.out, "layer wiring");
    VG_CHECK_ARG(L.k == m->buf[L.in].width && L.n == m->buf[L.out].width, "layer shape vs buffer width");
    VG_CHECK_ARG(L.k <= 32 * MLP_KSTEPS && L.k <= MLP_THREADS, "layer input width > 224");
    VG_CHECK_ARG(L.act == VG_ACT_NONE || L.act == VG_ACT_RELU, "activation");
    VG_CHECK_ARG(!(m->buf[L.out].role & VG_MLP_INPUT), "a layer writes a forward input");
  }
  return VG_OK;
}

template <typename K>
static int launch_mlp(K kernel, const VgMlp* m, const MlpOffsets& o, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) VG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<cdiv(m->rows, m->rows_per_cta), MLP_THREADS, smem, st>>>(*m, o);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_mlp_fwd(const VgMlp* m, void* stream) {
  MlpOffsets o;
  VG_TRY(check_mlp(m, o, false));
  const size_t smem = ((size_t)m->rows_per_cta * o.off[m->nbufs] + o.woff[m->nlayers] + o.boff[m->nlayers]) * sizeof(float);
  VG_CHECK_ARG(smem <= 224 * 1024, "weights + activations of rows_per_cta rows exceed shared memory");
  cudaStream_t st = as_stream(stream);
  switch (m->rows_per_cta) {
    case 1: return launch_mlp(mlp_fwd_kernel<1>, m, o, smem, st);
    case 2: return launch_mlp(mlp_fwd_kernel<2>, m, o, smem, st);
    case 4: return launch_mlp(mlp_fwd_kernel<4>, m, o, smem, st);
    default: return launch_mlp(mlp_fwd_kernel<8>, m, o, smem, st);
  }
}

extern "C" int vg_mlp_bwd(const VgMlp* m, void* stream) {
  MlpOffsets o;
  VG_TRY(check_mlp(m, o, true));
  const size_t smem = ((size_t)2 * m->rows_per_cta * o.off[m->nbufs] + o.woff[m->nlayers] + o.boff[m->nlayers]) * sizeof(float);
  VG_CHECK_ARG(smem <= 224 * 1024, "weights + activations of rows_per_cta rows exceed shared memory");
  cudaStream_t st = as_stream(stream);
  switch (m->rows_per_cta) {
    case 1: return launch_mlp(mlp_bwd_kernel<1>, m, o, smem, st);
    case 2: return launch_mlp(mlp_bwd_kernel<2>, m, o, smem, st);
    case 4: return launch_mlp(mlp_bwd_kernel<4>, m, o, smem, st);
    default: return launch_mlp(mlp_bwd_kernel<8>, m, o, smem, st);
  }
}
