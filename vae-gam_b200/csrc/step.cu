// Whole VAE-GAM step chained on one stream (vae_reg_GP.py:307-413 forward, :427-428 backward):
// encoder -> latent sample/KL -> 9 decoder passes run as ONE batch of 9B images with 9
// BatchNorm statistic groups -> gains -> fused reconstruction/likelihood/GLM pass, and the
// mirror-image backward.  No host synchronisation, no allocation: every intermediate lives
// in the caller's workspace, so the sequence can be captured into a CUDA graph.
#include <stdlib.h>

#include "common.cuh"
#include "conv_geom.cuh"
#include "gp_core.h"

namespace vg {

constexpr int V = 41 * 49 * 35;       // 70315 voxels (vae_reg_GP.py:32-33)
constexpr int VP = (V + 3) / 4 * 4;   // padded row
constexpr int L = 32, ZD = 41, KC = 8, NDEC = 9;

// parameter table indices: the reference's named_parameters() order
enum P {
  EPSILON = 0, SA_TASK = 1, LOGSTD_TASK = 2, GP0 = 3 /* 6 x {qu_m, qu_S, logkvar, logls, sa, logstd} */,
  SA_SEX = 39, LOGSTD_SEX = 40,
  CONV1 = 41, CONV2 = 43, CONV3 = 45, CONV4 = 47, CONV5 = 49,   // weight, bias
  BN1 = 51, BN3 = 53, BN5 = 55,
  FC1 = 57, FC2 = 59, FC31 = 61, FC32 = 63, FC33 = 65, FC41 = 67, FC42 = 69, FC43 = 71,
  FC5 = 73, FC6 = 75, FC7 = 77, FC8 = 79,
  CONVT1 = 81, CONVT2 = 83, CONVT3 = 85, CONVT4 = 87, CONVT5 = 89,
  BNT1 = 91, BNT3 = 93, BNT5 = 95
};

struct LayerShape { int transposed, cin, cout, k[3], stride, pad[3], opad[3], in[3], out[3]; };
// encoder (vae_reg_GP.py:189-193) and decoder (:211-215)
static const LayerShape kConv[5] = {
    {0, 1, 8, {3, 3, 3}, 1, {0, 0, 0}, {0, 0, 0}, {41, 49, 35}, {39, 47, 33}},
    {0, 8, 8, {3, 3, 3}, 2, {0, 0, 0}, {0, 0, 0}, {39, 47, 33}, {19, 23, 16}},
    {0, 8, 16, {3, 3, 3}, 1, {0, 0, 0}, {0, 0, 0}, {19, 23, 16}, {17, 21, 14}},
    {0, 16, 16, {3, 3, 3}, 2, {0, 0, 0}, {0, 0, 0}, {17, 21, 14}, {8, 10, 6}},
    {0, 16, 16, {3, 3, 3}, 1, {0, 0, 0}, {0, 0, 0}, {8, 10, 6}, {6, 8, 4}},
};
static const LayerShape kConvT[5] = {
    {1, 16, 16, {3, 3, 3}, 1, {0, 0, 0}, {0, 0, 0}, {6, 8, 5}, {8, 10, 7}},
    {1, 16, 16, {3, 3, 3}, 2, {1, 0, 1}, {1, 0, 1}, {8, 10, 7}, {16, 21, 14}},
    {1, 16, 8, {3, 3, 3}, 1, {0, 0, 0}, {0, 0, 0}, {16, 21, 14}, {18, 23, 16}},
    {1, 8, 8, {5, 3, 3}, 2, {0, 0, 0}, {0, 0, 0}, {18, 23, 16}, {39, 47, 33}},
    {1, 8, 1, {3, 3, 3}, 1, {0, 0, 0}, {0, 0, 0}, {39, 47, 33}, {41, 49, 35}},
};

// arith: VG_ARITH_FP32 / VG_ARITH_BF16 of this layer's calls (already resolved from the step config)
static VgConvDesc make_desc(const LayerShape& s, int n, int group, int arith, long long xs = 0, long long ys = 0) {
  VgConvDesc d{};
  d.arith = arith;
  d.transposed = s.transposed; d.cin = s.cin; d.cout = s.cout; d.stride = s.stride;
  for (int i = 0; i < 3; ++i) { d.k[i] = s.k[i]; d.pad[i] = s.pad[i]; d.opad[i] = s.opad[i]; d.in[i] = s.in[i]; d.out[i] = s.out[i]; }
  d.n = n; d.group_size = group; d.x_img_stride = xs; d.y_img_stride = ys;
  return d;
}
static long long vol(const int* g) { return (long long)g[0] * g[1] * g[2]; }

// ---- workspace carving -------------------------------------------------------------
struct Bump {
  char* base; size_t off;
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct BnBuf { double* stats; double* sums; float *scale, *shift, *istd, *mistd; };

struct EncWs {
  float *a1, *a2, *a3, *a4, *a5, *a5f, *h1, *h2, *h31, *h32, *h33, *heads;
  BnBuf bn1, bn3, bn5;
};
struct DecWs {
  float *f5, *f6, *f7, *f8, *t0, *t1, *t2, *t3, *t4;
  BnBuf bnt1, bnt3, bnt5;
};
struct StepWs {
  EncWs e; DecWs d;
  float *d_used, *klz, *zcat, *eps32, *logp, *norms;
  double* kl_terms; void* gain_ws; size_t gain_ws_bytes; void* recon_ws; size_t recon_ws_bytes;
  // backward
  uint16_t* d_t4h;
  double* j5_box; float *j5_raw, *j5_coef;
  float *dpre5, *d_t4, *d_t3, *d_t2, *d_t1, *d_t0, *d_f8, *d_f7, *d_f6, *d_f5, *d_zcat, *dheads, *d_h3, *d_h2,
      *d_h1, *d_a5f, *d_a5, *d_a4, *d_a3, *d_a2, *d_a1, *dg, *deps32, *dklz;
  double* zero_begin; size_t zero_bytes;   // contiguous region holding all BN stats / sums
};

static void take_bn(Bump& b, BnBuf& bn, int groups, int c) {
  bn.stats = b.take<double>((size_t)groups * c * 2);
  bn.sums = b.take<double>((size_t)groups * c * 2);
}
static void take_bn_coef(Bump& b, BnBuf& bn, int groups, int c) {
  bn.scale = b.take<float>((size_t)groups * c);
  bn.shift = b.take<float>((size_t)groups * c);
  bn.istd = b.take<float>((size_t)groups * c);
  bn.mistd = b.take<float>((size_t)groups * c);
}

static void carve_dec(Bump& b, DecWs& d, int nd) {
  d.f5 = b.take<float>((size_t)nd * 50);
  d.f6 = b.take<float>((size_t)nd * 100);
  d.f7 = b.take<float>((size_t)nd * 200);
  d.f8 = b.take<float>((size_t)nd * 3840);
  d.t0 = b.take<float>((size_t)nd * 240 * 16);
  d.t1 = b.take<float>((size_t)nd * 560 * 16);
  d.t2 = b.take<float>((size_t)nd * 4704 * 16);
  d.t3 = b.take<float>((size_t)nd * 6624 * 8);
  d.t4 = b.take<float>((size_t)nd * 60489 * 8);
}

static size_t carve_step(char* base, int B, int m, bool backward, StepWs& w) {
  Bump b{base, 0};
  const int nd = NDEC * B;
  // all BN statistics first, contiguous, so one memset clears them
  w.zero_begin = b.take<double>(0);
  take_bn(b, w.e.bn1, 1, 1); take_bn(b, w.e.bn3, 1, 8); take_bn(b, w.e.bn5, 1, 16);
  take_bn(b, w.d.bnt1, NDEC, 16); take_bn(b, w.d.bnt3, NDEC, 16); take_bn(b, w.d.bnt5, NDEC, 8);
  w.j5_box = b.take<double>((size_t)NDEC * 28);           // fused bnt5 junction: accumulated, so cleared with the statistics
  w.j5_raw = b.take<float>((size_t)NDEC * 27 * 8);
  b.off = (b.off + 255) & ~size_t(255);
  w.zero_bytes = b.off - (size_t)((char*)w.zero_begin - base);
  w.j5_coef = b.take<float>((size_t)NDEC * 8 * 3);
  take_bn_coef(b, w.e.bn1, 1, 1); take_bn_coef(b, w.e.bn3, 1, 8); take_bn_coef(b, w.e.bn5, 1, 16);
  take_bn_coef(b, w.d.bnt1, NDEC, 16); take_bn_coef(b, w.d.bnt3, NDEC, 16); take_bn_coef(b, w.d.bnt5, NDEC, 8);
  EncWs& e = w.e;
  e.a1 = b.take<float>((size_t)B * 60489 * 8);
  e.a2 = b.take<float>((size_t)B * 6992 * 8);
  e.a3 = b.take<float>((size_t)B * 4998 * 16);
  e.a4 = b.take<float>((size_t)B * 480 * 16);
  e.a5 = b.take<float>((size_t)B * 192 * 16);
  e.a5f = b.take<float>((size_t)B * 3072);
  e.h1 = b.take<float>((size_t)B * 200);
  e.h2 = b.take<float>((size_t)B * 100);
  e.h31 = b.take<float>((size_t)B * 50);
  e.h32 = b.take<float>((size_t)B * 50);
  e.h33 = b.take<float>((size_t)B * 50);
  e.heads = b.take<float>((size_t)3 * B * L);
  carve_dec(b, w.d, nd);
  w.d_used = b.take<float>((size_t)B * L);
  w.klz = b.take<float>(B);
  w.zcat = b.take<float>((size_t)nd * ZD);
  w.eps32 = b.take<float>(VP);
  w.logp = b.take<float>(B);
  w.norms = b.take<float>((size_t)KC * B);
  w.kl_terms = b.take<double>(2 * KC);
  w.gain_ws_bytes = vg_gain_workspace_bytes(B, m);
  w.gain_ws = b.take<char>(w.gain_ws_bytes);
  w.recon_ws_bytes = vg_recon_workspace_bytes(B, V);
  w.recon_ws = b.take<char>(w.recon_ws_bytes);
  w.dg = b.take<float>((size_t)KC * B);
  if (backward) {
    w.dpre5 = b.take<float>((size_t)nd * VP);
    w.d_t4 = b.take<float>((size_t)nd * 60489 * 8);
    w.d_t4h = b.take<uint16_t>((size_t)nd * 60489 * 8);      // bf16 copy of the final gradient (tensor-core arithmetic)
    w.d_t3 = b.take<float>((size_t)nd * 6624 * 8);
    w.d_t2 = b.take<float>((size_t)nd * 4704 * 16);
    w.d_t1 = b.take<float>((size_t)nd * 560 * 16);
    w.d_t0 = b.take<float>((size_t)nd * 240 * 16);
    w.d_f8 = b.take<float>((size_t)nd * 3840);
    w.d_f7 = b.take<float>((size_t)nd * 200);
    w.d_f6 = b.take<float>((size_t)nd * 100);
    w.d_f5 = b.take<float>((size_t)nd * 50);
    w.d_zcat = b.take<float>((size_t)nd * ZD);
    w.dheads = b.take<float>((size_t)3 * B * L);
    w.d_h3 = b.take<float>((size_t)3 * B * 50);
    w.d_h2 = b.take<float>((size_t)B * 100);
    w.d_h1 = b.take<float>((size_t)B * 200);
    w.d_a5f = b.take<float>((size_t)B * 3072);
    w.d_a5 = b.take<float>((size_t)B * 192 * 16);
    w.d_a4 = b.take<float>((size_t)B * 480 * 16);
    w.d_a3 = b.take<float>((size_t)B * 4998 * 16);
    w.d_a2 = b.take<float>((size_t)B * 6992 * 8);
    w.d_a1 = b.take<float>((size_t)B * 60489 * 8);
    w.deps32 = b.take<float>(VP);
    w.dklz = b.take<float>(B);
  }
  return b.off + 256;
}

// ---- small kernels -----------------------------------------------------------------
__global__ void cast_eps_kernel(const double* __restrict__ eps, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < VP) out[i] = i < V ? (float)eps[i] : 0.f;
}
__global__ void add_eps_grad_kernel(const float* __restrict__ d32, double* __restrict__ g64) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) g64[i] += (double)d32[i];
}
__global__ void fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// d[i] = a[i] > 0 ? d[i] : 0
__global__ void relu_mask_kernel(float* __restrict__ d, const float* __restrict__ a, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(a[i] > 0.f)) d[i] = 0.f;
}
// out = [tot, neg_elbo, gp_kl, glm_reg, mean logp, mean klz, 0, 0]   (vae_reg_GP.py:400-410)
__global__ void scalars_kernel(const float* __restrict__ logp, const float* __restrict__ klz,
                               const float* __restrict__ norms, const double* __restrict__ kl_terms, int B,
                               float gp_kl_scale, float glm_reg_scale, double* out) {
  double slp = 0, skl = 0, sn = 0, sgp = 0;
  for (int i = threadIdx.x; i < B; i += 32) { slp += logp[i]; skl += klz[i]; }
  for (int i = threadIdx.x; i < KC * B; i += 32) sn += norms[i];
  for (int i = threadIdx.x; i < 2 * KC; i += 32) sgp += kl_terms[i];
  slp = warp_sum(slp); skl = warp_sum(skl); sn = warp_sum(sn); sgp = warp_sum(sgp);
  if (threadIdx.x == 0) {
    const double neg_elbo = -(slp - skl) / B;
    const double glm = (double)B * sn;
    out[0] = neg_elbo + (double)gp_kl_scale * sgp + (double)glm_reg_scale * glm;
    out[1] = neg_elbo; out[2] = sgp; out[3] = glm; out[4] = slp / B; out[5] = skl / B; out[6] = 0; out[7] = 0;
  }
}

static int finalize_bn(const BnBuf& bn, const float* gamma, const float* beta, int groups, int c, double count,
                       cudaStream_t st) {
  return vg_bn_finalize(bn.stats, gamma, beta, groups, c, count, bn.scale, bn.shift, bn.istd, bn.mistd, st);
}

// Second stream for work that is off the critical path (the gain stage, every weight gradient): forked and
// joined with events, so the step is still one ordered unit of work on the caller's stream (and capturable).
// Disabled while per-operation profiling is on, so that the recorded times are those of un-overlapped kernels.
// One set of helper stream + events per host thread and per device: no state is shared between callers.
constexpr int kSideEvents = 24;
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[kSideEvents];
  cudaEvent_t join = nullptr, ready_main = nullptr, ready_side = nullptr;
  bool ok = false, tried = false;
  bool forked = false;      // a branch of this thread's step has not been joined into the main stream yet
  int n = 0;
};
static SideStream& side_stream() {
  constexpr int kMaxDev = 16;
  static thread_local SideStream per_dev[kMaxDev];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { (void)cudaGetLastError(); dev = 0; }
  SideStream& sd = per_dev[dev];
  if (!sd.tried) {
    sd.tried = true;
    const char* e = getenv("VAEGAM_SIDE_STREAM");
    if (!(e && e[0] == '0') && cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) == cudaSuccess) {
      sd.ok = true;
      for (int i = 0; i < kSideEvents; ++i) sd.ok = sd.ok && cudaEventCreateWithFlags(&sd.ev[i], cudaEventDisableTiming) == cudaSuccess;
      sd.ok = sd.ok && cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) == cudaSuccess;
      sd.ok = sd.ok && cudaEventCreateWithFlags(&sd.ready_main, cudaEventDisableTiming) == cudaSuccess;
      sd.ok = sd.ok && cudaEventCreateWithFlags(&sd.ready_side, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!sd.ok) (void)cudaGetLastError();
  }
  return sd;
}
struct Fork {
  SideStream& sd;
  cudaStream_t st;
  bool on;
  bool keep_open = false;   // phased backward: the branch stays open across calls, the last phase joins it
  int err = VG_OK;          // first CUDA error of a record / wait (returned by join())
  Fork(cudaStream_t main) : sd(side_stream()), st(main), on(sd.ok && !profiling()) {}
  // an early return between branch() and join() must not leave the helper stream un-joined (stream capture)
  ~Fork() { if (!keep_open) (void)join(); }
  void note(cudaError_t e, const char* what) {
    if (e != cudaSuccess && err == VG_OK) { set_error("step helper stream: %s: %s", what, cudaGetErrorString(e)); err = VG_ECUDA; }
  }
  // stream for an off-critical-path launch that depends on everything enqueued on the main stream so far
  cudaStream_t branch() {
    if (!on) return st;
    note(cudaEventRecord(sd.ev[sd.n], st), "cudaEventRecord");
    note(cudaStreamWaitEvent(sd.s, sd.ev[sd.n], 0), "cudaStreamWaitEvent");
    sd.n = (sd.n + 1) % kSideEvents;
    sd.forked = true;
    return err == VG_OK ? sd.s : st;
  }
  // the main stream waits for everything branched so far (by this or an earlier phase)
  int join() {
    if (on && sd.forked) {
      note(cudaEventRecord(sd.join, sd.s), "cudaEventRecord");
      note(cudaStreamWaitEvent(st, sd.join, 0), "cudaStreamWaitEvent");
      sd.forked = false;
    }
    return err;
  }
  // `rs` waits for everything enqueued so far on the main AND the helper stream; neither is joined
  int make_ready(cudaStream_t rs) {
    note(cudaEventRecord(sd.ready_main, st), "cudaEventRecord");
    note(cudaStreamWaitEvent(rs, sd.ready_main, 0), "cudaStreamWaitEvent");
    if (on && sd.forked) {
      note(cudaEventRecord(sd.ready_side, sd.s), "cudaEventRecord");
      note(cudaStreamWaitEvent(rs, sd.ready_side, 0), "cudaStreamWaitEvent");
    }
    return err;
  }
};

#define PF(idx) (reinterpret_cast<const float*>(io->params[idx]))
#define GF(idx) (reinterpret_cast<float*>(io->grads[idx]))

// The small fully-connected layers run as fused chains (csrc/mlp.cu): one launch per chain and direction.
static void mlp_layer(const VgStepIO* io, bool grads, VgMlpLayer& l, int param, int n, int k, int in, int out, int act) {
  l.w = PF(param); l.b = PF(param + 1);
  l.dw = grads ? GF(param) : nullptr; l.db = grads ? GF(param + 1) : nullptr;
  l.n = n; l.k = k; l.in = in; l.out = out; l.act = act; l.pad_ = 0;
}
static void mlp_buf(VgMlpBuf& b, float* act, float* grad, int width, int role) {
  b.act = act; b.grad = grad; b.width = width; b.role = role;
}
// fc2 -> fc31/32/33 -> fc41/42/43 (vae_reg_GP.py:245-251); heads = [mu | u | log d]
static VgMlp enc_head_mlp(const VgStepIO* io, const EncWs& e, int B, float* d_h1, float* dheads) {
  const bool g = dheads != nullptr;
  VgMlp m{};
  m.nlayers = 7; m.nbufs = 8; m.rows = B; m.rows_per_cta = B <= 64 ? 1 : (B <= 256 ? 2 : 4);   // >= ~32 CTAs
  mlp_buf(m.buf[0], e.h1, d_h1, 200, VG_MLP_INPUT | (g ? VG_MLP_GRAD_OUT : 0));
  mlp_buf(m.buf[1], e.h2, nullptr, 100, 0);
  mlp_buf(m.buf[2], e.h31, nullptr, 50, 0);
  mlp_buf(m.buf[3], e.h32, nullptr, 50, 0);
  mlp_buf(m.buf[4], e.h33, nullptr, 50, 0);
  for (int j = 0; j < 3; ++j)
    mlp_buf(m.buf[5 + j], e.heads + (size_t)j * B * L, g ? dheads + (size_t)j * B * L : nullptr, L, g ? VG_MLP_GRAD_IN : 0);
  mlp_layer(io, g, m.layer[0], FC2, 100, 200, 0, 1, VG_ACT_RELU);
  mlp_layer(io, g, m.layer[1], FC31, 50, 100, 1, 2, VG_ACT_RELU);
  mlp_layer(io, g, m.layer[2], FC32, 50, 100, 1, 3, VG_ACT_RELU);
  mlp_layer(io, g, m.layer[3], FC33, 50, 100, 1, 4, VG_ACT_RELU);
  mlp_layer(io, g, m.layer[4], FC41, L, 50, 2, 5, VG_ACT_NONE);
  mlp_layer(io, g, m.layer[5], FC42, L, 50, 3, 6, VG_ACT_NONE);
  mlp_layer(io, g, m.layer[6], FC43, L, 50, 4, 7, VG_ACT_NONE);
  return m;
}
// fc5 -> fc6 -> fc7 (vae_reg_GP.py:255-257)
static VgMlp dec_stem_mlp(const VgStepIO* io, const DecWs& d, const float* zcat, int nd, float* d_zcat, float* d_f7) {
  const bool g = d_f7 != nullptr;
  VgMlp m{};
  m.nlayers = 3; m.nbufs = 4; m.rows = nd; m.rows_per_cta = nd <= 320 ? 2 : (nd <= 1200 ? 4 : 8);   // ~one wave of CTAs
  mlp_buf(m.buf[0], const_cast<float*>(zcat), d_zcat, ZD, VG_MLP_INPUT | (g ? VG_MLP_GRAD_OUT : 0));
  mlp_buf(m.buf[1], d.f5, nullptr, 50, 0);
  mlp_buf(m.buf[2], d.f6, nullptr, 100, 0);
  mlp_buf(m.buf[3], d.f7, d_f7, 200, g ? VG_MLP_GRAD_IN : 0);
  mlp_layer(io, g, m.layer[0], FC5, 50, ZD, 0, 1, VG_ACT_RELU);
  mlp_layer(io, g, m.layer[1], FC6, 100, 50, 1, 2, VG_ACT_RELU);
  mlp_layer(io, g, m.layer[2], FC7, 200, 100, 2, 3, VG_ACT_RELU);
  return m;
}

// ---- forward pieces ----------------------------------------------------------------
static int run_encoder(const VgStepIO* io, const EncWs& e, int B, int arith, cudaStream_t st) {
  { VG_PROF("bn_stats", st);
  VG_TRY(vg_bn_stats(io->x, B, B, V, 1, e.bn1.stats, st));
  }
  { VG_PROF("bn1.bn_finalize", st);
  VG_TRY(finalize_bn(e.bn1, PF(BN1), PF(BN1 + 1), 1, 1, (double)B * V, st));
  }
  VgConvDesc d1 = make_desc(kConv[0], B, B, arith), d2 = make_desc(kConv[1], B, B, arith), d3 = make_desc(kConv[2], B, B, arith),
             d4 = make_desc(kConv[3], B, B, arith), d5 = make_desc(kConv[4], B, B, arith);
  { VG_PROF("conv1.fwd", st);
  VG_TRY(vg_conv_fwd(&d1, io->x, PF(CONV1), PF(CONV1 + 1), e.bn1.scale, e.bn1.shift, e.a1, VG_ACT_RELU, nullptr, st));
  }
  { VG_PROF("conv2.fwd", st);
  VG_TRY(vg_conv_fwd(&d2, e.a1, PF(CONV2), PF(CONV2 + 1), nullptr, nullptr, e.a2, VG_ACT_RELU, e.bn3.stats, st));
  }
  { VG_PROF("bn3.bn_finalize", st);
  VG_TRY(finalize_bn(e.bn3, PF(BN3), PF(BN3 + 1), 1, 8, (double)B * vol(kConv[1].out), st));
  }
  { VG_PROF("conv3.fwd", st);
  VG_TRY(vg_conv_fwd(&d3, e.a2, PF(CONV3), PF(CONV3 + 1), e.bn3.scale, e.bn3.shift, e.a3, VG_ACT_RELU, nullptr, st));
  }
  { VG_PROF("conv4.fwd", st);
  VG_TRY(vg_conv_fwd(&d4, e.a3, PF(CONV4), PF(CONV4 + 1), nullptr, nullptr, e.a4, VG_ACT_RELU, e.bn5.stats, st));
  }
  { VG_PROF("bn5.bn_finalize", st);
  VG_TRY(finalize_bn(e.bn5, PF(BN5), PF(BN5 + 1), 1, 16, (double)B * vol(kConv[3].out), st));
  }
  { VG_PROF("conv5.fwd", st);
  VG_TRY(vg_conv_fwd(&d5, e.a4, PF(CONV5), PF(CONV5 + 1), e.bn5.scale, e.bn5.shift, e.a5, VG_ACT_RELU, nullptr, st));
  }
  { VG_PROF("layout", st);
  VG_TRY(vg_nhwc_to_nchw(e.a5, e.a5f, B, 16, 192, st));   // h.view(-1, 3072) is channel-major
  }
  { VG_PROF("fc1.fwd", st);
  VG_TRY(vg_linear_fwd(e.a5f, PF(FC1), PF(FC1 + 1), e.h1, B, 200, 3072, VG_ACT_RELU, st));
  }
  { VG_PROF("fc2-fc43.fwd", st);
  const VgMlp m = enc_head_mlp(io, e, B, nullptr, nullptr);
  VG_TRY(vg_mlp_fwd(&m, st));
  }
  return VG_OK;
}

// nd images in groups of `group` share BatchNorm statistics; out rows have stride out_stride
static int run_decoder(const VgStepIO* io, const DecWs& d, const float* zcat, int nd, int group, float* out,
                       long long out_stride, int arith, cudaStream_t st) {
  const int groups = nd / group;
  { VG_PROF("fc5-fc7.fwd", st);
  const VgMlp m = dec_stem_mlp(io, d, zcat, nd, nullptr, nullptr);
  VG_TRY(vg_mlp_fwd(&m, st));
  }
  { VG_PROF("fc8.fwd", st);
  VG_TRY(vg_linear_fwd(d.f7, PF(FC8), PF(FC8 + 1), d.f8, nd, 3840, 200, VG_ACT_RELU, st));
  }
  { VG_PROF("layout", st);
  VG_TRY(vg_nchw_to_nhwc(d.f8, d.t0, nd, 16, 240, st));   // view(-1,16,6,8,5) -> channels-last
  }
  { VG_PROF("bnt1.bn_stats", st);
  VG_TRY(vg_bn_stats(d.t0, nd, group, 240, 16, d.bnt1.stats, st));
  }
  { VG_PROF("bnt1.bn_finalize", st);
  VG_TRY(finalize_bn(d.bnt1, PF(BNT1), PF(BNT1 + 1), groups, 16, (double)group * 240, st));
  }
  VgConvDesc c1 = make_desc(kConvT[0], nd, group, arith), c2 = make_desc(kConvT[1], nd, group, arith),
             c3 = make_desc(kConvT[2], nd, group, arith), c4 = make_desc(kConvT[3], nd, group, arith),
             c5 = make_desc(kConvT[4], nd, group, arith, 0, out_stride);
  // tensor-core arithmetic: t3 and t4 live in HBM as bf16 (they feed the TMA-staged convt4 / convt5 directly)
  if (arith == VG_ARITH_BF16) { c3.bf16_mask = VG_BF16_Y; c4.bf16_mask = VG_BF16_X | VG_BF16_Y; c5.bf16_mask = VG_BF16_X; }
  { VG_PROF("convt1.fwd", st);
  VG_TRY(vg_conv_fwd(&c1, d.t0, PF(CONVT1), PF(CONVT1 + 1), d.bnt1.scale, d.bnt1.shift, d.t1, VG_ACT_RELU, nullptr, st));
  }
  { VG_PROF("convt2.fwd", st);
  VG_TRY(vg_conv_fwd(&c2, d.t1, PF(CONVT2), PF(CONVT2 + 1), nullptr, nullptr, d.t2, VG_ACT_RELU, d.bnt3.stats, st));
  }
  { VG_PROF("bnt3.bn_finalize", st);
  VG_TRY(finalize_bn(d.bnt3, PF(BNT3), PF(BNT3 + 1), groups, 16, (double)group * vol(kConvT[1].out), st));
  }
  { VG_PROF("convt3.fwd", st);
  VG_TRY(vg_conv_fwd(&c3, d.t2, PF(CONVT3), PF(CONVT3 + 1), d.bnt3.scale, d.bnt3.shift, d.t3, VG_ACT_RELU, nullptr, st));
  }
  { VG_PROF("convt4.fwd", st);
  VG_TRY(vg_conv_fwd(&c4, d.t3, PF(CONVT4), PF(CONVT4 + 1), nullptr, nullptr, d.t4, VG_ACT_RELU, d.bnt5.stats, st));
  }
  { VG_PROF("bnt5.bn_finalize", st);
  VG_TRY(finalize_bn(d.bnt5, PF(BNT5), PF(BNT5 + 1), groups, 8, (double)group * vol(kConvT[3].out), st));
  }
  { VG_PROF("convt5.fwd", st);
  VG_TRY(vg_conv_fwd(&c5, d.t4, PF(CONVT5), PF(CONVT5 + 1), d.bnt5.scale, d.bnt5.shift, out, VG_ACT_SIGMOID, nullptr, st));
  }
  return VG_OK;
}

static void fill_gain_params(const VgStepConfig* cfg, const VgStepIO* io, VgGainParams& gp, VgGainGrads* gg) {
  for (int i = 0; i < KC; ++i) {
    gp.qu_m[i] = gp.qu_S[i] = gp.logkvar[i] = gp.logls[i] = gp.xu[i] = nullptr;
    gp.has_gp[i] = 0; gp.hrf[i] = 0;
    if (gg) gg->qu_m[i] = gg->qu_S[i] = gg->logkvar[i] = gg->logls[i] = nullptr;
  }
  gp.sa[0] = PF(SA_TASK); gp.logstd[0] = PF(LOGSTD_TASK);
  gp.sa[7] = PF(SA_SEX); gp.logstd[7] = PF(LOGSTD_SEX);
  gp.hrf[0] = cfg->neural_covariates ? 1 : 0;       // vae_reg_GP.py:377: i < num_covariates - 6
  if (gg) { gg->sa[0] = GF(SA_TASK); gg->logstd[0] = GF(LOGSTD_TASK); gg->sa[7] = GF(SA_SEX); gg->logstd[7] = GF(LOGSTD_SEX); }
  for (int k = 0; k < 6; ++k) {
    const int i = k + 1, base = GP0 + 6 * k;
    gp.qu_m[i] = PF(base); gp.qu_S[i] = PF(base + 1); gp.logkvar[i] = PF(base + 2); gp.logls[i] = PF(base + 3);
    gp.sa[i] = PF(base + 4); gp.logstd[i] = PF(base + 5); gp.xu[i] = io->xu[k]; gp.has_gp[i] = 1;   // :352
    if (gg) {
      gg->qu_m[i] = GF(base); gg->qu_S[i] = GF(base + 1); gg->logkvar[i] = GF(base + 2); gg->logls[i] = GF(base + 3);
      gg->sa[i] = GF(base + 4); gg->logstd[i] = GF(base + 5);
    }
  }
}

// Arithmetic of the step's convolutions from VgStepConfig.arith (include/vaegam.h):
// encoder forward | everything else
struct StepArith { int enc_fwd, rest; };
static StepArith step_arith(const VgStepConfig* cfg) {
  const int m = resolve_arith(cfg->arith);     // 0 fp32, 1 bf16, 2 mixed
  StepArith a;
  a.rest = m == 0 ? VG_ARITH_FP32 : VG_ARITH_BF16;
  a.enc_fwd = m == 1 ? VG_ARITH_BF16 : VG_ARITH_FP32;
  return a;
}

static int check_cfg(const VgStepConfig* cfg, const VgStepIO* io) {
  VG_CHECK_ARG(cfg && io, "null config / io");
  VG_CHECK_ARG(cfg->arith >= VG_ARITH_DEFAULT && cfg->arith <= VG_ARITH_MIXED, "arith must be a VG_ARITH_* value");
  VG_CHECK_ARG(cfg->b > 0 && cfg->b <= 4096, "batch out of range");
  VG_CHECK_ARG(cfg->m >= 2 && cfg->m <= kMaxInducing, "inducing points must be in [2,16]");
  for (int i = 0; i < VG_NUM_PARAMS; ++i) VG_CHECK_ARG(io->params[i] != nullptr, "null parameter pointer");
  return VG_OK;
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_step_workspace_bytes(const VgStepConfig* cfg) {
  if (!cfg || cfg->b <= 0) return 0;
  StepWs w;
  return carve_step(nullptr, cfg->b, cfg->m, true, w);
}

extern "C" int vg_step_fwd(const VgStepConfig* cfg, const VgStepIO* io, void* workspace, size_t workspace_bytes,
                           void* stream) {
  VG_TRY(check_cfg(cfg, io));
  VG_CHECK_ARG(io->x && io->covariates && io->eps_w && io->eps_d && io->eps_g && io->glm_t && io->taps, "null input");
  VG_CHECK_ARG(io->out_scalars && io->z && io->maps && io->g, "null output");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_step_workspace_bytes(cfg), "workspace too small");
  cudaStream_t st = as_stream(stream);
  const int B = cfg->b;
  StepWs w;
  carve_step((char*)workspace, B, cfg->m, true, w);
  VG_CUDA(cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st));
  cast_eps_kernel<<<cdiv(VP, 256), 256, 0, st>>>(reinterpret_cast<const double*>(io->params[EPSILON]), w.eps32);
  VG_LAUNCH_CHECK();
  Fork fk(st);
  VgGainParams gp;
  fill_gain_params(cfg, io, gp, nullptr);
  { cudaStream_t gs = fk.branch();          // the gains depend on the covariates and parameters only
  VG_PROF("gain.fwd", gs);
  VG_TRY(vg_gain_fwd(&gp, io->covariates, io->eps_g, io->taps, B, cfg->m, io->g, w.kl_terms, io->beta_mean,
                     io->beta_var, io->status, w.gain_ws, w.gain_ws_bytes, gs));
  }
  const StepArith ar = step_arith(cfg);
  VG_TRY(run_encoder(io, w.e, B, ar.enc_fwd, st));
  { VG_PROF("latent.fwd", st);
  VG_TRY(vg_latent_fwd(w.e.heads, io->eps_w, io->eps_d, B, io->z, w.klz, w.d_used, w.zcat,
                       io->status ? io->status + 8 : nullptr, st));
  }
  VG_TRY(run_decoder(io, w.d, w.zcat, NDEC * B, B, io->maps, VP, ar.rest, st));
  VG_TRY(fk.join());
  { VG_PROF("recon_loss.fwd", st);
  VG_TRY(vg_recon_loss_fwd(io->maps, io->g, io->x, w.eps32, io->glm_t, B, V, w.logp, w.norms,
                           cfg->want_maps ? io->cons : nullptr, cfg->want_maps ? io->x_rec : nullptr, w.recon_ws,
                           w.recon_ws_bytes, st));
  }
  scalars_kernel<<<1, 32, 0, st>>>(w.logp, w.klz, w.norms, w.kl_terms, B, cfg->gp_kl_scale, cfg->glm_reg_scale,
                                   io->out_scalars);
  VG_LAUNCH_CHECK();
  return VG_OK;
}

// phases: 1 << phase bits of what to run (see vg_step_bwd_phase in include/vaegam.h)
static int step_bwd_impl(const VgStepConfig* cfg, const VgStepIO* io, void* workspace, size_t workspace_bytes,
                         unsigned phases, cudaStream_t ready, cudaStream_t st) {
  VG_TRY(check_cfg(cfg, io));
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_step_workspace_bytes(cfg), "workspace too small");
  for (int i = 0; i < VG_NUM_PARAMS; ++i) VG_CHECK_ARG(io->grads[i] != nullptr, "null gradient pointer");
  const int B = cfg->b, nd = NDEC * B;
  StepWs w;
  carve_step((char*)workspace, B, cfg->m, true, w);
  const EncWs& e = w.e;
  const DecWs& d = w.d;
  const int ar = step_arith(cfg).rest;       // every backward pass uses the step's main arithmetic

  Fork fk(st);
  fk.keep_open = (phases & 4u) == 0;         // the last phase joins the helper stream
  if (phases & 1u) {
  // ---- objective
  { VG_PROF("recon_loss.bwd", st);
  VG_TRY(vg_recon_loss_bwd(io->maps, io->g, io->x, w.eps32, io->glm_t, w.norms, B, V, cfg->glm_reg_scale, w.dpre5,
                           w.dg, w.deps32, w.recon_ws, w.recon_ws_bytes, st));
  }
  add_eps_grad_kernel<<<cdiv(V, 256), 256, 0, st>>>(w.deps32, reinterpret_cast<double*>(io->grads[EPSILON]));
  VG_LAUNCH_CHECK();
  VgGainParams gp; VgGainGrads gg;
  fill_gain_params(cfg, io, gp, &gg);
  if (ar == VG_ARITH_BF16) {
    // fused bnt5 junction (below): the box sums of the reconstruction gradient need nothing but that gradient, so they
    // run on the helper stream beside convt5's grouped weight gradient; the finalize kernel joins the two
    cudaStream_t bs = fk.branch();
    VG_PROF("convt5.box_sums", bs);
    const int32_t ydims[3] = {kConvT[4].out[0], kConvT[4].out[1], kConvT[4].out[2]};
    const int32_t xdims[3] = {kConvT[4].in[0], kConvT[4].in[1], kConvT[4].in[2]};
    const int32_t pads[3] = {0, 0, 0};
    VG_TRY(vg_box_sums(w.dpre5, nd, B, ydims, VP, xdims, pads, w.j5_box, bs));
  }
  { cudaStream_t gs = fk.branch();
  VG_PROF("gain.bwd", gs);
  VG_TRY(vg_gain_bwd(&gp, &gg, io->covariates, io->eps_g, io->taps, w.dg, (double)cfg->gp_kl_scale, B, cfg->m,
                     w.gain_ws, w.gain_ws_bytes, gs));
  }

  // ---- decoder (9B images, groups of B)
  VgConvDesc c1 = make_desc(kConvT[0], nd, B, ar), c2 = make_desc(kConvT[1], nd, B, ar), c3 = make_desc(kConvT[2], nd, B, ar),
             c4 = make_desc(kConvT[3], nd, B, ar), c5 = make_desc(kConvT[4], nd, B, ar, 0, VP);
  const bool big16 = ar == VG_ARITH_BF16;
  // convt5: x = t4 (bf16); convt4: x = t3, y = t4 and both their gradients (bf16); convt3: y = t3 and its gradient
  if (big16) { c3.bf16_mask = VG_BF16_Y; c4.bf16_mask = VG_BF16_X | VG_BF16_Y | VG_BF16_DX; c5.bf16_mask = VG_BF16_X; }
  // tensor-core arithmetic: t4 (the largest activation) is stored as bf16 and so is its final gradient
  void* d_t4_final = big16 ? (void*)w.d_t4h : (void*)w.d_t4;
  if (big16) {
    // Fused bnt5 junction: the BatchNorm-backward statistics come from the weight-gradient products (which do not
    // depend on the data gradient), so the data gradient's epilogue applies BatchNorm backward + ReLU mask directly
    // and writes the final gradient once, as bf16 — no raw fp32 gradient, no separate pass over the largest tensor.
    { VG_PROF("convt5.wgrad", st);
    VG_TRY(vg_conv_wgrad_grouped(&c5, d.t4, w.dpre5, w.j5_raw, st));
    VG_TRY(fk.join());                            // the box sums (and the gain backward queued behind them)
    VG_TRY(vg_bn_fused_finalize(w.j5_raw, w.j5_box, PF(CONVT5), d.bnt5.scale, d.bnt5.shift, d.bnt5.istd, d.bnt5.mistd, NDEC, 8,
                                (double)B * vol(kConvT[3].out), GF(CONVT5), GF(CONVT5 + 1), GF(BNT5), GF(BNT5 + 1), w.j5_coef, st));
    }
    { VG_PROF("convt5.dgrad", st);
    VgConvDesc c5a = c5;
    c5a.bf16_mask = VG_BF16_X | VG_BF16_DX;
    VG_TRY(vg_conv_dgrad_bn_apply(&c5a, w.dpre5, PF(CONVT5), w.d_t4h, d.t4, w.j5_coef, GF(CONVT4 + 1), st));
    }
  } else {
  { cudaStream_t ws = fk.branch();
  VG_PROF("convt5.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&c5, d.t4, w.dpre5, d.bnt5.scale, d.bnt5.shift, GF(CONVT5), GF(CONVT5 + 1), ws));
  }
  { VG_PROF("convt5.dgrad", st);
  VG_TRY(vg_conv_dgrad(&c5, w.dpre5, PF(CONVT5), w.d_t4, nullptr, d.t4, d.bnt5.istd, d.bnt5.mistd, d.bnt5.sums, st));
  }
  { VG_PROF("bnt5.bn_bwd", st);
  VG_TRY(vg_bn_bwd_apply(w.d_t4, d.t4, d.bnt5.sums, d.bnt5.scale, d.bnt5.istd, d.bnt5.mistd, nd, B,
                         vol(kConvT[3].out), 8, (double)B * vol(kConvT[3].out), 1, 0,
                         w.d_t4, GF(BNT5), GF(BNT5 + 1), nullptr, st));
  }
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("convt4.wgrad", ws);
  // bf16 storage: convt4's bias gradient was summed by the BatchNorm backward above, so both operands of the
  // weight gradient are plain bf16 tensors that go to shared memory with asynchronous copies
  VG_TRY(vg_conv_wgrad(&c4, d.t3, d_t4_final, nullptr, nullptr, GF(CONVT4), big16 ? nullptr : GF(CONVT4 + 1), ws));
  }
  { VG_PROF("convt4.dgrad", st);
  VG_TRY(vg_conv_dgrad(&c4, d_t4_final, PF(CONVT4), w.d_t3, d.t3, nullptr, nullptr, nullptr, nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("convt3.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&c3, d.t2, w.d_t3, d.bnt3.scale, d.bnt3.shift, GF(CONVT3), GF(CONVT3 + 1), ws));
  }
  { VG_PROF("convt3.dgrad", st);
  VG_TRY(vg_conv_dgrad(&c3, w.d_t3, PF(CONVT3), w.d_t2, nullptr, d.t2, d.bnt3.istd, d.bnt3.mistd, d.bnt3.sums, st));
  }
  { VG_PROF("bnt3.bn_bwd", st);
  VG_TRY(vg_bn_bwd_apply(w.d_t2, d.t2, d.bnt3.sums, d.bnt3.scale, d.bnt3.istd, d.bnt3.mistd, nd, B,
                         vol(kConvT[1].out), 16, (double)B * vol(kConvT[1].out), 1, 0, w.d_t2, GF(BNT3), GF(BNT3 + 1), nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("convt2.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&c2, d.t1, w.d_t2, nullptr, nullptr, GF(CONVT2), GF(CONVT2 + 1), ws));
  }
  { VG_PROF("convt2.dgrad", st);
  VG_TRY(vg_conv_dgrad(&c2, w.d_t2, PF(CONVT2), w.d_t1, d.t1, nullptr, nullptr, nullptr, nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("convt1.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&c1, d.t0, w.d_t1, d.bnt1.scale, d.bnt1.shift, GF(CONVT1), GF(CONVT1 + 1), ws));
  }
  { VG_PROF("convt1.dgrad", st);
  VG_TRY(vg_conv_dgrad(&c1, w.d_t1, PF(CONVT1), w.d_t0, nullptr, d.t0, d.bnt1.istd, d.bnt1.mistd, d.bnt1.sums, st));
  }
  { VG_PROF("bnt1.bn_bwd", st);
  VG_TRY(vg_bn_bwd_apply(w.d_t0, d.t0, d.bnt1.sums, d.bnt1.scale, d.bnt1.istd, d.bnt1.mistd, nd, B, 240, 16,
                         (double)B * 240, 1, 0, w.d_t0, GF(BNT1), GF(BNT1 + 1), nullptr, st));
  }
  { VG_PROF("layout", st);
  VG_TRY(vg_nhwc_to_nchw(w.d_t0, w.d_f8, nd, 16, 240, st));     // gradient w.r.t. fc8 pre-activation
  }
  { cudaStream_t ws = fk.branch();                               // weight + bias gradient beside the data gradient
  VG_PROF("fc8.wgrad", ws);
  VG_TRY(vg_linear_bwd(w.d_f8, nullptr, d.f7, PF(FC8), nullptr, GF(FC8), GF(FC8 + 1), nd, 3840, 200, ws));
  }
  { VG_PROF("fc8.dgrad", st);
  VG_TRY(vg_linear_bwd(w.d_f8, nullptr, d.f7, PF(FC8), w.d_f7, nullptr, nullptr, nd, 3840, 200, st));
  }
  { VG_PROF("fc5-fc7.bwd", st);
  const VgMlp m = dec_stem_mlp(io, d, w.zcat, nd, w.d_zcat, w.d_f7);
  VG_TRY(vg_mlp_bwd(&m, st));
  }

  }
  if (phases & 2u) {
  // ---- latent
  fill_kernel<<<cdiv(B, 128), 128, 0, st>>>(w.dklz, B, 1.f / (float)B);   // tot = -mean(logp - klz)
  VG_LAUNCH_CHECK();
  { VG_PROF("latent.bwd", st);
  VG_TRY(vg_latent_bwd(e.heads, io->eps_w, io->eps_d, w.d_used, w.d_zcat, w.dklz, B, w.dheads, st));
  }

  // ---- encoder
  { VG_PROF("fc2-fc43.bwd", st);
  const VgMlp m = enc_head_mlp(io, e, B, w.d_h1, w.dheads);
  VG_TRY(vg_mlp_bwd(&m, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("fc1.wgrad", ws);
  VG_TRY(vg_linear_bwd(w.d_h1, e.h1, e.a5f, PF(FC1), nullptr, GF(FC1), GF(FC1 + 1), B, 200, 3072, ws));
  }
  { VG_PROF("fc1.dgrad", st);
  VG_TRY(vg_linear_bwd(w.d_h1, e.h1, e.a5f, PF(FC1), w.d_a5f, nullptr, nullptr, B, 200, 3072, st));
  }
  }
  if (phases & 4u) {
  { VG_PROF("layout", st);
  VG_TRY(vg_nchw_to_nhwc(w.d_a5f, w.d_a5, B, 16, 192, st));
  }
  relu_mask_kernel<<<cdiv((long long)B * 3072, 256), 256, 0, st>>>(w.d_a5, e.a5, (long long)B * 3072);
  VG_LAUNCH_CHECK();
  VgConvDesc d1 = make_desc(kConv[0], B, B, ar), d2 = make_desc(kConv[1], B, B, ar), d3 = make_desc(kConv[2], B, B, ar),
             d4 = make_desc(kConv[3], B, B, ar), d5 = make_desc(kConv[4], B, B, ar);
  { cudaStream_t ws = fk.branch();
  VG_PROF("conv5.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&d5, e.a4, w.d_a5, e.bn5.scale, e.bn5.shift, GF(CONV5), GF(CONV5 + 1), ws));
  }
  { VG_PROF("conv5.dgrad", st);
  VG_TRY(vg_conv_dgrad(&d5, w.d_a5, PF(CONV5), w.d_a4, nullptr, e.a4, e.bn5.istd, e.bn5.mistd, e.bn5.sums, st));
  }
  { VG_PROF("bn5.bn_bwd", st);
  VG_TRY(vg_bn_bwd_apply(w.d_a4, e.a4, e.bn5.sums, e.bn5.scale, e.bn5.istd, e.bn5.mistd, B, B, vol(kConv[3].out), 16,
                         (double)B * vol(kConv[3].out), 1, 0, w.d_a4, GF(BN5), GF(BN5 + 1), nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("conv4.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&d4, e.a3, w.d_a4, nullptr, nullptr, GF(CONV4), GF(CONV4 + 1), ws));
  }
  { VG_PROF("conv4.dgrad", st);
  VG_TRY(vg_conv_dgrad(&d4, w.d_a4, PF(CONV4), w.d_a3, e.a3, nullptr, nullptr, nullptr, nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("conv3.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&d3, e.a2, w.d_a3, e.bn3.scale, e.bn3.shift, GF(CONV3), GF(CONV3 + 1), ws));
  }
  { VG_PROF("conv3.dgrad", st);
  VG_TRY(vg_conv_dgrad(&d3, w.d_a3, PF(CONV3), w.d_a2, nullptr, e.a2, e.bn3.istd, e.bn3.mistd, e.bn3.sums, st));
  }
  { VG_PROF("bn3.bn_bwd", st);
  VG_TRY(vg_bn_bwd_apply(w.d_a2, e.a2, e.bn3.sums, e.bn3.scale, e.bn3.istd, e.bn3.mistd, B, B, vol(kConv[1].out), 8,
                         (double)B * vol(kConv[1].out), 1, 0, w.d_a2, GF(BN3), GF(BN3 + 1), nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("conv2.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&d2, e.a1, w.d_a2, nullptr, nullptr, GF(CONV2), GF(CONV2 + 1), ws));
  }
  { VG_PROF("conv2.dgrad", st);
  VG_TRY(vg_conv_dgrad(&d2, w.d_a2, PF(CONV2), w.d_a1, e.a1, nullptr, nullptr, nullptr, nullptr, st));
  }
  { cudaStream_t ws = fk.branch();
  VG_PROF("conv1.wgrad", ws);
  VG_TRY(vg_conv_wgrad(&d1, io->x, w.d_a1, e.bn1.scale, e.bn1.shift, GF(CONV1), GF(CONV1 + 1), ws));
  }
  // bn1 sits on the network input: only its affine parameters need a gradient
  { VG_PROF("conv1.dgrad", st);
  VG_TRY(vg_conv_dgrad(&d1, w.d_a1, PF(CONV1), nullptr, nullptr, io->x, e.bn1.istd, e.bn1.mistd, e.bn1.sums, st));
  }
  { VG_PROF("bn1.bn_bwd", st);
  VG_TRY(vg_bn_bwd_apply(nullptr, io->x, e.bn1.sums, nullptr, nullptr, nullptr, B, B, V, 1, (double)B * V, 0, 0, nullptr,
                         GF(BN1), GF(BN1 + 1), nullptr, st));
  }
  }
  if (ready) VG_TRY(fk.make_ready(ready));
  if (phases & 4u) VG_TRY(fk.join());
  return fk.err;
}

extern "C" int vg_step_bwd(const VgStepConfig* cfg, const VgStepIO* io, void* workspace, size_t workspace_bytes,
                           void* stream) {
  return step_bwd_impl(cfg, io, workspace, workspace_bytes, 7u, nullptr, as_stream(stream));
}

extern "C" int vg_step_bwd_phase(const VgStepConfig* cfg, const VgStepIO* io, void* workspace, size_t workspace_bytes,
                                 int phase, void* ready_stream, void* stream) {
  VG_CHECK_ARG(phase >= 0 && phase < VG_BWD_PHASES, "phase must be 0, 1 or 2");
  return step_bwd_impl(cfg, io, workspace, workspace_bytes, 1u << phase, as_stream(ready_stream), as_stream(stream));
}

extern "C" int vg_encode_fwd(const VgStepConfig* cfg, const VgStepIO* io, float* heads, void* workspace,
                             size_t workspace_bytes, void* stream) {
  VG_TRY(check_cfg(cfg, io));
  VG_CHECK_ARG(io->x && heads, "null input");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_step_workspace_bytes(cfg), "workspace too small");
  cudaStream_t st = as_stream(stream);
  StepWs w;
  carve_step((char*)workspace, cfg->b, cfg->m, true, w);
  VG_CUDA(cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st));
  VG_TRY(run_encoder(io, w.e, cfg->b, step_arith(cfg).enc_fwd, st));
  VG_CUDA(cudaMemcpyAsync(heads, w.e.heads, (size_t)3 * cfg->b * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return VG_OK;
}

extern "C" size_t vg_decode_workspace_bytes(int n) {
  if (n <= 0) return 0;
  Bump b{nullptr, 0};
  BnBuf bn;
  DecWs d;
  take_bn(b, bn, 1, 16); take_bn(b, bn, 1, 16); take_bn(b, bn, 1, 8);
  take_bn_coef(b, bn, 1, 16); take_bn_coef(b, bn, 1, 16); take_bn_coef(b, bn, 1, 8);
  carve_dec(b, d, n);
  return b.off + 512;
}

extern "C" int vg_decode_fwd(const VgStepIO* io, const float* zcat, int n, int arith, float* out, void* workspace,
                             size_t workspace_bytes, void* stream) {
  VG_CHECK_ARG(io && zcat && out && n > 0, "bad arguments");
  VG_CHECK_ARG(arith >= VG_ARITH_DEFAULT && arith <= VG_ARITH_MIXED, "arith must be a VG_ARITH_* value");
  VG_CHECK_ARG(workspace && workspace_bytes >= vg_decode_workspace_bytes(n), "workspace too small");
  cudaStream_t st = as_stream(stream);
  Bump b{(char*)workspace, 0};
  DecWs d;
  double* z0 = b.take<double>(0);
  take_bn(b, d.bnt1, 1, 16); take_bn(b, d.bnt3, 1, 16); take_bn(b, d.bnt5, 1, 8);
  b.off = (b.off + 255) & ~size_t(255);
  const size_t zbytes = b.off - (size_t)((char*)z0 - (char*)workspace);
  take_bn_coef(b, d.bnt1, 1, 16); take_bn_coef(b, d.bnt3, 1, 16); take_bn_coef(b, d.bnt5, 1, 8);
  carve_dec(b, d, n);
  VG_CUDA(cudaMemsetAsync(z0, 0, zbytes, st));
  // dense (n, V) output like VAE.decode
  return run_decoder(io, d, zcat, n, n, out, V, resolve_arith(arith) == 0 ? VG_ARITH_FP32 : VG_ARITH_BF16, st);
}
