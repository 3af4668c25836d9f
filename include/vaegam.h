/*
 * vaegam.h — C ABI of libvaegam_sm100.so: the B200 (sm_100a) hot path of VAE-GAM.
 *
 * The reference (dannyfa/VAE-GAM) has no FFI: its hot path is Python calling PyTorch
 * library kernels.  Each entry point below names the reference call site(s) it
 * replaces (file:line in the reference tree).  The drop-in Python modules in
 * vae-gam_b200/ (vae_reg_GP.py, gp.py) bind these through ctypes; INTEGRATION.md shows
 * the binding a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is DEVICE memory owned by the caller unless
 *     a parameter is documented as host memory; no allocation, no ownership transfer;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the host, so every call is CUDA-graph capturable;
 *   - return 0 on success or a negative VG_E* code; vg_last_error() gives a message
 *     (thread-local);
 *   - activations are fp32, channels-last (N, D, H, W, C); weights keep PyTorch's
 *     layouts (Conv3d: (Cout,Cin,kD,kH,kW); ConvTranspose3d: (Cin,Cout,kD,kH,kW);
 *     Linear: (out,in)), so checkpoints interchange with the reference byte for byte;
 *   - BatchNorm3d(track_running_stats=False) is never a kernel of its own: batch
 *     statistics are reduced per (group, channel) — group = image_index / group_size, so
 *     the 9 decoder passes keep 9 separate statistics (vae_reg_GP.py:326-343 calls
 *     decode 9 times) — and the normalisation is folded into the consumer's operand load.
 */
#ifndef VAEGAM_H_
#define VAEGAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_OK 0
#define VG_EINVAL (-1)  /* bad shape / null pointer / unsupported channel count */
#define VG_ECUDA (-2)   /* CUDA runtime error; see vg_last_error() */
#define VG_ENOTPD (-3)  /* reported through a device flag: non-positive-definite pivot */

#define VG_ACT_NONE 0
#define VG_ACT_RELU 1
#define VG_ACT_SIGMOID 2

/* library identity / diagnostics */
int vg_version(void);                 /* 100 * major + minor */
const char* vg_last_error(void);      /* thread-local, never NULL */
int vg_sm_count(void);                /* SMs of the current device (148 on B200) */
long long vg_launch_count(void);      /* kernels launched by this library so far (process-wide) */
/* Per-operation timing of the whole-step entry points with CUDA events on the launching stream
 * (bench.py's live roofline).  enable(1) clears and starts recording, enable(0) stops;
 * collect() waits for the recorded events and writes "name<TAB>ms\n" lines into buf (host memory),
 * returning the number of records.  Not for use inside CUDA-graph capture. */
int vg_profile_enable(int on);
long long vg_profile_collect(char* buf, size_t cap);

/* ------------------------------------------------------------------------------------
 * 3-D convolution family.  One descriptor covers Conv3d and ConvTranspose3d.
 * Replaces torch nn.Conv3d / nn.ConvTranspose3d forward + autograd backward at
 * vae_reg_GP.py:238-242 (conv1..5, layers :189-193) and :260-264 (convt1..5, :211-215).
 * ---------------------------------------------------------------------------------- */
typedef struct VgConvDesc {
  int32_t transposed;      /* 0: Conv3d, 1: ConvTranspose3d */
  int32_t cin, cout;       /* module in/out channels; each in {1, 8, 16} */
  int32_t k[3];            /* kernel (kD,kH,kW), each <= 5 */
  int32_t stride;          /* 1 or 2 (same in all dims) */
  int32_t pad[3];          /* ConvTranspose3d padding (0 for Conv3d) */
  int32_t opad[3];         /* ConvTranspose3d output_padding */
  int32_t in[3];           /* input grid (D,H,W) */
  int32_t out[3];          /* output grid (D,H,W) — must satisfy the PyTorch size formula */
  int32_t n;               /* images */
  int32_t group_size;      /* images per BatchNorm statistics group (n % group_size == 0) */
  int32_t arith;           /* VG_ARITH_*: arithmetic of THIS call (0 = the process default below) */
  int32_t bf16_mask;       /* storage of the big activations (tensor-core arithmetic only; 0 = all fp32):
                              VG_BF16_X  x / bn_x / mask_act hold bf16 (2 bytes per element, same layout),
                              VG_BF16_Y  y / dy hold bf16, VG_BF16_DX  dx is written as bf16.
                              Supported for 8- and 16-channel tensors (a voxel = one or two 16-byte words) */
  int32_t reserved_;
  int64_t x_img_stride;    /* floats between consecutive images of x / dx; 0 = dense (D*H*W*cin) */
  int64_t y_img_stride;    /* floats between consecutive images of y / dy; 0 = dense (D*H*W*cout) */
} VgConvDesc;

/* Arithmetic of the convolution kernels.  Every call carries its own choice — VgConvDesc.arith for
 * the per-layer entry points, VgStepConfig.arith for the whole-step ones — so callers that state it
 * are independent of any process state:
 *   VG_ARITH_FP32 : fp32 CUDA-core kernels everywhere ("check mode", 1e-5 parity with PyTorch fp32);
 *   VG_ARITH_BF16 : bf16 operands on the tcgen05 tensor cores with fp32 accumulation in TMEM for
 *      every geometry the implicit-GEMM kernels cover, mma.sync bf16 weight gradients; the
 *      remaining geometries use the fp32 kernels;
 *   VG_ARITH_MIXED (whole-step entry points; per-layer calls treat it as BF16): BF16 for the decoder
 *      and for every backward pass, FP32 for the five forward convolutions of the ENCODER.  The
 *      encoder's forward rounding is what dominates the gradient deviation of the BF16 mode
 *      (tools/precision_study.py: 25 % -> 2 % on the encoder's weight gradients at B = 32) while
 *      the encoder is 9 % of the convolution work;
 *   VG_ARITH_DEFAULT (0, what a zero-initialised struct says): the process default, which is
 *      read once from the environment variable VAEGAM_CONV_MODE ("0"/"fp32", "1"/"bf16",
 *      "2"/"mixed"; unset = mixed) and can be changed with vg_set_conv_mode(0 | 1 | 2).
 * vg_set_conv_mode / vg_set_conv_tuning / vg_recon_tune only move such process DEFAULTS and
 * benchmarking knobs; no entry point keeps per-call state in the library. */
#define VG_BF16_X 1
#define VG_BF16_Y 2
#define VG_BF16_DX 4
#define VG_ARITH_DEFAULT 0
#define VG_ARITH_FP32 1
#define VG_ARITH_BF16 2
#define VG_ARITH_MIXED 3
int vg_set_conv_mode(int mode);     /* 0 fp32, 1 bf16, 2 mixed */
int vg_get_conv_mode(void);
/* Dispatch tuning (benchmarking knob, process default).  "t2_min_voxels": output voxels per launch from
 * which the persistent plane-folded tcgen05 kernel is preferred over the per-tile kernels (default
 * 400000; env VAEGAM_T2_MIN_VOXELS). */
int vg_set_conv_tuning(const char* key, long long value);

/* Host-side description of the launches a layer maps to under d->arith (kind 0: forward,
 * 1: data gradient), one text line per launch ("tc2 ..." plane-folded tcgen05, "tc1 ..." per-tile
 * tcgen05, "fp32 ..." CUDA cores); returns the number of launches. */
int vg_conv_describe(const VgConvDesc* d, int kind, char* buf, size_t cap);

/* y = act(conv(x * in_scale[g,ci] + in_shift[g,ci]) + bias); zero padding is applied
 * AFTER the affine (it pads the normalised tensor).  in_scale/in_shift: (n/group_size, cin)
 * or NULL.  out_stats: (n/group_size, cout, 2) doubles, ACCUMULATED with sum(y), sum(y^2)
 * (caller zeroes) or NULL. */
int vg_conv_fwd(const VgConvDesc* d, const void* x, const float* w, const float* bias,
                const float* in_scale, const float* in_shift, void* y, int act,
                double* out_stats, void* stream);

/* dx = conv^T(dy) w.r.t. the (affine-folded) input.  dy must already be the gradient
 * w.r.t. the pre-activation.  Epilogue modes:
 *   mask_act != NULL                : dx *= (mask_act > 0)        (ReLU of the producer)
 *   bn_x != NULL                    : bn_sums (n/group_size, cin, 2) doubles accumulate
 *                                     sum(dx), sum(dx * xhat), xhat = bn_x*bn_istd - bn_mistd
 *                                     (bn_istd, bn_mistd: (groups, cin) = 1/std, mean/std)
 *   dx == NULL                      : nothing is stored (statistics only; bn1 of conv1). */
int vg_conv_dgrad(const VgConvDesc* d, const void* dy, const float* w, void* dx,
                  const void* mask_act, const void* bn_x, const float* bn_istd,
                  const float* bn_mistd, double* bn_sums, void* stream);

/* dw (PyTorch layout, ACCUMULATED — caller zeroes) and dbias (cout, accumulated) from
 * x (with the same affine fold as the forward) and dy (gradient w.r.t. pre-activation). */
int vg_conv_wgrad(const VgConvDesc* d, const void* x, const void* dy, const float* in_scale,
                  const float* in_shift, float* dw, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused backward of a BatchNorm -> ConvTranspose3d(k = 3, stride 1, one output channel) junction
 * (bnt5 -> convt5, vae_reg_GP.py:264; autograd of :215,218).  Instead of data gradient -> statistics ->
 * separate BatchNorm-backward pass over the largest tensor of the step, the statistics come from the
 * WEIGHT-gradient products, which do not depend on the data gradient:
 *   vg_box_sums            out (groups, 28) fp64 (caller zeroes): box sums of dy per tap + total
 *   vg_conv_wgrad_grouped  raw (groups, taps, cs, cu) fp32 (caller zeroes): per-group products of the
 *                          UN-normalised x with dy (x may be bf16: staged with asynchronous copies)
 *   vg_bn_fused_finalize   -> dw, dbias of the convolution, dgamma, dbeta of the BatchNorm (all
 *                          ACCUMULATED) and coef (groups, c, 3) for the apply
 *   vg_conv_dgrad_bn_apply dx = (bn_x > 0) * (A * conv^T(dy) + B * bn_x + C): the gradient w.r.t. the
 *                          pre-ReLU activation of the layer BEFORE the BatchNorm, written once (bf16
 *                          with VG_BF16_DX); dx_chan_sum (cin, accumulated, may be NULL) = that
 *                          layer's bias gradient.
 * ---------------------------------------------------------------------------------- */
int vg_box_sums(const float* dy, int n, int group_size, const int32_t* y_dims, long long y_img_stride,
                const int32_t* x_dims, const int32_t* pad, double* out, void* stream);
int vg_conv_wgrad_grouped(const VgConvDesc* d, const void* x, const void* dy, float* raw, void* stream);
int vg_bn_fused_finalize(const float* raw, const double* box, const float* w, const float* scale,
                         const float* shift, const float* istd, const float* mistd, int groups, int c,
                         double count, float* dw, float* dbias, float* dgamma, float* dbeta, float* coef,
                         void* stream);
int vg_conv_dgrad_bn_apply(const VgConvDesc* d, const void* dy, const float* w, void* dx, const void* bn_x,
                           const float* coef, float* dx_chan_sum, void* stream);

/* ------------------------------------------------------------------------------------
 * Batch-norm helpers (nn.BatchNorm3d(track_running_stats=False), vae_reg_GP.py:194-196,
 * 216-218; applied at :238,240,242,260,262,264).
 * ---------------------------------------------------------------------------------- */
/* stats (groups, c, 2) doubles accumulate sum(x), sum(x^2) over a channels-last tensor of
 * n images x `spatial` voxels x c channels. */
int vg_bn_stats(const float* x, int n, int group_size, long long spatial, int c, double* stats,
                void* stream);
/* From stats -> fold coefficients, per (group, channel):
 *   scale = gamma/std, shift = beta - mean*scale, istd = 1/std, mistd = mean/std,
 *   std = sqrt(var_biased + 1e-5), count = group_size * spatial. */
int vg_bn_finalize(const double* stats, const float* gamma, const float* beta, int groups, int c,
                   double count, float* scale, float* shift, float* istd, float* mistd,
                   void* stream);
/* dx = scale[g,c] * (dy - m1 - xhat*m2) [* (x > 0) if relu_mask], m1 = sums[g,c,0]/count,
 * m2 = sums[g,c,1]/count; dgamma[c] += sum_g sums[g,c,1]; dbeta[c] += sum_g sums[g,c,0]
 * (dgamma/dbeta may be NULL; they are accumulated by one thread block). In-place (dx == dy) ok.
 * bf16_mask: VG_BF16_X = x holds bf16, VG_BF16_DX = dx is written as bf16 (dy is always fp32, so
 * dx must then be a different buffer than dy); c must be 8 or 16 for bf16 storage.
 * dx_chan_sum (c floats, ACCUMULATED, may be NULL; bf16-storage variant only): sum of dx per channel =
 * the bias gradient of the layer whose ReLU output x is, so that its weight-gradient call can take
 * dbias == NULL and stage dx with plain asynchronous copies. */
int vg_bn_bwd_apply(const float* dy, const void* x, const double* sums, const float* scale,
                    const float* istd, const float* mistd, int n, int group_size,
                    long long spatial, int c, double count, int relu_mask, int bf16_mask, void* dx,
                    float* dgamma, float* dbeta, float* dx_chan_sum, void* stream);
/* (n, c, spatial) <-> (n, spatial, c) */
int vg_nchw_to_nhwc(const float* src, float* dst, int n, int c, long long spatial, void* stream);
int vg_nhwc_to_nchw(const float* src, float* dst, int n, int c, long long spatial, void* stream);

/* ------------------------------------------------------------------------------------
 * Linear layers (nn.Linear fc1..fc8, vae_reg_GP.py:197-210, used at :244-251, :255-258).
 * ---------------------------------------------------------------------------------- */
/* y (m,n) = act(x (m,k) @ w (n,k)^T + bias (n)) */
int vg_linear_fwd(const float* x, const float* w, const float* bias, float* y, int m, int n,
                  int k, int act, void* stream);
/* Given dy (m,n) w.r.t. the layer OUTPUT and, if relu_out != NULL, the saved output y
 * (gradient is masked by y > 0): dx (m,k) (may be NULL), dw (n,k) and db (n) ACCUMULATED. */
int vg_linear_bwd(const float* dy, const float* relu_out, const float* x, const float* w,
                  float* dx, float* dw, float* db, int m, int n, int k, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused chains of small fully-connected layers (encoder heads fc2 -> fc31/32/33 -> fc41/42/43,
 * vae_reg_GP.py:198-204,245-251; decoder stem fc5 -> fc6 -> fc7, :207-209,255-257): one launch
 * per chain and direction, a CTA owns rows_per_cta batch rows and walks the layer list with the
 * activations in shared memory.  buf[i]: (rows, width) row-major fp32 activations `act` and their
 * gradients `grad`; layer l: buf[out] = act(buf[in] @ w^T + b), w (n,k) row-major (nn.Linear), k <= 224.
 * Layers must be listed in a topological order; several layers may read the same buffer.
 *   vg_mlp_fwd: reads the VG_MLP_INPUT buffers, writes every layer output to its `act`.
 *   vg_mlp_bwd: reads every `act` (saved forward values) and the `grad` of VG_MLP_GRAD_IN buffers
 *     (dLoss/d post-activation output), writes the `grad` of VG_MLP_GRAD_OUT buffers and
 *     ACCUMULATES dw (n,k) / db (n) (fp32 RED; either may be NULL).
 * ---------------------------------------------------------------------------------- */
#define VG_MLP_MAX_LAYERS 8
#define VG_MLP_MAX_BUFS 10
#define VG_MLP_INPUT 1
#define VG_MLP_GRAD_IN 2
#define VG_MLP_GRAD_OUT 4
typedef struct VgMlpLayer {
  const float* w; const float* b; float* dw; float* db;
  int32_t n, k, in, out, act, pad_;
} VgMlpLayer;
typedef struct VgMlpBuf {
  float* act; float* grad;
  int32_t width, role;
} VgMlpBuf;
typedef struct VgMlp {
  int32_t nlayers, nbufs, rows, rows_per_cta;     /* rows_per_cta: 1, 2, 4 or 8 */
  VgMlpLayer layer[VG_MLP_MAX_LAYERS];
  VgMlpBuf buf[VG_MLP_MAX_BUFS];
} VgMlp;
int vg_mlp_fwd(const VgMlp* m, void* stream);
int vg_mlp_bwd(const VgMlp* m, void* stream);

/* ------------------------------------------------------------------------------------
 * Latent sample + KL (vae_reg_GP.py:321-325 LowRankMultivariateNormal(mu,u,d).rsample(),
 * :400 kl_divergence(latent_dist, z_prior); torch lowrank_multivariate_normal.py:214-223,
 * kl.py:342-372).  heads: (3, b, 32) = [mu | u | log d] (fc41/42/43 outputs, fc43 BEFORE exp).
 * zcat: (9, b, 41) rows [z | onehot(j)] for the 9 decoder passes (:326-330,339-343).
 * ---------------------------------------------------------------------------------- */
int vg_latent_fwd(const float* heads, const float* eps_w, const float* eps_d, int b, float* z,
                  float* klz, float* d_out, float* zcat, int* jitter_flag, void* stream);
/* dheads (3,b,32) from dzcat (9,b,41) (summed over the 9 passes) and dklz (b). */
int vg_latent_bwd(const float* heads, const float* eps_w, const float* eps_d, const float* d_used,
                  const float* dzcat, const float* dklz, int b, float* dheads, void* stream);

/* ------------------------------------------------------------------------------------
 * Gains: linear + sparse-GP posterior, Cholesky sample, HRF FIR, KLs, for all K covariates
 * in one launch (one warp per covariate GP).  Replaces vae_reg_GP.py:345-378, :266-281,
 * :283-305 and gp.py:41-65,67-110,113-136.  Internals are fp64.
 * ---------------------------------------------------------------------------------- */
typedef struct VgGainParams {       /* device pointers, one entry per covariate (K = 8) */
  const float* sa[8];               /* (1,1) */
  const float* logstd[8];           /* (1,1) */
  const float* qu_m[8];             /* (1,m)   NULL when !has_gp */
  const float* qu_S[8];             /* (m,m) */
  const float* logkvar[8];          /* () */
  const float* logls[8];            /* () */
  const float* xu[8];               /* (m) inducing locations (constant) */
  int32_t has_gp[8];                /* vae_reg_GP.py:352 -> {0,1,1,1,1,1,1,0} */
  int32_t hrf[8];                   /* vae_reg_GP.py:377 -> {neural_covariates,0,...} */
} VgGainParams;
typedef struct VgGainGrads {
  float* sa[8]; float* logstd[8]; float* qu_m[8]; float* qu_S[8]; float* logkvar[8]; float* logls[8];
} VgGainGrads;

size_t vg_gain_workspace_bytes(int b, int m);
/* covariates (b,8) fp32; eps (8,b); taps: 15 HRF taps (fp64, device).  Outputs: g (8,b) fp32
 * (post-HRF), kl_terms (8,2) fp64 = [linear-weight KL, GP KL] per covariate, beta_mean (8,b) and
 * beta_var (8,b) = diag(beta_cov) for the TensorBoard hook (may be NULL), status (8) int:
 * 0 ok, else 1-based index of the failed pivot (covariance / qu_S not PD). */
int vg_gain_fwd(const VgGainParams* p, const float* covariates, const float* eps,
                const double* taps, int b, int m, float* g, double* kl_terms, float* beta_mean,
                float* beta_var, int* status, void* workspace, size_t workspace_bytes,
                void* stream);
/* dg (8,b): dLoss/dg; kl_scale: gradient weight of the KL terms (gp_kl_scale).  Grads are
 * ACCUMULATED into the VgGainGrads pointers (fp32). */
int vg_gain_bwd(const VgGainParams* p, const VgGainGrads* grads, const float* covariates,
                const float* eps, const double* taps, const float* dg, double kl_scale, int b,
                int m, void* workspace, size_t workspace_bytes, void* stream);
/* GP posterior at arbitrary query points (plot_GPs, vae_reg_GP.py:655-666; gp.GP.evaluate_posterior,
 * gp.py:67-110).  k_var, ls: DEVICE scalars (the reference passes 0-d tensors).  f_bar (nq),
 * var (nq) = diag(Sigma) (may be NULL); sigma (nq,nq) optional — NULL skips the O(nq^2) matrix —
 * and then a_ws, an (nq, m) fp64 scratch, is required. */
int vg_gp_posterior(const float* xu, int m, const float* k_var, const float* ls, const float* qu_m,
                    const float* qu_S, const float* xq, int nq, float* f_bar, float* var,
                    float* sigma, double* a_ws, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused reconstruction + likelihood + GLM regulariser (vae_reg_GP.py:380, :388-390, :401-405).
 * maps (9,b,VP): decoder outputs (base, 8 covariate maps), every row padded to
 * VP = round_up(V,4) floats so rows are 16-byte aligned (V = 70315 is odd); g (8,b); x (b,V)
 * dense; eps (VP) fp32 copy of the fp64 epsilon parameter (.float() at :402); glm (8,VP) fp32
 * transposed GLM maps.  One HBM pass over warp items of 4 rows x 32 voxels dealt
 * in equal spans to a persistent grid; partial sums are reduced deterministically in a second
 * tiny kernel.  dpre has the (9,b,VP) layout of maps; deps is (VP).
 * ---------------------------------------------------------------------------------- */
size_t vg_recon_workspace_bytes(int b, long long v);
/* Launch shape of both passes (warps per CTA x CTAs per SM x cp.async ring stages):
 * 0: 4 x 5 x 2, 1: 4 x 3 x 3, 2: 4 x 4 x 2, 3 (default, also for any other value): 8 x 2 x 3. */
void vg_recon_tune(int variant);
/* Work decomposition a launch with this (b, v, variant) uses (host-side, no device needed):
 * out[0..5] = row groups of 4, warp items (4 rows x 32 voxels) per row group, items, items per CTA, CTAs,
 * warps per CTA. */
int vg_recon_plan(int b, long long v, int variant, int* out);
/* logp (b), norms (8,b) = ||g_i D_i[b] - G_i||_2.  Optional cons (8,b,V) and x_rec (b,V)
 * (R5 side outputs, NULL in training). */
int vg_recon_loss_fwd(const float* maps, const float* g, const float* x, const float* eps,
                      const float* glm, int b, long long v, float* logp, float* norms,
                      float* cons, float* x_rec, void* workspace, size_t workspace_bytes,
                      void* stream);
/* Backward of  tot = -mean_b(logp) + lam * b * sum_{i,b} norms  w.r.t. the decoder
 * PRE-sigmoid activations (dpre = dD * D * (1-D)), g and epsilon:
 *   dpre (9,b,V), dg (8,b), deps (V, fp32; caller adds into the fp64 grad). */
int vg_recon_loss_bwd(const float* maps, const float* g, const float* x, const float* eps,
                      const float* glm, const float* norms, int b, long long v, float lam,
                      float* dpre, float* dg, float* deps, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused Adam over a flat parameter buffer (torch.optim.Adam defaults, vae_reg_GP.py:179,429).
 * fp32 segment [0,n32) and fp64 segment (epsilon).  step_count: device int64 incremented here.
 * grad_scale multiplies the gradient (1/world_size after the NCCL sum).
 * skip_flags (n_flags device ints, may be NULL): when any of them is non-zero the whole update
 * (parameters, moments and step count) is skipped on the device.  The step passes VgStepIO.status,
 * so a minibatch that met a non-positive-definite covariance leaves the parameters untouched, as
 * the reference does by raising inside forward (vae_reg_GP.py:368, gp.py:51) before backward()/step().
 * ---------------------------------------------------------------------------------- */
int vg_adam_step(float* p32, const float* g32, float* m32, float* v32, long long n32, double* p64,
                 const double* g64, double* m64, double* v64, long long n64, double lr, double beta1,
                 double beta2, double eps, double grad_scale, long long* step_count,
                 const int32_t* skip_flags, int n_flags, void* stream);

/* ------------------------------------------------------------------------------------
 * Whole training step (vae_reg_GP.py:307-413 forward, :427-428 backward) chained on one
 * stream with no host synchronisation.  See vaegam/step.py for the buffer tables.
 * ---------------------------------------------------------------------------------- */
#define VG_NUM_PARAMS 97
typedef struct VgStepConfig {
  int32_t b;                 /* minibatch */
  int32_t m;                 /* inducing points */
  int32_t neural_covariates; /* HRF on covariate 1 */
  int32_t want_maps;         /* also emit cons / x_rec (return_latent_rec) */
  float gp_kl_scale;
  float glm_reg_scale;
  int32_t arith;             /* VG_ARITH_* for every convolution of this call (0 = process default) */
  int32_t reserved[3];       /* zero */
} VgStepConfig;

size_t vg_step_workspace_bytes(const VgStepConfig* cfg);
/* params: VG_NUM_PARAMS device pointers in the order of vaegam/step.py:PARAM_ORDER (the
 * reference's named_parameters() order; epsilon is fp64).  consts: xu (6 pointers, (m)),
 * glm_t (8,V) fp32, taps (15) fp64.  inputs: x (b,V), covariates (b,8), eps_w (b,1),
 * eps_d (b,32), eps_g (8,b).  outputs: out_scalars (8) fp64 = [tot, neg_elbo, gp_kl, glm_reg,
 * mean logp, mean klz, 0, 0]; z (b,32); maps (9,b,VP), VP = round_up(V,4); g (8,b); optional
 * cons (8,b,V) / x_rec (b,V) dense. */
typedef struct VgStepIO {
  const void* params[VG_NUM_PARAMS];
  void* grads[VG_NUM_PARAMS];          /* backward only; epsilon grad is fp64 */
  const float* xu[6];
  const float* glm_t;
  const double* taps;
  const float* x;
  const float* covariates;
  const float* eps_w;
  const float* eps_d;
  const float* eps_g;
  double* out_scalars;
  float* z;
  float* maps;
  float* g;
  float* cons;
  float* x_rec;
  float* beta_mean;
  float* beta_var;
  int32_t* status;                     /* (16) ints: [0..7] gain status, [8] latent jitter flag */
} VgStepIO;

int vg_step_fwd(const VgStepConfig* cfg, const VgStepIO* io, void* workspace,
                size_t workspace_bytes, void* stream);
/* Gradients of out_scalars[0] are ACCUMULATED into io->grads (caller zeroes).  Must follow
 * vg_step_fwd with the same cfg/io/workspace.  Weight gradients and the gain backward run on a
 * helper stream (fork / join through events; still capturable). */
int vg_step_bwd(const VgStepConfig* cfg, const VgStepIO* io, void* workspace,
                size_t workspace_bytes, void* stream);
/* The same backward in three phases, for data-parallel training that overlaps the gradient
 * all-reduce with the rest of the backward (the reference has no collective; its step is
 * vae_reg_GP.py:427-429).  Phases must be called in order 0, 1, 2 on `stream`:
 *   0: objective + the whole decoder      -> grads of epsilon, fc5..fc8, convt1..5, bnt1/3/5 complete
 *   1: latent + encoder fully connected   -> grads of fc1..fc43 complete
 *   2: encoder convolutions + gain stage  -> grads of conv1..5, bn1/3/5 and all gain parameters complete
 * If ready_stream != NULL it is made to wait (events only) for every kernel that writes this
 * phase's gradients, helper-stream kernels included, WITHOUT joining them into `stream`: the
 * caller enqueues the phase's all-reduce on ready_stream and joins ready_stream into `stream`
 * before the optimizer step.  vg_step_bwd == phases 0, 1, 2 with ready_stream NULL. */
#define VG_BWD_PHASES 3
int vg_step_bwd_phase(const VgStepConfig* cfg, const VgStepIO* io, void* workspace,
                      size_t workspace_bytes, int phase, void* ready_stream, void* stream);
/* Encoder only (VAE.encode, vae_reg_GP.py:236-252): heads (3,b,32) = [mu | u | log d]. */
int vg_encode_fwd(const VgStepConfig* cfg, const VgStepIO* io, float* heads, void* workspace,
                  size_t workspace_bytes, void* stream);
/* Decoder only (VAE.decode, :254-264) for n rows of zcat (n,41) treated as ONE BatchNorm batch;
 * arith: VG_ARITH_* (MIXED = BF16 here: the decoder is never the fp32 part). */
size_t vg_decode_workspace_bytes(int n);
int vg_decode_fwd(const VgStepIO* io, const float* zcat, int n, int arith, float* out, void* workspace,
                  size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEGAM_H_ */
