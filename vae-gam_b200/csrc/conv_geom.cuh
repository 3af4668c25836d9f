// Geometry / argument structs shared by the CUDA-core gather (conv.cu) and the tcgen05
// implicit-GEMM gather (tc_conv.cu).
#pragma once
#include "common.cuh"

namespace vg {

constexpr int kMaxTaps = 48;

struct Tap {
  int8_t dd, dh, dw, pad_;
  int32_t widx;
};

struct Geom {
  int N, group_size;
  int inD, inH, inW;
  int outD, outH, outW;
  int qD, qH, qW;
  int sin, sout;
  int rD, rH, rW;
  int ntaps, check;
  int wst_t, wst_ci, wst_co;
  long long in_img, out_img;   // floats between images of the tensor read / written
  Tap taps[kMaxTaps];
};

struct GatherArgs {
  // storage of the three big tensors of a gather: fp32 by default, bf16 (same layout, 2 bytes per element) when
  // the flag is set — tensor-core kernels only (VgConvDesc.bf16_mask)
  int in_bf16, out_bf16, aux_bf16;
  const float* in;
  const float* w;
  const float* bias;      // (COUT) or null
  const float* in_scale;  // (groups, CIN) or null
  const float* in_shift;
  float* out;             // null: statistics only
  int act;
  double* stats;          // (groups, COUT, 2): sum y, sum y^2   (forward)
  const float* aux;       // saved tensor at the output positions (backward)
  int aux_mode;           // 0 none, 1 relu mask, 2 batch-norm backward sums, 3 fused batch-norm backward apply + relu mask
  const float* aux_coef;  // mode 3: (groups, COUT, 3) = A, B, C of  out = (aux > 0) * (A y + B aux + C)
  float* chan_sum;        // mode 3: (COUT) accumulated sum of the output per channel (bias gradient of the producer layer) or null
  const float* aux_istd;  // (groups, COUT)
  const float* aux_mistd;
  double* aux_sums;       // (groups, COUT, 2): sum dy, sum dy*xhat
};


// tc_conv.cu
bool tc_supported(int cin, int cout, const Geom& g);
int launch_tc_gather(int cin, int cout, const Geom& g, const GatherArgs& a, cudaStream_t st);
// tc2_conv.cu (plane-folded tcgen05 kernel: unit-stride gathers with 1 or 8 input channels)
// gs[0..ng): gathers of one layer pass sharing their input (ng > 1: the output-parity phases of a stride-2 layer)
bool tc2_supported(int cin, int cout, const Geom* gs, int ng);
int launch_tc2_gather(int cin, int cout, const Geom* gs, int ng, const GatherArgs& a, cudaStream_t st);
int tc2_describe(int cin, int cout, const Geom* gs, int ng, char* buf, size_t cap, bool in_bf16 = false);
int conv_mode();   // process default: 0 fp32 CUDA cores ("check mode"), 1 bf16 tcgen05 where the geometry allows, 2 mixed
int resolve_arith(int arith);   // VG_ARITH_* -> 0 / 1 / 2 (VG_ARITH_DEFAULT -> conv_mode())

}  // namespace vg
