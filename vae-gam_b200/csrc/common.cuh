// Shared helpers for libvaegam_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vaegam.h"

namespace vg {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional per-operation timing with CUDA events on the launching stream (vg_profile_*).
// A scope costs nothing unless profiling was enabled.
bool profiling();
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(const char* name, cudaStream_t s) : st(s), on(profiling()) { if (on) prof_begin(name, st); }
  ~ProfScope() { if (on) prof_end(st); }
};
#define VG_PROF(name, st) vg::ProfScope prof_scope_##__LINE__(name, st)

#define VG_CHECK_ARG(cond, msg)                                        \
  do {                                                                 \
    if (!(cond)) {                                                     \
      vg::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
      return VG_EINVAL;                                                \
    }                                                                  \
  } while (0)

#define VG_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      vg::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VG_ECUDA;                                                             \
    }                                                                              \
  } while (0)

// after a kernel launch: catches launch-configuration errors without synchronising
#define VG_LAUNCH_CHECK()                 \
  do {                                    \
    vg::count_launch();                   \
    VG_CUDA(cudaPeekAtLastError());       \
  } while (0)

#define VG_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != VG_OK) return rc__; \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load: read once, do not keep in L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace vg
