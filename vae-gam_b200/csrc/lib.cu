// Library identity, thread-local error text, launch accounting.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vg {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace vg

extern "C" int vg_version(void) { return 100; }
extern "C" const char* vg_last_error(void) { return vg::g_err; }
extern "C" long long vg_launch_count(void) { return vg::g_launches.load(std::memory_order_relaxed); }
extern "C" int vg_sm_count(void) {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    (void)cudaGetLastError();
    return 148;  // B200; only reached without a visible device (build container)
  }
  cached = sms;
  return sms;
}
