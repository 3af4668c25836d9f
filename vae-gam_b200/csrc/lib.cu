// Library identity, thread-local error text, launch accounting.
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include <vector>

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

struct ProfRec { const char* name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
bool profiling() { return g_prof_on; }
static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(const char* name, cudaStream_t st) {
  ProfRec r{name, get_event(), get_event()};
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st);
}
}  // namespace vg

extern "C" int vg_profile_enable(int on) {
  for (auto& r : vg::g_prof) { vg::g_pool.push_back(r.a); vg::g_pool.push_back(r.b); }
  vg::g_prof.clear();
  vg::g_prof_on = on != 0;
  return VG_OK;
}
// Writes "name<TAB>milliseconds\n" per recorded operation into buf (synchronises on the events).
extern "C" long long vg_profile_collect(char* buf, size_t cap) {
  size_t off = 0;
  long long n = 0;
  for (auto& r : vg::g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
      (void)cudaGetLastError();
      continue;
    }
    if (buf && off + 96 < cap) off += (size_t)snprintf(buf + off, cap - off, "%s\t%.6f\n", r.name, ms);
    ++n;
  }
  if (buf && cap) buf[off < cap ? off : cap - 1] = 0;
  return n;
}

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
