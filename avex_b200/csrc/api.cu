// api.cu -- error reporting, launch counter and device queries shared by all libavexk entry points.
#include <stdarg.h>

#include <atomic>
#include <vector>

#include "common.cuh"

namespace avexk {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel CUDA-event profiler (bench.py roofline leg; off on the hot path) -----------------------
namespace {
struct ProfRec { int kid; double work; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace
bool prof_enabled() { return g_prof_on; }
void prof_begin(cudaStream_t st, int kid, double work) {
  if (!g_prof_on) return;
  ProfRec r{kid, work, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, st);
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}
}  // namespace avexk

extern "C" const char* avexk_last_error(void) { return avexk::g_err; }
extern "C" int avexk_version(void) { return 200; }
#ifndef AVEXK_BUILD_ID_STR
#define AVEXK_BUILD_ID_STR "unset"
#endif
// sha256 prefix of csrc/*.cu{,h} + include/avexk.h at build time (avex_b200/build.py); the marker makes it greppable in the file.
static const char g_build_id[] = "AVEXK_BUILD_ID=" AVEXK_BUILD_ID_STR;
extern "C" const char* avexk_build_id(void) { return g_build_id + 15; }
extern "C" long long avexk_launch_count(void) { return avexk::g_launches.load(std::memory_order_relaxed); }

extern "C" void avexk_profile_enable(int on) {
  for (auto& r : avexk::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  avexk::g_prof.clear();
  avexk::g_prof_on = on != 0;
}
extern "C" int avexk_profile_read(int kid, long long* launches, double* total_ms, double* total_work) {
  long long n = 0;
  double ms = 0.0, work = 0.0;
  for (auto& r : avexk::g_prof) {
    if (r.kid != kid) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) { avexk::set_error("avexk_profile_read: event sync failed"); return AVEXK_ECUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms += t; work += r.work; ++n;
  }
  if (launches) *launches = n;
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  return AVEXK_OK;
}
