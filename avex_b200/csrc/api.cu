// api.cu -- error reporting, launch counter and device queries shared by all libavexk entry points.
#include <stdarg.h>

#include <atomic>

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
extern "C" int avexk_version(void) { return 100; }
extern "C" long long avexk_launch_count(void) { return avexk::g_launches.load(std::memory_order_relaxed); }
