#include "common.h"

#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace vb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VACNIC_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int check_last(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(VACNIC_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return VACNIC_OK;
}

}  // namespace vb

extern "C" {
const char* vacnic_last_error(void) { return vb::g_err; }
int vacnic_version(void) { return 100; }
int64_t vacnic_launch_count(void) { return vb::g_launches.load(std::memory_order_relaxed); }
}
