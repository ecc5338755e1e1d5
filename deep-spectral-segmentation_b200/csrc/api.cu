// Library-wide plumbing: thread-local error message, version, device queries.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace dss {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached[dev] = n;
  return n;
}

}  // namespace dss

extern "C" const char* dss_last_error(void) { return dss::last_error(); }
extern "C" int dss_version(void) { return 1; }
extern "C" int dss_device_sm_count(void) {
  const int n = dss::device_sm_count();
  if (n < 0) dss::set_error("no CUDA device available");
  return n;
}
