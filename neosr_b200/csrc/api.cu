// Library-wide C-ABI plumbing: error string, version, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace nsr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace nsr

extern "C" const char* nsr_last_error(void) { return nsr::g_err; }
extern "C" int nsr_version(void) { return 100; }
extern "C" int nsr_device_supports_tcgen05(void) {
  // cached per device: cudaGetDeviceProperties costs milliseconds and this sits on the dispatch path
  static int cache[64] = {0};  // 0 = unknown, 1 = no, 2 = yes
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev] == 2;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) cache[dev] = major == 10 ? 2 : 1;
  return major == 10 ? 1 : 0;
}
