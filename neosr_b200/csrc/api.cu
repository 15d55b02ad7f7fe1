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
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}
