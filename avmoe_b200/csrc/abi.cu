// Library-level entry points of the C ABI: version, build info, error string, device probes.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace avmoe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace avmoe

extern "C" int avmoe_abi_version(void) { return 1; }
extern "C" const char* avmoe_build_info(void) {
  return "avmoe_b200 sm_100a (tcgen05/TMA) built " __DATE__ " " __TIME__;
}
extern "C" const char* avmoe_last_error(void) { return avmoe::g_err; }
