// Library-level entry points: ABI version, error string, device info.
#include "common.h"
#include <string.h>

namespace rt {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
double g_small_flops = 0.0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rt

extern "C" int rt_abi_version(void) { return RT_ABI_VERSION; }
extern "C" const char* rt_last_error(void) { return rt::g_err; }
extern "C" int rt_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  RT_CHECK_CUDA(cudaGetDevice(&dev));
  RT_CHECK_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  RT_CHECK_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  RT_CHECK_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}
extern "C" unsigned long long rt_launch_count(void) { return rt::g_launches; }
extern "C" double rt_small_flop_count(void) { return rt::g_small_flops; }
