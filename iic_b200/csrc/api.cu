// Library-level plumbing: error string, launch counter, device properties.
#include <stdarg.h>

#include "common.cuh"

namespace iic {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { ++g_launches; }

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace iic

extern "C" int iic_abi_version(void) { return 1; }
extern "C" const char* iic_last_error(void) { return iic::g_err; }
extern "C" long long iic_launch_count(int reset) {
  long long v = iic::g_launches;
  if (reset) iic::g_launches = 0;
  return v;
}
