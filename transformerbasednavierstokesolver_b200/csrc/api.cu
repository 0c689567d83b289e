// Library-level plumbing of libtbns: thread-local error string, version, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace tbns {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tbns

extern "C" const char* tbns_last_error(void) { return tbns::g_err; }
extern "C" int tbns_version(void) { return 200; }
extern "C" int tbns_sm_count(void) { return tbns::sm_count(); }
extern "C" int tbns_device_ok(void) {
  int dev = 0, n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    tbns::set_error("no CUDA device visible");
    return 0;
  }
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    tbns::set_error("cannot query CUDA device");
    return 0;
  }
  if (p.major != 10) {
    tbns::set_error("libtbns is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
    return 0;
  }
  return 1;
}
