// Library-level plumbing of libtbns: thread-local error string, version, device probe.
#include <stdarg.h>

#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace tbns {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int encode_tmap(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box, CUtensorMapSwizzle swz) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey k;
  memset(&k, 0, sizeof(k));
  int dev = 0;
  cudaGetDevice(&dev);
  k.v[0] = reinterpret_cast<uint64_t>(ptr);
  k.v[1] = ((uint64_t)dt << 32) | ((uint64_t)swz << 16) | ((uint64_t)rank << 8) | (uint64_t)(dev & 0xff);
  for (int i = 0; i < rank; ++i) {
    k.v[2 + i] = dims[i];
    k.v[10 + i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) k.v[6 + i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(k);
    if (it != cache.end()) {
      *m = it->second;
      return TBNS_OK;
    }
  }
  const int rc = encode_tmap_uncached(m, dt, ptr, rank, dims, strides_bytes, box, swz);
  if (rc == TBNS_OK) {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(k, *m);
  }
  return rc;
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
