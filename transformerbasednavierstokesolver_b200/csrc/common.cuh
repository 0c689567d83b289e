// Shared helpers for libtbns (sm_100a).  Error plumbing, launch checks, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tbns.h"

namespace tbns {

void set_error(const char* fmt, ...);
int splitk_reduce(const tbns_gemm_desc& d, cudaStream_t st);  // gemm_simt.cu
int gemm_x3_launch(const tbns_gemm_desc& d, int vecA, int vecB, int vecC, dim3 grid, cudaStream_t st);  // gemm_x3.cu

#define TBNS_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      tbns::set_error(__VA_ARGS__);      \
      return TBNS_ERR_INVALID;           \
    }                                    \
  } while (0)

#define TBNS_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      tbns::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TBNS_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define TBNS_LAUNCH_CHECK() TBNS_CUDA(cudaGetLastError())

// opt in to large dynamic shared memory once per kernel AND device (not a stream operation: keep it out of CUDA-graph
// capture).  The attribute is per device, so the cache is indexed by the current device ordinal.
#define TBNS_SMEM_OPT_IN(kernel, bytes)                                                                      \
  do {                                                                                                       \
    static int _cur[TBNS_MAX_DEVICES];                                                                       \
    int _dev = 0;                                                                                            \
    TBNS_CUDA(cudaGetDevice(&_dev));                                                                         \
    if (_dev < 0 || _dev >= TBNS_MAX_DEVICES) {                                                              \
      tbns::set_error("device ordinal %d out of range", _dev);                                               \
      return TBNS_ERR_INVALID;                                                                               \
    }                                                                                                        \
    if (_cur[_dev] < (int)(bytes)) {                                                                         \
      TBNS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));    \
      _cur[_dev] = (int)(bytes);                                                                             \
    }                                                                                                        \
  } while (0)

constexpr int TBNS_MAX_DEVICES = 64;

// multiProcessorCount of the current device (cached per device)
static inline int sm_count() {
  static int cache[TBNS_MAX_DEVICES];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TBNS_MAX_DEVICES) return 148;
  if (!cache[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The training step is a chain of several hundred short kernels.  Launched with the
// attribute below, a kernel's CTAs may be scheduled while the CTAs of its predecessor in the stream are still exiting, so
// launch latency and CTA rasterisation overlap the predecessor's tail instead of following its completion signal.
// Correctness rule: EVERY thread of EVERY kernel launched through launch_pdl() executes pdl_sync() before it reads or writes
// global memory that an earlier kernel may touch - griddepcontrol.wait returns only when all prerequisite grids have completed
// and their writes are visible, and because every kernel in the chain waits, completion is transitive.  Launched without the
// attribute (TBNS_PDL=0, or a predecessor that is not a kernel) the instruction is a no-op.  Under stream capture the edge
// becomes a programmatic graph edge.
// Measured (cfg-1 step, one B200): 9.88 -> 9.77 ms; unrolled look_ahead 10: 25.7 -> 25.2 ms.  The explicit early trigger
// (griddepcontrol.launch_dependents at kernel entry) is deliberately NOT used: it was 3 % SLOWER (9.87 vs 9.55 ms on the same
// box) - CTAs parked at the wait take the SM slots that the side stream's weight-gradient kernels otherwise fill.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("TBNS_PDL"); return !e || e[0] != '0'; }();
  return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * expf(-0.5f * x * x) * 0.39894228040143267794f;
}

}  // namespace tbns
