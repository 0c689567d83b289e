// Shared helpers for libtbns (sm_100a).  Error plumbing, launch checks, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tbns.h"

namespace tbns {

void set_error(const char* fmt, ...);
int splitk_reduce(const tbns_gemm_desc& d, cudaStream_t st);  // gemm_simt.cu

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

// opt in to large dynamic shared memory once per kernel (not a stream operation: keep it out of CUDA-graph capture)
#define TBNS_SMEM_OPT_IN(kernel, bytes)                                                                      \
  do {                                                                                                       \
    static int _cur = -1;                                                                                    \
    if (_cur < (int)(bytes)) {                                                                               \
      TBNS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));    \
      _cur = (int)(bytes);                                                                                   \
    }                                                                                                        \
  } while (0)

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
