// LayerNorm forward/backward (nn.LayerNorm(hidden_dim), eps 1e-5) and small deterministic reductions.
// Reference: model/Transolver_Structured_Mesh_2D.py:59,63,66 (ln_1/ln_2/ln_3) and forward :70-73.
// HBM-bound streaming kernels: one warp per token row, float4 accesses, fixed-order reductions.
#include "common.cuh"

namespace tbns {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_CTAS = 296;  // 2 x 148 SMs

__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float* __restrict__ y,
                                                                    float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                                    int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= rows) return;
  const float* xr = x + (long long)row * C;
  float* yr = y + (long long)row * C;
  const bool vec = (C & 3) == 0;
  float s = 0.f;
  if (vec) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int c = lane; c < C; c += 32) s += xr[c];
  }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
  if (vec) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float a = v.x - mu, b = v.y - mu, cc = v.z - mu, dd = v.w - mu;
      q += (a * a + b * b) + (cc * cc + dd * dd);
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float a = xr[c] - mu;
      q += a * a;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  if (vec) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 b = *reinterpret_cast<const float4*>(beta + c);
      float4 o;
      o.x = (v.x - mu) * rs * g.x + b.x;
      o.y = (v.y - mu) * rs * g.y + b.y;
      o.z = (v.z - mu) * rs * g.z + b.z;
      o.w = (v.w - mu) * rs * g.w + b.w;
      *reinterpret_cast<float4*>(yr + c) = o;
    }
  } else {
    for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mu) * rs * gamma[c] + beta[c];
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx = (g - mean(g) - xhat*mean(g*xhat)) * rstd (+ dres), g = dy*gamma.  Per-CTA partial dgamma/dbeta -> part[cta][2][C].
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma, const float* __restrict__ dres,
                                                                    float* __restrict__ dx, float* __restrict__ part, int rows, int C) {
  extern __shared__ float sm[];  // [LN_WARPS][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dg = sm + (long long)warp * 2 * C;
  float* db = dg + C;
  for (int c = lane; c < C; c += 32) {
    dg[c] = 0.f;
    db[c] = 0.f;
  }
  __syncwarp();
  const float invC = 1.f / (float)C;
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const float* xr = x + (long long)row * C;
    const float* dyr = dy + (long long)row * C;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = dyr[c] * gamma[c];
      const float xh = (xr[c] - mu) * rs;
      s1 += g;
      s2 += g * xh;
    }
    const float c1 = warp_sum(s1) * invC, c2 = warp_sum(s2) * invC;
    for (int c = lane; c < C; c += 32) {
      const float d = dyr[c];
      const float xh = (xr[c] - mu) * rs;
      float o = (d * gamma[c] - c1 - xh * c2) * rs;
      if (dres) o += dres[(long long)row * C + c];
      dx[(long long)row * C + c] = o;
      dg[c] += d * xh;  // lane-private columns: no race
      db[c] += d;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += sm[(long long)w * 2 * C + c];
    part[(long long)blockIdx.x * 2 * C + c] = s;
  }
}

__global__ void reduce_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, long long cols, long long ld) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < cols; j += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < rows; ++i) s += in[(long long)i * ld + j];
    out[j] = s;
  }
}

constexpr int COLSUM_ROWS = 64;  // row groups
__global__ void colsum_partial_kernel(const float* __restrict__ in, long long ld, float* __restrict__ ws, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) s += in[(long long)r * ld + c];
  ws[(long long)blockIdx.y * cols + c] = s;
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                  int rows, int C, float eps, void* stream) {
  TBNS_REQUIRE(x && gamma && beta && y && mean && rstd, "tbns_layernorm_fwd: null pointer");
  TBNS_REQUIRE(rows >= 0 && C > 0, "tbns_layernorm_fwd: bad dims");
  if (rows == 0) return TBNS_OK;
  layernorm_fwd_kernel<<<cdiv(rows, LN_WARPS), LN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, rows, C, eps);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" size_t tbns_layernorm_bwd_ws_floats(int C) { return (size_t)LN_MAX_CTAS * 2 * (size_t)C; }

extern "C" int tbns_reduce_rows(const float* in, float* out, int rows, long long cols, void* stream) {
  TBNS_REQUIRE(in && out && rows >= 0 && cols >= 0, "tbns_reduce_rows: bad args");
  if (cols == 0) return TBNS_OK;
  int blocks = (int)((cols + 127) / 128);
  if (blocks > 148 * 16) blocks = 148 * 16;
  reduce_rows_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(in, out, rows, cols, cols);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                  const float* dres, float* dx, float* dgamma, float* dbeta, float* ws, int rows, int C,
                                  void* stream) {
  TBNS_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws, "tbns_layernorm_bwd: null pointer");
  TBNS_REQUIRE(rows > 0 && C > 0, "tbns_layernorm_bwd: bad dims");
  const size_t smem = (size_t)LN_WARPS * 2 * C * sizeof(float);
  TBNS_REQUIRE(smem <= 200 * 1024, "tbns_layernorm_bwd: C=%d too large", C);
  static bool attr_set = false;
  if (!attr_set) {
    TBNS_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int ctas = cdiv(rows, LN_WARPS * 4);
  if (ctas > LN_MAX_CTAS) ctas = LN_MAX_CTAS;
  if (ctas < 1) ctas = 1;
  cudaStream_t st = (cudaStream_t)stream;
  layernorm_bwd_kernel<<<ctas, LN_WARPS * 32, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, ws, rows, C);
  TBNS_LAUNCH_CHECK();
  // ws rows are [dgamma | dbeta], width 2C
  reduce_rows_kernel<<<cdiv(C, 128), 128, 0, st>>>(ws, dgamma, ctas, C, 2LL * C);
  TBNS_LAUNCH_CHECK();
  reduce_rows_kernel<<<cdiv(C, 128), 128, 0, st>>>(ws + C, dbeta, ctas, C, 2LL * C);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" size_t tbns_colsum_ws_floats(long long cols) { return (size_t)COLSUM_ROWS * (size_t)cols; }

extern "C" int tbns_colsum(const float* in, long long ld, float* out, float* ws, int rows, int cols, void* stream) {
  TBNS_REQUIRE(in && out && ws && rows > 0 && cols > 0, "tbns_colsum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int groups = rows < COLSUM_ROWS ? rows : COLSUM_ROWS;
  dim3 grid(cdiv(cols, 128), groups);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(in, ld, ws, rows, cols);
  TBNS_LAUNCH_CHECK();
  return tbns_reduce_rows(ws, out, groups, cols, stream);
}
