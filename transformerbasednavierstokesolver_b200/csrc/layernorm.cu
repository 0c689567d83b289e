// LayerNorm forward/backward (nn.LayerNorm(hidden_dim), eps 1e-5) and small deterministic reductions.
// Reference: model/Transolver_Structured_Mesh_2D.py:59,63,66 (ln_1/ln_2/ln_3) and forward :70-73.
// HBM-bound streaming kernels: one warp per token row, float4 accesses, fixed-order reductions.
#include "common.cuh"

namespace tbns {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_CTAS = 296;  // 2 x 148 SMs
constexpr int LN_BWD_CTAS = 444;  // 3 x 148 SMs: one full wave at the register-limited occupancy of the backward kernel
constexpr int COLSUM_ROWS = 592;  // max row chunks of the column-sum kernel (4 x 148)

__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float* __restrict__ y,
                                                                    __nv_bfloat16* __restrict__ y16, float* __restrict__ mean,
                                                                    float* __restrict__ rstd, int rows, int C, float eps) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= rows) return;
  const float* xr = x + (long long)row * C;
  float* yr = y ? y + (long long)row * C : nullptr;
  __nv_bfloat16* yr16 = y16 ? y16 + (long long)row * C : nullptr;
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
      if (yr) *reinterpret_cast<float4*>(yr + c) = o;
      if (yr16) {
        __nv_bfloat162 h2[2] = {__floats2bfloat162_rn(o.x, o.y), __floats2bfloat162_rn(o.z, o.w)};
        *reinterpret_cast<uint2*>(yr16 + c) = *reinterpret_cast<uint2*>(h2);
      }
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float o = (xr[c] - mu) * rs * gamma[c] + beta[c];
      if (yr) yr[c] = o;
      if (yr16) yr16[c] = __float2bfloat16_rn(o);
    }
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// Register-resident forward for C = 128*NV: each warp keeps a whole row in registers (one HBM read, no re-reads through L1)
// and works on TWO rows per iteration so that 2*NV independent 16-byte loads are in flight per lane.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, float* __restrict__ y,
                                                                        __nv_bfloat16* __restrict__ y16, float* __restrict__ mean,
                                                                        float* __restrict__ rstd, int rows, float eps) {
  pdl_sync();
  constexpr int C = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 g4[NV], b4[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    g4[k] = *reinterpret_cast<const float4*>(gamma + lane * 4 + k * 128);
    b4[k] = *reinterpret_cast<const float4*>(beta + lane * 4 + k * 128);
  }
  const float invC = 1.f / (float)C;
  const int stride = gridDim.x * LN_WARPS * 2;
  for (int row0 = (blockIdx.x * LN_WARPS + warp) * 2; row0 < rows; row0 += stride) {
    float4 v[2][NV];
    const bool has2 = row0 + 1 < rows;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < NV; ++k)
        v[r][k] = (r == 0 || has2) ? *reinterpret_cast<const float4*>(x + (long long)(row0 + r) * C + lane * 4 + k * 128)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !has2) break;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) s += (v[r][k].x + v[r][k].y) + (v[r][k].z + v[r][k].w);
      const float mu = warp_sum(s) * invC;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const float a = v[r][k].x - mu, b = v[r][k].y - mu, c = v[r][k].z - mu, d = v[r][k].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
      }
      const float rs = rsqrtf(warp_sum(q) * invC + eps);
      const long long off = (long long)(row0 + r) * C + lane * 4;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float4 o;
        o.x = (v[r][k].x - mu) * rs * g4[k].x + b4[k].x;
        o.y = (v[r][k].y - mu) * rs * g4[k].y + b4[k].y;
        o.z = (v[r][k].z - mu) * rs * g4[k].z + b4[k].z;
        o.w = (v[r][k].w - mu) * rs * g4[k].w + b4[k].w;
        if (y) *reinterpret_cast<float4*>(y + off + k * 128) = o;
        if (y16) {
          __nv_bfloat162 h2[2] = {__floats2bfloat162_rn(o.x, o.y), __floats2bfloat162_rn(o.z, o.w)};
          *reinterpret_cast<uint2*>(y16 + off + k * 128) = *reinterpret_cast<uint2*>(h2);
        }
      }
      if (lane == 0) {
        mean[row0 + r] = mu;
        rstd[row0 + r] = rs;
      }
    }
  }
}

// dx = (g - mean(g) - xhat*mean(g*xhat)) * rstd (+ dres), g = dy*gamma.  Per-CTA partial dgamma/dbeta -> part[cta][2][C].
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma, const float* __restrict__ dres,
                                                                    float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16,
                                                                    float* __restrict__ part, int rows, int C) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];  // [LN_WARPS][2][C]
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
    if ((C & 3) == 0) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 d4 = *reinterpret_cast<const float4*>(dyr + c);
        const float4 x4 = *reinterpret_cast<const float4*>(xr + c);
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + c);
        const float gx = d4.x * g4.x, gy = d4.y * g4.y, gz = d4.z * g4.z, gw = d4.w * g4.w;
        s1 += (gx + gy) + (gz + gw);
        s2 += gx * ((x4.x - mu) * rs) + gy * ((x4.y - mu) * rs) + gz * ((x4.z - mu) * rs) + gw * ((x4.w - mu) * rs);
      }
      const float c1 = warp_sum(s1) * invC, c2 = warp_sum(s2) * invC;
      for (int c = lane * 4; c < C; c += 128) {
        const float4 d4 = *reinterpret_cast<const float4*>(dyr + c);
        const float4 x4 = *reinterpret_cast<const float4*>(xr + c);
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + c);
        const float xh0 = (x4.x - mu) * rs, xh1 = (x4.y - mu) * rs, xh2 = (x4.z - mu) * rs, xh3 = (x4.w - mu) * rs;
        float4 o;
        o.x = (d4.x * g4.x - c1 - xh0 * c2) * rs;
        o.y = (d4.y * g4.y - c1 - xh1 * c2) * rs;
        o.z = (d4.z * g4.z - c1 - xh2 * c2) * rs;
        o.w = (d4.w * g4.w - c1 - xh3 * c2) * rs;
        if (dres) {
          const float4 r4 = *reinterpret_cast<const float4*>(dres + (long long)row * C + c);
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        *reinterpret_cast<float4*>(dx + (long long)row * C + c) = o;
        if (dx16) {
          __nv_bfloat162 h2[2] = {__floats2bfloat162_rn(o.x, o.y), __floats2bfloat162_rn(o.z, o.w)};
          *reinterpret_cast<uint2*>(dx16 + (long long)row * C + c) = *reinterpret_cast<uint2*>(h2);
        }
        float4* pg = reinterpret_cast<float4*>(dg + c);  // lane-private columns: no race
        float4* pb = reinterpret_cast<float4*>(db + c);
        float4 ag = *pg, ab = *pb;
        ag.x += d4.x * xh0; ag.y += d4.y * xh1; ag.z += d4.z * xh2; ag.w += d4.w * xh3;
        ab.x += d4.x; ab.y += d4.y; ab.z += d4.z; ab.w += d4.w;
        *pg = ag;
        *pb = ab;
      }
      continue;
    }
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
      if (dx16) dx16[(long long)row * C + c] = __float2bfloat16_rn(o);
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

// Register-resident variant for C = 128*NV (NV float4 per lane): the row's dy / x values are read ONCE and kept in registers
// between the two passes, dgamma / dbeta / column-sum accumulators live in registers, one combine through shared memory at
// the end.  part: [cta][3][C] = dgamma | dbeta | column sums of the OUTPUT dx (= bias gradient of whatever produced the
// tensor this gradient belongs to; see ops.py).
//
// R1 (rank-one dy, the model's last layer mlp2 = Linear(C -> 1) after ln_3): dy[row][c] = dy[row] * wvec[c] is formed on the
// fly from the [rows] gradient of the scalar output, and the first two accumulators hold S_c = sum_r dy[r]*xhat[r][c] and
// D = sum_r dy[r] instead, from which dgamma = w*S, dbeta = w*D, dW = gamma*S + beta*D, db = D.
// DY16: dy arrives as bf16 (the data-gradient GEMM that produced it wrote bf16 only: 84 MB less traffic per launch at the
// benchmark shape; the residual gradient stream `dres` stays fp32).
template <int NV, bool R1, bool DY16 = false>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_reg_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                        const float* __restrict__ gamma, const float* __restrict__ dres,
                                                                        float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16,
                                                                        float* __restrict__ part, int rows,
                                                                        const float* __restrict__ wvec) {
  pdl_sync();
  constexpr int C = NV * 128;
  __shared__ __align__(16) float sm[LN_WARPS][3][C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 g4[NV], w4[R1 ? NV : 1], ag[NV], ab[NV], as[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    g4[k] = *reinterpret_cast<const float4*>(gamma + lane * 4 + k * 128);
    if (R1) w4[R1 ? k : 0] = *reinterpret_cast<const float4*>(wvec + lane * 4 + k * 128);
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = ag[k];
    as[k] = ag[k];
  }
  const float invC = 1.f / (float)C;
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const long long off = (long long)row * C + lane * 4;
    const float mu = mean[row], rs = rstd[row];
    float4 d4[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
    const float dsc = R1 ? dy[row] : 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (R1) {
        const float4 w = w4[R1 ? k : 0];
        d4[k] = make_float4(dsc * w.x, dsc * w.y, dsc * w.z, dsc * w.w);
      } else if (DY16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + off + k * 128);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        d4[k] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d4[k] = *reinterpret_cast<const float4*>(dy + off + k * 128);
      }
      const float4 xv = *reinterpret_cast<const float4*>(x + off + k * 128);
      xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      const float gx = d4[k].x * g4[k].x, gy = d4[k].y * g4[k].y, gz = d4[k].z * g4[k].z, gw = d4[k].w * g4[k].w;
      s1 += (gx + gy) + (gz + gw);
      s2 += gx * xh[k].x + gy * xh[k].y + gz * xh[k].z + gw * xh[k].w;
    }
    const float c1 = warp_sum(s1) * invC, c2 = warp_sum(s2) * invC;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float4 o;
      o.x = (d4[k].x * g4[k].x - c1 - xh[k].x * c2) * rs;
      o.y = (d4[k].y * g4[k].y - c1 - xh[k].y * c2) * rs;
      o.z = (d4[k].z * g4[k].z - c1 - xh[k].z * c2) * rs;
      o.w = (d4[k].w * g4[k].w - c1 - xh[k].w * c2) * rs;
      if (dres) {
        const float4 r4 = *reinterpret_cast<const float4*>(dres + off + k * 128);
        o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
      }
      *reinterpret_cast<float4*>(dx + off + k * 128) = o;
      if (dx16) {
        __nv_bfloat162 h2[2] = {__floats2bfloat162_rn(o.x, o.y), __floats2bfloat162_rn(o.z, o.w)};
        *reinterpret_cast<uint2*>(dx16 + off + k * 128) = *reinterpret_cast<uint2*>(h2);
      }
      if (R1) {
        ag[k].x += dsc * xh[k].x; ag[k].y += dsc * xh[k].y; ag[k].z += dsc * xh[k].z; ag[k].w += dsc * xh[k].w;
        ab[k].x += dsc; ab[k].y += dsc; ab[k].z += dsc; ab[k].w += dsc;
      } else {
        ag[k].x += d4[k].x * xh[k].x; ag[k].y += d4[k].y * xh[k].y; ag[k].z += d4[k].z * xh[k].z; ag[k].w += d4[k].w * xh[k].w;
        ab[k].x += d4[k].x; ab[k].y += d4[k].y; ab[k].z += d4[k].z; ab[k].w += d4[k].w;
      }
      as[k].x += o.x; as[k].y += o.y; as[k].z += o.z; as[k].w += o.w;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    *reinterpret_cast<float4*>(&sm[warp][0][lane * 4 + k * 128]) = ag[k];
    *reinterpret_cast<float4*>(&sm[warp][1][lane * 4 + k * 128]) = ab[k];
    *reinterpret_cast<float4*>(&sm[warp][2][lane * 4 + k * 128]) = as[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += (&sm[w][0][0])[c];
    part[(long long)blockIdx.x * 3 * C + c] = s;
  }
}

__global__ void colsum_partial_kernel(const float* __restrict__ in, long long ld, float* __restrict__ ws, int rows, int cols,
                                      int rows_per_chunk);

__global__ void reduce_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, long long cols, long long ld) {
  pdl_sync();
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < cols; j += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < rows; ++i) s += in[(long long)i * ld + j];
    out[j] = s;
  }
}

// Tall partial buffers (hundreds of rows, a few hundred columns): 32 row lanes x 32 column lanes per CTA.  Row lane l adds
// rows l, l+32, ... in ascending order, then the lanes are combined in lane order: fixed order, bit-reproducible, and the
// loads of a thread are independent so the pass is not a chain of exposed latencies.
__global__ void __launch_bounds__(1024) reduce_rows_par_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, long long ld) {
  pdl_sync();
  __shared__ float red[32][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float acc = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int r = rl; r < rows; r += 32) acc += in[(long long)r * ld + c];
  }
  red[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float s = red[0][cl];
#pragma unroll
    for (int i = 1; i < 32; ++i) s += red[i][cl];
    out[c] = s;
  }
}

// Column sums (bias gradients): HBM-bound streaming reduction.  Each CTA owns a 128-column panel and a contiguous chunk of
// rows; 32 float4-lanes x 8 row-lanes, fixed-order combine in shared memory, then a deterministic second stage.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ in, long long ld, float* __restrict__ ws, int rows,
                                                            int cols, int rows_per_chunk) {
  pdl_sync();
  __shared__ float4 red[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cl * 4;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (c + 3 < cols) && ((ld & 3) == 0);
  if (vec) {
    for (int r = r0 + rl; r < r1; r += 8) {
      const float4 v = *reinterpret_cast<const float4*>(in + (long long)r * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  } else {
    for (int r = r0 + rl; r < r1; r += 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < cols) (&acc.x)[j] += in[(long long)r * ld + c + j];
    }
  }
  red[rl][cl] = acc;
  __syncthreads();
  if (rl == 0) {
    float4 s = red[0][cl];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      s.x += red[i][cl].x; s.y += red[i][cl].y; s.z += red[i][cl].z; s.w += red[i][cl].w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < cols) ws[(long long)blockIdx.y * cols + c + j] = (&s.x)[j];
  }
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, void* y16, float* mean,
                                  float* rstd, int rows, int C, float eps, void* stream) {
  TBNS_REQUIRE(x && gamma && beta && (y || y16) && mean && rstd, "tbns_layernorm_fwd: null pointer");
  TBNS_REQUIRE(rows >= 0 && C > 0, "tbns_layernorm_fwd: bad dims");
  if (rows == 0) return TBNS_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                         reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y16)) & 15) == 0;
  if (aligned && (C == 128 || C == 256 || C == 512)) {
    int ctas = cdiv(rows, LN_WARPS * 2);
    if (ctas > 148 * 8) ctas = 148 * 8;   // 8 CTAs of 8 warps per SM, grid-stride over row pairs
    __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(y16);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 128) TBNS_CUDA(launch_pdl(layernorm_fwd_reg_kernel<1>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y, o16, mean, rstd, rows, eps));
    else if (C == 256) TBNS_CUDA(launch_pdl(layernorm_fwd_reg_kernel<2>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y, o16, mean, rstd, rows, eps));
    else TBNS_CUDA(launch_pdl(layernorm_fwd_reg_kernel<4>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y, o16, mean, rstd, rows, eps));
    TBNS_LAUNCH_CHECK();
    return TBNS_OK;
  }
  TBNS_CUDA(launch_pdl(layernorm_fwd_kernel, dim3(cdiv(rows, LN_WARPS)), dim3(LN_WARPS * 32), 0, (cudaStream_t)stream, x, gamma, beta, y, reinterpret_cast<__nv_bfloat16*>(y16), mean, rstd, rows, C, eps));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" size_t tbns_layernorm_bwd_ws_floats(int C) { return (size_t)LN_BWD_CTAS * 3 * (size_t)C; }

extern "C" int tbns_reduce_rows(const float* in, float* out, int rows, long long cols, void* stream) {
  TBNS_REQUIRE(in && out && rows >= 0 && cols >= 0, "tbns_reduce_rows: bad args");
  if (cols == 0) return TBNS_OK;
  if (rows >= 64 && cols <= (long long)sm_count() * 32 * 8) {
    TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3((unsigned)((cols + 31) / 32)), dim3(1024), 0, (cudaStream_t)stream, in, out, rows, (int)cols, cols));
    TBNS_LAUNCH_CHECK();
    return TBNS_OK;
  }
  if (cols <= 0x7fffffffLL && (cols % 4 == 0) && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && cols / 128 < 65535 * 32) {
    // 8 row-lanes x 32 float4 column-lanes per CTA, fixed summation order
    TBNS_CUDA(launch_pdl(colsum_partial_kernel, dim3(dim3((unsigned)((cols + 127) / 128), 1)), dim3(256), 0, (cudaStream_t)stream, in, cols, out, rows, (int)cols, rows));
  } else {
    int blocks = (int)((cols + 127) / 128);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    TBNS_CUDA(launch_pdl(reduce_rows_kernel, dim3(blocks), dim3(128), 0, (cudaStream_t)stream, in, out, rows, cols, cols));
  }
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_layernorm_bwd_ctas(int rows) {
  int ctas = cdiv(rows, LN_WARPS * 2);
  return ctas > LN_BWD_CTAS ? LN_BWD_CTAS : ctas;
}

extern "C" int tbns_layernorm_bwd_supported16(int C) { return (C == 128 || C == 256 || C == 512) ? 1 : 0; }

extern "C" int tbns_layernorm_bwd16(const void* dy16, const float* x, const float* mean, const float* rstd, const float* gamma,
                                    const float* dres, float* dx, void* dx16, float* sums /* [3][C] */, float* ws, int rows, int C,
                                    void* stream) {
  TBNS_REQUIRE(dy16 && x && mean && rstd && gamma && dx && ws, "tbns_layernorm_bwd16: null pointer");
  TBNS_REQUIRE(rows > 0 && tbns_layernorm_bwd_supported16(C), "tbns_layernorm_bwd16: C=%d unsupported (128, 256, 512)", C);
  TBNS_REQUIRE(((reinterpret_cast<uintptr_t>(dy16) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                 reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dres) | reinterpret_cast<uintptr_t>(dx16)) & 15) == 0,
               "tbns_layernorm_bwd16: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(dx16);
  const float* dy = reinterpret_cast<const float*>(dy16);
  int ctas = cdiv(rows, LN_WARPS * 2);
  if (ctas > LN_BWD_CTAS) ctas = LN_BWD_CTAS;
  if (C == 128) TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<1, false, true>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, nullptr));
  else if (C == 256) TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<2, false, true>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, nullptr));
  else TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<4, false, true>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, nullptr));
  TBNS_LAUNCH_CHECK();
  // sums == nullptr: the caller reduces the per-CTA partials ws[tbns_layernorm_bwd_ctas(rows)][3*C] itself (tbns_reduce_rows on
  // another stream: the column sums are parameter gradients, off the critical path of the data gradient)
  if (sums) TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3(cdiv(3 * C, 32)), dim3(1024), 0, st, ws, sums, ctas, 3 * C, 3LL * C));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                  const float* dres, float* dx, void* dx16, float* dgamma, float* dbeta, float* dsum, float* ws,
                                  int rows, int C, void* stream) {
  TBNS_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws, "tbns_layernorm_bwd: null pointer");
  TBNS_REQUIRE(rows > 0 && C > 0, "tbns_layernorm_bwd: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(dx16);
  const bool aligned = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                         reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dres) | reinterpret_cast<uintptr_t>(dx16)) & 15) == 0;
  if (aligned && (C == 128 || C == 256 || C == 512)) {
    int ctas = cdiv(rows, LN_WARPS * 2);
    if (ctas > LN_BWD_CTAS) ctas = LN_BWD_CTAS;
    if (C == 128) TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<1, false>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, nullptr));
    else if (C == 256) TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<2, false>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, nullptr));
    else TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<4, false>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, nullptr));
    TBNS_LAUNCH_CHECK();
    // ws rows are [dgamma | dbeta | colsum(dx)], width 3C: fixed-order reduction; a single launch when the caller's three
    // outputs are one contiguous [3][C] array
    if (dsum && dbeta == dgamma + C && dsum == dgamma + 2 * C) {
      TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3(cdiv(3 * C, 32)), dim3(1024), 0, st, ws, dgamma, ctas, 3 * C, 3LL * C));
      TBNS_LAUNCH_CHECK();
      return TBNS_OK;
    }
    TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3(cdiv(C, 32)), dim3(1024), 0, st, ws, dgamma, ctas, C, 3LL * C));
    TBNS_LAUNCH_CHECK();
    TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3(cdiv(C, 32)), dim3(1024), 0, st, ws + C, dbeta, ctas, C, 3LL * C));
    TBNS_LAUNCH_CHECK();
    if (dsum) {
      TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3(cdiv(C, 32)), dim3(1024), 0, st, ws + 2 * C, dsum, ctas, C, 3LL * C));
      TBNS_LAUNCH_CHECK();
    }
    return TBNS_OK;
  }
  // generic path (any C): per-warp shared-memory accumulators
  const size_t smem = (size_t)LN_WARPS * 2 * C * sizeof(float);
  TBNS_REQUIRE(smem <= 200 * 1024, "tbns_layernorm_bwd: C=%d too large", C);
  TBNS_SMEM_OPT_IN((layernorm_bwd_kernel), 200 * 1024);
  int ctas = cdiv(rows, LN_WARPS * 4);
  if (ctas > LN_MAX_CTAS) ctas = LN_MAX_CTAS;
  if (ctas < 1) ctas = 1;
  TBNS_CUDA(launch_pdl(layernorm_bwd_kernel, dim3(ctas), dim3(LN_WARPS * 32), smem, st, dy, x, mean, rstd, gamma, dres, dx, o16, ws, rows, C));
  TBNS_LAUNCH_CHECK();
  TBNS_CUDA(launch_pdl(reduce_rows_kernel, dim3(cdiv(C, 128)), dim3(128), 0, st, ws, dgamma, ctas, C, 2LL * C));
  TBNS_LAUNCH_CHECK();
  TBNS_CUDA(launch_pdl(reduce_rows_kernel, dim3(cdiv(C, 128)), dim3(128), 0, st, ws + C, dbeta, ctas, C, 2LL * C));
  TBNS_LAUNCH_CHECK();
  if (dsum) {   // column sums of dx through the generic reduction
    float* tmp = ws;  // ws is free again: reuse it as the colsum workspace (needs COLSUM_ROWS*C <= LN_BWD_CTAS*3*C floats)
    int chunks = cdiv(rows, 64);
    if (chunks > COLSUM_ROWS) chunks = COLSUM_ROWS;
    const int rpc = cdiv(rows, chunks);
    chunks = cdiv(rows, rpc);
    TBNS_CUDA(launch_pdl(colsum_partial_kernel, dim3(dim3(cdiv(C, 128), chunks)), dim3(256), 0, st, dx, C, tmp, rows, C, rpc));
    TBNS_LAUNCH_CHECK();
    TBNS_CUDA(launch_pdl(reduce_rows_kernel, dim3(cdiv(C, 128)), dim3(128), 0, st, tmp, dsum, chunks, C, C));
    TBNS_LAUNCH_CHECK();
  }
  return TBNS_OK;
}

// bf16 input variant: 16 column-lanes x 8 columns (one uint4) per row-lane
__global__ void __launch_bounds__(256) colsum16_partial_kernel(const __nv_bfloat16* __restrict__ in, long long ld, float* __restrict__ ws,
                                                              int rows, int cols, int rows_per_chunk) {
  pdl_sync();
  __shared__ float red[16][16][8];
  const int cl = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int c = blockIdx.x * 128 + cl * 8;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c + 7 < cols) {
    for (int r = r0 + rl; r < r1; r += 16) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (long long)r * ld + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cl][j] = acc[j];
  __syncthreads();
  if (rl == 0 && c + 7 < cols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += red[i][cl][j];
      ws[(long long)blockIdx.y * cols + c + j] = s;
    }
  }
}

extern "C" size_t tbns_colsum_ws_floats(long long cols) { return (size_t)COLSUM_ROWS * (size_t)cols; }

extern "C" int tbns_colsum(const float* in, long long ld, float* out, float* ws, int rows, int cols, void* stream) {
  TBNS_REQUIRE(in && out && ws && rows > 0 && cols > 0, "tbns_colsum: bad args");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, "tbns_colsum: input must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int chunks = cdiv(rows, 64);
  if (chunks > COLSUM_ROWS) chunks = COLSUM_ROWS;
  const int rows_per_chunk = cdiv(rows, chunks);
  chunks = cdiv(rows, rows_per_chunk);
  dim3 grid(cdiv(cols, 128), chunks);
  TBNS_CUDA(launch_pdl(colsum_partial_kernel, dim3(grid), dim3(256), 0, st, in, ld, ws, rows, cols, rows_per_chunk));
  TBNS_LAUNCH_CHECK();
  return tbns_reduce_rows(ws, out, chunks, cols, stream);
}

extern "C" int tbns_colsum_bf16(const void* in16, long long ld, float* out, float* ws, int rows, int cols, void* stream) {
  TBNS_REQUIRE(in16 && out && ws && rows > 0 && cols > 0, "tbns_colsum_bf16: bad args");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(in16) & 15) == 0 && cols % 8 == 0 && ld % 8 == 0, "tbns_colsum_bf16: needs 16-byte aligned rows, cols %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  int chunks = cdiv(rows, 128);
  if (chunks > COLSUM_ROWS) chunks = COLSUM_ROWS;
  const int rows_per_chunk = cdiv(rows, chunks);
  chunks = cdiv(rows, rows_per_chunk);
  dim3 grid(cdiv(cols, 128), chunks);
  TBNS_CUDA(launch_pdl(colsum16_partial_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(in16), ld, ws, rows, cols, rows_per_chunk));
  TBNS_LAUNCH_CHECK();
  return tbns_reduce_rows(ws, out, chunks, cols, stream);
}

// ---- last layer: mlp2(ln_3(fx)) with out_dim = 1  (model/Transolver_Structured_Mesh_2D.py:72-73) -------------------------
// forward: LayerNorm and the C -> 1 projection in one pass over the row (nothing but [rows] scalars is written);
// backward: the rank-one variant of the register-resident LayerNorm backward above.
namespace tbns {
__global__ void __launch_bounds__(LN_WARPS * 32) ln_linear1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ w,
                                                                     const float* __restrict__ b, float* __restrict__ out,
                                                                     long long ldo, float* __restrict__ mean,
                                                                     float* __restrict__ rstd, int rows, int C, float eps) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= rows) return;
  const float* xr = x + (long long)row * C;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a = v.x - mu, bb = v.y - mu, cc = v.z - mu, dd = v.w - mu;
    q += (a * a + bb * bb) + (cc * cc + dd * dd);
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  float dot = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    dot += (((v.x - mu) * rs * g.x + be.x) * ww.x + ((v.y - mu) * rs * g.y + be.y) * ww.y) +
           (((v.z - mu) * rs * g.z + be.z) * ww.z + ((v.w - mu) * rs * g.w + be.w) * ww.w);
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    out[(long long)row * ldo] = dot + b[0];
    mean[row] = mu;
    rstd[row] = rs;
  }
}
}  // namespace tbns

extern "C" int tbns_ln_linear1_supported(int C) { return (C == 128 || C == 256 || C == 512) ? 1 : 0; }

extern "C" int tbns_ln_linear1_fwd_strided(const float* x, const float* gamma, const float* beta, const float* w, const float* b,
                                           float* out, long long ldo, float* mean, float* rstd, int rows, int C, float eps,
                                           void* stream);

extern "C" int tbns_ln_linear1_fwd(const float* x, const float* gamma, const float* beta, const float* w, const float* b, float* out,
                                   float* mean, float* rstd, int rows, int C, float eps, void* stream) {
  return tbns_ln_linear1_fwd_strided(x, gamma, beta, w, b, out, 1, mean, rstd, rows, C, eps, stream);
}

extern "C" int tbns_ln_linear1_fwd_strided(const float* x, const float* gamma, const float* beta, const float* w, const float* b,
                                           float* out, long long ldo, float* mean, float* rstd, int rows, int C, float eps,
                                           void* stream) {
  TBNS_REQUIRE(x && gamma && beta && w && b && out && mean && rstd && ldo >= 1, "tbns_ln_linear1_fwd: null pointer / bad stride");
  TBNS_REQUIRE(rows >= 0 && tbns_ln_linear1_supported(C), "tbns_ln_linear1_fwd: C=%d unsupported (128, 256, 512)", C);
  TBNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                 reinterpret_cast<uintptr_t>(w)) & 15) == 0, "tbns_ln_linear1_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return TBNS_OK;
  TBNS_CUDA(launch_pdl(ln_linear1_fwd_kernel, dim3(cdiv(rows, LN_WARPS)), dim3(LN_WARPS * 32), 0, (cudaStream_t)stream, x, gamma, beta, w, b, out, ldo, mean, rstd, rows, C, eps));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

/* sums: [3][C] = S (sum_r dout[r]*xhat[r][c]) | D (sum_r dout[r], replicated over c) | column sums of dx */
extern "C" int tbns_ln_linear1_bwd(const float* dout, const float* w, const float* x, const float* mean, const float* rstd,
                                   const float* gamma, const float* dres, float* dx, void* dx16, float* sums, float* ws, int rows,
                                   int C, void* stream) {
  TBNS_REQUIRE(dout && w && x && mean && rstd && gamma && dx && sums && ws, "tbns_ln_linear1_bwd: null pointer");
  TBNS_REQUIRE(rows > 0 && tbns_ln_linear1_supported(C), "tbns_ln_linear1_bwd: C=%d unsupported (128, 256, 512)", C);
  TBNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(gamma) |
                 reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dres) | reinterpret_cast<uintptr_t>(dx16)) & 15) == 0,
               "tbns_ln_linear1_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(dx16);
  int ctas = cdiv(rows, LN_WARPS * 2);
  if (ctas > LN_BWD_CTAS) ctas = LN_BWD_CTAS;
  if (C == 128) TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<1, true>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dout, x, mean, rstd, gamma, dres, dx, o16, ws, rows, w));
  else if (C == 256) TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<2, true>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dout, x, mean, rstd, gamma, dres, dx, o16, ws, rows, w));
  else TBNS_CUDA(launch_pdl(layernorm_bwd_reg_kernel<4, true>, dim3(ctas), dim3(LN_WARPS * 32), 0, st, dout, x, mean, rstd, gamma, dres, dx, o16, ws, rows, w));
  TBNS_LAUNCH_CHECK();
  TBNS_CUDA(launch_pdl(reduce_rows_par_kernel, dim3(cdiv(3 * C, 32)), dim3(1024), 0, st, ws, sums, ctas, 3 * C, 3LL * C));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}
