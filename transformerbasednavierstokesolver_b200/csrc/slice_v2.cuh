// Templated slice-stage kernels (forward / backward) and their launchers; instantiated per dim_head in slice_inst_d*.cu
// so the 20 (dim_head, slice_num) specialisations compile in parallel.  See slice.cu for the stage description.
#pragma once
#include "common.cuh"

namespace tbns {

constexpr int TOK = 128;  // tokens per CTA in the slice kernels (== blockDim.x)

__device__ __forceinline__ float clamp_tau(float t, int clamp) { return clamp ? fminf(fmaxf(t, 0.1f), 5.0f) : t; }

// ------------------------------------------------------------------------------------------------
// Slice stage, forward and backward.  One CTA = 128 tokens x one head per chunk, looping over chunks
// (grid.x = groups <= nchunk) so token-reduced quantities accumulate in registers and only `groups`
// partials per (batch, head) reach HBM.  dim_head D and slice_num G are template parameters: the token's
// X / F rows and its G logits live in registers, weights are broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
template <int D, int G>
struct SliceCfg {
  static constexpr int XS = D + 4;                                   // tile row stride (floats): 16B aligned, conflict-free LDS.128
  static constexpr int LS = G + 1;
  static constexpr int NOUT = G * D;
  static constexpr int R = NOUT >= TOK ? NOUT / TOK : 1;              // partial-sum outputs per thread (same g, consecutive d)
  static_assert(D % 4 == 0 && R <= D && D % R == 0, "unsupported dim_head / slice_num combination");
};

template <int D>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, long long ld, int rows_valid, float* __restrict__ dst, int XS,
                                          int tid) {
  // 128 rows x D floats, float4 along the row; rows >= rows_valid are zero-filled
  constexpr int V = D / 4;
  for (int idx = tid; idx < TOK * V; idx += TOK) {
    const int t = idx / V, j = idx - t * V;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < rows_valid) v = *reinterpret_cast<const float4*>(src + t * ld + 4 * j);
    *reinterpret_cast<float4*>(dst + t * XS + 4 * j) = v;
  }
}

template <int D>
__device__ __forceinline__ void store_tile(float* __restrict__ dst32, __nv_bfloat16* __restrict__ dst16, long long ld, int rows_valid,
                                           const float* __restrict__ src, int XS, int tid) {
  constexpr int V = D / 4;
  for (int idx = tid; idx < TOK * V; idx += TOK) {
    const int t = idx / V, j = idx - t * V;
    if (t < rows_valid) {
      const float4 v = *reinterpret_cast<const float4*>(src + t * XS + 4 * j);
      if (dst32) *reinterpret_cast<float4*>(dst32 + t * ld + 4 * j) = v;
      if (dst16) {
        __nv_bfloat162 o[2] = {__floats2bfloat162_rn(v.x, v.y), __floats2bfloat162_rn(v.z, v.w)};
        *reinterpret_cast<uint2*>(dst16 + t * ld + 4 * j) = *reinterpret_cast<uint2*>(o);
      }
    }
  }
}

// grid (groups, H, B), block TOK
template <int D, int G>
__global__ void __launch_bounds__(TOK) slice_fwd_v2_kernel(const float* __restrict__ XF, const float* __restrict__ Ws,
                                                           const float* __restrict__ bs, const float* __restrict__ temperature,
                                                           float* __restrict__ w, __nv_bfloat16* __restrict__ w16,
                                                           float* __restrict__ part, int N, int H, int nchunk, int clamp) {
  using Cf = SliceCfg<D, G>;
  constexpr int XS = Cf::XS, LS = Cf::LS, R = Cf::R;
  extern __shared__ __align__(16) float sm[];
  float* Xs = sm;                 // [TOK][XS]
  float* Fs = Xs + TOK * XS;      // [TOK][XS]
  float* Ls = Fs + TOK * XS;      // [TOK][LS]
  float* Wsm = Ls + TOK * LS;     // [G][D]
  float* bsm = Wsm + G * D;       // [G]
  const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int I = H * D;
  for (int idx = tid; idx < G * D; idx += TOK) Wsm[idx] = Ws[idx];
  for (int idx = tid; idx < G; idx += TOK) bsm[idx] = bs[idx];
  const float inv_tau = 1.0f / clamp_tau(temperature[h], clamp);

  const bool active = tid * R < Cf::NOUT;
  const int og = (tid * R) / D, od = (tid * R) % D;
  float acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = 0.f;
  float sacc = 0.f;

  for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
    const int n0 = chunk * TOK;
    const int valid = min(TOK, N - n0);
    const long long rowbase = (long long)b * N + n0;
    const float* src = XF + rowbase * (2LL * I) + h * D;
    load_tile<D>(src, 2LL * I, valid, Xs, XS, tid);
    load_tile<D>(src + I, 2LL * I, valid, Fs, XS, tid);
    __syncthreads();
    {
      float x[D];
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Xs + tid * XS + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
      float l[G];
      float mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float a = bsm[g];
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + g * D + j);
          a = fmaf(x[j], wv.x, a); a = fmaf(x[j + 1], wv.y, a); a = fmaf(x[j + 2], wv.z, a); a = fmaf(x[j + 3], wv.w, a);
        }
        a *= inv_tau;
        l[g] = a;
        mx = fmaxf(mx, a);
      }
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        l[g] = expf(l[g] - mx);
        sum += l[g];
      }
      const float inv = tid < valid ? 1.0f / sum : 0.0f;  // tokens past N contribute nothing
#pragma unroll
      for (int g = 0; g < G; ++g) Ls[tid * LS + g] = l[g] * inv;
    }
    __syncthreads();
    // coalesced write of the slice weights (fp32 and/or bf16 copy for the tensor-core deslice)
    for (int idx = tid; idx < TOK * G; idx += TOK) {
      const int t = idx / G, g = idx - t * G;
      if (t < valid) {
        const float v = Ls[t * LS + g];
        const long long o = ((rowbase + t) * H + h) * G + g;
        if (w) w[o] = v;
        if (w16) w16[o] = __float2bfloat16_rn(v);
      }
    }
    // token-reduced partials:  Tt[g][d] += sum_t w[t][g] F[t][d],  s[g] += sum_t w[t][g]
    if (active) {
#pragma unroll 4
      for (int t = 0; t < TOK; ++t) {
        const float wv = Ls[t * LS + og];
        if (od == 0) sacc += wv;
        if constexpr (R % 4 == 0) {
#pragma unroll
          for (int j = 0; j < R; j += 4) {
            const float4 f = *reinterpret_cast<const float4*>(Fs + t * XS + od + j);
            acc[j] = fmaf(wv, f.x, acc[j]); acc[j + 1] = fmaf(wv, f.y, acc[j + 1]);
            acc[j + 2] = fmaf(wv, f.z, acc[j + 2]); acc[j + 3] = fmaf(wv, f.w, acc[j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = fmaf(wv, Fs[t * XS + od + j], acc[j]);
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    float* pout = part + (((long long)b * H + h) * gridDim.x + blockIdx.x) * G * (D + 1) + og * (D + 1);
#pragma unroll
    for (int j = 0; j < R; ++j) pout[od + j] = acc[j];
    if (od == 0) pout[D] = sacc;
  }
}

// grid (groups, H, B), block TOK
template <int D, int G>
__global__ void __launch_bounds__(TOK) slice_bwd_v2_kernel(const float* __restrict__ XF, const float* __restrict__ Ws,
                                                           const float* __restrict__ bs, const float* __restrict__ temperature,
                                                           const float* __restrict__ dw, const float* __restrict__ dTt,
                                                           const float* __restrict__ ds, float* __restrict__ dXF,
                                                           __nv_bfloat16* __restrict__ dXF16, float* __restrict__ dWs_part,
                                                           float* __restrict__ dtau_part, float* __restrict__ dbcat_part, int N, int H,
                                                           int nchunk, int clamp) {
  using Cf = SliceCfg<D, G>;
  constexpr int XS = Cf::XS, LS = Cf::LS, R = Cf::R;
  extern __shared__ __align__(16) float sm[];
  float* Xs = sm;                 // [TOK][XS]  X, later dX
  float* Fs = Xs + TOK * XS;      // [TOK][XS]  F, later dF
  float* Gs = Fs + TOK * XS;      // [TOK][LS]  dw (deslice gradient), later dL
  float* Ls = Gs + TOK * LS;      // [TOK][LS]  pre-temperature logits
  float* Wsm = Ls + TOK * LS;     // [G][D]
  float* dTs = Wsm + G * D;       // [G][D]
  float* bsm = dTs + G * D;       // [G]
  float* dss = bsm + G;           // [G]
  float* red = dss + G;           // [TOK/32]
  const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int I = H * D, HG = H * G;
  const long long bh = (long long)b * H + h;
  for (int idx = tid; idx < G * D; idx += TOK) {
    Wsm[idx] = Ws[idx];
    dTs[idx] = dTt[bh * G * D + idx];
  }
  for (int idx = tid; idx < G; idx += TOK) {
    bsm[idx] = bs[idx];
    dss[idx] = ds[bh * G + idx];
  }
  const float tau = clamp_tau(temperature[h], clamp);
  const float inv_tau = 1.0f / tau;

  const bool active = tid * R < Cf::NOUT;
  const int og = (tid * R) / D, od = (tid * R) % D;
  float acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = 0.f;
  float dbs_acc = 0.f, dtau_acc = 0.f, dbx_acc = 0.f, dbf_acc = 0.f;

  for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
    const int n0 = chunk * TOK;
    const int valid = min(TOK, N - n0);
    const long long rowbase = (long long)b * N + n0;
    const float* src = XF + rowbase * (2LL * I) + h * D;
    load_tile<D>(src, 2LL * I, valid, Xs, XS, tid);
    load_tile<D>(src + I, 2LL * I, valid, Fs, XS, tid);
    for (int idx = tid; idx < TOK * G; idx += TOK) {
      const int t = idx / G, g = idx - t * G;
      Gs[t * LS + g] = t < valid ? dw[(rowbase + t) * HG + h * G + g] : 0.f;
    }
    __syncthreads();
    float dX[D];
    {
      // per-token arrays indexed by the slice g live in this thread's own shared-memory rows (Ls: logits, Gs: dw -> dL);
      // only the dim_head-long vectors are kept in registers, so the g loops need not be unrolled
      float x[D];
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Xs + tid * XS + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
      float mx = -INFINITY;
#pragma unroll 2
      for (int g = 0; g < G; ++g) {
        float a = bsm[g];
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + g * D + j);
          a = fmaf(x[j], wv.x, a); a = fmaf(x[j + 1], wv.y, a); a = fmaf(x[j + 2], wv.z, a); a = fmaf(x[j + 3], wv.w, a);
        }
        Ls[tid * LS + g] = a;            // pre-temperature logit
        mx = fmaxf(mx, a * inv_tau);
      }
      // x <- F row ; total gradient wrt w:  Gs[g] = dw[g] + ds[g] + <F, dTt[g]>
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Fs + tid * XS + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
      float sum = 0.f;
#pragma unroll 2
      for (int g = 0; g < G; ++g) {
        float a = Gs[tid * LS + g] + dss[g];
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 tv = *reinterpret_cast<const float4*>(dTs + g * D + j);
          a = fmaf(x[j], tv.x, a); a = fmaf(x[j + 1], tv.y, a); a = fmaf(x[j + 2], tv.z, a); a = fmaf(x[j + 3], tv.w, a);
        }
        Gs[tid * LS + g] = a;
        sum += expf(Ls[tid * LS + g] * inv_tau - mx);
      }
      const float inv = tid < valid ? 1.0f / sum : 0.0f;
      float dot = 0.f;
#pragma unroll 2
      for (int g = 0; g < G; ++g) dot = fmaf(Gs[tid * LS + g], expf(Ls[tid * LS + g] * inv_tau - mx) * inv, dot);
      // softmax backward + temperature gradient, fused with dF = w.dTt (x registers) and dX = dL.Ws
#pragma unroll
      for (int j = 0; j < D; ++j) {
        x[j] = 0.f;
        dX[j] = 0.f;
      }
#pragma unroll 2
      for (int g = 0; g < G; ++g) {
        const float L = Ls[tid * LS + g];
        const float wv = expf(L * inv_tau - mx) * inv;
        const float dLp = wv * (Gs[tid * LS + g] - dot);
        dtau_acc = fmaf(dLp, L, dtau_acc);
        const float dl = dLp * inv_tau;
        Gs[tid * LS + g] = dl;
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 tv = *reinterpret_cast<const float4*>(dTs + g * D + j);
          const float4 wq = *reinterpret_cast<const float4*>(Wsm + g * D + j);
          x[j] = fmaf(wv, tv.x, x[j]); x[j + 1] = fmaf(wv, tv.y, x[j + 1]);
          x[j + 2] = fmaf(wv, tv.z, x[j + 2]); x[j + 3] = fmaf(wv, tv.w, x[j + 3]);
          dX[j] = fmaf(dl, wq.x, dX[j]); dX[j + 1] = fmaf(dl, wq.y, dX[j + 1]);
          dX[j + 2] = fmaf(dl, wq.z, dX[j + 2]); dX[j + 3] = fmaf(dl, wq.w, dX[j + 3]);
        }
      }
      // dF goes into this thread's own row of Fs (F is no longer needed by anyone else)
#pragma unroll
      for (int j = 0; j < D; j += 4) *reinterpret_cast<float4*>(Fs + tid * XS + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    }
    __syncthreads();
    // dWs[g][d] += sum_t dL[t][g] X[t][d] ; dbs[g] += sum_t dL[t][g]
    if (active) {
#pragma unroll 4
      for (int t = 0; t < TOK; ++t) {
        const float dl = Gs[t * LS + og];
        if (od == 0) dbs_acc += dl;
        if constexpr (R % 4 == 0) {
#pragma unroll
          for (int j = 0; j < R; j += 4) {
            const float4 f = *reinterpret_cast<const float4*>(Xs + t * XS + od + j);
            acc[j] = fmaf(dl, f.x, acc[j]); acc[j + 1] = fmaf(dl, f.y, acc[j + 1]);
            acc[j + 2] = fmaf(dl, f.z, acc[j + 2]); acc[j + 3] = fmaf(dl, f.w, acc[j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) acc[j] = fmaf(dl, Xs[t * XS + od + j], acc[j]);
        }
      }
    }
    if (tid < D) {  // bias gradient of in_project_fx: column sums of the dF tile
      float a = 0.f;
      for (int t = 0; t < TOK; ++t) a += Fs[t * XS + tid];
      dbf_acc += a;
    }
    float* dst32 = dXF ? dXF + rowbase * (2LL * I) + h * D : nullptr;
    __nv_bfloat16* dst16 = dXF16 ? dXF16 + rowbase * (2LL * I) + h * D : nullptr;
    store_tile<D>(dst32 ? dst32 + I : nullptr, dst16 ? dst16 + I : nullptr, 2LL * I, valid, Fs, XS, tid);
    __syncthreads();  // everyone is done reading Xs
#pragma unroll
    for (int j = 0; j < D; j += 4) *reinterpret_cast<float4*>(Xs + tid * XS + j) = make_float4(dX[j], dX[j + 1], dX[j + 2], dX[j + 3]);
    __syncthreads();
    if (tid < D) {  // bias gradient of in_project_x
      float a = 0.f;
      for (int t = 0; t < TOK; ++t) a += Xs[t * XS + tid];
      dbx_acc += a;
    }
    store_tile<D>(dst32, dst16, 2LL * I, valid, Xs, XS, tid);
    __syncthreads();
  }
  const long long slot = bh * gridDim.x + blockIdx.x;
  dtau_acc = warp_sum(dtau_acc);
  if ((tid & 31) == 0) red[tid >> 5] = dtau_acc;
  __syncthreads();
  if (tid == 0) {
    float sacc = 0.f;
    for (int i = 0; i < TOK / 32; ++i) sacc += red[i];
    dtau_part[slot] = -sacc * inv_tau * inv_tau;
  }
  if (active) {
    float* pout = dWs_part + slot * G * (D + 1) + og * (D + 1);
#pragma unroll
    for (int j = 0; j < R; ++j) pout[od + j] = acc[j];
    if (od == 0) pout[D] = dbs_acc;
  }
  if (tid < D) {
    float* pb = dbcat_part + (((long long)b * gridDim.x + blockIdx.x) * H + h) * 2 * D;
    pb[tid] = dbx_acc;
    pb[D + tid] = dbf_acc;
  }
}

template <int D, int G>
inline size_t slice_fwd_v2_smem() {
  return sizeof(float) * ((size_t)2 * TOK * SliceCfg<D, G>::XS + (size_t)TOK * SliceCfg<D, G>::LS + (size_t)G * D + G);
}
template <int D, int G>
inline size_t slice_bwd_v2_smem() {
  return sizeof(float) * ((size_t)2 * TOK * SliceCfg<D, G>::XS + (size_t)2 * TOK * SliceCfg<D, G>::LS + (size_t)2 * G * D + 2 * G + TOK / 32);
}

template <int D, int G>
int launch_slice_fwd(const float* XF, const float* Ws, const float* bs, const float* temperature, float* w, __nv_bfloat16* w16,
                            float* part, int B, int N, int H, int groups, int clamp, cudaStream_t st) {
  const size_t smem = slice_fwd_v2_smem<D, G>();
  TBNS_SMEM_OPT_IN((slice_fwd_v2_kernel<D, G>), (int)smem);
  dim3 grid(groups, H, B);
  slice_fwd_v2_kernel<D, G><<<grid, TOK, smem, st>>>(XF, Ws, bs, temperature, w, w16, part, N, H, cdiv(N, TOK), clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}
template <int D, int G>
int launch_slice_bwd(const float* XF, const float* Ws, const float* bs, const float* temperature, const float* dw,
                            const float* dTt, const float* ds, float* dXF, __nv_bfloat16* dXF16, float* dWs_part, float* dtau_part,
                            float* dbcat_part, int B, int N, int H, int groups, int clamp, cudaStream_t st) {
  const size_t smem = slice_bwd_v2_smem<D, G>();
  TBNS_SMEM_OPT_IN((slice_bwd_v2_kernel<D, G>), (int)smem);
  dim3 grid(groups, H, B);
  slice_bwd_v2_kernel<D, G><<<grid, TOK, smem, st>>>(XF, Ws, bs, temperature, dw, dTt, ds, dXF, dXF16, dWs_part, dtau_part, dbcat_part, N,
                                                     H, cdiv(N, TOK), clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}


#define TBNS_SLICE_INSTANTIATE(D_, G_)                                                                                              \
  template int launch_slice_fwd<D_, G_>(const float*, const float*, const float*, const float*, float*, __nv_bfloat16*, float*, int, \
                                        int, int, int, int, cudaStream_t);                                                          \
  template int launch_slice_bwd<D_, G_>(const float*, const float*, const float*, const float*, const float*, const float*,         \
                                        const float*, float*, __nv_bfloat16*, float*, float*, float*, int, int, int, int, int,      \
                                        cudaStream_t);
#define TBNS_SLICE_INSTANTIATE_D(D_) \
  TBNS_SLICE_INSTANTIATE(D_, 4) TBNS_SLICE_INSTANTIATE(D_, 8) TBNS_SLICE_INSTANTIATE(D_, 16) TBNS_SLICE_INSTANTIATE(D_, 32) TBNS_SLICE_INSTANTIATE(D_, 64)

}  // namespace tbns
