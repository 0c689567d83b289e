// Slice / token / deslice-side small stages of Physics-Attention, forward and backward (all fp32).
//
//   slice_fwd      model/Physics_Attention.py:40-42 (irregular) / :98-101 (structured):
//                  logits = X.Ws^T + bs, w = softmax(logits / tau), per-chunk partials of sum_n w and sum_n w (x) F
//   token_attn_fwd :43-52 / :102-111  (+ fold of to_out (:57/:119) into P = O.Wo_h^T)
//   *_bwd          SURVEY.md §8 (a-bwd), restated in oracle/physics_attention.py
//
// These stages are HBM/latency bound (K = dim_head or slice_num <= 64): CUDA-core fp32 with shared-memory
// staging and coalesced global access; reductions over tokens are done in a fixed order (per-chunk partials,
// then a deterministic second stage) so data-parallel replicas stay bitwise identical.
#include "common.cuh"

namespace tbns {

constexpr int TOK = 128;  // tokens per CTA in the slice kernels (== blockDim.x)
constexpr float EPS_NORM = 1e-5f;

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
  float* Wsm = Gs + TOK * LS;     // [G][D]
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
      float x[D];
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Xs + tid * XS + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
      float L[G];   // pre-temperature logits, later the slice weights
      float mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float a = bsm[g];
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + g * D + j);
          a = fmaf(x[j], wv.x, a); a = fmaf(x[j + 1], wv.y, a); a = fmaf(x[j + 2], wv.z, a); a = fmaf(x[j + 3], wv.w, a);
        }
        L[g] = a;
        mx = fmaxf(mx, a * inv_tau);
      }
      // x <- F row ; total gradient wrt w:  dwv[g] = dw[g] + ds[g] + <F, dTt[g]>
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(Fs + tid * XS + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
      float dwv[G];
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float a = Gs[tid * LS + g] + dss[g];
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 tv = *reinterpret_cast<const float4*>(dTs + g * D + j);
          a = fmaf(x[j], tv.x, a); a = fmaf(x[j + 1], tv.y, a); a = fmaf(x[j + 2], tv.z, a); a = fmaf(x[j + 3], tv.w, a);
        }
        dwv[g] = a;
        sum += expf(L[g] * inv_tau - mx);
      }
      const float inv = tid < valid ? 1.0f / sum : 0.0f;
      float dot = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) dot = fmaf(dwv[g], expf(L[g] * inv_tau - mx) * inv, dot);
      // softmax backward, temperature gradient; L <- w, dwv <- dL
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float wv = expf(L[g] * inv_tau - mx) * inv;
        const float dLp = wv * (dwv[g] - dot);
        dtau_acc = fmaf(dLp, L[g], dtau_acc);
        L[g] = wv;
        dwv[g] = dLp * inv_tau;
        Gs[tid * LS + g] = dwv[g];
      }
      // dF = w . dTt  (into this thread's own row of Fs: F is no longer needed by anyone else)
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 tv = *reinterpret_cast<const float4*>(dTs + g * D + j);
          x[j] = fmaf(L[g], tv.x, x[j]); x[j + 1] = fmaf(L[g], tv.y, x[j + 1]);
          x[j + 2] = fmaf(L[g], tv.z, x[j + 2]); x[j + 3] = fmaf(L[g], tv.w, x[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < D; j += 4) *reinterpret_cast<float4*>(Fs + tid * XS + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      // dX = dL . Ws  (kept in registers until every thread is done reading the X tile)
#pragma unroll
      for (int j = 0; j < D; ++j) dX[j] = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + g * D + j);
          dX[j] = fmaf(dwv[g], wv.x, dX[j]); dX[j + 1] = fmaf(dwv[g], wv.y, dX[j + 1]);
          dX[j + 2] = fmaf(dwv[g], wv.z, dX[j + 2]); dX[j + 3] = fmaf(dwv[g], wv.w, dX[j + 3]);
        }
      }
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

// grid (H, B), block 256.  All [rows][D] shared arrays use the padded stride DS = D+1 (bank-conflict free for both
// row-wise and column-wise thread mappings).
__global__ void __launch_bounds__(256) token_attn_fwd_kernel(const float* __restrict__ part, int nchunk, const float* __restrict__ Wq,
                                                             const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                             const float* __restrict__ Wo, float* __restrict__ s_out,
                                                             float* __restrict__ Tt_out, float* __restrict__ tok_out,
                                                             float* __restrict__ q_out, float* __restrict__ k_out,
                                                             float* __restrict__ v_out, float* __restrict__ A_out,
                                                             float* __restrict__ O_out, float* __restrict__ P,
                                                             __nv_bfloat16* __restrict__ P16, __nv_bfloat16* __restrict__ PT16, int H,
                                                             int D, int G, int Cout, int stage) {
  extern __shared__ float sm[];
  const int DS = D + 1, GD = G * D, GS = G * DS, AS = G + 1;
  float* tok = sm;            // [G][DS]
  float* q = tok + GS;
  float* k = q + GS;
  float* v = k + GS;
  float* O = v + GS;
  float* A = O + GS;          // [G][G+1]
  float* ssum = A + G * AS;   // [G]
  float* Wqs = ssum + G;      // [D][DS] x3
  float* Wks = Wqs + D * DS;
  float* Wvs = Wks + D * DS;
  float* Wos = Wvs + D * DS;  // [Cout][DS]    (stage != 0) this head's slice of to_out.weight
  float* Ps = Wos + Cout * DS;  // [G][Cout+1] (stage != 0)
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const long long bh = (long long)b * H + h;
  const int I = H * D;
  if (stage) {
    for (int idx = tid; idx < Cout * D; idx += nt) {
      const int c = idx / D, dd = idx - c * D;
      Wos[c * DS + dd] = Wo[(long long)c * I + h * D + dd];
    }
  }
  // 1. fixed-order reduction of the per-chunk partials
  const float* pin = part + bh * nchunk * G * (D + 1);
  for (int o = tid; o < G * (D + 1); o += nt) {
    float acc = 0.f;
    for (int c = 0; c < nchunk; ++c) acc += pin[(long long)c * G * (D + 1) + o];
    const int g = o / (D + 1), dd = o - g * (D + 1);
    if (dd == D) {
      ssum[g] = acc;
      s_out[bh * G + g] = acc;
    } else {
      q[g * DS + dd] = acc;  // q temporarily holds Tt
      Tt_out[bh * GD + g * D + dd] = acc;
    }
  }
  for (int i = tid; i < D * D; i += nt) {
    const int r = i / D, c = i - r * D;
    Wqs[r * DS + c] = Wq[i];
    Wks[r * DS + c] = Wk[i];
    Wvs[r * DS + c] = Wv[i];
  }
  __syncthreads();
  // 2. normalise
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    const float t = q[g * DS + dd] / (ssum[g] + EPS_NORM);
    tok[g * DS + dd] = t;
    tok_out[bh * GD + o] = t;
  }
  __syncthreads();
  // 3. q,k,v = tok W^T   (nn.Linear: y[i] = sum_j x[j] W[i][j])
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, i = o - g * D;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int j = 0; j < D; ++j) {
      const float t = tok[g * DS + j];
      aq = fmaf(t, Wqs[i * DS + j], aq);
      ak = fmaf(t, Wks[i * DS + j], ak);
      av = fmaf(t, Wvs[i * DS + j], av);
    }
    q[g * DS + i] = aq; k[g * DS + i] = ak; v[g * DS + i] = av;
    q_out[bh * GD + o] = aq; k_out[bh * GD + o] = ak; v_out[bh * GD + o] = av;
  }
  __syncthreads();
  // 4. dots
  const float scale = rsqrtf((float)D);
  for (int o = tid; o < G * G; o += nt) {
    const int g = o / G, g2 = o - g * G;
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(q[g * DS + dd], k[g2 * DS + dd], acc);
    A[g * AS + g2] = acc * scale;
  }
  __syncthreads();
  // 5. row softmax, one warp per row
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int g = warp; g < G; g += nw) {
      float mx = -INFINITY;
      for (int g2 = lane; g2 < G; g2 += 32) mx = fmaxf(mx, A[g * AS + g2]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int g2 = lane; g2 < G; g2 += 32) {
        const float e = expf(A[g * AS + g2] - mx);
        A[g * AS + g2] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int g2 = lane; g2 < G; g2 += 32) {
        const float a = A[g * AS + g2] * inv;
        A[g * AS + g2] = a;
        A_out[bh * G * G + g * G + g2] = a;
      }
    }
  }
  __syncthreads();
  // 6. O = A v
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    float acc = 0.f;
    for (int g2 = 0; g2 < G; ++g2) acc = fmaf(A[g * AS + g2], v[g2 * DS + dd], acc);
    O[g * DS + dd] = acc;
    O_out[bh * GD + o] = acc;
  }
  __syncthreads();
  // 7. P[b, h*G+g, c] = sum_d O[g,d] Wo[c, h*D+d]
  float* Pout = P + ((long long)b * H * G + (long long)h * G) * Cout;
  if (!stage) {
    for (int o = tid; o < G * Cout; o += nt) {
      const int g = o / Cout, c = o - g * Cout;
      const float* wr = Wo + (long long)c * I + h * D;
      float acc = 0.f;
      for (int dd = 0; dd < D; ++dd) acc = fmaf(O[g * DS + dd], wr[dd], acc);
      Pout[(long long)g * Cout + c] = acc;
      if (P16) P16[((long long)b * H * G + (long long)h * G + g) * Cout + c] = __float2bfloat16_rn(acc);
      if (PT16) PT16[((long long)b * Cout + c) * (H * G) + h * G + g] = __float2bfloat16_rn(acc);
    }
    return;
  }
  for (int o = tid; o < G * Cout; o += nt) {
    const int g = o / Cout, c = o - g * Cout;
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(O[g * DS + dd], Wos[c * DS + dd], acc);
    Pout[(long long)g * Cout + c] = acc;
    if (P16) P16[((long long)b * H * G + (long long)h * G + g) * Cout + c] = __float2bfloat16_rn(acc);
    Ps[g * (Cout + 1) + c] = acc;
  }
  if (PT16) {
    __syncthreads();
    for (int o = tid; o < G * Cout; o += nt) {  // g fastest: coalesced rows of the transposed (K-major) bf16 copy
      const int c = o / G, g = o - c * G;
      PT16[((long long)b * Cout + c) * (H * G) + h * G + g] = __float2bfloat16_rn(Ps[g * (Cout + 1) + c]);
    }
  }
}

// grid (H, B), block 256
__global__ void __launch_bounds__(256) token_attn_bwd_kernel(const float* __restrict__ dP, const float* __restrict__ Wq,
                                                             const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                             const float* __restrict__ Wo, const float* __restrict__ s_in,
                                                             const float* __restrict__ tok_in, const float* __restrict__ q_in,
                                                             const float* __restrict__ k_in, const float* __restrict__ v_in,
                                                             const float* __restrict__ A_in, const float* __restrict__ O_in,
                                                             float* __restrict__ dTt, float* __restrict__ ds,
                                                             float* __restrict__ dWqkv_part, float* __restrict__ dWo_part, int H, int D,
                                                             int G, int Cout) {
  extern __shared__ float sm[];
  const int DS = D + 1, GD = G * D, GS = G * DS, AS = G + 1;
  float* tok = sm;
  float* q = tok + GS;
  float* k = q + GS;
  float* v = k + GS;
  float* dO = v + GS;  // later reused as dtok
  float* dq = dO + GS;
  float* dk = dq + GS;
  float* dv = dk + GS;
  float* A = dv + GS;         // [G][G+1]
  float* dA = A + G * AS;     // [G][G+1], becomes dS
  float* ssum = dA + G * AS;  // [G]
  float* Wqs = ssum + G;      // [D][DS] x3
  float* Wks = Wqs + D * DS;
  float* Wvs = Wks + D * DS;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const long long bh = (long long)b * H + h;
  const int I = H * D;

  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    tok[g * DS + dd] = tok_in[bh * GD + o];
    q[g * DS + dd] = q_in[bh * GD + o];
    k[g * DS + dd] = k_in[bh * GD + o];
    v[g * DS + dd] = v_in[bh * GD + o];
  }
  for (int o = tid; o < G * G; o += nt) A[(o / G) * AS + (o % G)] = A_in[bh * G * G + o];
  for (int o = tid; o < G; o += nt) ssum[o] = s_in[bh * G + o];
  for (int i = tid; i < D * D; i += nt) {
    const int r = i / D, c = i - r * D;
    Wqs[r * DS + c] = Wq[i];
    Wks[r * DS + c] = Wk[i];
    Wvs[r * DS + c] = Wv[i];
  }
  const float* dPh = dP + ((long long)b * H * G + (long long)h * G) * Cout;  // [G][Cout]
  // dO[g,d] = sum_c dP[g,c] Wo[c,h*D+d]
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    float acc = 0.f;
    for (int c = 0; c < Cout; ++c) acc = fmaf(dPh[(long long)g * Cout + c], Wo[(long long)c * I + h * D + dd], acc);
    dO[g * DS + dd] = acc;
  }
  // dWo_part[b, c, h*D+d] = sum_g dP[g,c] O[g,d]
  const float* Oh = O_in + bh * GD;
  for (int o = tid; o < Cout * D; o += nt) {
    const int c = o / D, dd = o - c * D;
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc = fmaf(dPh[(long long)g * Cout + c], Oh[g * D + dd], acc);
    dWo_part[((long long)b * Cout + c) * I + h * D + dd] = acc;
  }
  __syncthreads();
  // dA = dO v^T ; dv = A^T dO
  for (int o = tid; o < G * G; o += nt) {
    const int g = o / G, g2 = o - g * G;
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(dO[g * DS + dd], v[g2 * DS + dd], acc);
    dA[g * AS + g2] = acc;
  }
  for (int o = tid; o < GD; o += nt) {
    const int g2 = o / D, dd = o - g2 * D;
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc = fmaf(A[g * AS + g2], dO[g * DS + dd], acc);
    dv[g2 * DS + dd] = acc;
  }
  __syncthreads();
  // dS = A o (dA - rowsum(dA o A))
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int g = warp; g < G; g += nw) {
      float r = 0.f;
      for (int g2 = lane; g2 < G; g2 += 32) r = fmaf(dA[g * AS + g2], A[g * AS + g2], r);
      r = warp_sum(r);
      for (int g2 = lane; g2 < G; g2 += 32) dA[g * AS + g2] = A[g * AS + g2] * (dA[g * AS + g2] - r);
    }
  }
  __syncthreads();
  const float scale = rsqrtf((float)D);
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    float aq = 0.f, ak = 0.f;
    for (int g2 = 0; g2 < G; ++g2) {
      aq = fmaf(dA[g * AS + g2], k[g2 * DS + dd], aq);   // dq[g] = sum_g2 dS[g,g2] k[g2]
      ak = fmaf(dA[g2 * AS + g], q[g2 * DS + dd], ak);   // dk[g] = sum_g2 dS[g2,g] q[g2]
    }
    dq[g * DS + dd] = aq * scale;
    dk[g * DS + dd] = ak * scale;
  }
  __syncthreads();
  // dtok = dq Wq + dk Wk + dv Wv   (dO buffer reused)
  float* dtok = dO;
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, j = o - g * D;
    float acc = 0.f;
    for (int i = 0; i < D; ++i) {
      acc = fmaf(dq[g * DS + i], Wqs[i * DS + j], acc);
      acc = fmaf(dk[g * DS + i], Wks[i * DS + j], acc);
      acc = fmaf(dv[g * DS + i], Wvs[i * DS + j], acc);
    }
    dtok[g * DS + j] = acc;
  }
  // dW{q,k,v}_part[i][j] = sum_g d{q,k,v}[g,i] tok[g,j]
  float* dWp = dWqkv_part + bh * 3 * D * D;
  for (int o = tid; o < D * D; o += nt) {
    const int i = o / D, j = o - i * D;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int g = 0; g < G; ++g) {
      const float t = tok[g * DS + j];
      aq = fmaf(dq[g * DS + i], t, aq);
      ak = fmaf(dk[g * DS + i], t, ak);
      av = fmaf(dv[g * DS + i], t, av);
    }
    dWp[o] = aq;
    dWp[D * D + o] = ak;
    dWp[2 * D * D + o] = av;
  }
  __syncthreads();
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    dTt[bh * GD + o] = dtok[g * DS + dd] / (ssum[g] + EPS_NORM);
  }
  for (int g = tid; g < G; g += nt) {
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(dtok[g * DS + dd], tok[g * DS + dd], acc);
    ds[bh * G + g] = -acc / (ssum[g] + EPS_NORM);
  }
}

__global__ void dtau_finish_kernel(const float* __restrict__ dtau_part, const float* __restrict__ temperature,
                                   float* __restrict__ dtemperature, int B, int H, int nchunk, int clamp) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < nchunk; ++c) acc += dtau_part[((long long)b * H + h) * nchunk + c];
  const float t = temperature[h];
  if (clamp && !(t >= 0.1f && t <= 5.0f)) acc = 0.f;
  dtemperature[h] = acc;
}

// Wf[n][tap*C+ci], Wd[ci][tap*2I+n], bcat[n]  from nn.Conv2d/Linear weights [I][C][taps]
__global__ void pack_proj_weights_kernel(const float* __restrict__ Wx, const float* __restrict__ bx, const float* __restrict__ Wfx,
                                         const float* __restrict__ bfx, float* __restrict__ Wf, float* __restrict__ Wd,
                                         float* __restrict__ bcat, int I, int C, int taps) {
  const long long total = 2LL * I * C * taps;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates the packed fprop layout (coalesced writes to Wf)
    const int n = (int)(idx / ((long long)taps * C));
    const int r = (int)(idx - (long long)n * taps * C);
    const int tap = r / C, ci = r - tap * C;
    const float* src = n < I ? Wx : Wfx;
    const int co = n < I ? n : n - I;
    const float v = src[((long long)co * C + ci) * taps + tap];
    Wf[idx] = v;
    if (Wd) Wd[(long long)ci * taps * 2 * I + (long long)tap * 2 * I + n] = v;
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < 2 * I; n += gridDim.x * blockDim.x) bcat[n] = n < I ? bx[n] : bfx[n - I];
}

template <int D, int G>
static size_t slice_fwd_v2_smem() {
  return sizeof(float) * ((size_t)2 * TOK * SliceCfg<D, G>::XS + (size_t)TOK * SliceCfg<D, G>::LS + (size_t)G * D + G);
}
template <int D, int G>
static size_t slice_bwd_v2_smem() {
  return sizeof(float) * ((size_t)2 * TOK * SliceCfg<D, G>::XS + (size_t)TOK * SliceCfg<D, G>::LS + (size_t)2 * G * D + 2 * G + TOK / 32);
}

template <int D, int G>
static int launch_slice_fwd(const float* XF, const float* Ws, const float* bs, const float* temperature, float* w, __nv_bfloat16* w16,
                            float* part, int B, int N, int H, int groups, int clamp, cudaStream_t st) {
  const size_t smem = slice_fwd_v2_smem<D, G>();
  TBNS_CUDA(cudaFuncSetAttribute(slice_fwd_v2_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(groups, H, B);
  slice_fwd_v2_kernel<D, G><<<grid, TOK, smem, st>>>(XF, Ws, bs, temperature, w, w16, part, N, H, cdiv(N, TOK), clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}
template <int D, int G>
static int launch_slice_bwd(const float* XF, const float* Ws, const float* bs, const float* temperature, const float* dw,
                            const float* dTt, const float* ds, float* dXF, __nv_bfloat16* dXF16, float* dWs_part, float* dtau_part,
                            float* dbcat_part, int B, int N, int H, int groups, int clamp, cudaStream_t st) {
  const size_t smem = slice_bwd_v2_smem<D, G>();
  TBNS_CUDA(cudaFuncSetAttribute(slice_bwd_v2_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(groups, H, B);
  slice_bwd_v2_kernel<D, G><<<grid, TOK, smem, st>>>(XF, Ws, bs, temperature, dw, dTt, ds, dXF, dXF16, dWs_part, dtau_part, dbcat_part, N,
                                                     H, cdiv(N, TOK), clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

// supported (dim_head, slice_num) pairs: powers of two, dim_head 8..64, slice_num 4..64
#define TBNS_SLICE_SHAPES(X) \
  X(8, 4) X(8, 8) X(8, 16) X(8, 32) X(8, 64) X(16, 4) X(16, 8) X(16, 16) X(16, 32) X(16, 64) \
  X(32, 4) X(32, 8) X(32, 16) X(32, 32) X(32, 64) X(64, 4) X(64, 8) X(64, 16) X(64, 32) X(64, 64)

static bool slice_shape_ok(int D, int G) {
#define X(d, g) if (D == d && G == g) return true;
  TBNS_SLICE_SHAPES(X)
#undef X
  return false;
}

static size_t token_fwd_smem(int D, int G) { return sizeof(float) * ((size_t)5 * G * (D + 1) + (size_t)G * (G + 1) + G + (size_t)3 * D * (D + 1)); }
static size_t token_fwd_stage_smem(int D, int G, int Cout) { return sizeof(float) * ((size_t)Cout * (D + 1) + (size_t)G * (Cout + 1)); }
static size_t token_bwd_smem(int D, int G) { return sizeof(float) * ((size_t)8 * G * (D + 1) + (size_t)2 * G * (G + 1) + G + (size_t)3 * D * (D + 1)); }
constexpr size_t SMEM_LIMIT = 227 * 1024;

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_slice_groups(int B, int N, int H) {
  const int nchunk = cdiv(N, TOK);
  const int bh = B * H > 0 ? B * H : 1;
  int target = cdiv(4 * 148, bh);
  if (target < 1) target = 1;
  return nchunk < target ? nchunk : target;
}

extern "C" int tbns_pa_slice_fwd(const float* XF, const float* Ws, const float* bs, const float* temperature, float* w, void* w16,
                                 float* part, int B, int N, int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF && Ws && bs && temperature && (w || w16) && part, "tbns_pa_slice_fwd: null pointer");
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0, "tbns_pa_slice_fwd: bad dims");
  TBNS_REQUIRE(slice_shape_ok(D, G), "tbns_pa_slice_fwd: dim_head=%d / slice_num=%d unsupported (powers of two, 8..64 / 4..64)", D, G);
  TBNS_REQUIRE(H <= 65535 && B <= 65535, "tbns_pa_slice_fwd: grid too large");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(XF) & 15) == 0, "tbns_pa_slice_fwd: XF must be 16-byte aligned");
  const int groups = tbns_slice_groups(B, N, H);
  cudaStream_t st = (cudaStream_t)stream;
#define X(d, g) if (D == d && G == g) return launch_slice_fwd<d, g>(XF, Ws, bs, temperature, w, reinterpret_cast<__nv_bfloat16*>(w16), part, B, N, H, groups, clamp, st);
  TBNS_SLICE_SHAPES(X)
#undef X
  return TBNS_ERR_UNSUPPORTED;
}

extern "C" int tbns_pa_token_attn_fwd(const float* part, int nchunk, const float* Wq, const float* Wk, const float* Wv,
                                      const float* Wo, float* s, float* Tt, float* tok, float* q, float* k, float* v, float* A,
                                      float* O, float* P, void* P16, void* PT16, int B, int H, int D, int G, int Cout,
                                      void* stream) {
  TBNS_REQUIRE(part && Wq && Wk && Wv && Wo && s && Tt && tok && q && k && v && A && O && P, "tbns_pa_token_attn_fwd: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && G > 0 && Cout > 0 && nchunk > 0, "tbns_pa_token_attn_fwd: bad dims");
  size_t smem = token_fwd_smem(D, G);
  TBNS_REQUIRE(smem <= SMEM_LIMIT, "tbns_pa_token_attn_fwd: dim_head=%d slice_num=%d exceed shared memory", D, G);
  int stage = 0;
  if (smem + token_fwd_stage_smem(D, G, Cout) <= SMEM_LIMIT) {
    stage = 1;
    smem += token_fwd_stage_smem(D, G, Cout);
  }
  TBNS_CUDA(cudaFuncSetAttribute(token_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  dim3 grid(H, B);
  token_attn_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(part, nchunk, Wq, Wk, Wv, Wo, s, Tt, tok, q, k, v, A, O, P,
                                                                   reinterpret_cast<__nv_bfloat16*>(P16),
                                                                   reinterpret_cast<__nv_bfloat16*>(PT16), H, D, G, Cout, stage);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_pa_token_attn_bwd(const float* dP, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                                      const float* s, const float* tok, const float* q, const float* k, const float* v,
                                      const float* A, const float* O, float* dTt, float* ds, float* dWqkv_part, float* dWo_part,
                                      int B, int H, int D, int G, int Cout, void* stream) {
  TBNS_REQUIRE(dP && Wq && Wk && Wv && Wo && s && tok && q && k && v && A && O && dTt && ds && dWqkv_part && dWo_part,
               "tbns_pa_token_attn_bwd: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && G > 0 && Cout > 0, "tbns_pa_token_attn_bwd: bad dims");
  const size_t smem = token_bwd_smem(D, G);
  TBNS_REQUIRE(smem <= SMEM_LIMIT, "tbns_pa_token_attn_bwd: dim_head=%d slice_num=%d exceed shared memory", D, G);
  TBNS_CUDA(cudaFuncSetAttribute(token_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  dim3 grid(H, B);
  token_attn_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dP, Wq, Wk, Wv, Wo, s, tok, q, k, v, A, O, dTt, ds, dWqkv_part,
                                                                   dWo_part, H, D, G, Cout);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_pa_slice_bwd(const float* XF, const float* Ws, const float* bs, const float* temperature, const float* dw,
                                 const float* dTt, const float* ds, float* dXF, void* dXF16, float* dWs_part, float* dtau_part,
                                 float* dbcat_part, int B, int N, int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF && Ws && bs && temperature && dw && dTt && ds && (dXF || dXF16) && dWs_part && dtau_part && dbcat_part,
               "tbns_pa_slice_bwd: null pointer");
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0, "tbns_pa_slice_bwd: bad dims");
  TBNS_REQUIRE(slice_shape_ok(D, G), "tbns_pa_slice_bwd: dim_head=%d / slice_num=%d unsupported (powers of two, 8..64 / 4..64)", D, G);
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(XF) & 15) == 0 && (!dXF || (reinterpret_cast<uintptr_t>(dXF) & 15) == 0) &&
                   (!dXF16 || (reinterpret_cast<uintptr_t>(dXF16) & 15) == 0),
               "tbns_pa_slice_bwd: XF / dXF must be 16-byte aligned");
  const int groups = tbns_slice_groups(B, N, H);
  cudaStream_t st = (cudaStream_t)stream;
#define X(d, g) if (D == d && G == g) return launch_slice_bwd<d, g>(XF, Ws, bs, temperature, dw, dTt, ds, dXF, reinterpret_cast<__nv_bfloat16*>(dXF16), dWs_part, dtau_part, dbcat_part, B, N, H, groups, clamp, st);
  TBNS_SLICE_SHAPES(X)
#undef X
  return TBNS_ERR_UNSUPPORTED;
}

extern "C" int tbns_pa_dtau_finish(const float* dtau_part, const float* temperature, float* dtemperature, int B, int H, int nchunk,
                                   int clamp, void* stream) {
  TBNS_REQUIRE(dtau_part && temperature && dtemperature && B > 0 && H > 0 && nchunk > 0, "tbns_pa_dtau_finish: bad args");
  dtau_finish_kernel<<<cdiv(H, 64), 64, 0, (cudaStream_t)stream>>>(dtau_part, temperature, dtemperature, B, H, nchunk, clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_pack_proj_weights(const float* Wx, const float* bx, const float* Wfx, const float* bfx, float* Wf, float* Wd,
                                      float* bcat, int I, int C, int taps, void* stream) {
  TBNS_REQUIRE(Wx && bx && Wfx && bfx && Wf && bcat, "tbns_pack_proj_weights: null pointer");
  TBNS_REQUIRE(I > 0 && C > 0 && (taps == 1 || taps == 9), "tbns_pack_proj_weights: bad dims");
  const long long total = 2LL * I * C * taps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_proj_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Wx, bx, Wfx, bfx, Wf, Wd, bcat, I, C, taps);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}
