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

// grid (nchunk, H, B), block TOK
__global__ void __launch_bounds__(TOK) slice_fwd_kernel(const float* __restrict__ XF, const float* __restrict__ Ws,
                                                        const float* __restrict__ bs, const float* __restrict__ temperature,
                                                        float* __restrict__ w, float* __restrict__ part, int N, int H, int D, int G,
                                                        int clamp) {
  extern __shared__ float sm[];
  const int XS = D + 1, LS = G + 1;
  float* Xs = sm;                 // [TOK][D+1]
  float* Fs = Xs + TOK * XS;      // [TOK][D+1]  (column D == 1 -> sum_n w comes out of the same contraction)
  float* Ls = Fs + TOK * XS;      // [TOK][G+1]
  float* Wsm = Ls + TOK * LS;     // [G][D]
  float* bsm = Wsm + G * D;       // [G]
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int nchunk = gridDim.x;
  const int I = H * D;
  const int n0 = chunk * TOK;
  const long long rowbase = (long long)b * N + n0;

  for (int idx = tid; idx < TOK * D; idx += TOK) {
    const int t = idx / D, dd = idx - t * D;
    float xv = 0.f, fv = 0.f;
    if (n0 + t < N) {
      const float* r = XF + (rowbase + t) * (2LL * I) + h * D + dd;
      xv = r[0];
      fv = r[I];
    }
    Xs[t * XS + dd] = xv;
    Fs[t * XS + dd] = fv;
  }
  Fs[tid * XS + D] = 1.0f;
  for (int idx = tid; idx < G * D; idx += TOK) Wsm[idx] = Ws[idx];
  for (int idx = tid; idx < G; idx += TOK) bsm[idx] = bs[idx];
  __syncthreads();

  const float inv_tau = 1.0f / clamp_tau(temperature[h], clamp);
  {
    const int t = tid;
    float mx = -INFINITY;
    for (int g = 0; g < G; ++g) {
      float acc = bsm[g];
      for (int dd = 0; dd < D; ++dd) acc = fmaf(Xs[t * XS + dd], Wsm[g * D + dd], acc);
      acc *= inv_tau;
      Ls[t * LS + g] = acc;
      mx = fmaxf(mx, acc);
    }
    float sum = 0.f;
    for (int g = 0; g < G; ++g) {
      const float e = expf(Ls[t * LS + g] - mx);
      Ls[t * LS + g] = e;
      sum += e;
    }
    const float inv = (n0 + t < N) ? 1.0f / sum : 0.0f;  // tokens past N contribute nothing
    for (int g = 0; g < G; ++g) Ls[t * LS + g] *= inv;
  }
  __syncthreads();

  for (int idx = tid; idx < TOK * G; idx += TOK) {
    const int t = idx / G, g = idx - t * G;
    if (n0 + t < N) w[((rowbase + t) * H + h) * G + g] = Ls[t * LS + g];
  }
  float* pout = part + ((long long)(b * H + h) * nchunk + chunk) * G * (D + 1);
  for (int o = tid; o < G * (D + 1); o += TOK) {
    const int g = o / (D + 1), dd = o - g * (D + 1);
    float acc = 0.f;
#pragma unroll 4
    for (int t = 0; t < TOK; ++t) acc = fmaf(Ls[t * LS + g], Fs[t * XS + dd], acc);
    pout[o] = acc;
  }
}

// grid (H, B), block 256
__global__ void __launch_bounds__(256) token_attn_fwd_kernel(const float* __restrict__ part, int nchunk, const float* __restrict__ Wq,
                                                             const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                             const float* __restrict__ Wo, float* __restrict__ s_out,
                                                             float* __restrict__ Tt_out, float* __restrict__ tok_out,
                                                             float* __restrict__ q_out, float* __restrict__ k_out,
                                                             float* __restrict__ v_out, float* __restrict__ A_out,
                                                             float* __restrict__ O_out, float* __restrict__ P, int H, int D, int G,
                                                             int Cout) {
  extern __shared__ float sm[];
  const int GD = G * D, AS = G + 1;
  float* tok = sm;            // [G][D]
  float* q = tok + GD;
  float* k = q + GD;
  float* v = k + GD;
  float* O = v + GD;
  float* A = O + GD;          // [G][G+1]
  float* ssum = A + G * AS;   // [G]
  float* Wqs = ssum + G;      // [D][D] x3
  float* Wks = Wqs + D * D;
  float* Wvs = Wks + D * D;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const long long bh = (long long)b * H + h;
  const int I = H * D;

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
      q[g * D + dd] = acc;  // q temporarily holds Tt
      Tt_out[bh * GD + g * D + dd] = acc;
    }
  }
  for (int i = tid; i < D * D; i += nt) {
    Wqs[i] = Wq[i];
    Wks[i] = Wk[i];
    Wvs[i] = Wv[i];
  }
  __syncthreads();
  // 2. normalise
  for (int o = tid; o < GD; o += nt) {
    const float t = q[o] / (ssum[o / D] + EPS_NORM);
    tok[o] = t;
    tok_out[bh * GD + o] = t;
  }
  __syncthreads();
  // 3. q,k,v = tok W^T   (nn.Linear: y[i] = sum_j x[j] W[i][j])
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, i = o - g * D;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int j = 0; j < D; ++j) {
      const float t = tok[g * D + j];
      aq = fmaf(t, Wqs[i * D + j], aq);
      ak = fmaf(t, Wks[i * D + j], ak);
      av = fmaf(t, Wvs[i * D + j], av);
    }
    q[o] = aq; k[o] = ak; v[o] = av;
    q_out[bh * GD + o] = aq; k_out[bh * GD + o] = ak; v_out[bh * GD + o] = av;
  }
  __syncthreads();
  // 4. dots
  const float scale = rsqrtf((float)D);
  for (int o = tid; o < G * G; o += nt) {
    const int g = o / G, g2 = o - g * G;
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(q[g * D + dd], k[g2 * D + dd], acc);
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
    for (int g2 = 0; g2 < G; ++g2) acc = fmaf(A[g * AS + g2], v[g2 * D + dd], acc);
    O[o] = acc;
    O_out[bh * GD + o] = acc;
  }
  __syncthreads();
  // 7. P[b, h*G+g, c] = sum_d O[g,d] Wo[c, h*D+d]
  float* Pout = P + ((long long)b * H * G + (long long)h * G) * Cout;
  for (int o = tid; o < G * Cout; o += nt) {
    const int g = o / Cout, c = o - g * Cout;
    const float* wr = Wo + (long long)c * I + h * D;
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(O[g * D + dd], wr[dd], acc);
    Pout[(long long)g * Cout + c] = acc;
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
  const int GD = G * D, AS = G + 1;
  float* tok = sm;
  float* q = tok + GD;
  float* k = q + GD;
  float* v = k + GD;
  float* dO = v + GD;  // later reused as dtok
  float* dq = dO + GD;
  float* dk = dq + GD;
  float* dv = dk + GD;
  float* A = dv + GD;         // [G][G+1]
  float* dA = A + G * AS;     // [G][G+1], becomes dS
  float* ssum = dA + G * AS;  // [G]
  float* Wqs = ssum + G;
  float* Wks = Wqs + D * D;
  float* Wvs = Wks + D * D;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const long long bh = (long long)b * H + h;
  const int I = H * D;

  for (int o = tid; o < GD; o += nt) {
    tok[o] = tok_in[bh * GD + o];
    q[o] = q_in[bh * GD + o];
    k[o] = k_in[bh * GD + o];
    v[o] = v_in[bh * GD + o];
  }
  for (int o = tid; o < G * G; o += nt) A[(o / G) * AS + (o % G)] = A_in[bh * G * G + o];
  for (int o = tid; o < G; o += nt) ssum[o] = s_in[bh * G + o];
  for (int i = tid; i < D * D; i += nt) {
    Wqs[i] = Wq[i];
    Wks[i] = Wk[i];
    Wvs[i] = Wv[i];
  }
  const float* dPh = dP + ((long long)b * H * G + (long long)h * G) * Cout;  // [G][Cout]
  // dO[g,d] = sum_c dP[g,c] Wo[c,h*D+d]
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, dd = o - g * D;
    float acc = 0.f;
    for (int c = 0; c < Cout; ++c) acc = fmaf(dPh[(long long)g * Cout + c], Wo[(long long)c * I + h * D + dd], acc);
    dO[o] = acc;
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
    for (int dd = 0; dd < D; ++dd) acc = fmaf(dO[g * D + dd], v[g2 * D + dd], acc);
    dA[g * AS + g2] = acc;
  }
  for (int o = tid; o < GD; o += nt) {
    const int g2 = o / D, dd = o - g2 * D;
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc = fmaf(A[g * AS + g2], dO[g * D + dd], acc);
    dv[o] = acc;
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
      aq = fmaf(dA[g * AS + g2], k[g2 * D + dd], aq);   // dq[g] = sum_g2 dS[g,g2] k[g2]
      ak = fmaf(dA[g2 * AS + g], q[g2 * D + dd], ak);   // dk[g] = sum_g2 dS[g2,g] q[g2]
    }
    dq[o] = aq * scale;
    dk[o] = ak * scale;
  }
  __syncthreads();
  // dtok = dq Wq + dk Wk + dv Wv   (dO buffer reused)
  float* dtok = dO;
  for (int o = tid; o < GD; o += nt) {
    const int g = o / D, j = o - g * D;
    float acc = 0.f;
    for (int i = 0; i < D; ++i) {
      acc = fmaf(dq[g * D + i], Wqs[i * D + j], acc);
      acc = fmaf(dk[g * D + i], Wks[i * D + j], acc);
      acc = fmaf(dv[g * D + i], Wvs[i * D + j], acc);
    }
    dtok[o] = acc;
  }
  // dW{q,k,v}_part[i][j] = sum_g d{q,k,v}[g,i] tok[g,j]
  float* dWp = dWqkv_part + bh * 3 * D * D;
  for (int o = tid; o < D * D; o += nt) {
    const int i = o / D, j = o - i * D;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int g = 0; g < G; ++g) {
      const float t = tok[g * D + j];
      aq = fmaf(dq[g * D + i], t, aq);
      ak = fmaf(dk[g * D + i], t, ak);
      av = fmaf(dv[g * D + i], t, av);
    }
    dWp[o] = aq;
    dWp[D * D + o] = ak;
    dWp[2 * D * D + o] = av;
  }
  __syncthreads();
  for (int o = tid; o < GD; o += nt) dTt[bh * GD + o] = dtok[o] / (ssum[o / D] + EPS_NORM);
  for (int g = tid; g < G; g += nt) {
    float acc = 0.f;
    for (int dd = 0; dd < D; ++dd) acc = fmaf(dtok[g * D + dd], tok[g * D + dd], acc);
    ds[bh * G + g] = -acc / (ssum[g] + EPS_NORM);
  }
}

// grid (nchunk, H, B), block TOK
__global__ void __launch_bounds__(TOK) slice_bwd_kernel(const float* __restrict__ XF, const float* __restrict__ Ws,
                                                        const float* __restrict__ bs, const float* __restrict__ temperature,
                                                        const float* __restrict__ dw, const float* __restrict__ dTt,
                                                        const float* __restrict__ ds, float* __restrict__ dXF,
                                                        float* __restrict__ dWs_part, float* __restrict__ dtau_part, int N, int H, int D,
                                                        int G, int clamp) {
  extern __shared__ float sm[];
  const int XS = D + 1, LS = G + 1;
  float* Xs = sm;               // [TOK][D+1]  (column D == 1 for dbs), later dX
  float* Fs = Xs + TOK * XS;    // [TOK][D+1]  F, later dF
  float* Ls = Fs + TOK * XS;    // [TOK][G+1]  logits L, later w
  float* Gs = Ls + TOK * LS;    // [TOK][G+1]  dw_flat, later dL
  float* Wsm = Gs + TOK * LS;   // [G][D]
  float* dTs = Wsm + G * D;     // [G][D]
  float* bsm = dTs + G * D;     // [G]
  float* dss = bsm + G;         // [G]
  float* red = dss + G;         // [TOK/32]
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int nchunk = gridDim.x;
  const int I = H * D;
  const int n0 = chunk * TOK;
  const long long rowbase = (long long)b * N + n0;
  const long long bh = (long long)b * H + h;

  for (int idx = tid; idx < TOK * D; idx += TOK) {
    const int t = idx / D, dd = idx - t * D;
    float xv = 0.f, fv = 0.f;
    if (n0 + t < N) {
      const float* r = XF + (rowbase + t) * (2LL * I) + h * D + dd;
      xv = r[0];
      fv = r[I];
    }
    Xs[t * XS + dd] = xv;
    Fs[t * XS + dd] = fv;
  }
  Xs[tid * XS + D] = 1.0f;
  for (int idx = tid; idx < TOK * G; idx += TOK) {
    const int t = idx / G, g = idx - t * G;
    Gs[t * LS + g] = (n0 + t < N) ? dw[((rowbase + t) * H + h) * G + g] : 0.f;
  }
  for (int idx = tid; idx < G * D; idx += TOK) {
    Wsm[idx] = Ws[idx];
    dTs[idx] = dTt[bh * G * D + idx];
  }
  for (int idx = tid; idx < G; idx += TOK) {
    bsm[idx] = bs[idx];
    dss[idx] = ds[bh * G + idx];
  }
  __syncthreads();

  const float tau = clamp_tau(temperature[h], clamp);
  const float inv_tau = 1.0f / tau;
  const int t = tid;
  const bool valid = n0 + t < N;
  float dtau_acc = 0.f;
  {
    // pass 1: logits L (pre-temperature), running max of L/tau
    float mx = -INFINITY;
    for (int g = 0; g < G; ++g) {
      float acc = bsm[g];
      for (int dd = 0; dd < D; ++dd) acc = fmaf(Xs[t * XS + dd], Wsm[g * D + dd], acc);
      Ls[t * LS + g] = acc;
      mx = fmaxf(mx, acc * inv_tau);
    }
    float sum = 0.f;
    for (int g = 0; g < G; ++g) sum += expf(Ls[t * LS + g] * inv_tau - mx);
    const float inv = valid ? 1.0f / sum : 0.f;
    // pass 2: total gradient wrt w, and <dw, w>
    float dot = 0.f;
    for (int g = 0; g < G; ++g) {
      float acc = Gs[t * LS + g] + dss[g];
      for (int dd = 0; dd < D; ++dd) acc = fmaf(Fs[t * XS + dd], dTs[g * D + dd], acc);
      Gs[t * LS + g] = acc;
      const float wv = expf(Ls[t * LS + g] * inv_tau - mx) * inv;
      dot = fmaf(acc, wv, dot);
    }
    // pass 3: softmax backward, temperature gradient; Ls <- w, Gs <- dL
    for (int g = 0; g < G; ++g) {
      const float L = Ls[t * LS + g];
      const float wv = expf(L * inv_tau - mx) * inv;
      const float dLp = wv * (Gs[t * LS + g] - dot);
      dtau_acc = fmaf(dLp, L, dtau_acc);
      Ls[t * LS + g] = wv;
      Gs[t * LS + g] = dLp * inv_tau;
    }
    // dF (own row of Fs is no longer needed by this thread)
    for (int dd = 0; dd < D; ++dd) {
      float acc = 0.f;
      for (int g = 0; g < G; ++g) acc = fmaf(Ls[t * LS + g], dTs[g * D + dd], acc);
      Fs[t * XS + dd] = acc;
    }
  }
  // block-reduce the temperature gradient:  dtau = -sum dL' * L / tau^2
  dtau_acc = warp_sum(dtau_acc);
  if ((tid & 31) == 0) red[tid >> 5] = dtau_acc;
  __syncthreads();
  if (tid == 0) {
    float sacc = 0.f;
    for (int i = 0; i < TOK / 32; ++i) sacc += red[i];
    dtau_part[bh * nchunk + chunk] = -sacc * inv_tau * inv_tau;
  }
  // dWs partial (+ dbs through the ones column): [G][D+1]
  float* pout = dWs_part + (bh * nchunk + chunk) * G * (D + 1);
  for (int o = tid; o < G * (D + 1); o += TOK) {
    const int g = o / (D + 1), dd = o - g * (D + 1);
    float acc = 0.f;
#pragma unroll 4
    for (int tt = 0; tt < TOK; ++tt) acc = fmaf(Gs[tt * LS + g], Xs[tt * XS + dd], acc);
    pout[o] = acc;
  }
  // coalesced write of dF
  for (int idx = tid; idx < TOK * D; idx += TOK) {
    const int tt = idx / D, dd = idx - tt * D;
    if (n0 + tt < N) dXF[(rowbase + tt) * (2LL * I) + I + h * D + dd] = Fs[tt * XS + dd];
  }
  __syncthreads();  // everyone is done reading Xs
  for (int dd = 0; dd < D; ++dd) {
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc = fmaf(Gs[t * LS + g], Wsm[g * D + dd], acc);
    Xs[t * XS + dd] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < TOK * D; idx += TOK) {
    const int tt = idx / D, dd = idx - tt * D;
    if (n0 + tt < N) dXF[(rowbase + tt) * (2LL * I) + h * D + dd] = Xs[tt * XS + dd];
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

static size_t slice_fwd_smem(int D, int G) { return sizeof(float) * ((size_t)TOK * (D + 1) * 2 + (size_t)TOK * (G + 1) + (size_t)G * D + G); }
static size_t slice_bwd_smem(int D, int G) {
  return sizeof(float) * ((size_t)TOK * (D + 1) * 2 + (size_t)TOK * (G + 1) * 2 + (size_t)G * D * 2 + 2 * G + TOK / 32);
}
static size_t token_fwd_smem(int D, int G) { return sizeof(float) * ((size_t)5 * G * D + (size_t)G * (G + 1) + G + (size_t)3 * D * D); }
static size_t token_bwd_smem(int D, int G) { return sizeof(float) * ((size_t)8 * G * D + (size_t)2 * G * (G + 1) + G + (size_t)3 * D * D); }
constexpr size_t SMEM_LIMIT = 227 * 1024;

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_slice_nchunk(int N) { return cdiv(N, TOK); }

extern "C" int tbns_pa_slice_fwd(const float* XF, const float* Ws, const float* bs, const float* temperature, float* w, float* part,
                                 int B, int N, int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF && Ws && bs && temperature && w && part, "tbns_pa_slice_fwd: null pointer");
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0 && D > 0 && G > 0, "tbns_pa_slice_fwd: bad dims");
  const size_t smem = slice_fwd_smem(D, G);
  TBNS_REQUIRE(smem <= SMEM_LIMIT, "tbns_pa_slice_fwd: dim_head=%d slice_num=%d exceed shared memory", D, G);
  TBNS_REQUIRE(H <= 65535 && B <= 65535, "tbns_pa_slice_fwd: grid too large");
  TBNS_CUDA(cudaFuncSetAttribute(slice_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  dim3 grid(cdiv(N, TOK), H, B);
  slice_fwd_kernel<<<grid, TOK, smem, (cudaStream_t)stream>>>(XF, Ws, bs, temperature, w, part, N, H, D, G, clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_pa_token_attn_fwd(const float* part, int nchunk, const float* Wq, const float* Wk, const float* Wv,
                                      const float* Wo, float* s, float* Tt, float* tok, float* q, float* k, float* v, float* A,
                                      float* O, float* P, int B, int H, int D, int G, int Cout, void* stream) {
  TBNS_REQUIRE(part && Wq && Wk && Wv && Wo && s && Tt && tok && q && k && v && A && O && P, "tbns_pa_token_attn_fwd: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && G > 0 && Cout > 0 && nchunk > 0, "tbns_pa_token_attn_fwd: bad dims");
  const size_t smem = token_fwd_smem(D, G);
  TBNS_REQUIRE(smem <= SMEM_LIMIT, "tbns_pa_token_attn_fwd: dim_head=%d slice_num=%d exceed shared memory", D, G);
  TBNS_CUDA(cudaFuncSetAttribute(token_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  dim3 grid(H, B);
  token_attn_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(part, nchunk, Wq, Wk, Wv, Wo, s, Tt, tok, q, k, v, A, O, P, H, D, G, Cout);
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
                                 const float* dTt, const float* ds, float* dXF, float* dWs_part, float* dtau_part, int B, int N,
                                 int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF && Ws && bs && temperature && dw && dTt && ds && dXF && dWs_part && dtau_part, "tbns_pa_slice_bwd: null pointer");
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0 && D > 0 && G > 0, "tbns_pa_slice_bwd: bad dims");
  const size_t smem = slice_bwd_smem(D, G);
  TBNS_REQUIRE(smem <= SMEM_LIMIT, "tbns_pa_slice_bwd: dim_head=%d slice_num=%d exceed shared memory", D, G);
  TBNS_CUDA(cudaFuncSetAttribute(slice_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
  dim3 grid(cdiv(N, TOK), H, B);
  slice_bwd_kernel<<<grid, TOK, smem, (cudaStream_t)stream>>>(XF, Ws, bs, temperature, dw, dTt, ds, dXF, dWs_part, dtau_part, N, H, D, G, clamp);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
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
