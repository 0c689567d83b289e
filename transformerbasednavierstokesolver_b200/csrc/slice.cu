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
#include "slice_v2.cuh"

namespace tbns {

constexpr float EPS_NORM = 1e-5f;
// one CTA per (batch, head) and only B*H of them: latency-bound chains of tiny contractions -> as many warps as a CTA can hold
// 512 threads and <= 64 registers: two CTAs share an SM, so the B*H = 160 CTAs of the benchmark shape are ONE wave on 148 SMs
// (1024-thread CTAs ran 148 + 12: two waves of a latency-bound kernel)
constexpr int TOKEN_THREADS = 512;

// grid (H, B), block 256.  All [rows][D] shared arrays use the padded stride DS = D+1 (bank-conflict free for both
// row-wise and column-wise thread mappings).
template <int DT, int GT>
__global__ void __launch_bounds__(TOKEN_THREADS, 2) token_attn_fwd_kernel(const float* __restrict__ part, int nchunk, const float* __restrict__ Wq,
                                                             const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                             const float* __restrict__ Wo, float* __restrict__ s_out,
                                                             float* __restrict__ Tt_out, float* __restrict__ tok_out,
                                                             float* __restrict__ q_out, float* __restrict__ k_out,
                                                             float* __restrict__ v_out, float* __restrict__ A_out,
                                                             float* __restrict__ O_out, float* __restrict__ P,
                                                             __nv_bfloat16* __restrict__ P16, __nv_bfloat16* __restrict__ PT16, int H,
                                                             int D_rt, int G_rt, int Cout, int stage) {
  const int D = DT > 0 ? DT : D_rt, G = GT > 0 ? GT : G_rt;   // compile-time for the common head shapes: loops unroll, index math folds
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
  float* Ou = sm + (((Ps - sm) + G * (Cout + 1) + 3) & ~3);  // [G][D] unpadded, 16-byte aligned copy of O for step 7 (stage != 0)
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
    int c = 0;
    for (; c + 3 < nchunk; c += 4) {   // four independent loads in flight, same (ascending) summation order
      const float p0 = pin[(long long)c * G * (D + 1) + o], p1 = pin[(long long)(c + 1) * G * (D + 1) + o];
      const float p2 = pin[(long long)(c + 2) * G * (D + 1) + o], p3 = pin[(long long)(c + 3) * G * (D + 1) + o];
      acc += p0; acc += p1; acc += p2; acc += p3;
    }
    for (; c < nchunk; ++c) acc += pin[(long long)c * G * (D + 1) + o];
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
    if (stage) Ou[o] = acc;
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
      if (PT16) PT16[((long long)b * Cout + c) * (H * G) + h * G + g] = __float2bfloat16_rn(acc);
    }
    if (P16) {   // centred copy (see below), slow path: P is read back from global memory
      __syncthreads();
      for (int o = tid; o < G * Cout; o += nt) {
        const int g = o / Cout, c = o - g * Cout;
        float m = 0.f;
        for (int g2 = 0; g2 < G; ++g2) m += Pout[(long long)g2 * Cout + c];
        P16[((long long)b * H * G + (long long)h * G + g) * Cout + c] = __float2bfloat16_rn(Pout[(long long)g * Cout + c] - m / (float)G);
      }
    }
    return;
  }
  bool done7 = false;
  if constexpr (DT > 0 && (DT & 3) == 0) {
    if (nt % Cout == 0) {
      // thread = one output channel c for the slice rows g0, g0 + nt/Cout, ...: its to_out row lives in registers and every
      // O row is read with 16-byte broadcast loads (all lanes of a warp share g) - 4x fewer shared-memory instructions
      const int c = tid % Cout, g0 = tid / Cout, gs = nt / Cout;
      float w[DT];
#pragma unroll
      for (int dd = 0; dd < DT; ++dd) w[dd] = Wos[c * DS + dd];
      for (int g = g0; g < G; g += gs) {
        const float4* ov = reinterpret_cast<const float4*>(Ou + g * DT);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int q4 = 0; q4 < DT / 4; ++q4) {
          const float4 o4 = ov[q4];
          a0 = fmaf(o4.x, w[4 * q4], a0);
          a1 = fmaf(o4.y, w[4 * q4 + 1], a1);
          a0 = fmaf(o4.z, w[4 * q4 + 2], a0);
          a1 = fmaf(o4.w, w[4 * q4 + 3], a1);
        }
        const float acc = a0 + a1;
        Pout[(long long)g * Cout + c] = acc;
        Ps[g * (Cout + 1) + c] = acc;
      }
      done7 = true;
    }
  }
  if (!done7) {
    for (int o = tid; o < G * Cout; o += nt) {
      const int g = o / Cout, c = o - g * Cout;
      const float* ov = O + g * DS;
      const float* wv = Wos + c * DS;
      float acc = 0.f;
#pragma unroll 8
      for (int dd = 0; dd < D; ++dd) acc = fmaf(ov[dd], wv[dd], acc);
      Pout[(long long)g * Cout + c] = acc;
      Ps[g * (Cout + 1) + c] = acc;
    }
  }
  if (P16 || PT16) __syncthreads();
  if (P16) {
    // P16 is the weight operand of the deslice gradient dw = dOut.P^T (backward).  The slice tokens of a head are all close
    // to the field's mean, so the rows P[g,:] share a large common component; the softmax backward annihilates anything
    // constant over g (dL' = w o (dw - sum_g w dw), sum_g w = 1), which would leave bf16 rounding noise of the common part
    // over a small signal.  Storing P[g,:] - mean_g P[g,:] removes it at the operand level: same gradient, ~8x less error.
    const float invG = 1.0f / (float)G;
    if (nt % Cout == 0) {   // the thread's column is fixed: one pass for the mean, one for its rows
      const int c = tid % Cout;
      float m = 0.f;
      for (int g2 = 0; g2 < G; ++g2) m += Ps[g2 * (Cout + 1) + c];
      m *= invG;
      for (int g = tid / Cout; g < G; g += nt / Cout)
        P16[((long long)b * H * G + (long long)h * G + g) * Cout + c] = __float2bfloat16_rn(Ps[g * (Cout + 1) + c] - m);
    } else {
      for (int o = tid; o < G * Cout; o += nt) {
        const int c = o % Cout, g = o / Cout;
        float m = 0.f;
        for (int g2 = 0; g2 < G; ++g2) m += Ps[g2 * (Cout + 1) + c];
        P16[((long long)b * H * G + (long long)h * G + g) * Cout + c] = __float2bfloat16_rn(Ps[g * (Cout + 1) + c] - m * invG);
      }
    }
  }
  if (PT16) {
    // transposed (K-major) bf16 copy: row c of image b holds this head's G slices contiguously
    __nv_bfloat16* ptb = PT16 + (long long)b * Cout * (H * G) + h * G;
    if ((G & 7) == 0 && ((H * G) & 7) == 0 && (reinterpret_cast<uintptr_t>(PT16) & 15) == 0) {
      // one 16-byte store of 8 consecutive slices per thread; lanes walk consecutive c (stride Cout+1 in Ps: conflict-free)
      const int G8 = G >> 3;
      for (int o = tid; o < G8 * Cout; o += nt) {
        const int c = o % Cout, g8 = (o / Cout) * 8;
        const float* ps = Ps + g8 * (Cout + 1) + c;
        __nv_bfloat162 q4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q4[j] = __floats2bfloat162_rn(ps[(2 * j) * (Cout + 1)], ps[(2 * j + 1) * (Cout + 1)]);
        *reinterpret_cast<uint4*>(ptb + (long long)c * (H * G) + g8) = *reinterpret_cast<uint4*>(q4);
      }
    } else {
      for (int o = tid; o < G * Cout; o += nt) {
        const int c = o / G, g = o - c * G;
        ptb[(long long)c * (H * G) + g] = __float2bfloat16_rn(Ps[g * (Cout + 1) + c]);
      }
    }
  }
}

// grid (H, B), block 256
template <int DT, int GT>
__global__ void __launch_bounds__(TOKEN_THREADS, 2) token_attn_bwd_kernel(const float* __restrict__ dP, const float* __restrict__ Wq,
                                                             const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                             const float* __restrict__ Wo, const float* __restrict__ s_in,
                                                             const float* __restrict__ tok_in, const float* __restrict__ q_in,
                                                             const float* __restrict__ k_in, const float* __restrict__ v_in,
                                                             const float* __restrict__ A_in, const float* __restrict__ O_in,
                                                             float* __restrict__ dTt, float* __restrict__ ds,
                                                             float* __restrict__ dWqkv_part, float* __restrict__ dWo_part, int H,
                                                             int D_rt, int G_rt, int Cout, int stage) {
  const int D = DT > 0 ? DT : D_rt, G = GT > 0 ? GT : G_rt;
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
  float* dPs = Wvs + D * DS;  // [G][Cout]  (stage != 0) this (b,h)'s rows of dP
  float* Os = dPs + G * Cout; // [G][D]     (stage != 0)
  // (to_out.weight is read straight from L2: staging it too would cost 32 KB and drop the kernel to one CTA per SM,
  //  i.e. two waves for the B*H = 160 CTAs of the benchmark shape)
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
  const float* Oh = O_in + bh * GD;
  if (stage) {
    for (int o = tid; o < G * Cout; o += nt) dPs[o] = dPh[o];
    for (int o = tid; o < GD; o += nt) Os[o] = Oh[o];
    __syncthreads();
    // dO[g,d] = sum_c dP[g,c] Wo[c,h*D+d].  The head's slice of to_out.weight streams through a double-buffered window of
    // 32 output channels in shared memory: every 128-byte line is fetched from L2 once per CTA (not once per warp - with
    // 20 batches x 32 warps hammering the same 32 KB per head that was an L2-bandwidth bound of ~50 us), and the loads of
    // window k+1 are in flight while window k is consumed.
    {
      constexpr int CH = 32;
      float* Wc = Os + GD;   // [2][CH][D]
      const int per = CH * D;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float pre[2];
      const int nck = (Cout + CH - 1) / CH;
      auto fetch = [&](int ck) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int idx = tid + r * nt;
          const int c = ck * CH + idx / D;
          pre[r] = (idx < per && c < Cout) ? __ldg(Wo + (long long)c * I + h * D + (idx % D)) : 0.f;
        }
      };
      auto commit = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int idx = tid + r * nt;
          if (idx < per) Wc[buf * per + idx] = pre[r];
        }
      };
      fetch(0);
      commit(0);
      __syncthreads();
      for (int ck = 0; ck < nck; ++ck) {
        if (ck + 1 < nck) fetch(ck + 1);
        const float* wc = Wc + (ck & 1) * per;
        const int c0 = ck * CH;
        const int cn = min(CH, Cout - c0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int o = tid + r * nt;
          if (o < GD) {
            const int g = o / D, dd = o - g * D;
            const float* a = dPs + g * Cout + c0;
            float sacc = acc[r];
            if (cn == CH && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {   // dP row: one 16-byte broadcast load per four channels
                const float4 a4 = *reinterpret_cast<const float4*>(a + j);
                sacc = fmaf(a4.x, wc[j * D + dd], sacc);
                sacc = fmaf(a4.y, wc[(j + 1) * D + dd], sacc);
                sacc = fmaf(a4.z, wc[(j + 2) * D + dd], sacc);
                sacc = fmaf(a4.w, wc[(j + 3) * D + dd], sacc);
              }
            } else if (cn == CH) {
#pragma unroll
              for (int j = 0; j < CH; ++j) sacc = fmaf(a[j], wc[j * D + dd], sacc);
            } else {
              for (int j = 0; j < cn; ++j) sacc = fmaf(a[j], wc[j * D + dd], sacc);
            }
            acc[r] = sacc;
          }
        }
        if (ck + 1 < nck) commit((ck + 1) & 1);   // the buffer read in iteration ck-1: every thread is past that barrier
        __syncthreads();
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int o = tid + r * nt;
        if (o < GD) {
          const int g = o / D, dd = o - g * D;
          dO[g * DS + dd] = acc[r];
        }
      }
    }
    // dWo_part[b, c, h*D+d] = sum_g dP[g,c] O[g,d]
    if ((D & 3) == 0) {
      // 4 consecutive dim_head columns per thread: one LDS + one LDS.128 per 4 FMAs, 16-byte stores
      const int D4 = D >> 2;
      for (int o = tid; o < Cout * D4; o += nt) {
        const int c = o / D4, d4 = (o - c * D4) * 4;
        const float* a = dPs + c;
        const float* ov = Os + d4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int g = 0; g < G; ++g, a += Cout, ov += D) {
          const float av = *a;
          const float4 o4 = *reinterpret_cast<const float4*>(ov);
          acc.x = fmaf(av, o4.x, acc.x); acc.y = fmaf(av, o4.y, acc.y); acc.z = fmaf(av, o4.z, acc.z); acc.w = fmaf(av, o4.w, acc.w);
        }
        *reinterpret_cast<float4*>(dWo_part + ((long long)b * Cout + c) * I + h * D + d4) = acc;
      }
    } else {
      for (int o = tid; o < Cout * D; o += nt) {
        const int c = o / D, dd = o - c * D;
        float acc = 0.f;
        for (int g = 0; g < G; ++g) acc = fmaf(dPs[g * Cout + c], Os[g * D + dd], acc);
        dWo_part[((long long)b * Cout + c) * I + h * D + dd] = acc;
      }
    }
  } else {
    // dO[g,d] = sum_c dP[g,c] Wo[c,h*D+d]
    for (int o = tid; o < GD; o += nt) {
      const int g = o / D, dd = o - g * D;
      float acc = 0.f;
      for (int c = 0; c < Cout; ++c) acc = fmaf(dPh[(long long)g * Cout + c], Wo[(long long)c * I + h * D + dd], acc);
      dO[g * DS + dd] = acc;
    }
    // dWo_part[b, c, h*D+d] = sum_g dP[g,c] O[g,d]
    for (int o = tid; o < Cout * D; o += nt) {
      const int c = o / D, dd = o - c * D;
      float acc = 0.f;
      for (int g = 0; g < G; ++g) acc = fmaf(dPh[(long long)g * Cout + c], Oh[g * D + dd], acc);
      dWo_part[((long long)b * Cout + c) * I + h * D + dd] = acc;
    }
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
  pdl_sync();
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < nchunk; ++c) acc += dtau_part[((long long)b * H + h) * nchunk + c];
  const float t = temperature[h];
  if (clamp && !(t >= 0.1f && t <= 5.0f)) acc = 0.f;
  dtemperature[h] = acc;
}

// Projection-bias gradients of the tensor-core route (model/Physics_Attention.py:94-97, biases of in_project_x / in_project_fx)
// from token-reduced quantities: db_x = (sum_t dL).Ws with sum_t dL = the bias column of the slice backward's partials,
// db_fx = (sum_t w).dTt.  One CTA per head, fixed summation order (replicas stay bitwise identical).
// block 256; dynamic shared memory: (G + 8 * max(D, G)) floats
__global__ void proj_bias_grad_kernel(const float* __restrict__ dWs_part, const float* __restrict__ Ws, const float* __restrict__ s,
                                      const float* __restrict__ dTt, float* __restrict__ dbx, float* __restrict__ dbfx, int B, int H,
                                      int D, int G, int groups) {
  pdl_sync();
  extern __shared__ float pbg_sm[];
  float* dbs = pbg_sm;              // [G]  sum over (batch, chunk) of the logit-bias partials
  float* red = pbg_sm + G;          // [8][max(D, G)]
  const int h = blockIdx.x, tid = threadIdx.x;
  const int W = D > G ? D : G;
  // (1) dbs[g]: the B*groups partials are split over up to 8 thread groups, then combined in a fixed order
  {
    const int g = tid % G, part = tid / G, P = min(8, (int)blockDim.x / G);
    if (part < P) {
      float a = 0.f;
      const int n = B * groups;
#pragma unroll 4
      for (int i = part; i < n; i += P) {
        const int b = i / groups, c = i - b * groups;
        a += dWs_part[((((long long)b * H + h) * groups + c) * G + g) * (D + 1) + D];
      }
      red[part * W + g] = a;
    }
    __syncthreads();
    if (tid < G) {
      float a = 0.f;
      for (int q = 0; q < P; ++q) a += red[q * W + tid];
      dbs[tid] = a;
    }
    __syncthreads();
  }
  // (2) db_fx[d] = sum_{b,g} s[b,h,g] dTt[b,h,g,d], same split; db_x[d] = sum_g dbs[g] Ws[g,d]
  {
    const int d = tid % D, part = tid / D, P = min(8, (int)blockDim.x / D);
    if (part < P) {
      float acc = 0.f;
      const int n = B * G;
#pragma unroll 4
      for (int bg = part; bg < n; bg += P) {
        const int b = bg / G, g = bg - b * G;
        acc = fmaf(s[((long long)b * H + h) * G + g], dTt[(((long long)b * H + h) * G + g) * D + d], acc);
      }
      red[part * W + d] = acc;
    }
    __syncthreads();
    if (tid < D) {
      float f = 0.f;
      for (int q = 0; q < P; ++q) f += red[q * W + tid];
      dbfx[h * D + tid] = f;
      float x = 0.f;
      for (int g = 0; g < G; ++g) x = fmaf(dbs[g], Ws[g * D + tid], x);
      dbx[h * D + tid] = x;
    }
  }
}

// Wf[n][tap*C+ci], Wd[ci][tap*2I+n], bcat[n]  from nn.Conv2d/Linear weights [I][C][taps]; fp32 and/or bf16 (K-major tensor-core
// operand) outputs: the bf16 copies are written directly, so a weight refresh after an optimizer step is one launch per layer
__global__ void pack_proj_weights_kernel(const float* __restrict__ Wx, const float* __restrict__ bx, const float* __restrict__ Wfx,
                                         const float* __restrict__ bfx, float* __restrict__ Wf, float* __restrict__ Wd,
                                         __nv_bfloat16* __restrict__ Wf16, __nv_bfloat16* __restrict__ Wd16,
                                         float* __restrict__ bcat, int I, int C, int taps) {
  pdl_sync();
  const long long total = 2LL * I * C * taps;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates the packed fprop layout (coalesced writes to Wf)
    const int n = (int)(idx / ((long long)taps * C));
    const int r = (int)(idx - (long long)n * taps * C);
    const int tap = r / C, ci = r - tap * C;
    const float* src = n < I ? Wx : Wfx;
    const int co = n < I ? n : n - I;
    const float v = src[((long long)co * C + ci) * taps + tap];
    const long long di = (long long)ci * taps * 2 * I + (long long)tap * 2 * I + n;
    if (Wf) Wf[idx] = v;
    if (Wd) Wd[di] = v;
    if (Wf16) Wf16[idx] = __float2bfloat16_rn(v);
    if (Wd16) Wd16[di] = __float2bfloat16_rn(v);
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < 2 * I; n += gridDim.x * blockDim.x) bcat[n] = n < I ? bx[n] : bfx[n - I];
}

// bf16 copy [R][Kp] and / or transposed bf16 copy [Kp][R] of an fp32 matrix [R][K] (K zero-padded to Kp): both K-major
// operand layouts of a Linear weight (forward / data-gradient contraction) in ONE launch.  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) cast_bf16_pair_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out,
                                                            __nv_bfloat16* __restrict__ outT, int R, int K, int Kp) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + ty + 8 * j, k = k0 + tx;
    const float v = (r < R && k < K) ? W[(long long)r * K + k] : 0.f;
    tile[ty + 8 * j][tx] = v;
    if (out && r < R && k < Kp) out[(long long)r * Kp + k] = __float2bfloat16_rn(v);
  }
  if (!outT) return;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + ty + 8 * j, r = r0 + tx;
    if (k < Kp && r < R) outT[(long long)k * R + r] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
  }
}

#define TBNS_SLICE_EXTERN(D_, G_)                                                                                                          \
  extern template int launch_slice_fwd<D_, G_>(const float*, const float*, const float*, const float*, float*, __nv_bfloat16*, float*, int, \
                                               int, int, int, int, cudaStream_t);                                                          \
  extern template int launch_slice_bwd<D_, G_>(const float*, const float*, const float*, const float*, const float*, const float*,         \
                                               const float*, float*, __nv_bfloat16*, float*, float*, float*, int, int, int, int, int,      \
                                               cudaStream_t);

// supported (dim_head, slice_num) pairs: powers of two, dim_head 8..64, slice_num 4..64
#define TBNS_SLICE_SHAPES(X) \
  X(8, 4) X(8, 8) X(8, 16) X(8, 32) X(8, 64) X(16, 4) X(16, 8) X(16, 16) X(16, 32) X(16, 64) \
  X(32, 4) X(32, 8) X(32, 16) X(32, 32) X(32, 64) X(64, 4) X(64, 8) X(64, 16) X(64, 32) X(64, 64)

#define X(d, g) TBNS_SLICE_EXTERN(d, g)
TBNS_SLICE_SHAPES(X)
#undef X

static bool slice_shape_ok(int D, int G) {
#define X(d, g) if (D == d && G == g) return true;
  TBNS_SLICE_SHAPES(X)
#undef X
  return false;
}

static size_t token_fwd_smem(int D, int G) { return sizeof(float) * ((size_t)5 * G * (D + 1) + (size_t)G * (G + 1) + G + (size_t)3 * D * (D + 1)); }
static size_t token_fwd_stage_smem(int D, int G, int Cout) { return sizeof(float) * ((size_t)Cout * (D + 1) + (size_t)G * (Cout + 1) + (size_t)G * D + 4); }
static size_t token_bwd_smem(int D, int G) { return sizeof(float) * ((size_t)8 * G * (D + 1) + (size_t)2 * G * (G + 1) + G + (size_t)3 * D * (D + 1)); }
constexpr size_t SMEM_LIMIT = 227 * 1024;

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_slice_groups(int B, int N, int H) {
  // CTAs per (batch, head): each CTA walks ceil(nchunk / groups) 128-token chunks.  Pick the count that minimises the number
  // of chunk-times on the critical path, rounds(B*H*groups over the resident-CTA slots) x chunks per CTA, for the backward
  // kernel (2 CTAs / SM, weighted double) and the forward kernel (4 CTAs / SM); ties go to fewer partials.
  const int nchunk = cdiv(N, TOK);
  static const int forced = [] { const char* e = getenv("TBNS_SLICE_GROUPS"); return e ? atoi(e) : 0; }();   // tuning knob
  if (forced > 0) return forced < nchunk ? forced : nchunk;
  const long long bh = (long long)(B > 0 ? B : 1) * (H > 0 ? H : 1);
  long long best_cost = -1;
  int best = 1;
  for (int g = 1; g <= nchunk; ++g) {
    const long long per = cdiv(nchunk, g);
    const long long s2 = 2LL * sm_count(), s4 = 4LL * sm_count();
    const long long cost = 2 * ((bh * g + s2 - 1) / s2) * per + ((bh * g + s4 - 1) / s4) * per;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = g;
    }
  }
  return best;
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
  set_error("tbns_pa_slice_fwd: dim_head=%d / slice_num=%d unsupported", D, G);
  return TBNS_ERR_UNSUPPORTED;
}

// SIMT token kernels: any (dim_head, slice_num); the extern "C" entry points live in token.cu, which takes the warp-MMA
// kernels for dim_head 32 and falls back to these
namespace tbns {
int token_attn_fwd_simt(const float* part, int nchunk, const float* Wq, const float* Wk, const float* Wv,
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
  dim3 grid(H, B);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* p16 = reinterpret_cast<__nv_bfloat16*>(P16);
  __nv_bfloat16* pt16 = reinterpret_cast<__nv_bfloat16*>(PT16);
#define TBNS_TOKEN_FWD(DT_, GT_)                                                                                                     \
  do {                                                                                                                               \
    TBNS_SMEM_OPT_IN((token_attn_fwd_kernel<DT_, GT_>), (int)SMEM_LIMIT);                                                            \
    token_attn_fwd_kernel<DT_, GT_><<<grid, TOKEN_THREADS, smem, st>>>(part, nchunk, Wq, Wk, Wv, Wo, s, Tt, tok, q, k, v, A, O, P, p16, \
                                                                       pt16, H, D, G, Cout, stage);                                  \
  } while (0)
  if (D == 32 && G == 32) TBNS_TOKEN_FWD(32, 32);
  else if (D == 32 && G == 64) TBNS_TOKEN_FWD(32, 64);
  else if (D == 16 && G == 64) TBNS_TOKEN_FWD(16, 64);
  else if (D == 8 && G == 32) TBNS_TOKEN_FWD(8, 32);
  else TBNS_TOKEN_FWD(0, 0);
#undef TBNS_TOKEN_FWD
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

int token_attn_bwd_simt(const float* dP, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                                      const float* s, const float* tok, const float* q, const float* k, const float* v,
                                      const float* A, const float* O, float* dTt, float* ds, float* dWqkv_part, float* dWo_part,
                                      int B, int H, int D, int G, int Cout, void* stream) {
  TBNS_REQUIRE(dP && Wq && Wk && Wv && Wo && s && tok && q && k && v && A && O && dTt && ds && dWqkv_part && dWo_part,
               "tbns_pa_token_attn_bwd: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && G > 0 && Cout > 0, "tbns_pa_token_attn_bwd: bad dims");
  size_t smem = token_bwd_smem(D, G);
  TBNS_REQUIRE(smem <= SMEM_LIMIT, "tbns_pa_token_attn_bwd: dim_head=%d slice_num=%d exceed shared memory", D, G);
  int stage = 0;
  const size_t extra = sizeof(float) * ((size_t)G * Cout + (size_t)G * D + (size_t)2 * 32 * D);   // dP tile, O, to_out window x2
  if (smem + extra <= SMEM_LIMIT && G * D <= 4 * TOKEN_THREADS && 32 * D <= 2 * TOKEN_THREADS) {
    stage = 1;
    smem += extra;
  }
  dim3 grid(H, B);
  cudaStream_t st = (cudaStream_t)stream;
#define TBNS_TOKEN_BWD(DT_, GT_)                                                                                                       \
  do {                                                                                                                                 \
    TBNS_SMEM_OPT_IN((token_attn_bwd_kernel<DT_, GT_>), (int)SMEM_LIMIT);                                                              \
    token_attn_bwd_kernel<DT_, GT_><<<grid, TOKEN_THREADS, smem, st>>>(dP, Wq, Wk, Wv, Wo, s, tok, q, k, v, A, O, dTt, ds, dWqkv_part,  \
                                                                       dWo_part, H, D, G, Cout, stage);                                \
  } while (0)
  if (D == 32 && G == 32) TBNS_TOKEN_BWD(32, 32);
  else if (D == 32 && G == 64) TBNS_TOKEN_BWD(32, 64);
  else if (D == 16 && G == 64) TBNS_TOKEN_BWD(16, 64);
  else if (D == 8 && G == 32) TBNS_TOKEN_BWD(8, 32);
  else TBNS_TOKEN_BWD(0, 0);
#undef TBNS_TOKEN_BWD
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}
}  // namespace tbns

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
  set_error("tbns_pa_slice_bwd: dim_head=%d / slice_num=%d unsupported", D, G);
  return TBNS_ERR_UNSUPPORTED;
}

extern "C" int tbns_pa_dtau_finish(const float* dtau_part, const float* temperature, float* dtemperature, int B, int H, int nchunk,
                                   int clamp, void* stream) {
  TBNS_REQUIRE(dtau_part && temperature && dtemperature && B > 0 && H > 0 && nchunk > 0, "tbns_pa_dtau_finish: bad args");
  TBNS_CUDA(launch_pdl(dtau_finish_kernel, dim3(cdiv(H, 64)), dim3(64), 0, (cudaStream_t)stream, dtau_part, temperature, dtemperature, B, H, nchunk, clamp));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_pa_proj_bias_grad(const float* dWs_part, const float* Ws, const float* s, const float* dTt, float* dbx, float* dbfx,
                                      int B, int H, int D, int G, int groups, void* stream) {
  TBNS_REQUIRE(dWs_part && Ws && s && dTt && dbx && dbfx, "tbns_pa_proj_bias_grad: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && D <= 256 && G > 0 && G <= 256 && groups > 0, "tbns_pa_proj_bias_grad: bad dims");
  TBNS_CUDA(launch_pdl(proj_bias_grad_kernel, dim3(H), dim3(256), (size_t)(G + 8 * (D > G ? D : G)) * sizeof(float), (cudaStream_t)stream, dWs_part, Ws,
                       s, dTt, dbx, dbfx, B, H, D, G, groups));
  return TBNS_OK;
}

extern "C" int tbns_pack_proj_weights16(const float* Wx, const float* bx, const float* Wfx, const float* bfx, float* Wf, float* Wd,
                                        void* Wf16, void* Wd16, float* bcat, int I, int C, int taps, void* stream) {
  TBNS_REQUIRE(Wx && bx && Wfx && bfx && (Wf || Wf16) && bcat, "tbns_pack_proj_weights: null pointer");
  TBNS_REQUIRE(I > 0 && C > 0 && (taps == 1 || taps == 9), "tbns_pack_proj_weights: bad dims");
  const long long total = 2LL * I * C * taps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  TBNS_CUDA(launch_pdl(pack_proj_weights_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, Wx, bx, Wfx, bfx, Wf, Wd, reinterpret_cast<__nv_bfloat16*>(Wf16),
                                                                     reinterpret_cast<__nv_bfloat16*>(Wd16), bcat, I, C, taps));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_pack_proj_weights(const float* Wx, const float* bx, const float* Wfx, const float* bfx, float* Wf, float* Wd,
                                      float* bcat, int I, int C, int taps, void* stream) {
  TBNS_REQUIRE(Wf != nullptr, "tbns_pack_proj_weights: null pointer");
  return tbns_pack_proj_weights16(Wx, bx, Wfx, bfx, Wf, Wd, nullptr, nullptr, bcat, I, C, taps, stream);
}

extern "C" int tbns_cast_bf16_pair(const float* W, void* out16, void* outT16, int R, int K, int Kp, void* stream) {
  TBNS_REQUIRE(W && (out16 || outT16) && R > 0 && K > 0 && Kp >= K, "tbns_cast_bf16_pair: bad args");
  dim3 grid(cdiv(Kp, 32), cdiv(R, 32));
  TBNS_CUDA(launch_pdl(cast_bf16_pair_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, W, reinterpret_cast<__nv_bfloat16*>(out16),
                                                                reinterpret_cast<__nv_bfloat16*>(outT16), R, K, Kp));
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}
