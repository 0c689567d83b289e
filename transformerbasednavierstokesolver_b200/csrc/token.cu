// Token stage of Physics-Attention on warp-level tensor-core MMAs (dim_head 32, slice_num 32 / 64).
//
//   forward  (model/Physics_Attention.py:43-52 / :102-111, + fold of to_out (:57 / :119) into P = O.Wo_h^T)
//   backward (SURVEY.md §8 a-bwd; oracle/physics_attention.py: token_attn_bwd)
//
// Per (batch, head) the stage is a chain of ~10 contractions of 32..64 x 32 x 32..256 elements - far too small for
// tcgen05 (M >= 64, TMEM round trips) and, written as shared-memory FMA loops, bound by LDS issue: the SIMT kernels in
// slice.cu took 35 / 58 us per launch against ~5 us of memory time.  Here every contraction is a set of
// mma.sync.m16n8k8 tf32 tiles with the 3xTF32 split (hi*hi + hi*lo + lo*hi: fp32-level accuracy, so fp32 mode keeps its
// 1e-5 bound), one 16 x 8 output tile per warp and step, operands read straight from padded shared-memory arrays.
// One CTA of 256 threads per (batch, head); everything between the first global load and the last store stays on chip.
#include <stdlib.h>

#include "common.cuh"

namespace tbns {

// SIMT kernels of slice.cu: every other (dim_head, slice_num) and oversized Cout
int token_attn_fwd_simt(const float* part, int nchunk, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* s,
                        float* Tt, float* tok, float* q, float* k, float* v, float* A, float* O, float* P, void* P16, void* PT16, int B,
                        int H, int D, int G, int Cout, void* stream);
int token_attn_bwd_simt(const float* dP, const float* Wq, const float* Wk, const float* Wv, const float* Wo, const float* s,
                        const float* tok, const float* q, const float* k, const float* v, const float* A, const float* O, float* dTt,
                        float* ds, float* dWqkv_part, float* dWo_part, int B, int H, int D, int G, int Cout, void* stream);

constexpr int TK_THREADS = 256;
constexpr int TK_WARPS = TK_THREADS / 32;
constexpr int TK_D = 32;
constexpr int TK_LD = 36;          // row stride of [rows][32] operand arrays: K-contiguous fragment loads are conflict-free
constexpr float TK_EPS = 1e-5f;    // `slice_norm + 1e-5`, Physics_Attention.py:43 / :102

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// c (16 x 8 tile, fragment layout below) += A (16 x K) . B (K x 8), K a multiple of 8, 3xTF32.
//   AK: A(m, k) = A[m * lda + k]   else A(m, k) = A[k * lda + m]        (A points at the tile's first row m0)
//   BK: B(k, n) = B[n * ldb + k]   else B(k, n) = B[k * ldb + n]        (B points at the tile's first column n0)
// fragment: lane = 4 g + t;  c[0] = (g, 2t)  c[1] = (g, 2t+1)  c[2] = (g+8, 2t)  c[3] = (g+8, 2t+1)
template <bool AK, bool BK>
__device__ __forceinline__ void warp_mma(float (&c)[4], const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int K,
                                         int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 8) {
    float af[4], bf[2];
    if (AK) {
      af[0] = A[g * lda + k0 + t];
      af[1] = A[(g + 8) * lda + k0 + t];
      af[2] = A[g * lda + k0 + t + 4];
      af[3] = A[(g + 8) * lda + k0 + t + 4];
    } else {
      af[0] = A[(k0 + t) * lda + g];
      af[1] = A[(k0 + t) * lda + g + 8];
      af[2] = A[(k0 + t + 4) * lda + g];
      af[3] = A[(k0 + t + 4) * lda + g + 8];
    }
    if (BK) {
      bf[0] = B[g * ldb + k0 + t];
      bf[1] = B[g * ldb + k0 + t + 4];
    } else {
      bf[0] = B[(k0 + t) * ldb + g];
      bf[1] = B[(k0 + t + 4) * ldb + g];
    }
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(af[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(bf[i], bh[i], bl[i]);
    mma_tf32(c, al, bh);   // small terms first
    mma_tf32(c, ah, bl);
    mma_tf32(c, ah, bh);
  }
}

// write a 16 x 8 accumulator tile (times `scale`) to a row-major array (tile origin included in dst)
__device__ __forceinline__ void store_tile(float* __restrict__ dst, int ld, const float (&c)[4], int lane, float scale = 1.0f) {
  const int g = lane >> 2, t = lane & 3;
  *reinterpret_cast<float2*>(dst + g * ld + 2 * t) = make_float2(c[0] * scale, c[1] * scale);
  *reinterpret_cast<float2*>(dst + (g + 8) * ld + 2 * t) = make_float2(c[2] * scale, c[3] * scale);
}

// Wo slice of head h as the [k = channel][n = dim_head] operand, unpadded rows of 32 floats with the 8-float groups of a row
// XOR-swizzled by (row & 3): the N-contiguous fragment reads (4 consecutive rows x 8 consecutive floats) are conflict-free
__device__ __forceinline__ int wo_sw(int c, int d) { return c * TK_D + (d ^ ((c & 3) << 3)); }

template <int G>
struct TkFwdSmem {
  static constexpr int LA = G + 4;
  static constexpr int TOK = 0, Q = TOK + G * TK_LD, K = Q + G * TK_LD, V = K + G * TK_LD, O = V + G * TK_LD;
  static constexpr int A = O + G * TK_LD;
  static constexpr int WQ = A + G * LA, WK = WQ + TK_D * TK_LD, WV = WK + TK_D * TK_LD;
  static constexpr int SS = WV + TK_D * TK_LD;
  static constexpr int WO = SS + G;                 // [Cout][TK_LD]  (K-contiguous B operand of P)
  static constexpr int floats(int Cout) { return WO + Cout * TK_LD; }
};

// grid (H, B), block 256
template <int G>
__global__ void __launch_bounds__(TK_THREADS) token_attn_fwd_mma_kernel(
    const float* __restrict__ part, int nchunk, const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
    const float* __restrict__ Wo, float* __restrict__ s_out, float* __restrict__ Tt_out, float* __restrict__ tok_out,
    float* __restrict__ q_out, float* __restrict__ k_out, float* __restrict__ v_out, float* __restrict__ A_out, float* __restrict__ O_out,
    float* __restrict__ P, __nv_bfloat16* __restrict__ P16, __nv_bfloat16* __restrict__ PT16, int H, int Cout) {
  pdl_sync();
  using S = TkFwdSmem<G>;
  constexpr int D = TK_D, LD = TK_LD, LA = S::LA, GD = G * D, MT = G / 16;
  extern __shared__ float sm[];
  float* tok = sm + S::TOK;
  float* q = sm + S::Q;
  float* k = sm + S::K;
  float* v = sm + S::V;
  float* O = sm + S::O;
  float* A = sm + S::A;
  float* Wqs = sm + S::WQ;
  float* Wks = sm + S::WK;
  float* Wvs = sm + S::WV;
  float* ssum = sm + S::SS;
  float* Wos = sm + S::WO;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long bh = (long long)b * H + h;
  const int I = H * D, HG = H * G;

  // shared q/k/v weights (128-byte rows, coalesced)
  for (int idx = tid; idx < D * (D / 4); idx += TK_THREADS) {
    const int r = idx >> 3, d4 = (idx & 7) * 4;
    *reinterpret_cast<float4*>(Wqs + r * LD + d4) = *reinterpret_cast<const float4*>(Wq + r * D + d4);
    *reinterpret_cast<float4*>(Wks + r * LD + d4) = *reinterpret_cast<const float4*>(Wk + r * D + d4);
    *reinterpret_cast<float4*>(Wvs + r * LD + d4) = *reinterpret_cast<const float4*>(Wv + r * D + d4);
  }
  // this head's slice of to_out.weight: requested now (registers), parked in shared memory once the reduction scratch is free
  constexpr int WR = 8;
  float4 wreg[WR];
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const int idx = tid + r * TK_THREADS;
    if (idx < Cout * (D / 4)) wreg[r] = *reinterpret_cast<const float4*>(Wo + (long long)(idx >> 3) * I + h * D + (idx & 7) * 4);
  }
  // 1. fixed-order reduction of the per-chunk partials [nchunk][G][D+1].  The kernel's critical path is memory latency, so the
  //    loads are spread for maximal parallelism: warp w sums chunks w, w+8, ... for ALL outputs (33 independent loads per lane
  //    and chunk), the eight warp sums are combined in warp order through shared memory (the to_out slice is staged into that
  //    region afterwards).  Order of additions is fixed -> bitwise reproducible.
  const float* pin = part + bh * nchunk * G * (D + 1);
  constexpr int PS = G * (D + 1);       // 1056 (G = 32) / 2112 (G = 64): a multiple of 32
  float* red = Wos;                      // [8 warps][PS/2 or PS] floats <= Cout * LD (checked on the host)
  constexpr int HALF = 33 * 32;          // outputs per pass (33 per lane)
  for (int base_o = 0; base_o < PS; base_o += HALF) {
    float acc[33];
#pragma unroll
    for (int i = 0; i < 33; ++i) acc[i] = 0.f;
    for (int c = warp; c < nchunk; c += TK_WARPS) {
      const float* pc = pin + (long long)c * PS + base_o + lane;
#pragma unroll
      for (int i = 0; i < 33; ++i) acc[i] += pc[32 * i];
    }
#pragma unroll
    for (int i = 0; i < 33; ++i) red[warp * HALF + lane + 32 * i] = acc[i];
    __syncthreads();
    for (int o = tid; o < HALF; o += TK_THREADS) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < TK_WARPS; ++w) a += red[w * HALF + o];
      const int oo = base_o + o, g = oo / (D + 1), dd = oo - g * (D + 1);
      if (dd == D) {
        ssum[g] = a;
        s_out[bh * G + g] = a;
      } else {
        q[g * LD + dd] = a;   // q temporarily holds Tt
        Tt_out[bh * GD + g * D + dd] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const int idx = tid + r * TK_THREADS;
    if (idx < Cout * (D / 4)) *reinterpret_cast<float4*>(Wos + (idx >> 3) * LD + (idx & 7) * 4) = wreg[r];
  }
  for (int idx = tid + WR * TK_THREADS; idx < Cout * (D / 4); idx += TK_THREADS) {   // Cout > 256
    const int c = idx >> 3, d4 = (idx & 7) * 4;
    *reinterpret_cast<float4*>(Wos + c * LD + d4) = *reinterpret_cast<const float4*>(Wo + (long long)c * I + h * D + d4);
  }
  // 2. normalise
  for (int o = tid; o < GD; o += TK_THREADS) {
    const int g = o >> 5, dd = o & 31;
    const float t = q[g * LD + dd] / (ssum[g] + TK_EPS);
    tok[g * LD + dd] = t;
    tok_out[bh * GD + o] = t;
  }
  __syncthreads();
  // 3. q, k, v = tok W^T   (nn.Linear: y[i] = sum_j x[j] W[i][j]  ->  B(k = j, n = i) = W[i][j], K-contiguous)
  for (int job = warp; job < 3 * MT * 4; job += TK_WARPS) {
    const int which = job / (MT * 4), tile = job - which * (MT * 4), mt = tile >> 2, nt = tile & 3;
    const float* W = which == 0 ? Wqs : (which == 1 ? Wks : Wvs);
    float* dst = which == 0 ? q : (which == 1 ? k : v);
    float* gout = (which == 0 ? q_out : (which == 1 ? k_out : v_out)) + bh * GD;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    warp_mma<true, true>(c, tok + mt * 16 * LD, LD, W + nt * 8 * LD, LD, D, lane);
    __syncwarp();
    store_tile(gout + mt * 16 * D + nt * 8, D, c, lane);
    // q still holds Tt until every warp has read tok (tok is a separate array: safe to overwrite q now)
    store_tile(dst + mt * 16 * LD + nt * 8, LD, c, lane);
  }
  __syncthreads();
  // 4. dots = q k^T * D^-1/2        B(k = d, n = g2) = k[g2][d], K-contiguous
  const float scale = rsqrtf((float)D);
  for (int tile = warp; tile < MT * (G / 8); tile += TK_WARPS) {
    const int mt = tile / (G / 8), nt = tile - mt * (G / 8);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    warp_mma<true, true>(c, q + mt * 16 * LD, LD, k + nt * 8 * LD, LD, D, lane);
    store_tile(A + mt * 16 * LA + nt * 8, LA, c, lane, scale);
  }
  __syncthreads();
  // 5. row softmax, one warp per row
  for (int g = warp; g < G; g += TK_WARPS) {
    float mx = -INFINITY;
    for (int g2 = lane; g2 < G; g2 += 32) mx = fmaxf(mx, A[g * LA + g2]);
    mx = warp_max(mx);
    float e[G / 32], sum = 0.f;
#pragma unroll
    for (int u = 0; u < G / 32; ++u) {
      e[u] = expf(A[g * LA + lane + 32 * u] - mx);
      sum += e[u];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int u = 0; u < G / 32; ++u) {
      const float a = e[u] * inv;
      A[g * LA + lane + 32 * u] = a;
      A_out[bh * G * G + g * G + lane + 32 * u] = a;
    }
  }
  __syncthreads();
  // 6. O = A v                      B(k = g2, n = d) = v[g2][d], N-contiguous
  for (int tile = warp; tile < MT * 4; tile += TK_WARPS) {
    const int mt = tile >> 2, nt = tile & 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    warp_mma<true, false>(c, A + mt * 16 * LA, LA, v + nt * 8, LD, G, lane);
    store_tile(O + mt * 16 * LD + nt * 8, LD, c, lane);
    store_tile(O_out + bh * GD + mt * 16 * D + nt * 8, D, c, lane);
  }
  __syncthreads();
  // 7. P[b, h*G + g, c] = sum_d O[g, d] Wo[c, h*D + d]: a warp owns 8-channel column tiles and computes ALL slice rows of
  //    them, so the column means over g (for the centred bf16 copy) are warp-local.
  const int g = lane >> 2, t = lane & 3;
  for (int nt = warp; nt < Cout / 8; nt += TK_WARPS) {
    float c[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      c[mt][0] = c[mt][1] = c[mt][2] = c[mt][3] = 0.f;
      warp_mma<true, true>(c[mt], O + mt * 16 * LD, LD, Wos + nt * 8 * LD, LD, D, lane);
    }
    const int col = nt * 8 + 2 * t;
    if (P) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) store_tile(P + ((long long)b * HG + (long long)h * G + mt * 16) * Cout + nt * 8, Cout, c[mt], lane);
    }
    if (PT16) {   // K-major operand of out = w.P: row c of image b holds this head's G slices contiguously
      __nv_bfloat16* pt = PT16 + ((long long)b * Cout + col) * HG + h * G;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        pt[mt * 16 + g] = __float2bfloat16_rn(c[mt][0]);
        pt[HG + mt * 16 + g] = __float2bfloat16_rn(c[mt][1]);
        pt[mt * 16 + g + 8] = __float2bfloat16_rn(c[mt][2]);
        pt[HG + mt * 16 + g + 8] = __float2bfloat16_rn(c[mt][3]);
      }
    }
    if (P16) {
      // weight operand of the deslice gradient dw = dOut.P^T.  The slice tokens of a head are all close to the field's mean,
      // so the rows P[g,:] share a large common component; the softmax backward annihilates anything constant over g
      // (dL' = w o (dw - sum_g w dw), sum_g w = 1), which would leave bf16 rounding noise of the common part over a small
      // signal.  Storing P[g,:] - mean_g P[g,:] removes it at the operand level: same gradient, ~8x less error.
      float m0 = 0.f, m1 = 0.f;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        m0 += c[mt][0] + c[mt][2];
        m1 += c[mt][1] + c[mt][3];
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        m0 += __shfl_xor_sync(0xffffffffu, m0, o);
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      }
      m0 *= 1.0f / (float)G;
      m1 *= 1.0f / (float)G;
      __nv_bfloat16* pr = P16 + ((long long)b * HG + (long long)h * G) * Cout + col;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        *reinterpret_cast<__nv_bfloat162*>(pr + (long long)(mt * 16 + g) * Cout) = __floats2bfloat162_rn(c[mt][0] - m0, c[mt][1] - m1);
        *reinterpret_cast<__nv_bfloat162*>(pr + (long long)(mt * 16 + g + 8) * Cout) = __floats2bfloat162_rn(c[mt][2] - m0, c[mt][3] - m1);
      }
    }
  }
}

template <int G>
struct TkBwdSmem {
  static constexpr int LA = G + 4;
  static constexpr int TOK = 0, Q = TOK + G * TK_LD, K = Q + G * TK_LD, V = K + G * TK_LD, DO = V + G * TK_LD;   // dO, later dtok
  static constexpr int A = DO + G * TK_LD;
  static constexpr int WQ = A + G * LA, WK = WQ + TK_D * TK_LD, WV = WK + TK_D * TK_LD;
  static constexpr int SS = WV + TK_D * TK_LD;
  static constexpr int OS = SS + G;                       // O [G][TK_LD]
  static constexpr int R0 = OS + G * TK_LD;               // region: { dP tile [G][Cout+4] | Wo slice [Cout][32] swizzled } then,
                                                          // once dO and dWo are done, { dA | dq | dk | dv }
  static constexpr int floats(int Cout) {
    const int a = G * (Cout + 4) + Cout * TK_D, b = G * LA + 3 * G * TK_LD;
    return R0 + (a > b ? a : b);
  }
};

// grid (H, B), block 256
template <int G>
__global__ void __launch_bounds__(TK_THREADS) token_attn_bwd_mma_kernel(
    const float* __restrict__ dP, const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
    const float* __restrict__ Wo, const float* __restrict__ s_in, const float* __restrict__ tok_in, const float* __restrict__ q_in,
    const float* __restrict__ k_in, const float* __restrict__ v_in, const float* __restrict__ A_in, const float* __restrict__ O_in,
    float* __restrict__ dTt, float* __restrict__ ds, float* __restrict__ dWqkv_part, float* __restrict__ dWo_part, int H, int Cout) {
  pdl_sync();
  using S = TkBwdSmem<G>;
  constexpr int D = TK_D, LD = TK_LD, LA = S::LA, GD = G * D, MT = G / 16;
  extern __shared__ float sm[];
  float* tok = sm + S::TOK;
  float* q = sm + S::Q;
  float* k = sm + S::K;
  float* v = sm + S::V;
  float* dO = sm + S::DO;
  float* A = sm + S::A;
  float* Wqs = sm + S::WQ;
  float* Wks = sm + S::WK;
  float* Wvs = sm + S::WV;
  float* ssum = sm + S::SS;
  float* Os = sm + S::OS;
  const int LP = Cout + 4;                 // dP tile row stride: conflict-free as the K-contiguous A operand of the long (K = Cout) contraction
  float* dPs = sm + S::R0;                 // [G][LP]
  float* Wos = dPs + G * LP;               // [Cout][32], swizzled (wo_sw)
  float* dA = sm + S::R0;                  // after the barrier below: [G][LA]
  float* dq = dA + G * LA;
  float* dk = dq + G * LD;
  float* dv = dk + G * LD;
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long bh = (long long)b * H + h;
  const int I = H * D;

  const float* dPh = dP + ((long long)b * H * G + (long long)h * G) * Cout;   // [G][Cout]
  for (int idx = tid; idx < G * (Cout / 4); idx += TK_THREADS) {
    const int g = idx / (Cout / 4), c4 = (idx - g * (Cout / 4)) * 4;
    *reinterpret_cast<float4*>(dPs + g * LP + c4) = *reinterpret_cast<const float4*>(dPh + (long long)g * Cout + c4);
  }
  for (int idx = tid; idx < Cout * (D / 4); idx += TK_THREADS) {
    const int c = idx >> 3, d4 = (idx & 7) * 4;
    *reinterpret_cast<float4*>(Wos + wo_sw(c, d4)) = *reinterpret_cast<const float4*>(Wo + (long long)c * I + h * D + d4);
  }
  for (int idx = tid; idx < G * (D / 4); idx += TK_THREADS) {
    const int g = idx >> 3, d4 = (idx & 7) * 4;
    const long long off = bh * GD + g * D + d4;
    *reinterpret_cast<float4*>(Os + g * LD + d4) = *reinterpret_cast<const float4*>(O_in + off);
    *reinterpret_cast<float4*>(tok + g * LD + d4) = *reinterpret_cast<const float4*>(tok_in + off);
    *reinterpret_cast<float4*>(q + g * LD + d4) = *reinterpret_cast<const float4*>(q_in + off);
    *reinterpret_cast<float4*>(k + g * LD + d4) = *reinterpret_cast<const float4*>(k_in + off);
    *reinterpret_cast<float4*>(v + g * LD + d4) = *reinterpret_cast<const float4*>(v_in + off);
  }
  for (int idx = tid; idx < G * (G / 4); idx += TK_THREADS) {
    const int g = idx / (G / 4), c4 = (idx - g * (G / 4)) * 4;
    *reinterpret_cast<float4*>(A + g * LA + c4) = *reinterpret_cast<const float4*>(A_in + bh * G * G + g * G + c4);
  }
  for (int idx = tid; idx < D * (D / 4); idx += TK_THREADS) {
    const int r = idx >> 3, d4 = (idx & 7) * 4;
    *reinterpret_cast<float4*>(Wqs + r * LD + d4) = *reinterpret_cast<const float4*>(Wq + r * D + d4);
    *reinterpret_cast<float4*>(Wks + r * LD + d4) = *reinterpret_cast<const float4*>(Wk + r * D + d4);
    *reinterpret_cast<float4*>(Wvs + r * LD + d4) = *reinterpret_cast<const float4*>(Wv + r * D + d4);
  }
  for (int o = tid; o < G; o += TK_THREADS) ssum[o] = s_in[bh * G + o];
  __syncthreads();

  // a. dO[g, d] = sum_c dP[g, c] Wo[c, hD + d]           A K-contiguous, B(k = c, n = d) N-contiguous (swizzled rows of 32)
  for (int tile = warp; tile < MT * 4; tile += TK_WARPS) {
    const int mt = tile >> 2, nt = tile & 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    const int g = lane >> 2, t = lane & 3;
    const float* Ap = dPs + mt * 16 * LP;
#pragma unroll 4
    for (int k0 = 0; k0 < Cout; k0 += 8) {
      float af[4], bf[2];
      af[0] = Ap[g * LP + k0 + t];
      af[1] = Ap[(g + 8) * LP + k0 + t];
      af[2] = Ap[g * LP + k0 + t + 4];
      af[3] = Ap[(g + 8) * LP + k0 + t + 4];
      bf[0] = Wos[wo_sw(k0 + t, nt * 8 + g)];
      bf[1] = Wos[wo_sw(k0 + t + 4, nt * 8 + g)];
      uint32_t ah[4], al[4], bh2[2], bl[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_tf32(af[i], ah[i], al[i]);
#pragma unroll
      for (int i = 0; i < 2; ++i) split_tf32(bf[i], bh2[i], bl[i]);
      mma_tf32(c, al, bh2);
      mma_tf32(c, ah, bl);
      mma_tf32(c, ah, bh2);
    }
    store_tile(dO + mt * 16 * LD + nt * 8, LD, c, lane);
  }
  // b. dWo_part[b, c, hD + d] = sum_g dP[g, c] O[g, d]    A(m = c, k = g) M-contiguous, B(k = g, n = d) N-contiguous
  for (int tile = warp; tile < (Cout / 16) * 4; tile += TK_WARPS) {
    const int mt = tile >> 2, nt = tile & 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    warp_mma<false, false>(c, dPs + mt * 16, LP, Os + nt * 8, LD, G, lane);
    store_tile(dWo_part + ((long long)b * Cout + mt * 16) * I + h * D + nt * 8, I, c, lane);
  }
  __syncthreads();   // dP tile and Wo slice are dead: the region becomes dA | dq | dk | dv
  // c. dA = dO v^T (B(k = d, n = g2) = v[g2][d], K-contiguous) ; d. dv = A^T dO (A(m = g2, k = g) M-contiguous)
  for (int job = warp; job < MT * (G / 8) + MT * 4; job += TK_WARPS) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (job < MT * (G / 8)) {
      const int mt = job / (G / 8), nt = job - mt * (G / 8);
      warp_mma<true, true>(c, dO + mt * 16 * LD, LD, v + nt * 8 * LD, LD, D, lane);
      store_tile(dA + mt * 16 * LA + nt * 8, LA, c, lane);
    } else {
      const int tile = job - MT * (G / 8), mt = tile >> 2, nt = tile & 3;
      warp_mma<false, false>(c, A + mt * 16, LA, dO + nt * 8, LD, G, lane);
      store_tile(dv + mt * 16 * LD + nt * 8, LD, c, lane);
    }
  }
  __syncthreads();
  // e. dS = A o (dA - rowsum(dA o A))   (in place in dA)
  for (int g = warp; g < G; g += TK_WARPS) {
    float r = 0.f;
    for (int g2 = lane; g2 < G; g2 += 32) r = fmaf(dA[g * LA + g2], A[g * LA + g2], r);
    r = warp_sum(r);
    for (int g2 = lane; g2 < G; g2 += 32) dA[g * LA + g2] = A[g * LA + g2] * (dA[g * LA + g2] - r);
  }
  __syncthreads();
  // f. dq = dS k * D^-1/2 (B(k = g2, n = d) N-contiguous) ; dk = dS^T q * D^-1/2 (A(m = g2, k = g) M-contiguous)
  const float scale = rsqrtf((float)D);
  for (int job = warp; job < 2 * MT * 4; job += TK_WARPS) {
    const int which = job / (MT * 4), tile = job - which * (MT * 4), mt = tile >> 2, nt = tile & 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (which == 0) {
      warp_mma<true, false>(c, dA + mt * 16 * LA, LA, k + nt * 8, LD, G, lane);
      store_tile(dq + mt * 16 * LD + nt * 8, LD, c, lane, scale);
    } else {
      warp_mma<false, false>(c, dA + mt * 16, LA, q + nt * 8, LD, G, lane);
      store_tile(dk + mt * 16 * LD + nt * 8, LD, c, lane, scale);
    }
  }
  __syncthreads();
  // g. dtok = dq Wq + dk Wk + dv Wv (B(k = i, n = j) = W[i][j], N-contiguous; over the dO array)
  // h. dW{q,k,v}_part[i][j] = sum_g d{q,k,v}[g, i] tok[g, j]   (A(m = i, k = g) M-contiguous, B(k = g, n = j) N-contiguous)
  float* dtok = dO;
  float* dWp = dWqkv_part + bh * 3 * D * D;
  for (int job = warp; job < MT * 4 + 3 * 2 * 4; job += TK_WARPS) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (job < MT * 4) {
      const int mt = job >> 2, nt = job & 3;
      warp_mma<true, false>(c, dq + mt * 16 * LD, LD, Wqs + nt * 8, LD, D, lane);
      warp_mma<true, false>(c, dk + mt * 16 * LD, LD, Wks + nt * 8, LD, D, lane);
      warp_mma<true, false>(c, dv + mt * 16 * LD, LD, Wvs + nt * 8, LD, D, lane);
      store_tile(dtok + mt * 16 * LD + nt * 8, LD, c, lane);
    } else {
      const int j2 = job - MT * 4, which = j2 >> 3, tile = j2 & 7, mt = tile >> 2, nt = tile & 3;
      const float* src = which == 0 ? dq : (which == 1 ? dk : dv);
      warp_mma<false, false>(c, src + mt * 16, LD, tok + nt * 8, LD, G, lane);
      store_tile(dWp + which * D * D + mt * 16 * D + nt * 8, D, c, lane);
    }
  }
  __syncthreads();
  // i. dTt = dtok / (s + eps) ; ds = -sum_d dtok tok / (s + eps)
  for (int o = tid; o < GD; o += TK_THREADS) {
    const int g = o >> 5, dd = o & 31;
    dTt[bh * GD + o] = dtok[g * LD + dd] / (ssum[g] + TK_EPS);
  }
  for (int g = warp; g < G; g += TK_WARPS) {
    const float pr = warp_sum(dtok[g * LD + lane] * tok[g * LD + lane]);
    if (lane == 0) ds[bh * G + g] = -pr / (ssum[g] + TK_EPS);
  }
}

static bool token_mma_enabled() {
  static const bool on = [] { const char* e = getenv("TBNS_TOKEN_MMA"); return !e || atoi(e) != 0; }();
  return on;
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_pa_token_attn_fwd(const float* part, int nchunk, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                                      float* s, float* Tt, float* tok, float* q, float* k, float* v, float* A, float* O, float* P,
                                      void* P16, void* PT16, int B, int H, int D, int G, int Cout, void* stream) {
  TBNS_REQUIRE(part && Wq && Wk && Wv && Wo && s && Tt && tok && q && k && v && A && O && (P || (P16 && PT16)),
               "tbns_pa_token_attn_fwd: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && G > 0 && Cout > 0 && nchunk > 0, "tbns_pa_token_attn_fwd: bad dims");
  const bool al = ((reinterpret_cast<uintptr_t>(Wo) | reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(Wk) |
                    reinterpret_cast<uintptr_t>(Wv) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                    reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(O) | reinterpret_cast<uintptr_t>(P) |
                    reinterpret_cast<uintptr_t>(P16)) & 15) == 0;
  if (token_mma_enabled() && D == TK_D && (G == 32 || G == 64) && Cout % 16 == 0 && Cout * TK_LD >= TK_WARPS * 33 * 32 && al &&
      H <= 65535 && B <= 65535) {
    const size_t smem = sizeof(float) * (size_t)(G == 32 ? TkFwdSmem<32>::floats(Cout) : TkFwdSmem<64>::floats(Cout));
    if (smem <= 227 * 1024) {
      dim3 grid(H, B);
      cudaStream_t st = (cudaStream_t)stream;
      __nv_bfloat16* p16 = reinterpret_cast<__nv_bfloat16*>(P16);
      __nv_bfloat16* pt16 = reinterpret_cast<__nv_bfloat16*>(PT16);
      if (G == 32) {
        TBNS_SMEM_OPT_IN((token_attn_fwd_mma_kernel<32>), 227 * 1024);
        TBNS_CUDA(launch_pdl(token_attn_fwd_mma_kernel<32>, dim3(grid), dim3(TK_THREADS), smem, st, part, nchunk, Wq, Wk, Wv, Wo, s, Tt, tok, q, k, v, A, O, P, p16, pt16, H, Cout));
      } else {
        TBNS_SMEM_OPT_IN((token_attn_fwd_mma_kernel<64>), 227 * 1024);
        TBNS_CUDA(launch_pdl(token_attn_fwd_mma_kernel<64>, dim3(grid), dim3(TK_THREADS), smem, st, part, nchunk, Wq, Wk, Wv, Wo, s, Tt, tok, q, k, v, A, O, P, p16, pt16, H, Cout));
      }
      TBNS_LAUNCH_CHECK();
      return TBNS_OK;
    }
  }
  TBNS_REQUIRE(P != nullptr, "tbns_pa_token_attn_fwd: this shape needs the fp32 P output buffer");
  return token_attn_fwd_simt(part, nchunk, Wq, Wk, Wv, Wo, s, Tt, tok, q, k, v, A, O, P, P16, PT16, B, H, D, G, Cout, stream);
}

extern "C" int tbns_pa_token_attn_bwd(const float* dP, const float* Wq, const float* Wk, const float* Wv, const float* Wo,
                                      const float* s, const float* tok, const float* q, const float* k, const float* v, const float* A,
                                      const float* O, float* dTt, float* ds, float* dWqkv_part, float* dWo_part, int B, int H, int D,
                                      int G, int Cout, void* stream) {
  TBNS_REQUIRE(dP && Wq && Wk && Wv && Wo && s && tok && q && k && v && A && O && dTt && ds && dWqkv_part && dWo_part,
               "tbns_pa_token_attn_bwd: null pointer");
  TBNS_REQUIRE(B > 0 && H > 0 && D > 0 && G > 0 && Cout > 0, "tbns_pa_token_attn_bwd: bad dims");
  const bool al = ((reinterpret_cast<uintptr_t>(dP) | reinterpret_cast<uintptr_t>(Wo) | reinterpret_cast<uintptr_t>(Wq) |
                    reinterpret_cast<uintptr_t>(Wk) | reinterpret_cast<uintptr_t>(Wv) | reinterpret_cast<uintptr_t>(tok) |
                    reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                    reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(O) | reinterpret_cast<uintptr_t>(dWqkv_part) |
                    reinterpret_cast<uintptr_t>(dWo_part)) & 15) == 0;
  if (token_mma_enabled() && D == TK_D && (G == 32 || G == 64) && Cout % 16 == 0 && al && H <= 65535 && B <= 65535) {
    const size_t smem = sizeof(float) * (size_t)(G == 32 ? TkBwdSmem<32>::floats(Cout) : TkBwdSmem<64>::floats(Cout));
    if (smem <= 227 * 1024) {
      dim3 grid(H, B);
      cudaStream_t st = (cudaStream_t)stream;
      if (G == 32) {
        TBNS_SMEM_OPT_IN((token_attn_bwd_mma_kernel<32>), 227 * 1024);
        TBNS_CUDA(launch_pdl(token_attn_bwd_mma_kernel<32>, dim3(grid), dim3(TK_THREADS), smem, st, dP, Wq, Wk, Wv, Wo, s, tok, q, k, v, A, O, dTt, ds, dWqkv_part, dWo_part, H, Cout));
      } else {
        TBNS_SMEM_OPT_IN((token_attn_bwd_mma_kernel<64>), 227 * 1024);
        TBNS_CUDA(launch_pdl(token_attn_bwd_mma_kernel<64>, dim3(grid), dim3(TK_THREADS), smem, st, dP, Wq, Wk, Wv, Wo, s, tok, q, k, v, A, O, dTt, ds, dWqkv_part, dWo_part, H, Cout));
      }
      TBNS_LAUNCH_CHECK();
      return TBNS_OK;
    }
  }
  return token_attn_bwd_simt(dP, Wq, Wk, Wv, Wo, s, tok, q, k, v, A, O, dTt, ds, dWqkv_part, dWo_part, B, H, D, G, Cout, stream);
}
