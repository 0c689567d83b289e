// Operand gathers and the epilogue shared by the generic contraction kernels behind `tbns_gemm` (include/tbns.h):
// the exact fp32 SIMT engine (gemm_simt.cu) and the 3xTF32 tcgen05 engine (gemm_x3.cu).
#pragma once
#include "common.cuh"

namespace tbns {

// pointer to A(m,k) or nullptr for a structural zero (conv padding). Caller bounds-checks m,k.
__device__ __forceinline__ const float* a_src(const tbns_gemm_desc& d, const float* A, int m, int k) {
  if (d.conv_mode == 0) return d.a_kind == 0 ? A + (long long)m * d.lda + k : A + (long long)k * d.lda + m;
  int token = d.conv_mode == 1 ? m : k;
  int feat = d.conv_mode == 1 ? k : m;
  int tap = feat / d.Cin, ci = feat - tap * d.Cin;
  int dy = tap / 3 - 1, dx = tap % 3 - 1;
  if (d.flip) { dy = -dy; dx = -dx; }
  int hw = d.Hg * d.Wg;
  int b = token / hw, r = token - b * hw;
  int i = r / d.Wg, j = r - i * d.Wg;
  int ii = i + dy, jj = j + dx;
  if (ii < 0 || ii >= d.Hg || jj < 0 || jj >= d.Wg) return nullptr;
  return A + (long long)(b * hw + ii * d.Wg + jj) * d.lda + ci;
}
__device__ __forceinline__ const float* b_src(const tbns_gemm_desc& d, const float* B, int k, int n) {
  return d.b_kind == 0 ? B + (long long)n * d.ldb + k : B + (long long)k * d.ldb + n;
}

// KIND 0: elements (m, k..k+3) ; KIND 1: elements (m..m+3, k)
template <int KIND>
__device__ __forceinline__ float4 load_a4(const tbns_gemm_desc& d, const float* A, int m, int k, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m >= d.M || k >= d.K) return v;
  if (KIND == 0) {
    if (vec) {
      const float* p = a_src(d, A, m, k);
      if (p) v = *reinterpret_cast<const float4*>(p);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < d.K) {
          const float* p = a_src(d, A, m, k + j);
          if (p) (&v.x)[j] = *p;
        }
    }
  } else {
    if (vec && m + 3 < d.M) {
      const float* p = a_src(d, A, m, k);
      if (p) v = *reinterpret_cast<const float4*>(p);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (m + j < d.M) {
          const float* p = a_src(d, A, m + j, k);
          if (p) (&v.x)[j] = *p;
        }
    }
  }
  return v;
}
// KIND 0: elements (k..k+3, n) ; KIND 1: elements (k, n..n+3)
template <int KIND>
__device__ __forceinline__ float4 load_b4(const tbns_gemm_desc& d, const float* B, int k, int n, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n >= d.N || k >= d.K) return v;
  if (KIND == 0) {
    if (vec) {
      v = *reinterpret_cast<const float4*>(b_src(d, B, k, n));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < d.K) (&v.x)[j] = *b_src(d, B, k + j, n);
    }
  } else {
    if (vec && n + 3 < d.N) {
      v = *reinterpret_cast<const float4*>(b_src(d, B, k, n));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < d.N) (&v.x)[j] = *b_src(d, B, k, n + j);
    }
  }
  return v;
}

__device__ __forceinline__ void epi_store1(const tbns_gemm_desc& d, int bidx, int m, int n, float v) {
  if (d.bias) v += d.bias[n];
  if (d.act == 1) {
    if (d.aux_out) d.aux_out[bidx * d.sAux + (long long)m * d.ldaux + n] = v;
    v = gelu_erf(v);
  } else if (d.act == 2) {
    v *= gelu_erf_grad(d.aux_in[bidx * d.sAux + (long long)m * d.ldaux + n]);
  }
  if (d.residual) v += d.residual[bidx * d.sR + (long long)m * d.ldr + n];
  if (d.scatter) {
    int tap = m / d.Cin, ci = m - tap * d.Cin;
    float* base = n < d.I ? d.Cx : d.Cfx;
    int co = n < d.I ? n : n - d.I;
    base[((long long)co * d.Cin + ci) * d.taps + tap] = v;
  } else {
    d.C[bidx * d.sC + (long long)m * d.ldc + n] = v;
  }
}

// 4 consecutive n. vec => all of bias/aux/residual/C rows are 16B aligned and n+3 < N.
__device__ __forceinline__ void epi_store4(const tbns_gemm_desc& d, int bidx, int m, int n, float4 v, bool vec) {
  if (m >= d.M || n >= d.N) return;
  if (!vec || d.scatter) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < d.N) epi_store1(d, bidx, m, n + j, (&v.x)[j]);
    return;
  }
  if (d.bias) {
    float4 b = *reinterpret_cast<const float4*>(d.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (d.act == 1) {
    if (d.aux_out) *reinterpret_cast<float4*>(d.aux_out + bidx * d.sAux + (long long)m * d.ldaux + n) = v;
    v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
  } else if (d.act == 2) {
    float4 a = *reinterpret_cast<const float4*>(d.aux_in + bidx * d.sAux + (long long)m * d.ldaux + n);
    v.x *= gelu_erf_grad(a.x); v.y *= gelu_erf_grad(a.y); v.z *= gelu_erf_grad(a.z); v.w *= gelu_erf_grad(a.w);
  }
  if (d.residual) {
    float4 r = *reinterpret_cast<const float4*>(d.residual + bidx * d.sR + (long long)m * d.ldr + n);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  *reinterpret_cast<float4*>(d.C + bidx * d.sC + (long long)m * d.ldc + n) = v;
}

}  // namespace tbns
