// Generic fp32 SIMT GEMM with the gathers and epilogues of the Physics-Attention path.
//
// This is the exact-arithmetic engine of the path: fp32 FMA accumulation (north_star "fp32 mode",
// per-layer rel-L2 <= 1e-5) and, with precision == TBNS_PREC_BF16, operands rounded to bf16 on the way
// into shared memory so its results equal the tensor-core (tcgen05) path up to fp32 summation order.
// It also serves every contraction the tcgen05 kernels do not cover (odd shapes, wgrad scatter).
//
// Implements, depending on the descriptor (include/tbns.h):
//   conv 3x3 fprop / Linear projections   model/Physics_Attention.py:94-97, :36-39
//   deslice (+) to_out                    model/Physics_Attention.py:116-119, :55-57
//   MLP Linear/GELU/Linear                model/Transolver_Structured_Mesh_2D.py:26-37
//   and all dgrad / wgrad contractions of SURVEY.md §8 (a-bwd).
#include <stdlib.h>

#include "gemm_generic.cuh"

namespace tbns {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

template <int AK, int BKIND>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const tbns_gemm_desc d, int vecA, int vecB, int vecC) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int split = blockIdx.z % d.split_k;
  const int bidx = blockIdx.z / d.split_k;
  const float* A = d.A + bidx * d.sA;
  const float* B = d.B + bidx * d.sB;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = (d.K + BK - 1) / BK;
  const int per = (nk + d.split_k - 1) / d.split_k;
  const int kb0 = split * per;
  const int kb1 = min(nk, kb0 + per);
  const bool rb16 = d.precision == TBNS_PREC_BF16;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int kb) {
    const int k0 = kb * BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * NT;
      if (AK == 0) ra[i] = load_a4<0>(d, A, m0 + (idx >> 2), k0 + (idx & 3) * 4, vecA);
      else         ra[i] = load_a4<1>(d, A, m0 + (idx & 31) * 4, k0 + (idx >> 5), vecA);
      if (BKIND == 0) rb[i] = load_b4<0>(d, B, k0 + (idx & 3) * 4, n0 + (idx >> 2), vecB);
      else            rb[i] = load_b4<1>(d, B, k0 + (idx >> 5), n0 + (idx & 31) * 4, vecB);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * NT;
      float4 a = ra[i], b = rb[i];
      if (rb16) {
        a.x = round_bf16(a.x); a.y = round_bf16(a.y); a.z = round_bf16(a.z); a.w = round_bf16(a.w);
        b.x = round_bf16(b.x); b.y = round_bf16(b.y); b.z = round_bf16(b.z); b.w = round_bf16(b.w);
      }
      if (AK == 0) {
        const int row = idx >> 2, kc = (idx & 3) * 4;
        As[buf][kc + 0][row] = a.x; As[buf][kc + 1][row] = a.y; As[buf][kc + 2][row] = a.z; As[buf][kc + 3][row] = a.w;
      } else {
        *reinterpret_cast<float4*>(&As[buf][idx >> 5][(idx & 31) * 4]) = a;
      }
      if (BKIND == 0) {
        const int col = idx >> 2, kc = (idx & 3) * 4;
        Bs[buf][kc + 0][col] = b.x; Bs[buf][kc + 1][col] = b.y; Bs[buf][kc + 2][col] = b.z; Bs[buf][kc + 3][col] = b.w;
      } else {
        *reinterpret_cast<float4*>(&Bs[buf][idx >> 5][(idx & 31) * 4]) = b;
      }
    }
  };

  const int tx = tid & 15, ty = tid >> 4;  // thread owns rows {ty*4..+3, 64+ty*4..+3} x cols {tx*4..+3, 64+tx*4..+3}
  if (kb0 < kb1) {
    gload(kb0);
    sstore(0);
  }
  __syncthreads();
  for (int kb = kb0; kb < kb1; ++kb) {
    const int buf = (kb - kb0) & 1;
    if (kb + 1 < kb1) gload(kb + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < kb1) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      const float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (d.split_k > 1) {
        if (m < d.M && n < d.N) {
          float* p = d.ws + ((long long)(split * d.batch + bidx) * d.M + m) * d.N + n;
          if ((d.N & 3) == 0) {
            *reinterpret_cast<float4*>(p) = v;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < d.N) p[j] = (&v.x)[j];
          }
        }
      } else {
        epi_store4(d, bidx, m, n, v, vecC);
      }
    }
  }
}

// SL "split lanes" share one float4 of output: lane l adds partials s = l, l+SL, ... in ascending order and the lanes are
// combined in lane order through shared memory - a fixed summation order, so results are bit-reproducible.  SL = 8 turns the
// reduce of a 256x256 weight gradient with ~70 partials from 64 serial CTAs into 512 CTAs with 9 loads per thread.
template <int SL>
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const tbns_gemm_desc d, int vecC) {
  pdl_sync();
  constexpr int OUTS = 256 / SL;
  __shared__ float4 red[SL > 1 ? 256 : 1];
  const long long n4 = (d.N + 3) / 4;
  const long long total = (long long)d.batch * d.M * n4;
  const int ol = threadIdx.x % OUTS, sl = threadIdx.x / OUTS;
  for (long long t0 = blockIdx.x * (long long)OUTS; t0 < total; t0 += (long long)gridDim.x * OUTS) {
    const long long t = t0 + ol;
    const bool live = t < total;
    int nq = 0, m = 0, bidx = 0;
    if (live) {
      nq = (int)(t % n4);
      const long long rm = t / n4;
      m = (int)(rm % d.M);
      bidx = (int)(rm / d.M);
    }
    const int n = nq * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const long long sstride = (long long)d.batch * d.M * d.N;
      const float* p = d.ws + ((long long)bidx * d.M + m) * d.N + n + (long long)sl * sstride;
      if ((d.N & 3) == 0) {
#pragma unroll 4
        for (int s = sl; s < d.split_k; s += SL, p += SL * sstride) {
          const float4 u = *reinterpret_cast<const float4*>(p);
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
      } else {
        for (int s = sl; s < d.split_k; s += SL, p += SL * sstride) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < d.N) (&v.x)[j] += p[j];
        }
      }
    }
    if (SL > 1) {
      red[threadIdx.x] = v;
      __syncthreads();
      if (sl == 0) {
#pragma unroll
        for (int l = 1; l < SL; ++l) {
          const float4 u = red[l * OUTS + ol];
          v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
      }
      __syncthreads();
    }
    if (live && sl == 0) epi_store4(d, bidx, m, n, v, vecC);
  }
}

static void launch_splitk_reduce(const tbns_gemm_desc& d, int vecC, cudaStream_t st) {
  const long long total = (long long)d.batch * d.M * ((d.N + 3) / 4);
  if (d.split_k >= 16 && total <= (long long)sm_count() * 8 * 32) {
    int blocks = (int)((total + 31) / 32);
    (void)launch_pdl(gemm_splitk_reduce_kernel<8>, dim3(blocks), dim3(256), 0, st, d, vecC);   // the caller checks cudaGetLastError
  } else {
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    (void)launch_pdl(gemm_splitk_reduce_kernel<1>, dim3(blocks), dim3(256), 0, st, d, vecC);
  }
}

static bool fp32_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TBNS_FP32_TC");
    on = !(e && e[0] == '0');
  }
  return on != 0;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// fixed-order reduction of split-K partials ws[split][batch][M][N] followed by the descriptor's epilogue / scatter
int splitk_reduce(const tbns_gemm_desc& d, cudaStream_t st) {
  const int vecC = !d.scatter && al16(d.C) && (d.ldc % 4 == 0) && (d.sC % 4 == 0) && (d.N % 4 == 0) && (!d.bias || al16(d.bias)) &&
                   (!d.residual || (al16(d.residual) && d.ldr % 4 == 0 && d.sR % 4 == 0)) && !(d.aux_out || d.aux_in);
  launch_splitk_reduce(d, vecC, st);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_gemm(const tbns_gemm_desc* dp, void* stream) {
  TBNS_REQUIRE(dp != nullptr, "tbns_gemm: null descriptor");
  tbns_gemm_desc d = *dp;
  if (d.batch < 1) d.batch = 1;
  if (d.split_k < 1) d.split_k = 1;
  if (d.M == 0 || d.N == 0) return TBNS_OK;
  TBNS_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "tbns_gemm: bad dims M=%d N=%d K=%d", d.M, d.N, d.K);
  TBNS_REQUIRE(d.A && d.B, "tbns_gemm: null operand");
  TBNS_REQUIRE(d.scatter ? (d.Cx && d.Cfx && d.I > 0 && d.taps > 0 && d.Cin > 0) : (d.C != nullptr), "tbns_gemm: null output");
  TBNS_REQUIRE(d.a_kind == 0 || d.a_kind == 1, "tbns_gemm: a_kind");
  TBNS_REQUIRE(d.b_kind == 0 || d.b_kind == 1, "tbns_gemm: b_kind");
  TBNS_REQUIRE(d.conv_mode >= 0 && d.conv_mode <= 2, "tbns_gemm: conv_mode");
  if (d.conv_mode == 1) TBNS_REQUIRE(d.a_kind == 0 && d.K == 9 * d.Cin && d.M % (d.Hg * d.Wg) == 0 && d.batch == 1, "tbns_gemm: conv_mode 1 needs a_kind 0, K=9*Cin, M multiple of Hg*Wg");
  if (d.conv_mode == 2) TBNS_REQUIRE(d.a_kind == 1 && d.M == 9 * d.Cin && d.K % (d.Hg * d.Wg) == 0 && d.batch == 1, "tbns_gemm: conv_mode 2 needs a_kind 1, M=9*Cin, K multiple of Hg*Wg");
  TBNS_REQUIRE(d.split_k == 1 || d.ws != nullptr, "tbns_gemm: split_k needs a workspace");
  TBNS_REQUIRE(d.act != 2 || d.aux_in, "tbns_gemm: act 2 needs aux_in");

  const bool convA = d.conv_mode != 0;
  int vecA = al16(d.A) && (d.lda % 4 == 0) && (d.sA % 4 == 0) && (!convA || d.Cin % 4 == 0) && (d.a_kind == 1 || d.K % 4 == 0);
  int vecB = al16(d.B) && (d.ldb % 4 == 0) && (d.sB % 4 == 0) && (d.b_kind == 1 || d.K % 4 == 0);
  int vecC = !d.scatter && al16(d.C) && (d.ldc % 4 == 0) && (d.sC % 4 == 0) && (d.N % 4 == 0) &&
             (!d.bias || al16(d.bias)) && (!d.residual || (al16(d.residual) && d.ldr % 4 == 0 && d.sR % 4 == 0)) &&
             (!(d.aux_out || d.aux_in) || (d.ldaux % 4 == 0 && d.sAux % 4 == 0 && (!d.aux_out || al16(d.aux_out)) && (!d.aux_in || al16(d.aux_in))));
  if (d.split_k > 1) TBNS_REQUIRE(al16(d.ws), "tbns_gemm: workspace must be 16B aligned");

  dim3 grid(cdiv(d.N, BN), cdiv(d.M, BM), d.batch * d.split_k);
  TBNS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "tbns_gemm: grid too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // fp32 mode: 3xTF32 on tcgen05 (gemm_x3.cu) unless the tile would be mostly padding; TBNS_PREC_FP32_EXACT (or the
  // environment switch TBNS_FP32_TC=0) keeps the fp32 FMA engine below
  if (d.precision == TBNS_PREC_FP32 && fp32_tc_enabled() && d.M >= 64 && d.N >= 32 && d.K >= 32) {
    int rc = gemm_x3_launch(d, vecA, vecB, vecC, grid, st);
    if (rc != TBNS_OK) return rc;
    if (d.split_k > 1) {
      launch_splitk_reduce(d, vecC, st);
      TBNS_LAUNCH_CHECK();
    }
    return TBNS_OK;
  }
  if (d.a_kind == 0 && d.b_kind == 0) gemm_simt_kernel<0, 0><<<grid, NT, 0, st>>>(d, vecA, vecB, vecC);
  else if (d.a_kind == 0 && d.b_kind == 1) gemm_simt_kernel<0, 1><<<grid, NT, 0, st>>>(d, vecA, vecB, vecC);
  else if (d.a_kind == 1 && d.b_kind == 0) gemm_simt_kernel<1, 0><<<grid, NT, 0, st>>>(d, vecA, vecB, vecC);
  else gemm_simt_kernel<1, 1><<<grid, NT, 0, st>>>(d, vecA, vecB, vecC);
  TBNS_LAUNCH_CHECK();
  if (d.split_k > 1) {
    launch_splitk_reduce(d, vecC, st);
    TBNS_LAUNCH_CHECK();
  }
  return TBNS_OK;
}
