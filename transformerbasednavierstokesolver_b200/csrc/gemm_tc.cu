// tcgen05 / TMA / TMEM implicit-GEMM kernel for the dense token contractions of Physics-Attention (sm_100a).
//
//   C[m, n] = sum_{tap, ci} A[shift(m, tap), ci] * W[n, tap*Cin + ci]  (+ bias[n])
//
// with A a bf16 NHWC activation tensor [Bimg, Hg, Wg, Cin] (== the reference's [B, N, C] token layout, so the
// reshape/permute/contiguous copies of model/Physics_Attention.py:90-97 disappear), W a bf16 K-major packed weight
// [N, taps*Cin] and fp32 accumulation in tensor memory.  taps = 9 gives the 3x3/pad-1 convolution pair
// in_project_x | in_project_fx as ONE GEMM (N = 2*inner_dim) — zero padding comes for free from TMA out-of-bounds
// fill; flip = 1 gives the transposed convolution of the backward pass (dgrad); taps = 1 is a plain Linear.
//
// Kernels in this file (all warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// the remaining warps = epilogue):
//   gemm_tc_persistent_kernel  one CTA per SM loops over 128 x BN tiles, TMEM accumulator double-buffered (short-K contractions)
//   gemm_tc_2cta_kernel        CTA pairs on 256 x 256 tiles, tcgen05.mma.cta_group::2 (the long-K projection conv / dgrad)
//   gemm_tc_wgrad_kernel       token contraction (weight gradients), MN-major operands, split-K
#include <stdlib.h>

#include "tc_common.cuh"

namespace tbns {

constexpr int TC_THREADS = 192;

struct TcParams {
  float* C;                 // fp32 output (or nullptr)
  long long ldc;
  __nv_bfloat16* C16;       // bf16 output (or nullptr)
  long long ldc16;
  const float* bias;        // [N]
  const float* residual;    // [rows, ldr]
  long long ldr;
  int act;                  // 0 none, 1 GELU(erf) (pre-activation -> aux_out), 2 multiply by GELU'(aux_in)
  float* aux_out;
  const float* aux_in;
  long long ldaux;
  int aux_bf16;             // aux_out / aux_in hold bf16 instead of fp32 (persistent kernel only)
  int Bimg, Hg, Wg, Cin, taps, flip;
  int BW, BH;               // the 128-token M tile is a BH x BW patch of the grid (BW*BH == 128)
  int tiles_w, tiles_h;
  int N, w_batched;
  int round_tf32;           // round the fp32 output to TF32 (round-to-nearest): it is the operand of tf32 MMAs downstream
  // fused LayerNorm of the OUTPUT rows (N == BN: a tile holds whole rows): the consumer's LayerNorm runs in this epilogue
  const float* ln_gamma;    // [N] or nullptr
  const float* ln_beta;     // [N]
  __nv_bfloat16* ln_out16;  // [rows, N] bf16: LN(C) - the TMA operand of the next contraction
  float* ln_mean;           // [rows]
  float* ln_rstd;           // [rows]
  float ln_eps;
};

struct WgParams {
  float* ws;                // [split][batch][Mtot][Nb] fp32 partials
  int Mtot, Ma, Nb;
  int Bimg, Hg, Wg, taps;
  int BW, BH, tiles_w, tiles_h;  // the 64-token K block is a BH x BW patch (BW*BH == 64)
  int m_chunks, n_chunks;
  int batched, split_k;
};

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

// shared prologue: barrier init (one thread), TMEM allocation (warp 1), returns the TMEM base address
template <int BN, int STAGES>
__device__ __forceinline__ uint32_t tc_prologue(uint32_t bar_full, uint32_t bar_empty, uint32_t bar_acc, volatile uint32_t* tmem_slot,
                                                const CUtensorMap* tmA, const CUtensorMap* tmB, int warp, int lane) {
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();   // set-up above touched no global data: it overlapped the previous kernel (common.cuh: PDL)
  return *tmem_slot;
}

template <int BN>
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent TN kernel: one CTA per SM loops over output tiles; the fp32 accumulator is double-buffered in TMEM
// (2 x BN columns) so the fused epilogue of tile i (8 warps: tcgen05.ld -> bias / GELU / GELU' / residual -> fp32 / bf16
// stores) overlaps the TMA + tcgen05.mma main loop of tile i+1.  Roles: warp 0 TMA producer, warp 1 TMEM allocator +
// MMA issuer, warps 2-17 epilogue (warp%4 = TMEM lane quadrant, (warp-2)/4 = which quarter of the 32-column chunks): the
// fused epilogue is instruction-bound (GELU, conversions, 2-3 output streams), so it gets as many warps as registers allow.
// ------------------------------------------------------------------------------------------------
constexpr int TCP_EPI_WARPS = 16;                     // 4 TMEM lane quadrants x 4 column slots
constexpr int TCP_THREADS = 64 + 32 * TCP_EPI_WARPS;

// GELU / GELU' for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16
// rounding applied to the result), one exp shared between erf and the Gaussian density.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  const float e = __expf(-z * z);
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * e;
  cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  pdf = e * 0.39894228040143267794f;   // exp(-x^2/2)/sqrt(2 pi)
}
__device__ __forceinline__ float gelu_fast(float x) {
  float c, d;
  gelu_parts(x, c, d);
  return x * c;
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float c, d;
  gelu_parts(x, c, d);
  return fmaf(x, d, c);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// tcgen05.ld 16x256b.x4: 16 TMEM lanes x 32 columns per warp instruction.  Thread t receives, for column group j = 0..3,
// v[4j+0..1] = (lane t/4,     columns 8j + 2(t%4) + {0,1})  and  v[4j+2..3] = (lane t/4 + 8, same columns)
// (layout probed on hardware: profiles/probes/tmem_ld_shapes.cu).  Four threads cover one 32-byte sector of a row, so the
// epilogue's global accesses touch 8 full sectors per warp request instead of 32 half-used ones (thread-per-row layout).
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// fused epilogue on the two 16x256b fragments of a warp's 32-lane quadrant: v[0] = lanes 0..15, v[1] = lanes 16..31 (layout
// above); grow[k] = global row of tile row (t/4 + 8k) or -1; n = first of the 32 columns; t = lane id.
// All global loads (GELU' input, residual) are issued before any arithmetic so their latencies overlap.
// EPI < 0: every option is a runtime test on TcParams.  EPI >= 0: a bit set of EPI_* flags fixed at compile time, so the
// short-K contractions (whose time is all epilogue) run straight-line code for exactly the options they use.
enum : int { EPI_BIAS = 1, EPI_GELU = 2, EPI_DGELU = 4, EPI_RES = 8, EPI_ROUND = 16, EPI_C = 32, EPI_C16 = 64, EPI_AUX16 = 128,
             EPI_AUXOUT = 256, EPI_DERIV = 512 /* the aux stream holds GELU'(pre) instead of pre (act 3 / 4) */,
             EPI_LN = 1024 /* LayerNorm of the output rows fused in (persistent kernel, N == BN) */ };
static int epi_code(const TcParams& p) {
  const bool gelu = p.act == 1 || p.act == 3, mul = p.act == 2 || p.act == 4;
  return (p.bias ? EPI_BIAS : 0) | (gelu ? EPI_GELU : 0) | (mul ? EPI_DGELU : 0) | (p.act >= 3 ? EPI_DERIV : 0) |
         (p.residual ? EPI_RES : 0) | (p.round_tf32 ? EPI_ROUND : 0) | (p.C ? EPI_C : 0) | (p.C16 ? EPI_C16 : 0) |
         (p.aux_bf16 ? EPI_AUX16 : 0) | (gelu && p.aux_out ? EPI_AUXOUT : 0) | (p.ln_gamma ? EPI_LN : 0);
}

// The fused epilogue works on HALF fragments (one tcgen05.ld 16x256b.x4 = 16 TMEM lanes x 32 columns): v[16] holds tile rows
// (t/4 + 8k) for the half's two k values, grow2[kk] = global row of tile row k0 + kk or -1, n = first of the 32 columns.
// tc_epi_load_half issues the global loads the half needs (residual, or the GELU' stream of dpre) into `ld`; the persistent
// kernel calls it one half AHEAD of tc_epi_half, so the loads of half h+1 are in flight while half h is converted and stored
// (the short-K contractions were bound by exactly this latency: ncu long_scoreboard on the first use of the residual).
template <int EPI>
__device__ __forceinline__ constexpr int tc_epi_act(int runtime_act) {
  return EPI < 0 ? runtime_act : ((EPI & EPI_GELU) ? ((EPI & EPI_DERIV) ? 3 : 1) : (EPI & EPI_DGELU) ? ((EPI & EPI_DERIV) ? 4 : 2) : 0);
}
template <int EPI>
__device__ __forceinline__ bool tc_epi_has_loads(const TcParams& pp) {
  const int act = tc_epi_act<EPI>(pp.act);
  return act == 2 || act == 4 || (EPI < 0 ? pp.residual != nullptr : (EPI & EPI_RES) != 0);
}

template <int EPI>
__device__ __forceinline__ void tc_epi_load_half(const TcParams& pp, const int* grow2, int n, int t, float2 (&ld)[2][4]) {
  const int act = tc_epi_act<EPI>(pp.act);
  const bool mulaux = act == 2 || act == 4;   // multiply by GELU'(aux_in) / by aux_in itself
  if (!tc_epi_has_loads<EPI>(pp)) return;
  const float* src = mulaux ? pp.aux_in : pp.residual;
  const long long lds = mulaux ? pp.ldaux : pp.ldr;
  const bool h16 = mulaux && (EPI < 0 ? pp.aux_bf16 != 0 : (EPI & EPI_AUX16) != 0);
  const int cb = n + 2 * (t & 3);
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (grow2[kk] < 0) {
        ld[kk][j] = make_float2(0.f, 0.f);
      } else if (h16) {
        ld[kk][j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(src) + (long long)grow2[kk] * lds + cb + 8 * j));
      } else {
        ld[kk][j] = *reinterpret_cast<const float2*>(src + (long long)grow2[kk] * lds + cb + 8 * j);
      }
    }
}

// EPI < 0: every option is a runtime test on TcParams.  EPI >= 0: a bit set of EPI_* flags fixed at compile time, so the
// short-K contractions (whose time is all epilogue) run straight-line code for exactly the options they use.
template <int EPI>
__device__ __forceinline__ void tc_epi_half(const TcParams& pp, float (&v)[16], const int* grow2, int n, int t,
                                            const float2 (&ld)[2][4], float* st1 = nullptr, float* st2 = nullptr) {
  // compile-time view of the options (constant-folded when EPI >= 0)
  struct {
    const float* bias; int act; float* aux_out; long long ldaux; const float* residual;
    float* C; long long ldc; __nv_bfloat16* C16; long long ldc16; int round_tf32; int aux_bf16;
  } p;
  constexpr bool GEN = EPI < 0;
  p.bias = (GEN || (EPI & EPI_BIAS)) ? pp.bias : nullptr;
  p.act = tc_epi_act<EPI>(pp.act);
  p.aux_out = (GEN || (EPI & EPI_AUXOUT)) ? pp.aux_out : nullptr;
  p.ldaux = pp.ldaux;
  p.residual = (GEN || (EPI & EPI_RES)) ? pp.residual : nullptr;
  p.C = (GEN || (EPI & EPI_C)) ? pp.C : nullptr;
  p.ldc = pp.ldc;
  p.C16 = (GEN || (EPI & EPI_C16)) ? pp.C16 : nullptr;
  p.ldc16 = pp.ldc16;
  p.round_tf32 = GEN ? pp.round_tf32 : ((EPI & EPI_ROUND) ? 1 : 0);
  p.aux_bf16 = GEN ? pp.aux_bf16 : ((EPI & EPI_AUX16) ? 1 : 0);
  constexpr bool K_BIAS = !GEN && (EPI & EPI_BIAS), K_RES = !GEN && (EPI & EPI_RES), K_C = !GEN && (EPI & EPI_C),
                 K_C16 = !GEN && (EPI & EPI_C16), K_AUXOUT = !GEN && (EPI & EPI_AUXOUT);
  const int cb = n + 2 * (t & 3);
  float2 bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    bias[j] = (K_BIAS || (GEN && p.bias)) ? *reinterpret_cast<const float2*>(p.bias + cb + 8 * j) : make_float2(0.f, 0.f);
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const long long gr = grow2[kk];   // 64-bit from here on: gr * ld overflows 32 bits
    const bool ok = gr >= 0;   // rows past the end of the tensor: no stores, but every lane stays in the warp shuffles below
    uint32_t mine[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cb + 8 * j;
      float x0 = v[4 * j + 2 * kk] + bias[j].x, x1 = v[4 * j + 2 * kk + 1] + bias[j].y;
      if (p.act == 1 || p.act == 3) {
        float c0, d0, c1, d1;
        gelu_parts(x0, c0, d0);
        gelu_parts(x1, c1, d1);
        if (ok && (K_AUXOUT || (GEN && p.aux_out))) {
          // side stream for backward: the pre-activation (act 1) or, cheaper for the consumer, GELU'(pre) itself (act 3)
          const float a0 = p.act == 3 ? fmaf(x0, d0, c0) : x0, a1 = p.act == 3 ? fmaf(x1, d1, c1) : x1;
          if (p.aux_bf16) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.aux_out) + gr * p.ldaux + c) = __floats2bfloat162_rn(a0, a1);
          else *reinterpret_cast<float2*>(p.aux_out + gr * p.ldaux + c) = make_float2(a0, a1);
        }
        x0 *= c0;
        x1 *= c1;
      } else if (p.act == 2) {
        x0 *= gelu_grad_fast(ld[kk][j].x);
        x1 *= gelu_grad_fast(ld[kk][j].y);
      } else if (p.act == 4) {
        x0 *= ld[kk][j].x;
        x1 *= ld[kk][j].y;
      } else if (K_RES || (GEN && p.residual)) {
        x0 += ld[kk][j].x;
        x1 += ld[kk][j].y;
      }
      if (p.round_tf32) {
        uint32_t u0, u1;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u0) : "f"(x0));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u1) : "f"(x1));
        x0 = __uint_as_float(u0);
        x1 = __uint_as_float(u1);
      }
      if (ok && (K_C || (GEN && p.C))) *reinterpret_cast<float2*>(p.C + gr * p.ldc + c) = make_float2(x0, x1);
      if (EPI >= 0 && (EPI & EPI_LN)) {   // row statistics of the values just written (fused LayerNorm, pass 1); the values
        st1[kk] += x0 + x1;              // themselves go back into the fragment: the caller parks them in TMEM for pass 2
        st2[kk] = fmaf(x0, x0, fmaf(x1, x1, st2[kk]));
        v[4 * j + 2 * kk] = x0;
        v[4 * j + 2 * kk + 1] = x1;
      }
      if (K_C16 || (GEN && p.C16)) {
        // bf16 pairs are only 4 bytes: trade pairs with the neighbouring lane so that every thread owns 4 consecutive columns
        // (8 bytes) of column group j or j+1 and four threads fill one 32-byte sector of the row
        __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
        mine[j & 1] = *reinterpret_cast<uint32_t*>(&h2);
        if (j & 1) {
          const bool even = (t & 1) == 0;
          const uint32_t got = __shfl_xor_sync(0xffffffffu, even ? mine[1] : mine[0], 1);
          uint2 o;
          o.x = even ? mine[0] : got;
          o.y = even ? got : mine[1];
          const int cc = even ? (c - 8) : (c - 2);   // even: group j-1 columns 2a..2a+3 ; odd: group j columns 2a-2..2a+1
          if (ok) *reinterpret_cast<uint2*>(p.C16 + gr * p.ldc16 + cc) = o;
        }
      }
    }
  }
}

// whole 32-lane quadrant fragment at once (CTA-pair kernel): v[0] = lanes 0..15, v[1] = lanes 16..31; all global loads are
// issued before any arithmetic so their latencies overlap
template <int EPI>
__device__ __forceinline__ void tc_epi_frag2(const TcParams& pp, float (&v)[2][16], const int (&grow)[4], int n, int t,
                                             float* st1 = nullptr, float* st2 = nullptr) {
  float2 ld[2][2][4];
  tc_epi_load_half<EPI>(pp, grow, n, t, ld[0]);
  tc_epi_load_half<EPI>(pp, grow + 2, n, t, ld[1]);
  tc_epi_half<EPI>(pp, v[0], grow, n, t, ld[0], st1, st2);
  tc_epi_half<EPI>(pp, v[1], grow + 2, n, t, ld[1], st1 ? st1 + 2 : nullptr, st2 ? st2 + 2 : nullptr);
}

// fused LayerNorm, pass 2: the post-epilogue values of one 32-column chunk come back from TMEM (fragment layout of
// tmem_ld_16x256b_x4), are normalised and stored as bf16 (same lane-pair exchange as the C16 path: four threads fill one
// 32-byte sector of a row)
__device__ __forceinline__ void tc_epi_ln_store(const TcParams& p, const float (&v)[2][16], const int (&grow)[4], const float (&mu)[4],
                                                const float (&rs)[4], int n, int t) {
  const int cb = n + 2 * (t & 3);
  float2 g[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    g[j] = *reinterpret_cast<const float2*>(p.ln_gamma + cb + 8 * j);
    b[j] = *reinterpret_cast<const float2*>(p.ln_beta + cb + 8 * j);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool ok = grow[k] >= 0;
    uint32_t mine[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = v[k >> 1][4 * j + 2 * (k & 1)], x1 = v[k >> 1][4 * j + 2 * (k & 1) + 1];
      const float y0 = fmaf((x0 - mu[k]) * rs[k], g[j].x, b[j].x), y1 = fmaf((x1 - mu[k]) * rs[k], g[j].y, b[j].y);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
      mine[j & 1] = *reinterpret_cast<uint32_t*>(&h2);
      if (j & 1) {
        const bool even = (t & 1) == 0;
        const uint32_t got = __shfl_xor_sync(0xffffffffu, even ? mine[1] : mine[0], 1);
        uint2 o;
        o.x = even ? mine[0] : got;
        o.y = even ? got : mine[1];
        const int c = cb + 8 * j;
        const int cc = even ? (c - 8) : (c - 2);
        if (ok) *reinterpret_cast<uint2*>(p.ln_out16 + (long long)grow[k] * p.N + cc) = o;
      }
    }
  }
}

template <int BN, int STAGES>
struct TcpSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int NBAR = 2 * STAGES + 4;  // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
  static constexpr int LN_OFF = BAR_OFF + NBAR * 8 + 16;        // fused LayerNorm: [2 tile parities][4 column slots][128 rows] float2
  static constexpr int LN_BYTES = 2 * 4 * TC_BM * 8;
  static constexpr int TOTAL = LN_OFF + LN_BYTES + 1024;
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p,
                          int total_tiles) {
  using S = TcpSmem<BN, STAGES>;
  constexpr int TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_full = base + S::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_accf = bar_empty + STAGES * 8;   // [2] accumulator ready (MMA -> epilogue)
  const uint32_t bar_acce = bar_accf + 16;            // [2] accumulator drained (epilogue -> MMA)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR_OFF + S::NBAR * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.N / BN;
  const int kc_per_tap = p.Cin / TC_BK;
  const int nkb = p.taps * kc_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf + a * 8, 1);
      mbar_init(bar_acce + a * 8, TCP_THREADS - 64);  // every epilogue thread arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;   // running k-block counter across tiles -> ring slot / phase
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles, mt = tile / n_tiles;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, bimg = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, n0 = nt * BN;
        const int wb = p.w_batched ? bimg : 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + s * 8, ph ^ 1);
          const int tap = kb / kc_per_tap, kc = kb - tap * kc_per_tap;
          int dy = 0, dx = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
            if (p.flip) { dy = -dy; dx = -dx; }
          }
          const uint32_t sa = base + s * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          mbar_expect_tx(bar_full + s * 8, S::STAGE_BYTES);
          tma_load_4d(sa, &tmA, bar_full + s * 8, kc * TC_BK, w0 + dx, h0 + dy, bimg);
          tma_load_3d(sb, &tmB, bar_full + s * 8, kb * TC_BK, n0, wb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(BN);
      uint32_t it = 0, j = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
        const uint32_t as = j & 1;
        mbar_wait(bar_acce + as * 8, ((j >> 1) & 1) ^ 1);   // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t dtmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_full + s * 8, ph);
          tc_fence_after();
          const uint32_t sa = base + s * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UK; ++k) {
            const uint64_t ad = umma_desc_kmajor_sw128(sa + k * TC_UK * 2);
            const uint64_t bd = umma_desc_kmajor_sw128(sb + k * TC_UK * 2);
            umma_bf16(dtmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(bar_empty + s * 8);
        }
        umma_commit(bar_accf + as * 8);
      }
    }
  } else {
    // ===== epilogue: warps 2..17 =====
    const int q = warp & 3;
    const int slot = (warp - 2) >> 2;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
      const int nt = tile % n_tiles, mt = tile / n_tiles;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, bimg = mt / (p.tiles_w * p.tiles_h);
      const int n0 = nt * BN;
      // the four tile rows this thread touches: q*32 + lane/4 + 8k
      int grow[4];   // global row index (< 2^31, checked on the host) or -1
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = q * 32 + (lane >> 2) + 8 * k;
        const int hh = r / p.BW, ww = r - hh * p.BW;
        const int h = th * p.BH + hh, w = tw * p.BW + ww;
        grow[k] = (h < p.Hg && w < p.Wg) ? (bimg * p.Hg + h) * p.Wg + w : -1;
      }
      const uint32_t as = j & 1;
      constexpr bool LN = EPI >= 0 && (EPI & EPI_LN);
      constexpr int SL = TCP_EPI_WARPS / 4;                       // column slots
      constexpr int NCH = BN / (32 * SL) > 0 ? BN / (32 * SL) : 1;   // 32-column chunks per warp
      float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
      if (slot * 32 < BN) {
        // half-steps hs = 2 * chunk + half; the global loads of half-step hs + 1 are issued before half-step hs is processed
        const bool has_ld = tc_epi_has_loads<EPI>(p);
        float2 ld[2][2][4];
        if (has_ld) tc_epi_load_half<EPI>(p, grow, n0 + slot * 32, lane, ld[0]);   // in flight while the accumulator completes
        mbar_wait(bar_accf + as * 8, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int hs = 0; hs < 2 * NCH; ++hs) {
          const int c0 = slot * 32 + (hs >> 1) * 32 * SL, half = hs & 1;
          if (has_ld && hs + 1 < 2 * NCH)
            tc_epi_load_half<EPI>(p, grow + 2 * ((hs + 1) & 1), n0 + slot * 32 + ((hs + 1) >> 1) * 32 * SL, lane, ld[(hs + 1) & 1]);
          float v[16];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32 + 16 * half) << 16) + (uint32_t)(as * BN + c0);
          tmem_ld_16x256b_x4(ta, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (!LN && hs + 1 == 2 * NCH) {
            // last TMEM read of this tile by this thread: hand the accumulator buffer back to the MMA warp
            tc_fence_before();
            mbar_arrive(bar_acce + as * 8);
          }
          tc_epi_half<EPI>(p, v, grow + 2 * half, n0 + c0, lane, ld[hs & 1], st1 + 2 * half, st2 + 2 * half);
          if (LN) tmem_st_16x256b_x4(ta, v);   // park the post-epilogue values where the accumulator was: pass 2 reads them back
        }
      } else {
        mbar_wait(bar_accf + as * 8, (j >> 1) & 1);
        tc_fence_after();
        if (!LN) mbar_arrive(bar_acce + as * 8);   // narrow tiles: this warp's column slot does not exist
      }
      if (LN) {
        // fused LayerNorm of the output rows (the tile holds whole rows: N == BN).  Pass 1 above accumulated sum / sum of
        // squares of this thread's columns; combine the four lanes that share a row, then the four column-slot warps through
        // shared memory (double-buffered by tile parity: one named barrier per tile), then normalise the parked values.
        float2* red = reinterpret_cast<float2*>(gen + S::LN_OFF) + (j & 1) * (4 * TC_BM);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          st1[k] += __shfl_xor_sync(0xffffffffu, st1[k], 1);
          st2[k] += __shfl_xor_sync(0xffffffffu, st2[k], 1);
          st1[k] += __shfl_xor_sync(0xffffffffu, st1[k], 2);
          st2[k] += __shfl_xor_sync(0xffffffffu, st2[k], 2);
          if ((lane & 3) == 0) red[slot * TC_BM + q * 32 + (lane >> 2) + 8 * k] = make_float2(st1[k], st2[k]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(TCP_THREADS - 64) : "memory");
        float mu[4], rs[4];
        const float invN = 1.0f / (float)BN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = q * 32 + (lane >> 2) + 8 * k;
          float a = 0.f, b2 = 0.f;
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {   // fixed order: identical in all four slot warps
            const float2 pr = red[sl * TC_BM + r];
            a += pr.x;
            b2 += pr.y;
          }
          mu[k] = a * invN;
          rs[k] = rsqrtf(fmaxf(b2 * invN - mu[k] * mu[k], 0.f) + p.ln_eps);
          if (slot == 0 && (lane & 3) == 0 && grow[k] >= 0) {
            p.ln_mean[grow[k]] = mu[k];
            p.ln_rstd[grow[k]] = rs[k];
          }
        }
#pragma unroll 1
        for (int c0 = slot * 32; c0 < BN; c0 += 32 * (TCP_EPI_WARPS / 4)) {
          float v[2][16];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0);
          tmem_ld_16x256b_x4(ta, v[0]);
          tmem_ld_16x256b_x4(ta + (16u << 16), v[1]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_epi_ln_store(p, v, grow, mu, rs, n0 + c0, lane);
        }
        tc_fence_before();
        mbar_arrive(bar_acce + as * 8);   // the buffer (accumulator, then parked values) is free for the MMA of tile j + 2
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent TN kernel (long-K contractions: the 3x3 projection conv and its dgrad).
// Two CTAs of a cluster (one TPC) own a 256 x 256 output tile: each CTA stages ITS 128 token rows of A and HALF of the
// weight tile (128 of the 256 output channels) per k-block, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which
// reads both halves of B from the two shared memories, and each CTA finds its 128 x 256 accumulator rows in its own TMEM.
// Per CTA and k-block that is 32 KB of TMA traffic and 8 KB of operand reads per MMA instead of 48 KB / 12 KB, and the
// ring gets 6 stages instead of 4 in the same shared memory - the single-CTA kernel's tensor pipe sat at 77 % waiting for
// operands.  Protocol (CUTLASS sm100 2-SM pipelines): both producers signal the LEADER's full barrier (its expect_tx covers
// both CTAs' bytes); tcgen05.commit multicasts the slot-free / accumulator-ready arrivals to both CTAs; the epilogue warps of
// both CTAs hand the accumulator buffer back on the leader's barrier (remote arrive); cluster barriers fence set-up and
// teardown.
// ------------------------------------------------------------------------------------------------
constexpr int TC2_BN = 256;
constexpr int TC2_STAGES = 6;
struct Tc2Smem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;            // this CTA's 128 rows
  static constexpr int B_BYTES = (TC2_BN / 2) * TC_BK * 2;     // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = TC2_STAGES * STAGE_BYTES;
  static constexpr int NBAR = 2 * TC2_STAGES + 4;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 16 + 1024;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TCP_THREADS, 1)
gemm_tc_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p,
                    int total_pair_tiles, int m_tiles) {
  using S = Tc2Smem;
  constexpr int BN = TC2_BN, STAGES = TC2_STAGES;
  constexpr int TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_full = base + S::BAR_OFF;        // used in the leader CTA only
  const uint32_t bar_empty = bar_full + STAGES * 8;   // per CTA (multicast commit)
  const uint32_t bar_accf = bar_empty + STAGES * 8;   // [2] per CTA (multicast commit)
  const uint32_t bar_acce = bar_accf + 16;            // [2] used in the leader CTA only (both CTAs' epilogue warps arrive)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR_OFF + S::NBAR * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tiles = p.N / BN;
  const int kc_per_tap = p.Cin / TC_BK;
  const int nkb = p.taps * kc_per_tap;
  const int first = (int)cluster_id_x(), stride = (int)cluster_count_x();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf + a * 8, 1);
      mbar_init(bar_acce + a * 8, 2 * TCP_EPI_WARPS);   // one arrival per epilogue warp of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anything signals them
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own A rows + own half of B; bytes are counted on the leader's full barrier =====
      uint32_t it = 0;
      for (int tile = first; tile < total_pair_tiles; tile += stride) {
        const int nt = tile % n_tiles, mt = 2 * (tile / n_tiles) + (int)rank;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, bimg = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, n0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + s * 8, ph ^ 1);
          const int tap = kb / kc_per_tap, kc = kb - tap * kc_per_tap;
          int dy = 0, dx = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
            if (p.flip) { dy = -dy; dx = -dx; }
          }
          const uint32_t sa = base + s * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          const uint32_t full_leader = mapa_shared(bar_full + s * 8, 0);
          if (rank == 0) mbar_expect_tx(bar_full + s * 8, 2 * S::STAGE_BYTES);
          // an M tile past the end (odd tile count) has bimg == Bimg: the box is out of bounds and TMA fills zeros
          tma_load_4d_2sm(sa, &tmA, full_leader, kc * TC_BK, w0 + dx, h0 + dy, bimg);
          tma_load_3d_2sm(sb, &tmB, full_leader, kb * TC_BK, n0, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader CTA only) =====
      constexpr uint32_t idesc = umma_idesc_bf16_2sm(BN);
      uint32_t it = 0, j = 0;
      for (int tile = first; tile < total_pair_tiles; tile += stride, ++j) {
        const uint32_t as = j & 1;
        mbar_wait(bar_acce + as * 8, ((j >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained this accumulator buffer
        tc_fence_after();
        const uint32_t dtmem = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(bar_full + s * 8, ph);
          tc_fence_after();
          const uint32_t sa = base + s * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UK; ++k) {
            const uint64_t ad = umma_desc_kmajor_sw128(sa + k * TC_UK * 2);
            const uint64_t bd = umma_desc_kmajor_sw128(sb + k * TC_UK * 2);
            umma_bf16_2sm(dtmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(bar_empty + s * 8, 3);   // the slot is free in both CTAs
        }
        umma_commit_2sm(bar_accf + as * 8, 3);     // accumulator complete: both CTAs' epilogues
      }
    }
  } else {
    // ===== epilogue: warps 2..17 of both CTAs, each on its own 128 accumulator rows =====
    const int q = warp & 3;
    const int slot = (warp - 2) >> 2;
    const uint32_t acce_leader = mapa_shared(bar_acce, 0);
    uint32_t j = 0;
    for (int tile = first; tile < total_pair_tiles; tile += stride, ++j) {
      const int nt = tile % n_tiles, mt = 2 * (tile / n_tiles) + (int)rank;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, bimg = mt / (p.tiles_w * p.tiles_h);
      const int n0 = nt * BN;
      int grow[4];   // global row index (< 2^31, checked on the host) or -1
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = q * 32 + (lane >> 2) + 8 * k;
        const int hh = r / p.BW, ww = r - hh * p.BW;
        const int h = th * p.BH + hh, w = tw * p.BW + ww;
        grow[k] = (mt < m_tiles && h < p.Hg && w < p.Wg) ? (bimg * p.Hg + h) * p.Wg + w : -1;
      }
      const uint32_t as = j & 1;
      mbar_wait(bar_accf + as * 8, (j >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = slot * 32; c0 < BN; c0 += 32 * (TCP_EPI_WARPS / 4)) {
        float v[2][16];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0);
        tmem_ld_16x256b_x4(ta, v[0]);
        tmem_ld_16x256b_x4(ta + (16u << 16), v[1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 * (TCP_EPI_WARPS / 4) >= BN) {
          // last TMEM read of this tile by this warp: hand the buffer back to the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acce_leader + as * 8);
        }
        tc_epi_frag2<EPI>(p, v, grow, n0 + c0, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal or read it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// NT ("wgrad") kernel: D[(tap, a), n] = sum_token A[shift(token, tap), a] * B[token, n]
// Both operands are token-major in HBM, i.e. MN-major for the MMA: the SAME TMA boxes as above ([tokens][64 feat],
// 128B swizzle) are consumed through MN-major UMMA descriptors, so no transposed copy of activations is ever made.
// One CTA: 128 (a) x BN (n) output tile of one tap, over a contiguous range of 64-token K blocks (split-K);
// fp32 partials go to ws[split][batch][Mtot][Nb] and are reduced in a fixed order by gemm_splitk_reduce_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int WG_BKT = 64;  // tokens per K block

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgParams p) {
  using S = TcSmem<BN, STAGES>;  // A: 2 panels x [64 tok][64 a] = 16 KB ; B: BN/64 panels x 8 KB
  constexpr uint32_t PANEL = WG_BKT * 128;  // bytes of one [64 tokens][64 features] panel
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_full = base + S::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_acc = bar_empty + STAGES * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR_OFF + (2 * STAGES + 1) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  const int ni = tile % p.n_chunks;
  const int mi = (tile / p.n_chunks) % p.m_chunks;
  const int tap = tile / (p.n_chunks * p.m_chunks);
  const int kbpi = p.tiles_w * p.tiles_h;  // K blocks per image
  int e0, e1, bidx, split;
  if (p.batched) {
    bidx = blockIdx.y / p.split_k;
    split = blockIdx.y - bidx * p.split_k;
    const int per = (kbpi + p.split_k - 1) / p.split_k;
    e0 = bidx * kbpi + min(kbpi, split * per);
    e1 = bidx * kbpi + min(kbpi, (split + 1) * per);
  } else {
    bidx = 0;
    split = blockIdx.y;
    const int total = p.Bimg * kbpi;
    const int per = (total + p.split_k - 1) / p.split_k;
    e0 = min(total, split * per);
    e1 = min(total, (split + 1) * per);
  }
  const int nkb = e1 - e0;
  int dy = 0, dx = 0;
  if (p.taps == 9) {
    dy = tap / 3 - 1;
    dx = tap % 3 - 1;
  }

  const uint32_t tmem_base = tc_prologue<BN, STAGES>(bar_full, bar_empty, bar_acc, tmem_slot, &tmA, &tmB, warp, lane);

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        const int e = e0 + i;
        const int b = e / kbpi, rem = e - b * kbpi;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int h0 = th * p.BH, w0 = tw * p.BW;
        const uint32_t sa = base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        mbar_expect_tx(bar_full + s * 8, S::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_4d(sa + j * PANEL, &tmA, bar_full + s * 8, mi * TC_BM + j * 64, w0 + dx, h0 + dy, b);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * PANEL, &tmB, bar_full + s * 8, ni * BN + j * 64, w0, h0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN, 1);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_BKT / TC_UK; ++k) {
          const uint64_t ad = umma_desc_mnmajor_sw128(sa + k * TC_UK * 128, PANEL);
          const uint64_t bd = umma_desc_mnmajor_sw128(sb + k * TC_UK * 128, PANEL);
          umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty + s * 8);
      }
      if (nkb > 0) umma_commit(bar_acc);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int batch = p.batched ? p.Bimg : 1;
    float* orow = p.ws + (((long long)split * batch + bidx) * p.Mtot + (long long)tap * p.Ma + mi * TC_BM + r) * p.Nb + ni * BN;
    if (nkb > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  tc_teardown<BN>(tmem_base, warp);
}

// CTA-pair variant of the token-contraction kernel: a cluster of two CTAs owns a 256 (a) x 256 (n) output tile of one tap.
// Each CTA stages its 128 a-features of A and its 128 n-features of B per 64-token k-block (32 KB instead of 48 KB, 6 ring
// stages instead of 4); the leader issues tcgen05.mma.cta_group::2 with MN-major descriptors; each CTA drains its own 128
// accumulator rows.  Same barrier protocol as gemm_tc_2cta_kernel, one tile per cluster (no accumulator hand-back).
constexpr int WG2_STAGES = 6;
struct Wg2Smem {
  static constexpr int A_BYTES = 2 * WG_BKT * 128;          // 2 panels [64 tok][64 a]
  static constexpr int B_BYTES = 2 * WG_BKT * 128;          // 2 panels [64 tok][64 n] (this CTA's half of the 256-wide tile)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = WG2_STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * WG2_STAGES + 1) * 8 + 16 + 1024;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc_wgrad2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgParams p) {
  using S = Wg2Smem;
  constexpr int STAGES = WG2_STAGES, BN = 256;
  constexpr uint32_t PANEL = WG_BKT * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_full = base + S::BAR_OFF;       // leader only
  const uint32_t bar_empty = bar_full + STAGES * 8;  // per CTA
  const uint32_t bar_acc = bar_empty + STAGES * 8;   // per CTA
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR_OFF + (2 * STAGES + 1) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  // grid.x = 2 * pair tiles (cluster = consecutive CTA pair), grid.y = batch * split_k
  const int tile = blockIdx.x >> 1;
  const int ni = tile % p.n_chunks;                  // 256-wide n chunk
  const int mi = (tile / p.n_chunks) % p.m_chunks;   // 256-tall a chunk (pair)
  const int tap = tile / (p.n_chunks * p.m_chunks);
  const int kbpi = p.tiles_w * p.tiles_h;
  int e0, e1, bidx, split;
  if (p.batched) {
    bidx = blockIdx.y / p.split_k;
    split = blockIdx.y - bidx * p.split_k;
    const int per = (kbpi + p.split_k - 1) / p.split_k;
    e0 = bidx * kbpi + min(kbpi, split * per);
    e1 = bidx * kbpi + min(kbpi, (split + 1) * per);
  } else {
    bidx = 0;
    split = blockIdx.y;
    const int total = p.Bimg * kbpi;
    const int per = (total + p.split_k - 1) / p.split_k;
    e0 = min(total, split * per);
    e1 = min(total, (split + 1) * per);
  }
  const int nkb = e1 - e0;
  int dy = 0, dx = 0;
  if (p.taps == 9) {
    dy = tap / 3 - 1;
    dx = tap % 3 - 1;
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int a0 = mi * 256 + (int)rank * 128, nn0 = ni * BN + (int)rank * 128;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        const int e = e0 + i;
        const int b = e / kbpi, rem = e - b * kbpi;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int h0 = th * p.BH, w0 = tw * p.BW;
        const uint32_t sa = base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        const uint32_t full_leader = mapa_shared(bar_full + s * 8, 0);
        if (rank == 0) mbar_expect_tx(bar_full + s * 8, 2 * S::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_4d_2sm(sa + j * PANEL, &tmA, full_leader, a0 + j * 64, w0 + dx, h0 + dy, b);
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_4d_2sm(sb + j * PANEL, &tmB, full_leader, nn0 + j * 64, w0, h0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_2sm(BN) | (1u << 15) | (1u << 16);   // both operands MN-major
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_BKT / TC_UK; ++k) {
          const uint64_t ad = umma_desc_mnmajor_sw128(sa + k * TC_UK * 128, PANEL);
          const uint64_t bd = umma_desc_mnmajor_sw128(sb + k * TC_UK * 128, PANEL);
          umma_bf16_2sm(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit_2sm(bar_empty + s * 8, 3);
      }
      if (nkb > 0) umma_commit_2sm(bar_acc, 3);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int batch = p.batched ? p.Bimg : 1;
    float* orow = p.ws + (((long long)split * batch + bidx) * p.Mtot + (long long)tap * p.Ma + mi * 256 + (int)rank * 128 + r) * p.Nb + ni * BN;
    if (nkb > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  if (i + 7 < n) {
    const float4 a = *reinterpret_cast<const float4*>(in + i);
    const float4 b = *reinterpret_cast<const float4*>(in + i + 4);
    __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w), __floats2bfloat162_rn(b.x, b.y),
                           __floats2bfloat162_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<uint4*>(o);
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

// NHWC activation map [Bimg, Hg, Wg, F] with a {64, BW, BH, 1} box
static int encode_act(CUtensorMap* m, const void* ptr, int Bimg, int Hg, int Wg, int F, int BW, int BH) {
  cuuint64_t dims[4] = {(cuuint64_t)F, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)Bimg};
  cuuint64_t str[3] = {(cuuint64_t)F * 2, (cuuint64_t)Wg * F * 2, (cuuint64_t)Hg * Wg * F * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)BW, (cuuint32_t)BH, 1u};
  return encode_bf16(m, ptr, 4, dims, str, box);
}

// choose the BH x BW token patch (BW*BH == tokens) that wastes the fewest rows
static int pick_bw(int Hg, int Wg, int tokens) {
  int best = tokens;
  double bestU = -1.0;
  for (int bw = (tokens >= 128 ? 8 : 4); bw <= tokens; bw *= 2) {
    const int bh = tokens / bw;
    const double u = (double)Wg * Hg / ((double)cdiv(Wg, bw) * bw * (double)cdiv(Hg, bh) * bh);
    if (u > bestU + 1e-9) { bestU = u; best = bw; }
  }
  return best;
}

template <int BN, int STAGES, int EPI>
static int launch_tcp_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int m_tiles, cudaStream_t st) {
  using S = TcpSmem<BN, STAGES>;
  TBNS_SMEM_OPT_IN((gemm_tc_persistent_kernel<BN, STAGES, EPI>), S::TOTAL);
  const int sms = sm_count();
  const int total = m_tiles * (p.N / BN);
  const int grid = total < sms ? total : sms;
  TBNS_CUDA(launch_pdl(gemm_tc_persistent_kernel<BN, STAGES, EPI>, dim3(grid), dim3(TCP_THREADS), S::TOTAL, st, tmA, tmB, p, total));
  return TBNS_OK;
}

// the option sets of the Transolver block's contractions get straight-line epilogues; anything else takes the generic one
template <int BN, int STAGES>
static int launch_tcp(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int m_tiles, cudaStream_t st) {
  static const bool generic_only = [] { const char* e = getenv("TBNS_TC_EPI_GENERIC"); return e && e[0] == '1'; }();
  if (!generic_only) {
    switch (epi_code(p)) {
#define TBNS_EPI_CASE(code) case (code): return launch_tcp_epi<BN, STAGES, (code)>(tmA, tmB, p, m_tiles, st);
      TBNS_EPI_CASE(EPI_C)                                                          // dgrad, dx2
      TBNS_EPI_CASE(EPI_BIAS | EPI_C)                                               // projection fprop
      TBNS_EPI_CASE(EPI_BIAS | EPI_C16)                                             // projection fprop feeding the bf16 slice stage
      TBNS_EPI_CASE(EPI_BIAS | EPI_RES | EPI_C)                                     // deslice (+) to_out + residual, fc2 + residual
      TBNS_EPI_CASE(EPI_BIAS | EPI_RES | EPI_C | EPI_LN)                            // ... + the next stage's LayerNorm
      TBNS_EPI_CASE(EPI_BIAS | EPI_C | EPI_LN)                                      // preprocess fc2 + the first block's ln_1
      TBNS_EPI_CASE(EPI_C | EPI_C16)                                                // dw (fp32 + bf16)
      TBNS_EPI_CASE(EPI_C16)                                                        // dw (bf16 only)
      TBNS_EPI_CASE(EPI_BIAS | EPI_GELU | EPI_AUXOUT | EPI_AUX16 | EPI_C16)         // fc1: pre (bf16) + gelu (bf16)
      TBNS_EPI_CASE(EPI_DGELU | EPI_AUX16 | EPI_C16)                                // dpre = (dy W2) * gelu'(pre)
      TBNS_EPI_CASE(EPI_BIAS | EPI_GELU | EPI_AUXOUT | EPI_AUX16 | EPI_C16 | EPI_DERIV)   // fc1: gelu'(pre) (bf16) + gelu (bf16)
      TBNS_EPI_CASE(EPI_DGELU | EPI_AUX16 | EPI_C16 | EPI_DERIV)                    // dpre = (dy W2) * stored gelu' 
#undef TBNS_EPI_CASE
      default: break;
    }
  }
  if (p.ln_gamma) {
    set_error("tbns_gemm_tc: the fused LayerNorm exists for the option sets bias(+residual)+fp32 output only");
    return TBNS_ERR_UNSUPPORTED;
  }
  return launch_tcp_epi<BN, STAGES, -1>(tmA, tmB, p, m_tiles, st);
}

template <int EPI>
static int launch_tc2_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int m_tiles, cudaStream_t st) {
  TBNS_SMEM_OPT_IN((gemm_tc_2cta_kernel<EPI>), Tc2Smem::TOTAL);
  const int pairs = (m_tiles + 1) / 2 * (p.N / TC2_BN);
  int clusters = sm_count() / 2;
  if (pairs < clusters) clusters = pairs;
  TBNS_CUDA(launch_pdl(gemm_tc_2cta_kernel<EPI>, dim3(2 * clusters), dim3(TCP_THREADS), Tc2Smem::TOTAL, st, tmA, tmB, p, pairs, m_tiles));
  return TBNS_OK;
}

// CTA-pair kernel: the long-K option sets only (projection fprop / dgrad); everything else stays on the single-CTA kernel
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int m_tiles, cudaStream_t st, bool* handled) {
  *handled = true;
  switch (epi_code(p)) {
    case (EPI_C): return launch_tc2_epi<EPI_C>(tmA, tmB, p, m_tiles, st);
    case (EPI_C16): return launch_tc2_epi<EPI_C16>(tmA, tmB, p, m_tiles, st);
    case (EPI_BIAS | EPI_C): return launch_tc2_epi<EPI_BIAS | EPI_C>(tmA, tmB, p, m_tiles, st);
    case (EPI_BIAS | EPI_C16): return launch_tc2_epi<EPI_BIAS | EPI_C16>(tmA, tmB, p, m_tiles, st);
    default: break;
  }
  *handled = false;
  return TBNS_OK;
}

template <int BN, int STAGES>
static int launch_wg(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgParams& p, int tiles, int gy, cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  TBNS_SMEM_OPT_IN((gemm_tc_wgrad_kernel<BN, STAGES>), S::TOTAL);
  dim3 grid(tiles, gy);
  TBNS_CUDA(launch_pdl(gemm_tc_wgrad_kernel<BN, STAGES>, grid, dim3(TC_THREADS), S::TOTAL, st, tmA, tmB, p));
  return TBNS_OK;
}

static bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_cast_bf16(const float* in, void* out, long long n, void* stream) {
  TBNS_REQUIRE(in && out && n >= 0, "tbns_cast_bf16: bad args");
  if (n == 0) return TBNS_OK;
  TBNS_REQUIRE(al16p(in) && al16p(out), "tbns_cast_bf16: unaligned");
  const long long threads = (n + 7) / 8;
  TBNS_CUDA(launch_pdl(cast_bf16_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, in, reinterpret_cast<__nv_bfloat16*>(out), n));
  return TBNS_OK;
}

extern "C" int tbns_gemm_tc_supported(int Cin, int N, int taps) {
  return (Cin > 0 && Cin % TC_BK == 0 && N >= 64 && N % 64 == 0 && (taps == 1 || taps == 9)) ? 1 : 0;
}

extern "C" int tbns_gemm_tc(const tbns_tc_desc* dp, void* stream) {
  TBNS_REQUIRE(dp != nullptr, "tbns_gemm_tc: null descriptor");
  const tbns_tc_desc& d = *dp;
  TBNS_REQUIRE(d.A16 && d.W16 && (d.C || d.C16), "tbns_gemm_tc: null pointer");
  TBNS_REQUIRE(tbns_gemm_tc_supported(d.Cin, d.N, d.taps), "tbns_gemm_tc: unsupported shape Cin=%d N=%d taps=%d (need Cin%%64==0, N%%64==0)",
               d.Cin, d.N, d.taps);
  TBNS_REQUIRE(d.Bimg > 0 && d.Hg > 0 && d.Wg > 0 && (long long)d.Bimg * d.Hg * d.Wg < 0x7fffffffLL, "tbns_gemm_tc: bad dims");
  TBNS_REQUIRE(d.act >= 0 && d.act <= 4 && ((d.act != 2 && d.act != 4) || d.aux_in), "tbns_gemm_tc: bad activation spec");
  TBNS_REQUIRE(al16p(d.A16) && al16p(d.W16) && (!d.C || (al16p(d.C) && d.ldc % 4 == 0)) && (!d.C16 || (al16p(d.C16) && d.ldc16 % 8 == 0)) &&
                   (!d.bias || al16p(d.bias)) && (!d.residual || (al16p(d.residual) && d.ldr % 4 == 0)) &&
                   (!(d.aux_out || d.aux_in) || d.ldaux % 4 == 0) && (!d.aux_out || al16p(d.aux_out)) && (!d.aux_in || al16p(d.aux_in)),
               "tbns_gemm_tc: operands must be 16-byte aligned");
  TcParams p;
  p.C = d.C; p.ldc = d.ldc; p.C16 = reinterpret_cast<__nv_bfloat16*>(d.C16); p.ldc16 = d.ldc16;
  p.bias = d.bias; p.residual = d.residual; p.ldr = d.ldr;
  p.act = d.act; p.aux_out = d.aux_out; p.aux_in = d.aux_in; p.ldaux = d.ldaux; p.aux_bf16 = d.aux_bf16;
  p.Bimg = d.Bimg; p.Hg = d.Hg; p.Wg = d.Wg; p.Cin = d.Cin; p.taps = d.taps; p.flip = d.flip;
  p.BW = pick_bw(d.Hg, d.Wg, TC_BM); p.BH = TC_BM / p.BW;
  p.tiles_w = cdiv(d.Wg, p.BW); p.tiles_h = cdiv(d.Hg, p.BH);
  p.N = d.N; p.w_batched = d.w_batched; p.round_tf32 = d.round_tf32;
  p.ln_gamma = d.ln_gamma; p.ln_beta = d.ln_beta; p.ln_out16 = reinterpret_cast<__nv_bfloat16*>(d.ln_out16);
  p.ln_mean = d.ln_mean; p.ln_rstd = d.ln_rstd; p.ln_eps = d.ln_eps;
  if (d.ln_gamma) {
    TBNS_REQUIRE((d.N == 128 || d.N == 256) && d.C && d.ldc == d.N && d.ln_beta && d.ln_out16 && d.ln_mean && d.ln_rstd &&
                     al16p(d.ln_out16) && al16p(d.ln_gamma) && al16p(d.ln_beta),
                 "tbns_gemm_tc: fused LayerNorm needs N in {128, 256} (a tile holds whole rows), a dense fp32 output and all ln_* buffers");
  }
  const long long m_tiles = (long long)d.Bimg * p.tiles_w * p.tiles_h;
  TBNS_REQUIRE(m_tiles <= 0x7fffffffLL, "tbns_gemm_tc: too many tiles");

  CUtensorMap tmA, tmB;
  int rc = encode_act(&tmA, d.A16, d.Bimg, d.Hg, d.Wg, d.Cin, p.BW, p.BH);
  if (rc) return rc;
  const int N = d.N;
  const cuuint64_t K = (cuuint64_t)d.taps * d.Cin;
  cudaStream_t st = (cudaStream_t)stream;
  auto encode_w = [&](CUtensorMap* m, int rows_per_box) {
    cuuint64_t dims[3] = {K, (cuuint64_t)N, (cuuint64_t)(d.w_batched ? d.Bimg : 1)};
    cuuint64_t str[2] = {K * 2, K * 2 * (cuuint64_t)N};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)rows_per_box, 1u};
    return encode_bf16(m, d.W16, 3, dims, str, box);
  };
  static const bool two_cta = [] { const char* e = getenv("TBNS_TC_2CTA"); return !e || atoi(e) != 0; }();
  if (two_cta && N % TC2_BN == 0 && !d.w_batched && K / TC_BK >= 16 && m_tiles >= 2) {
    // long K (the 3x3 projections and their dgrad): CTA pairs, tcgen05.mma.cta_group::2; each CTA loads half of the weight tile
    rc = encode_w(&tmB, TC2_BN / 2);
    if (rc) return rc;
    bool handled = false;
    rc = launch_tc2(tmA, tmB, p, (int)m_tiles, st, &handled);
    if (handled) return rc;
  }
  // persistent single-CTA kernel: one CTA per SM, double-buffered TMEM accumulator, 16 epilogue warps.  Short-K contractions
  // (Linear / deslice / MLP: the fused epilogue outweighs the main loop) may run 128-wide tiles (TBNS_TC_SHORTK_BN128=1).
  static const bool bn128 = [] { const char* e = getenv("TBNS_TC_SHORTK_BN128"); return e && atoi(e) != 0; }();
  const bool short_k = K / TC_BK <= 8 && N % 128 == 0;
  // small M (one literal / unrolled model call): 256-wide tiles would leave SMs idle, 128-wide ones double the CTA count
  // (not with the fused LayerNorm, whose tile must hold whole rows)
  const bool few_tiles = short_k && !d.ln_gamma && m_tiles * (N / 256) < sm_count();
  const int BN = (N % 256 == 0 && !(short_k && (bn128 || few_tiles))) ? 256 : (N % 128 == 0 ? 128 : 64);
  rc = encode_w(&tmB, BN);
  if (rc) return rc;
  if (BN == 256) return launch_tcp<256, 4>(tmA, tmB, p, (int)m_tiles, st);
  if (BN == 128) return launch_tcp<128, 6>(tmA, tmB, p, (int)m_tiles, st);
  return launch_tcp<64, 8>(tmA, tmB, p, (int)m_tiles, st);
}

extern "C" int tbns_gemm_tc_wgrad_supported(int Ma, int Nb, int taps) {
  return (Ma > 0 && Ma % TC_BM == 0 && Nb >= 64 && Nb % 64 == 0 && (taps == 1 || taps == 9)) ? 1 : 0;
}

extern "C" int tbns_gemm_tc_wgrad(const tbns_tc_wgrad_desc* dp, void* stream) {
  TBNS_REQUIRE(dp != nullptr, "tbns_gemm_tc_wgrad: null descriptor");
  const tbns_tc_wgrad_desc& d = *dp;
  TBNS_REQUIRE(d.A16 && d.B16 && d.ws, "tbns_gemm_tc_wgrad: null pointer");
  TBNS_REQUIRE(tbns_gemm_tc_wgrad_supported(d.Ma, d.Nb, d.taps), "tbns_gemm_tc_wgrad: unsupported shape Ma=%d Nb=%d taps=%d", d.Ma, d.Nb, d.taps);
  TBNS_REQUIRE(d.Bimg > 0 && d.Hg > 0 && d.Wg > 0 && d.split_k >= 1, "tbns_gemm_tc_wgrad: bad dims");
  TBNS_REQUIRE(d.scatter ? (d.Cx && d.Cfx && d.I > 0 && !d.batched) : (d.C != nullptr), "tbns_gemm_tc_wgrad: null output");
  TBNS_REQUIRE(al16p(d.A16) && al16p(d.B16) && al16p(d.ws), "tbns_gemm_tc_wgrad: operands must be 16-byte aligned");
  WgParams p;
  p.ws = d.ws; p.Ma = d.Ma; p.Nb = d.Nb; p.Mtot = d.taps * d.Ma;
  p.Bimg = d.Bimg; p.Hg = d.Hg; p.Wg = d.Wg; p.taps = d.taps;
  p.BW = pick_bw(d.Hg, d.Wg, WG_BKT); p.BH = WG_BKT / p.BW;
  p.tiles_w = cdiv(d.Wg, p.BW); p.tiles_h = cdiv(d.Hg, p.BH);
  const int BN = d.Nb % 256 == 0 ? 256 : (d.Nb % 128 == 0 ? 128 : 64);
  static const bool two_cta = [] { const char* e = getenv("TBNS_TC_2CTA"); return !e || atoi(e) != 0; }();
  const bool pair = two_cta && BN == 256 && d.Ma % 256 == 0;   // CTA pairs on 256 x 256 tiles
  p.m_chunks = d.Ma / (pair ? 256 : TC_BM); p.n_chunks = d.Nb / BN;
  p.batched = d.batched; p.split_k = d.split_k;
  const int batch = d.batched ? d.Bimg : 1;
  const int tiles = d.taps * p.m_chunks * p.n_chunks;
  const int gy = batch * d.split_k;
  TBNS_REQUIRE(gy <= 65535, "tbns_gemm_tc_wgrad: grid too large");
  CUtensorMap tmA, tmB;
  int rc = encode_act(&tmA, d.A16, d.Bimg, d.Hg, d.Wg, d.Ma, p.BW, p.BH);
  if (rc) return rc;
  rc = encode_act(&tmB, d.B16, d.Bimg, d.Hg, d.Wg, d.Nb, p.BW, p.BH);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (pair) {
    TBNS_SMEM_OPT_IN((gemm_tc_wgrad2_kernel), Wg2Smem::TOTAL);
    dim3 grid(2 * tiles, gy);
    rc = launch_pdl(gemm_tc_wgrad2_kernel, grid, dim3(TC_THREADS), Wg2Smem::TOTAL, st, tmA, tmB, p) == cudaSuccess ? TBNS_OK : TBNS_ERR_CUDA;
    if (rc) set_error("gemm_tc_wgrad2_kernel launch failed");
  } else if (BN == 256) rc = launch_wg<256, 4>(tmA, tmB, p, tiles, gy, st);
  else if (BN == 128) rc = launch_wg<128, 6>(tmA, tmB, p, tiles, gy, st);
  else rc = launch_wg<64, 8>(tmA, tmB, p, tiles, gy, st);
  if (rc) return rc;
  // fixed-order reduction over the splits (+ scatter into weight.grad layout)
  tbns_gemm_desc g;
  memset(&g, 0, sizeof(g));
  g.M = p.Mtot; g.N = d.Nb; g.K = 1; g.batch = batch;
  g.C = d.C; g.ldc = d.ldc; g.sC = d.sC;
  g.split_k = d.split_k; g.ws = d.ws;
  g.scatter = d.scatter; g.I = d.I; g.taps = d.taps; g.Cin = d.Ma; g.Cx = d.Cx; g.Cfx = d.Cfx;
  return splitk_reduce(g, st);
}
