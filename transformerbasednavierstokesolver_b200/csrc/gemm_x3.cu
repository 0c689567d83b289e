// fp32 mode on the tensor cores: 3xTF32 split products on tcgen05 (north_star "fp32 mode", per-layer rel-L2 <= 1e-5).
//
// Every fp32 operand is split on its way into shared memory, a = hi + lo with hi = tf32(a) (cvt.rna) and lo = tf32(a - hi);
// the contraction is accumulated in TMEM (fp32) as  lo.hi + hi.lo + hi.hi  with three kind::tf32 tcgen05.mma chains per
// k-block.  The dropped lo.lo term and the rounding of lo are both <= 2^-22 relative to |a||b|: fp32-class results at
// about one sixth of the bf16 tensor rate instead of the SIMT FMA rate.
//
// The kernel serves the same descriptor as the SIMT engine (include/tbns.h `tbns_gemm_desc`: both operand orientations,
// the 3x3 gathers of conv fprop / dgrad / wgrad, batch, split-K, bias / GELU / residual / scatter epilogues), so every dense
// contraction of the path in fp32 mode runs on it:
//   conv 3x3 fprop / Linear projections   model/Physics_Attention.py:94-97, :36-39
//   deslice (+) to_out                    model/Physics_Attention.py:116-119, :55-57
//   MLP Linear/GELU/Linear                model/Transolver_Structured_Mesh_2D.py:26-37
//   and all dgrad / wgrad contractions of SURVEY.md §8 (a-bwd).
// Operands are gathered by the CTA's threads (coalesced 16-byte loads), not by TMA: the gathers are irregular and the split
// has to happen between global and shared memory anyway.  Shared-memory tiles are K-major, 128-byte swizzled.
#include "gemm_generic.cuh"
#include "tc_common.cuh"

namespace tbns {

namespace x3 {
constexpr int BM = 128, BN = 128, BK = 32, NT = 256;
constexpr int FLUSH = 2;                           // k-blocks accumulated in TMEM before the partial sum moves to registers
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KB: one operand part (hi or lo) of one k-block
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGE_BYTES + 1024 /* alignment slack */ + 64 /* barrier + TMEM slot */;
constexpr int TMEM_COLS = 2 * BN;                  // two accumulators, used alternately by consecutive chunks
}  // namespace x3

// a = hi + lo.  hi = a rounded to TF32 (nearest, ties away: add half an ulp to the magnitude bits and clear the 13 low bits -
// what cvt.rna.tf32.f32 computes, in two integer instructions instead of the ~8 the conversion expands to on sm_100);
// lo = a - hi is exact in fp32 and gets the same half-ulp bias: kind::tf32 ignores the 13 low bits of its operands, so the
// tensor core's truncation completes the rounding of lo.  (Inf/NaN inputs stay Inf/NaN: hi keeps the exponent.)
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
  lo = __uint_as_float(__float_as_uint(a - hi) + 0x1000u);
}

// KIND 0 (K contiguous in global memory): v = elements (row, k4*4 .. k4*4+3): one 16-byte chunk of the row
__device__ __forceinline__ void store_rowchunk(uint8_t* hi_tile, uint8_t* lo_tile, int row, int k4, const float4& v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  const uint32_t off = sw128_off(row, k4);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}
// KIND 1 (M / N contiguous in global memory): v = elements (row .. row+3, k)
__device__ __forceinline__ void store_colquad(uint8_t* hi_tile, uint8_t* lo_tile, int row, int k, const float4& v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float h, l;
    split_tf32((&v.x)[j], h, l);
    const uint32_t off = sw128_off(row + j, k >> 2) + (uint32_t)(k & 3) * 4u;
    *reinterpret_cast<float*>(hi_tile + off) = h;
    *reinterpret_cast<float*>(lo_tile + off) = l;
  }
}

// n / d and n % d for 0 <= n < 2^24, 0 < d: the float quotient is off by at most one
__device__ __forceinline__ void divmod24(int n, int d, float rd, int& q, int& r) {
  q = __float2int_rz(__int2float_rn(n) * rd);
  r = n - q * d;
  if (r < 0) { r += d; --q; }
  else if (r >= d) { r -= d; ++q; }
}

// FASTA: the 3x3 gathers with the index arithmetic hoisted out of the k loop (conv_mode 1 needs Cin % 32 == 0 so that a k-block
// lies inside one tap; conv_mode 2 needs < 2^24 tokens); every other case goes through the generic load_a4 / load_b4.
template <int AK, int BKIND, int FASTA>
__global__ void __launch_bounds__(x3::NT, 2) gemm_x3_kernel(const tbns_gemm_desc d, int vecA, int vecB, int vecC) {
  using namespace x3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tiles + STAGE_BYTES);   // "the MMAs of the last k-block have retired"
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.z % d.split_k;
  const int bidx = blockIdx.z / d.split_k;
  const float* A = d.A + bidx * d.sA;
  const float* B = d.B + bidx * d.sB;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = (d.K + BK - 1) / BK;
  const int per = (nk + d.split_k - 1) / d.split_k;
  const int kb0 = split * per;
  const int kb1 = min(nk, kb0 + per);
  const int nkb = max(0, kb1 - kb0);

  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // per-thread slots of a k-block: 4 float4 of A, 4 of B.
  //   KIND 0: slot i -> row = (tid >> 3) + 32*i, chunk = tid & 7: a quarter warp covers one 128-byte row (coalesced, and the
  //           eight swizzled 16-byte stores of a row hit eight different bank groups)
  //   KIND 1: rows r1..r1+3 with r1 = 4*((warp & 3)*8 + (lane & 7)); k is skewed across the lanes,
  //           k = (4*((warp >> 2)*4 + i) + 4*((lane & 7) >> 1) + (lane >> 3)) & 31, so that the 32 scalar stores of one
  //           instruction land in 32 different banks of the swizzled tile (a plain mapping would be 16-way conflicted)
  float4 ra[4], rb[4];
  auto k1 = [&](int i) { return (4 * ((warp >> 2) * 4 + i) + 4 * ((lane & 7) >> 1) + (lane >> 3)) & 31; };
  const int r1 = ((warp & 3) * 8 + (lane & 7)) * 4;
  const int row0 = tid >> 3, c4 = (tid & 7) * 4;

  // hoisted gather state (FASTA)
  const int hw = d.Hg * d.Wg;
  const float* arow[4];   // conv_mode 1: &A[token(row_i), 0] or nullptr beyond M
  int aij[4];             //              grid coordinates of the token, i << 16 | j
  int tap_dy = 0, tap_dx = 0, tap_ci = 0;       // conv_mode 2: tap offsets and channel of this thread's four rows
  bool a_live = true;
  float rhw = 0.f, rwg = 0.f;
  if (FASTA && AK == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + row0 + 32 * i;
      arow[i] = m < d.M ? A + (long long)m * d.lda : nullptr;
      const int r = m % hw, gi = r / d.Wg;
      aij[i] = (gi << 16) | (r - gi * d.Wg);
    }
  }
  if (FASTA && AK == 1) {
    const int m = m0 + r1;
    a_live = m < d.M;    // Cin % 4 == 0 and M = 9*Cin: a quad never straddles the edge
    const int tap = m / d.Cin;
    tap_ci = m - tap * d.Cin;
    tap_dy = tap / 3 - 1;
    tap_dx = tap % 3 - 1;
    rhw = 1.0f / (float)hw;
    rwg = 1.0f / (float)d.Wg;
  }
  auto gload = [&](int kb) {
    const int k0 = kb * BK;
    if (FASTA && AK == 0) {
      const int tap = k0 / d.Cin, ci = k0 - tap * d.Cin + c4;
      int dy = tap / 3 - 1, dx = tap % 3 - 1;
      if (d.flip) { dy = -dy; dx = -dx; }
      const long long shift = (long long)(dy * d.Wg + dx) * d.lda + ci;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ii = (aij[i] >> 16) + dy, jj = (aij[i] & 0xFFFF) + dx;
        const bool ok = arow[i] != nullptr && k0 < d.K && ii >= 0 && ii < d.Hg && jj >= 0 && jj < d.Wg;
        ra[i] = ok ? *reinterpret_cast<const float4*>(arow[i] + shift) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (FASTA && AK == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int token = k0 + k1(i);
        int b, r, ti, tj;
        divmod24(token, hw, rhw, b, r);
        divmod24(r, d.Wg, rwg, ti, tj);
        const int ii = ti + tap_dy, jj = tj + tap_dx;
        const bool ok = a_live && token < d.K && ii >= 0 && ii < d.Hg && jj >= 0 && jj < d.Wg;
        ra[i] = ok ? *reinterpret_cast<const float4*>(A + (long long)(token + tap_dy * d.Wg + tap_dx) * d.lda + tap_ci)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (AK == 0) ra[i] = load_a4<0>(d, A, m0 + row0 + 32 * i, k0 + c4, vecA);
        else         ra[i] = load_a4<1>(d, A, m0 + r1, k0 + k1(i), vecA);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (BKIND == 0) rb[i] = load_b4<0>(d, B, k0 + c4, n0 + row0 + 32 * i, vecB);
      else            rb[i] = load_b4<1>(d, B, k0 + k1(i), n0 + r1, vecB);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (AK == 0) store_rowchunk(tiles, tiles + TILE_BYTES, row0 + 32 * i, tid & 7, ra[i]);
      else         store_colquad(tiles, tiles + TILE_BYTES, r1, k1(i), ra[i]);
      if (BKIND == 0) store_rowchunk(tiles + 2 * TILE_BYTES, tiles + 3 * TILE_BYTES, row0 + 32 * i, tid & 7, rb[i]);
      else            store_colquad(tiles + 2 * TILE_BYTES, tiles + 3 * TILE_BYTES, r1, k1(i), rb[i]);
    }
  };

  // fp32 partial sums of this thread: row (warp & 3)*32 + lane of the tile, columns (warp >> 2)*64 .. +63.  Chunks of FLUSH
  // k-blocks are accumulated inside TMEM (the tensor core adds with truncation: short chains only) and moved here with
  // round-to-nearest adds, so the long-K error stays fp32-class.
  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;
  const uint32_t tmem_mine = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  auto flush = [&](int buf) {
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tmem_mine + (uint32_t)(buf * BN + c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[c * 32 + j] += v[j];
    }
    tc_fence_before();
  };

  constexpr uint32_t idesc = umma_idesc(2, BN, 0, 0);
  const uint32_t sbase = smem_u32(tiles);
  if (nkb > 0) gload(kb0);
  for (int it = 0; it < nkb; ++it) {
    if (it > 0) mbar_wait(smem_u32(bar), (it - 1) & 1);     // the MMAs of k-block it-1 have retired: the tiles are free
    sstore();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t dacc = tmem_base + (uint32_t)(((it / FLUSH) & 1) * BN);
      const uint64_t ahi = umma_desc_kmajor_sw128(sbase), alo = umma_desc_kmajor_sw128(sbase + TILE_BYTES);
      const uint64_t bhi = umma_desc_kmajor_sw128(sbase + 2 * TILE_BYTES), blo = umma_desc_kmajor_sw128(sbase + 3 * TILE_BYTES);
      // small terms first; one k-step = 8 tf32 = 32 bytes = +2 in the descriptor's (>>4) address field
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) umma_tf32(dacc, alo + 2 * ks, bhi + 2 * ks, idesc, ((it % FLUSH) | ks) != 0);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) umma_tf32(dacc, ahi + 2 * ks, blo + 2 * ks, idesc, 1);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) umma_tf32(dacc, ahi + 2 * ks, bhi + 2 * ks, idesc, 1);
      umma_commit(smem_u32(bar));
    }
    // the chunk that ended with k-block it-1 (its completion was observed above) moves to registers while the tensor core
    // works on k-block it; its TMEM half is not written again before the next __syncthreads
    if (it > 0 && it % FLUSH == 0) flush(((it / FLUSH) - 1) & 1);
    if (it + 1 < nkb) gload(kb0 + it + 1);
  }
  if (nkb > 0) {
    mbar_wait(smem_u32(bar), (nkb - 1) & 1);
    flush(((nkb - 1) / FLUSH) & 1);
  }

  const int row = m0 + (warp & 3) * 32 + lane;
  const int chalf = (warp >> 2) * 64;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int n = n0 + chalf + q * 4;
    const float4 o = make_float4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
    if (d.split_k > 1) {
      if (row < d.M && n < d.N) {
        float* p = d.ws + ((long long)(split * d.batch + bidx) * d.M + row) * d.N + n;
        if ((d.N & 3) == 0) {
          *reinterpret_cast<float4*>(p) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < d.N) p[j] = (&o.x)[j];
        }
      }
    } else {
      epi_store4(d, bidx, row, n, o, vecC);
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int AK, int BKIND, int FASTA>
static int launch_x3(const tbns_gemm_desc& d, int vecA, int vecB, int vecC, dim3 grid, cudaStream_t st) {
  using namespace x3;
  TBNS_SMEM_OPT_IN((gemm_x3_kernel<AK, BKIND, FASTA>), SMEM_BYTES);
  gemm_x3_kernel<AK, BKIND, FASTA><<<grid, NT, SMEM_BYTES, st>>>(d, vecA, vecB, vecC);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

// launched by tbns_gemm (gemm_simt.cu) for precision == TBNS_PREC_FP32 when the contraction is large enough to pay for a
// 128x128 tensor-core tile; the split-K reduce (fixed order) is shared with the SIMT engine.
int gemm_x3_launch(const tbns_gemm_desc& d, int vecA, int vecB, int vecC, dim3 grid, cudaStream_t st) {
  const bool fast1 = d.conv_mode == 1 && vecA && d.Cin % x3::BK == 0 && d.Hg < (1 << 15) && d.Wg < (1 << 16);
  const bool fast2 = d.conv_mode == 2 && vecA && !d.flip && d.K < (1 << 24) && d.Hg * d.Wg < (1 << 24);
  if (d.a_kind == 0 && d.b_kind == 0) return fast1 ? launch_x3<0, 0, 1>(d, vecA, vecB, vecC, grid, st) : launch_x3<0, 0, 0>(d, vecA, vecB, vecC, grid, st);
  if (d.a_kind == 0 && d.b_kind == 1) return fast1 ? launch_x3<0, 1, 1>(d, vecA, vecB, vecC, grid, st) : launch_x3<0, 1, 0>(d, vecA, vecB, vecC, grid, st);
  if (d.a_kind == 1 && d.b_kind == 0) return fast2 ? launch_x3<1, 0, 1>(d, vecA, vecB, vecC, grid, st) : launch_x3<1, 0, 0>(d, vecA, vecB, vecC, grid, st);
  return fast2 ? launch_x3<1, 1, 1>(d, vecA, vecB, vecC, grid, st) : launch_x3<1, 1, 0>(d, vecA, vecB, vecC, grid, st);
}

}  // namespace tbns
