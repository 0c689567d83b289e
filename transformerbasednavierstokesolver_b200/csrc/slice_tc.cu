// Tensor-core slice stage (bf16 mode, dim_head == 32, slice_num in {32, 64}): the small contractions of
// Physics-Attention run as tcgen05.mma.kind::f16 (bf16 operands, fp32 accumulation in TMEM) on 128-token x one-head tiles,
// fused with the temperature softmax.  The projections XF = [x_mid | fx_mid] arrive in bf16 - the projection GEMM's epilogue
// writes nothing else - so the stage reads 2 bytes per projected value instead of 4 and the fp32 copy of XF no longer exists.
//
//   forward  (model/Physics_Attention.py:98-101 / :40-42)
//     logits[128 x G]   = X[128 x 32] . Ws^T           A = X tile (TMA, K-major, 64B swizzle), B = Ws (K-major)  -> TMEM
//     w = softmax(logits / tau)                        one thread per token reads its TMEM lane (tcgen05.ld), registers
//     Tt^T[32 x G]     += F^T[32 x 128] . w[128 x G]   both operands K-major with K = tokens: the threads scatter F^T and w^T
//                                                      into transposed swizzled tiles; accumulates in TMEM over all chunks
//   backward (SURVEY.md §8 a-bwd; oracle/physics_attention.py: slice_bwd)
//     logits, dwv = F.dTt^T  -> softmax backward in registers -> dF = w.dTt, dX = dL.Ws, dWs^T += X^T.dL
//
// bf16 operands only touch these K<=128 contractions (7% of the block's FLOPs); softmax, normalisation and all accumulation
// stay fp32.  fp32 mode and other head shapes use the exact SIMT kernels in slice_v2.cuh (fp32 XF).
#include "tc_common.cuh"

namespace tbns {

constexpr int ST_TOK = 128;                 // tokens per chunk == threads per CTA == TMEM lanes
constexpr int ST_D = 32;                    // dim_head handled here: one 64-byte swizzle row of bf16
constexpr uint32_t ST_TILE = ST_TOK * 64;   // bytes of a [128 tokens][32 bf16] tile (64B swizzle)
constexpr uint32_t ST_TP = 32 * 128;        // panel of a transposed [32 rows][64 tokens] bf16 operand (2 panels = 128 tokens)

__device__ __forceinline__ float st_clamp_tau(float t, int clamp) { return clamp ? fminf(fmaxf(t, 0.1f), 5.0f) : t; }

// K-major bf16 tile whose rows are ROWB bytes: 64 (64B swizzle) or 128 (128B swizzle)
template <int ROWB>
__device__ __forceinline__ uint32_t st_off(int row, int chunk16) { return ROWB == 64 ? sw64_off(row, chunk16) : sw128_off(row, chunk16); }
template <int ROWB>
__device__ __forceinline__ uint64_t st_desc(uint32_t saddr) { return ROWB == 64 ? umma_desc_kmajor_sw64(saddr) : umma_desc_kmajor_sw128(saddr); }

// thread `t` owns token t of the chunk: scatter its vector as column t of a transposed K-major bf16 operand
// ([rows][tokens], 2 panels of 64 tokens, rows of 128 bytes, 128B swizzle): the operands of the contractions over tokens
template <int ROWS>
__device__ __forceinline__ void st_scatter_col16(uint8_t* tile, uint32_t panel_stride, int t, const float (&v)[ROWS]) {
  uint8_t* p = tile + (t >> 6) * panel_stride + (t & 7) * 2;
  const int c = (t & 63) >> 3;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) *reinterpret_cast<__nv_bfloat16*>(p + sw128_off(r, c)) = __float2bfloat16_rn(v[r]);
}
// same for a row that already is bf16 (this token's 32 projected values, 4 x 16 bytes as read from the TMA tile)
__device__ __forceinline__ void st_scatter_col16_raw(uint8_t* tile, uint32_t panel_stride, int t, const uint4 (&row)[4]) {
  uint8_t* p = tile + (t >> 6) * panel_stride + (t & 7) * 2;
  const int c = (t & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t w[4] = {row[q].x, row[q].y, row[q].z, row[q].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<uint16_t*>(p + sw128_off(8 * q + 2 * j, c)) = (uint16_t)(w[j] & 0xffffu);
      *reinterpret_cast<uint16_t*>(p + sw128_off(8 * q + 2 * j + 1, c)) = (uint16_t)(w[j] >> 16);
    }
  }
}

__device__ __forceinline__ uint4 st_pack8(const float* v) {
  __nv_bfloat162 o[4] = {__floats2bfloat162_rn(v[0], v[1]), __floats2bfloat162_rn(v[2], v[3]), __floats2bfloat162_rn(v[4], v[5]),
                         __floats2bfloat162_rn(v[6], v[7])};
  return *reinterpret_cast<uint4*>(o);
}
__device__ __forceinline__ float st_sum8(const uint4& u) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h2[j]);
    a += f.x + f.y;
  }
  return a;
}

// [rows][32] fp32 matrix (row-major) -> K-major bf16 operand with 64-byte rows (64B swizzle)
__device__ __forceinline__ void st_stage_k32(uint8_t* dst, const float* __restrict__ src, int rows, int tid) {
  for (int idx = tid; idx < rows * 4; idx += ST_TOK) {
    const int r = idx >> 2, c = idx & 3;
    const float4 a = *reinterpret_cast<const float4*>(src + r * ST_D + c * 8), b = *reinterpret_cast<const float4*>(src + r * ST_D + c * 8 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<uint4*>(dst + sw64_off(r, c)) = st_pack8(v);
  }
}

// Shared memory per CTA: 28 KB for G = 32 (the fp32 / tf32 version needed 53 KB), so residency is bounded by the CTA slots
// and TMEM columns, not by shared memory: each chunk is a chain of short dependent phases and the SM hides them by running
// many CTAs.
template <int G>
struct StFwdSmem {
  static constexpr int FT = 0;                        // F^T bf16: 2 panels x [32 d][64 tokens]; the MMA has M = 128 rows, so every
                                                      // panel is read 16 KB deep: rows >= 32 are don't-care lanes of the
                                                      // accumulator and simply alias the tiles that follow
  static constexpr int XS = FT + 2 * ST_TP;           // X tile [128 tokens][32] bf16 (TMA, SW64)
  static constexpr int FS = XS + ST_TILE;             // F tile; dead once F^T is written, so
  static constexpr int WT = FS;                       // w^T (2 panels x [G][64 tokens]) is built over it
  static constexpr int WS = WT + (2 * G * 128 > (int)ST_TILE ? 2 * G * 128 : (int)ST_TILE);   // Ws [G][32] bf16 K-major (SW64)
  static constexpr int BAR = WS + G * 64;             // mbarriers: X tile, F tile, mma ; tmem slot
  static constexpr int BS = BAR + 32;                 // bias [G]
  static constexpr int TOTAL = BS + G * 4 + 1024;
  static_assert(FT + ST_TP + 16384 <= BS, "the 128-row reads of the F^T operand must stay inside the allocation");
};

// grid (groups, H, B), block 128
template <int G>
__global__ void __launch_bounds__(ST_TOK, (G == 32 ? 8 : 3)) slice_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmXF, const float* __restrict__ Ws,
                                                              const float* __restrict__ bs, const float* __restrict__ temperature,
                                                              __nv_bfloat16* __restrict__ w16, float* __restrict__ part, int N, int H,
                                                              int nchunk, int clamp) {
  using S = StFwdSmem<G>;
  constexpr int TMEM_COLS = 2 * G;   // logits [G] | token accumulator [G] (lane = dim_head index)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_x = base + S::BAR, bar_f = bar_x + 8, bar_mma = bar_x + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR + 24);
  float* bsm = reinterpret_cast<float*>(gen + S::BS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int I = H * ST_D;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXF) : "memory");
    mbar_init(bar_x, 1);
    mbar_init(bar_f, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_sync();   // barrier init / TMEM allocation overlapped the previous kernel (common.cuh: PDL)
  st_stage_k32(gen + S::WS, Ws, G, tid);   // Ws -> K-major swizzled B operand
  for (int idx = tid; idx < G; idx += ST_TOK) bsm[idx] = bs[idx];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float inv_tau = 1.0f / st_clamp_tau(temperature[h], clamp);
  constexpr uint32_t idesc = umma_idesc(1, G, 0, 0);   // bf16, M = 128, N = G, both operands K-major

  uint32_t ph_mma = 0;
  float sacc = 0.f;
  int iter = 0;
  auto issue_x = [&](int chunk) {
    mbar_expect_tx(bar_x, ST_TILE);
    tma_load_3d(base + S::XS, &tmXF, bar_x, h * ST_D, chunk * ST_TOK, b);
  };
  auto issue_f = [&](int chunk) {
    mbar_expect_tx(bar_f, ST_TILE);
    tma_load_3d(base + S::FS, &tmXF, bar_f, I + h * ST_D, chunk * ST_TOK, b);
  };
  if (tid == 0 && blockIdx.x < nchunk) issue_x(blockIdx.x);
  for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x, ++iter) {
    const int n0 = chunk * ST_TOK;
    const uint32_t ph = iter & 1;
    if (tid == 0) issue_f(chunk);   // the F / w^T region was last read by the previous chunk's token MMA (waited for below)
    mbar_wait(bar_x, ph);
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < ST_D / 16; ++k)
        umma_bf16(tmem, umma_desc_kmajor_sw64(base + S::XS + k * 32), umma_desc_kmajor_sw64(base + S::WS + k * 32), idesc, k != 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_f, ph);
    {
      // meanwhile: column `tid` of F^T (A operand of the token contraction)
      uint4 f[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) f[c] = *reinterpret_cast<const uint4*>(gen + S::FS + sw64_off(tid, c));
      st_scatter_col16_raw(gen + S::FT, ST_TP, tid, f);
    }
    __syncthreads();   // every F row has been read: w^T may be written over the F tile
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    if (tid == 0 && chunk + (int)gridDim.x < nchunk) issue_x(chunk + gridDim.x);   // the logits MMA was the X tile's only reader
    {
      float l[G];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), l);
      if (G == 64) tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 32, l + 32);
      float mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        l[g] = (l[g] + bsm[g]) * inv_tau;
        mx = fmaxf(mx, l[g]);
      }
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        l[g] = __expf(l[g] - mx);
        sum += l[g];
      }
      const bool valid = n0 + tid < N;
      const float inv = valid ? 1.0f / sum : 0.0f;   // tokens past N contribute nothing
#pragma unroll
      for (int g = 0; g < G; ++g) l[g] *= inv;
      st_scatter_col16<G>(gen + S::WT, G * 128, tid, l);   // column `tid` of w^T (B operand of the token contraction)
      if (valid) {                                          // the same bf16 values in HBM (operand of the tensor-core deslice)
        __nv_bfloat16* wr = w16 + (((long long)b * N + n0 + tid) * H + h) * G;
#pragma unroll
        for (int c = 0; c < G / 8; ++c) *reinterpret_cast<uint4*>(wr + 8 * c) = st_pack8(l + 8 * c);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // Tt^T[d][g] += sum_t F^T[d][t] w^T[g][t]: M = 128 (lanes 0..31 = dim_head index are meaningful), N = G, K = 16 tokens / MMA
#pragma unroll
      for (int kp = 0; kp < 2; ++kp)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + G, umma_desc_kmajor_sw128(base + S::FT + kp * ST_TP + k * 32),
                    umma_desc_kmajor_sw128(base + S::WT + kp * (G * 128) + k * 32), idesc, (iter | kp | k) != 0);
      umma_commit(bar_mma);
    }
    if (tid < G) {   // s[g] += sum_t w[t][g]: row g of w^T (the bf16 values the contraction sums), 2 panels x 64 tokens
      float a = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp)
#pragma unroll
        for (int c = 0; c < 8; ++c) a += st_sum8(*reinterpret_cast<const uint4*>(gen + S::WT + kp * (G * 128) + sw128_off(tid, c)));
      sacc += a;
    }
    mbar_wait(bar_mma, ph_mma);   // operand tiles may be overwritten by the next chunk
    ph_mma ^= 1;
    __syncthreads();              // ... once the threads summing rows of w^T are done with it as well
  }
  tc_fence_after();
  float* pbase = part + (((long long)b * H + h) * gridDim.x + blockIdx.x) * G * (ST_D + 1);
  if (warp == 0) {
    // TMEM lane d (= lane) holds Tt[g][d] for g = 0..G-1
    float v[G];
    tmem_ld32(tmem + G, v);
    if (G == 64) tmem_ld32(tmem + G + 32, v + 32);
#pragma unroll
    for (int g = 0; g < G; ++g) pbase[g * (ST_D + 1) + lane] = v[g];
  }
  if (tid < G) pbase[tid * (ST_D + 1) + ST_D] = sacc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// XF16 [B][N][2I] bf16 viewed as a 3-D tensor with a {32 values, 128 tokens, 1} box, 64B swizzle
static int encode_xf(CUtensorMap* m, const void* XF16, int B, int N, int I2) {
  cuuint64_t dims[3] = {(cuuint64_t)I2, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)I2 * 2, (cuuint64_t)N * I2 * 2};
  cuuint32_t box[3] = {32u, (cuuint32_t)ST_TOK, 1u};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, XF16, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int G>
static int launch_slice_fwd_tc(const void* XF16, const float* Ws, const float* bs, const float* temperature, __nv_bfloat16* w16,
                               float* part, int B, int N, int H, int groups, int clamp, cudaStream_t st) {
  CUtensorMap tm;
  int rc = encode_xf(&tm, XF16, B, N, 2 * H * ST_D);
  if (rc) return rc;
  TBNS_SMEM_OPT_IN((slice_fwd_tc_kernel<G>), StFwdSmem<G>::TOTAL);
  dim3 grid(groups, H, B);
  TBNS_CUDA(launch_pdl(slice_fwd_tc_kernel<G>, grid, dim3(ST_TOK), StFwdSmem<G>::TOTAL, st, tm, Ws, bs, temperature, w16, part, N, H,
                       cdiv(N, ST_TOK), clamp));
  return TBNS_OK;
}

template <int G>
struct StBwdSmem {
  static constexpr int RB = G * 2;                     // row bytes of a [.][G] bf16 K-major operand: 64 (SW64) or 128 (SW128)
  static constexpr int XS = 0;                         // { X tile, F tile } (TMA, SW64)
  static constexpr int WL = G == 32 ? 0 : 2 * (int)ST_TILE;   // { w tile, dL tile } [128 tokens][G] bf16.  G = 32: same size as the
                                                       // X / F tiles and built over them (each thread overwrites only the row it
                                                       // has read, after the first MMA group - the other reader - has completed)
  static constexpr int WL_TILE = ST_TOK * RB;
  static constexpr int XT = G == 32 ? 2 * (int)ST_TILE : WL + 2 * WL_TILE;   // X^T bf16: 2 panels x [32 d][64 tokens]; the MMA reads
                                                       // M = 128 rows per panel: the don't-care tail aliases what follows
  static constexpr int LT = XT + 2 * (int)ST_TP;       // dL^T bf16: 2 panels x [G][64 tokens]
  static constexpr int WS = LT + 2 * G * 128;          // Ws   [G][32]   K-major (B of the logits)
  static constexpr int DT = WS + G * 64;               // dTt  [G][32]   K-major (B of dwv)
  static constexpr int WST = DT + G * 64;              // Ws^T [32][G]   K-major (B of dX)
  static constexpr int DTT = WST + 32 * RB;            // dTt^T[32][G]   (B of dF)
  static constexpr int BAR = DTT + 32 * RB;            // mbarriers: tma, mma ; tmem slot
  static constexpr int MISC = BAR + 32;                // bias [G], ds [G], red[4]
  static constexpr int TOTAL = MISC + (2 * G + 8) * 4; // no alignment slack: the kernel declares its dynamic smem 1024-aligned
                                                       // (G = 32: 41 256 B and 128 TMEM columns -> four CTAs per SM)
  static_assert(XT + (int)ST_TP + 16384 <= TOTAL, "the 128-row reads of the X^T operand must stay inside the allocation");
};

// grid (groups, H, B), block 128
template <int G>
__global__ void __launch_bounds__(ST_TOK, (G == 32 ? 4 : 1)) slice_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmXF, const float* __restrict__ Ws,
                                                              const float* __restrict__ bs, const float* __restrict__ temperature,
                                                              const __nv_bfloat16* __restrict__ dw, const float* __restrict__ dTt,
                                                              const float* __restrict__ ds, __nv_bfloat16* __restrict__ dXF16,
                                                              float* __restrict__ dWs_part, float* __restrict__ dtau_part, int N, int H,
                                                              int nchunk, int clamp) {
  using S = StBwdSmem<G>;
  constexpr int RB = S::RB;
  // L [G] | dwv [G] | dWs^T [G]; dF [32] and dX [32] reuse the L / dwv columns (read into registers before the second MMA
  // group is issued): 96 columns -> 128 allocated for G = 32, so four CTAs share the SM's 512 columns
  constexpr int TMEM_COLS = G == 32 ? 128 : 256;
  constexpr uint32_t C_L = 0, C_DW = G, C_DF = 0, C_DX = G, C_WS = 2 * G;
  extern __shared__ __align__(1024) uint8_t smem_al[];   // swizzled tiles need 1024-byte aligned bases
  const uint32_t base = smem_u32(smem_al);
  if (base & 1023u) __trap();
  uint8_t* gen = smem_al;
  const uint32_t bar_tma = base + S::BAR, bar_mma = bar_tma + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR + 24);
  float* bsm = reinterpret_cast<float*>(gen + S::MISC);
  float* dss = bsm + G;
  float* red = dss + G;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int I = H * ST_D, HG = H * G;
  const long long bh = (long long)b * H + h;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXF) : "memory");
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_sync();
  // per-(b,h) constant operands: Ws, dTt (K-major over dim_head) and their transposes (K-major over the slice index)
  const float* dTh = dTt + bh * G * ST_D;
  st_stage_k32(gen + S::WS, Ws, G, tid);
  st_stage_k32(gen + S::DT, dTh, G, tid);
  for (int idx = tid; idx < G * ST_D; idx += ST_TOK) {
    const int g = idx / ST_D, d = idx - g * ST_D;   // element (d, g) of the transposes: row d, column g
    const uint32_t off = st_off<RB>(d, g >> 3) + (g & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(gen + S::WST + off) = __float2bfloat16_rn(Ws[idx]);
    *reinterpret_cast<__nv_bfloat16*>(gen + S::DTT + off) = __float2bfloat16_rn(dTh[idx]);
  }
  for (int idx = tid; idx < G; idx += ST_TOK) {
    bsm[idx] = bs[idx];
    dss[idx] = ds[bh * G + idx];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float tau = st_clamp_tau(temperature[h], clamp);
  const float inv_tau = 1.0f / tau;
  constexpr uint32_t idesc_g = umma_idesc(1, G, 0, 0);    // bf16, N = G
  constexpr uint32_t idesc_d = umma_idesc(1, 32, 0, 0);   // bf16, N = dim_head

  uint32_t ph_mma = 0;
  float dbs_acc = 0.f, dtau_acc = 0.f;
  int iter = 0;
  // one X / F stage per CTA (four CTAs per SM cover each other's latencies); the next chunk's tiles are requested as soon as
  // the second MMA group - the last reader of the w / dL tiles that may lie over them - has completed
  auto issue = [&](int chunk) {
    mbar_expect_tx(bar_tma, 2 * ST_TILE);
    tma_load_3d(base + S::XS, &tmXF, bar_tma, h * ST_D, chunk * ST_TOK, b);
    tma_load_3d(base + S::XS + ST_TILE, &tmXF, bar_tma, I + h * ST_D, chunk * ST_TOK, b);
  };
  if (tid == 0 && blockIdx.x < nchunk) issue(blockIdx.x);
  for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x, ++iter) {
    const int n0 = chunk * ST_TOK;
    const bool valid = n0 + tid < N;
    const long long row = (long long)b * N + n0 + tid;
    const uint32_t xs = S::XS, fs = xs + ST_TILE;
    const uint32_t wt = S::WL, dl = S::WL + S::WL_TILE;
    // this token's deslice gradient row (issued early: overlaps the TMA / MMA latency)
    float dwv[G];
    if (valid) {
      const __nv_bfloat16* dr = dw + row * HG + h * G;
#pragma unroll
      for (int c = 0; c < G / 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(dr + 8 * c);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          dwv[8 * c + 2 * j] = f.x;
          dwv[8 * c + 2 * j + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) dwv[g] = 0.f;
    }
    mbar_wait(bar_tma, iter & 1);
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 2; ++k)   // L = X Ws^T
        umma_bf16(tmem + C_L, umma_desc_kmajor_sw64(base + xs + k * 32), umma_desc_kmajor_sw64(base + S::WS + k * 32), idesc_g, k != 0);
#pragma unroll
      for (int k = 0; k < 2; ++k)   // F dTt^T
        umma_bf16(tmem + C_DW, umma_desc_kmajor_sw64(base + fs + k * 32), umma_desc_kmajor_sw64(base + S::DT + k * 32), idesc_g, k != 0);
      umma_commit(bar_mma);
    }
    {
      // meanwhile: column `tid` of X^T (A operand of dWs^T += X^T dL)
      uint4 x[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const uint4*>(gen + xs + sw64_off(tid, c));
      st_scatter_col16_raw(gen + S::XT, ST_TP, tid, x);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    {
      float L[G];
      {
        float t[G];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + C_L, L);
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + C_DW, t);
        if (G == 64) {
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + C_L + 32, L + 32);
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + C_DW + 32, t + 32);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          L[g] += bsm[g];                 // pre-temperature logit
          dwv[g] += t[g] + dss[g];        // total gradient wrt the slice weight
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < G; ++g) mx = fmaxf(mx, L[g] * inv_tau);
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) sum += __expf(L[g] * inv_tau - mx);
      const float inv = valid ? 1.0f / sum : 0.0f;
      float dot = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) dot = fmaf(dwv[g], __expf(L[g] * inv_tau - mx) * inv, dot);
      // softmax backward, temperature gradient; L <- w, dwv <- dL
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float wv = __expf(L[g] * inv_tau - mx) * inv;
        const float dLp = wv * (dwv[g] - dot);
        dtau_acc = fmaf(dLp, L[g], dtau_acc);
        L[g] = wv;
        dwv[g] = dLp * inv_tau;
      }
      // w row -> K-major A operand of dF = w dTt (G = 32: over this thread's own row of the X tile);
      // dL row -> K-major A operand of dX = dL Ws; dL column -> dL^T (B operand of dWs^T += X^T dL)
#pragma unroll
      for (int c = 0; c < G / 8; ++c) {
        *reinterpret_cast<uint4*>(gen + wt + st_off<RB>(tid, c)) = st_pack8(L + 8 * c);
        *reinterpret_cast<uint4*>(gen + dl + st_off<RB>(tid, c)) = st_pack8(dwv + 8 * c);
      }
      st_scatter_col16<G>(gen + S::LT, G * 128, tid, dwv);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < G / 16; ++k) {   // dF = w dTt ; dX = dL Ws   (K = G)
        umma_bf16(tmem + C_DF, st_desc<RB>(base + wt + k * 32), st_desc<RB>(base + S::DTT + k * 32), idesc_d, k != 0);
        umma_bf16(tmem + C_DX, st_desc<RB>(base + dl + k * 32), st_desc<RB>(base + S::WST + k * 32), idesc_d, k != 0);
      }
#pragma unroll
      for (int kp = 0; kp < 2; ++kp)     // dWs^T[d][g] += sum_t X^T[d][t] dL^T[g][t]   (16 tokens per MMA)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + C_WS, umma_desc_kmajor_sw128(base + S::XT + kp * ST_TP + k * 32),
                    umma_desc_kmajor_sw128(base + S::LT + kp * (G * 128) + k * 32), idesc_g, (iter | kp | k) != 0);
      umma_commit(bar_mma);
    }
    if (tid < G) {   // dbs[g] += sum_t dL[t][g]: row g of dL^T (bf16, 2 panels of 64 tokens)
      float a = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp)
#pragma unroll
        for (int c = 0; c < 8; ++c) a += st_sum8(*reinterpret_cast<const uint4*>(gen + S::LT + kp * (G * 128) + sw128_off(tid, c)));
      dbs_acc += a;
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    if (tid == 0 && chunk + (int)gridDim.x < nchunk) issue(chunk + gridDim.x);   // overlaps the stores below
    {
      // this token's dX / dF rows -> bf16 operand of the projection dgrad / wgrad GEMMs
      float o[ST_D];
      __nv_bfloat16* orow = dXF16 + row * (2LL * I) + h * ST_D;
#pragma unroll
      for (int part2 = 0; part2 < 2; ++part2) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (part2 == 0 ? C_DX : C_DF), o);
        if (valid) {
          __nv_bfloat16* dst = orow + (part2 == 0 ? 0 : I);
#pragma unroll
          for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + 8 * c) = st_pack8(o + 8 * c);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // all TMEM reads of this chunk done before the next chunk's MMAs overwrite dF / dX / L
    tc_fence_after();
  }
  const long long slot = bh * gridDim.x + blockIdx.x;
  dtau_acc = warp_sum(dtau_acc);
  if (lane == 0) red[warp] = dtau_acc;
  float* pout = dWs_part + slot * G * (ST_D + 1);
  if (warp == 0) {
    // TMEM lane d (= lane) holds dWs[g][d] for g = 0..G-1
    float v[G];
    tmem_ld32(tmem + C_WS, v);
    if (G == 64) tmem_ld32(tmem + C_WS + 32, v + 32);
#pragma unroll
    for (int g = 0; g < G; ++g) pout[g * (ST_D + 1) + lane] = v[g];
  }
  if (tid < G) pout[tid * (ST_D + 1) + ST_D] = dbs_acc;
  tc_fence_before();
  __syncthreads();
  if (tid == 0) dtau_part[slot] = -(red[0] + red[1] + red[2] + red[3]) * inv_tau * inv_tau;
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

template <int G>
static int launch_slice_bwd_tc(const void* XF16, const float* Ws, const float* bs, const float* temperature, const __nv_bfloat16* dw,
                               const float* dTt, const float* ds, __nv_bfloat16* dXF16, float* dWs_part, float* dtau_part, int B, int N,
                               int H, int groups, int clamp, cudaStream_t st) {
  CUtensorMap tm;
  int rc = encode_xf(&tm, XF16, B, N, 2 * H * ST_D);
  if (rc) return rc;
  TBNS_SMEM_OPT_IN((slice_bwd_tc_kernel<G>), StBwdSmem<G>::TOTAL);
  dim3 grid(groups, H, B);
  TBNS_CUDA(launch_pdl(slice_bwd_tc_kernel<G>, grid, dim3(ST_TOK), StBwdSmem<G>::TOTAL, st, tm, Ws, bs, temperature, dw, dTt, ds, dXF16,
                       dWs_part, dtau_part, N, H, cdiv(N, ST_TOK), clamp));
  return TBNS_OK;
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_slice_groups(int B, int N, int H);

extern "C" int tbns_pa_slice_tc_supported(int D, int G) { return (D == ST_D && (G == 32 || G == 64)) ? 1 : 0; }

extern "C" int tbns_pa_slice_fwd_tc(const void* XF16, const float* Ws, const float* bs, const float* temperature, void* w16, float* part,
                                    int B, int N, int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF16 && Ws && bs && temperature && w16 && part, "tbns_pa_slice_fwd_tc: null pointer");
  TBNS_REQUIRE(tbns_pa_slice_tc_supported(D, G), "tbns_pa_slice_fwd_tc: needs dim_head 32 and slice_num 32 or 64 (got %d, %d)", D, G);
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "tbns_pa_slice_fwd_tc: bad dims");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(XF16) & 15) == 0 && (reinterpret_cast<uintptr_t>(w16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(Ws) & 15) == 0,
               "tbns_pa_slice_fwd_tc: operands must be 16-byte aligned");
  const int groups = tbns_slice_groups(B, N, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (G == 32) return launch_slice_fwd_tc<32>(XF16, Ws, bs, temperature, reinterpret_cast<__nv_bfloat16*>(w16), part, B, N, H, groups, clamp, st);
  return launch_slice_fwd_tc<64>(XF16, Ws, bs, temperature, reinterpret_cast<__nv_bfloat16*>(w16), part, B, N, H, groups, clamp, st);
}

extern "C" int tbns_pa_slice_bwd_tc(const void* XF16, const float* Ws, const float* bs, const float* temperature, const void* dw16,
                                    const float* dTt, const float* ds, void* dXF16, float* dWs_part, float* dtau_part, int B, int N,
                                    int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF16 && Ws && bs && temperature && dw16 && dTt && ds && dXF16 && dWs_part && dtau_part, "tbns_pa_slice_bwd_tc: null pointer");
  const __nv_bfloat16* dw = reinterpret_cast<const __nv_bfloat16*>(dw16);
  TBNS_REQUIRE(tbns_pa_slice_tc_supported(D, G), "tbns_pa_slice_bwd_tc: needs dim_head 32 and slice_num 32 or 64 (got %d, %d)", D, G);
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "tbns_pa_slice_bwd_tc: bad dims");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(XF16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dXF16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(Ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dTt) & 15) == 0,
               "tbns_pa_slice_bwd_tc: operands must be 16-byte aligned");
  const int groups = tbns_slice_groups(B, N, H);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dXF16);
  if (G == 32) return launch_slice_bwd_tc<32>(XF16, Ws, bs, temperature, dw, dTt, ds, o, dWs_part, dtau_part, B, N, H, groups, clamp, st);
  return launch_slice_bwd_tc<64>(XF16, Ws, bs, temperature, dw, dTt, ds, o, dWs_part, dtau_part, B, N, H, groups, clamp, st);
}
