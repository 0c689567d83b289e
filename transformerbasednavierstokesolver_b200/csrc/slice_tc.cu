// Tensor-core slice stage (bf16 mode, dim_head == 32, slice_num in {32, 64}): the small contractions of
// Physics-Attention run as tcgen05.mma.kind::tf32 on 128-token x one-head tiles, fused with the temperature softmax.
//
//   forward  (model/Physics_Attention.py:98-101 / :40-42)
//     logits[128 x G]   = X[128 x 32] . Ws^T           A = X tile (TMA, K-major, 128B swizzle), B = Ws (K-major)  -> TMEM
//     w = softmax(logits / tau)                        one thread per token reads its TMEM lane (tcgen05.ld), registers
//     Tt^T[32 x G]     += F^T[32 x 128] . w[128 x G]   both operands K-major with K = tokens: the threads scatter F^T and w^T
//                                                      into transposed swizzled tiles; accumulates in TMEM over all chunks
//   backward (SURVEY.md §8 a-bwd; oracle/physics_attention.py: slice_bwd)
//     logits, dwv = F.dTt^T  -> softmax backward in registers -> dF = w.dTt, dX = dL.Ws, dWs^T += X^T.dL
//
// TF32 operands (10-bit mantissa) only touch these K<=128 contractions (7% of the block's FLOPs); softmax, normalisation
// and all accumulation stay fp32.  fp32 mode and other head shapes use the exact SIMT kernels in slice_v2.cuh.
#include "tc_common.cuh"

namespace tbns {

constexpr int ST_TOK = 128;                 // tokens per chunk == threads per CTA == TMEM lanes
constexpr int ST_D = 32;                    // dim_head handled here: one 128-byte swizzle row of fp32
constexpr uint32_t ST_TILE = ST_TOK * 128;  // bytes of a [128 tokens][32 fp32] tile
constexpr uint32_t ST_TP = 4096;            // panel stride of a transposed [32 rows][32 tokens] operand (4 panels = 128 tokens)

__device__ __forceinline__ float st_clamp_tau(float t, int clamp) { return clamp ? fminf(fmaxf(t, 0.1f), 5.0f) : t; }

// thread `t` owns token t of the chunk: scatter its D-vector as column t of a transposed K-major operand
// ([rows][tokens], 4 panels of 32 tokens, rows of 128 bytes, 128B swizzle).  Conflict-free: a warp fills one 128-byte row per store.
template <int ROWS>
__device__ __forceinline__ void st_scatter_col(uint8_t* tile, uint32_t panel_stride, int t, const float (&v)[ROWS]) {
  uint8_t* p = tile + (t >> 5) * panel_stride + (t & 3) * 4;
  const int c = (t & 31) >> 2;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) *reinterpret_cast<float*>(p + sw128_off(r, c)) = v[r];
}

// same for a bf16 K-major operand ([rows][tokens], 2 panels of 64 tokens): the operands of the token contraction in the
// backward kernel (dWs^T += X^T dL) - half the shared memory of the tf32 tiles, which is what lets three CTAs share an SM
template <int ROWS>
__device__ __forceinline__ void st_scatter_col16(uint8_t* tile, uint32_t panel_stride, int t, const float (&v)[ROWS]) {
  uint8_t* p = tile + (t >> 6) * panel_stride + (t & 7) * 2;
  const int c = (t & 63) >> 3;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) *reinterpret_cast<__nv_bfloat16*>(p + sw128_off(r, c)) = __float2bfloat16_rn(v[r]);
}

// Shared memory is what limits residency here (each chunk is a chain of short dependent phases, so the SM wants many CTAs
// in flight): 53 KB per CTA for G = 32, i.e. four CTAs per SM.
template <int G>
struct StFwdSmem {
  static constexpr int FT = 0;                        // F^T: 4 panels x [32 d][32 tokens]; the MMA has M = 128 rows, so every panel
                                                      // is read 16 KB deep: rows >= 32 are don't-care lanes of the accumulator and
                                                      // simply alias the tiles that follow
  static constexpr int XS = FT + 4 * ST_TP;           // X tile [128 tokens][32] fp32 (TMA, SW128)
  static constexpr int FS = XS + ST_TILE;             // F tile; dead once F^T is written, so
  static constexpr int WT = FS;                       // w^T (4 panels x [G][32 tokens]) is built over it
  static constexpr int WS = WT + (4 * G * 128 > (int)ST_TILE ? 4 * G * 128 : (int)ST_TILE);   // Ws [G][32] K-major
  static constexpr int BAR = WS + G * 128;            // mbarriers: X tile, F tile, mma ; tmem slot
  static constexpr int BS = BAR + 32;                 // bias [G]
  static constexpr int TOTAL = BS + G * 4 + 1024;
};

// grid (groups, H, B), block 128
template <int G>
__global__ void __launch_bounds__(ST_TOK) slice_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmXF, const float* __restrict__ Ws,
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
  for (int idx = tid; idx < G * 8; idx += ST_TOK) {   // Ws -> K-major swizzled B operand
    const int g = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(gen + S::WS + sw128_off(g, c)) = *reinterpret_cast<const float4*>(Ws + g * ST_D + c * 4);
  }
  for (int idx = tid; idx < G; idx += ST_TOK) bsm[idx] = bs[idx];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float inv_tau = 1.0f / st_clamp_tau(temperature[h], clamp);
  constexpr uint32_t idesc = umma_idesc(2, G, 0, 0);   // tf32, M = 128, N = G, both operands K-major

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
      for (int k = 0; k < ST_D / 8; ++k)
        umma_tf32(tmem, umma_desc_kmajor_sw128(base + S::XS + k * 32), umma_desc_kmajor_sw128(base + S::WS + k * 32), idesc, k != 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_f, ph);
    {
      // meanwhile: column `tid` of F^T (A operand of the token contraction)
      float f[ST_D];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(gen + S::FS + sw128_off(tid, c));
        f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
      }
      st_scatter_col<ST_D>(gen + S::FT, ST_TP, tid, f);
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
      st_scatter_col<G>(gen + S::WT, G * 128, tid, l);   // column `tid` of w^T (B operand of the token contraction)
      if (valid) {                                        // bf16 copy in HBM (operand of the tensor-core deslice)
        __nv_bfloat16* wr = w16 + (((long long)b * N + n0 + tid) * H + h) * G;
#pragma unroll
        for (int c = 0; c < G / 8; ++c) {
          __nv_bfloat162 o[4] = {__floats2bfloat162_rn(l[8 * c], l[8 * c + 1]), __floats2bfloat162_rn(l[8 * c + 2], l[8 * c + 3]),
                                 __floats2bfloat162_rn(l[8 * c + 4], l[8 * c + 5]), __floats2bfloat162_rn(l[8 * c + 6], l[8 * c + 7])};
          *reinterpret_cast<uint4*>(wr + 8 * c) = *reinterpret_cast<uint4*>(o);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // Tt^T[d][g] += sum_t F^T[d][t] w^T[g][t]: M = 128 (lanes 0..31 = dim_head index are meaningful), N = G, K = 8 tokens / MMA
#pragma unroll
      for (int kp = 0; kp < 4; ++kp)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem + G, umma_desc_kmajor_sw128(base + S::FT + kp * ST_TP + k * 32),
                    umma_desc_kmajor_sw128(base + S::WT + kp * (G * 128) + k * 32), idesc, (iter | kp | k) != 0);
      umma_commit(bar_mma);
    }
    if (tid < G) {   // s[g] += sum_t w[t][g]: row g of w^T, 4 panels x 32 tokens
      float a = 0.f;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(gen + S::WT + kp * (G * 128) + sw128_off(tid, c));
          a += (v.x + v.y) + (v.z + v.w);
        }
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

// XF [B][N][2I] fp32 viewed as a 3-D tensor with a {32 floats, 128 tokens, 1} box
static int encode_xf(CUtensorMap* m, const float* XF, int B, int N, int I2) {
  cuuint64_t dims[3] = {(cuuint64_t)I2, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)I2 * 4, (cuuint64_t)N * I2 * 4};
  cuuint32_t box[3] = {32u, (cuuint32_t)ST_TOK, 1u};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, XF, 3, dims, str, box);
}

template <int G>
static int launch_slice_fwd_tc(const float* XF, const float* Ws, const float* bs, const float* temperature, __nv_bfloat16* w16,
                               float* part, int B, int N, int H, int groups, int clamp, cudaStream_t st) {
  CUtensorMap tm;
  int rc = encode_xf(&tm, XF, B, N, 2 * H * ST_D);
  if (rc) return rc;
  TBNS_SMEM_OPT_IN((slice_fwd_tc_kernel<G>), StFwdSmem<G>::TOTAL);
  dim3 grid(groups, H, B);
  TBNS_CUDA(launch_pdl(slice_fwd_tc_kernel<G>, grid, dim3(ST_TOK), StFwdSmem<G>::TOTAL, st, tm, Ws, bs, temperature, w16, part, N, H,
                       cdiv(N, ST_TOK), clamp));
  return TBNS_OK;
}


// round-to-nearest TF32 (the MMA itself truncates fp32 operands: pre-rounding keeps the operand error unbiased)
__device__ __forceinline__ float st_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <int G>
struct StBwdSmem {
  static constexpr int KP = G / 32;                    // 128-byte k-panels of a K = G operand
  static constexpr int XS = 0;                         // { X tile, F tile } (TMA): the X tile later becomes k-panel 0 of the w
  static constexpr int STAGE = 2 * ST_TILE;            // tile and the F tile k-panel 1 (G = 64) or, for G = 32, the dL tile
                                                       // (F is dead once the first MMA group has read it)
  static constexpr int DL = STAGE;                     // dL tile [128 tokens][G] K-major, KP panels (own region only for G = 64)
  static constexpr int XT = DL + (KP == 1 ? 0 : KP * (int)ST_TILE);   // X^T bf16: 2 panels x [32 d][64 tokens]; the MMA reads
                                                       // M = 128 rows per panel: the don't-care tail aliases what follows
  static constexpr int XT_P = 32 * 128;                // panel stride of X^T
  static constexpr int LT = XT + 2 * XT_P;             // dL^T bf16: 2 panels x [G][64 tokens]
  static constexpr int WS = LT + 2 * G * 128;          // Ws   [G][32]   K-major (B of the logits)
  static constexpr int DT = WS + G * 128;              // dTt  [G][32]   K-major (B of dwv)
  static constexpr int WST = DT + G * 128;             // Ws^T [32][G]   K-major, KP panels of [32][128 B] (B of dX)
  static constexpr int DTT = WST + KP * 4096;          // dTt^T[32][G]   (B of dF)
  static constexpr int BAR = DTT + KP * 4096;          // mbarriers: tma, mma ; tmem slot
  static constexpr int MISC = BAR + 32;                // bias [G], ds [G], red[4]
  static constexpr int TOTAL = MISC + (2 * G + 8) * 4; // no alignment slack: the kernel declares its dynamic smem 1024-aligned
                                                       // (G = 32: 65 832 B and 128 TMEM columns -> three CTAs per SM)
  static_assert(XT + XT_P + 16384 <= TOTAL, "the 128-row reads of the X^T operand must stay inside the allocation");
};

// grid (groups, H, B), block 128
template <int G>
__global__ void __launch_bounds__(ST_TOK) slice_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmXF, const float* __restrict__ Ws,
                                                              const float* __restrict__ bs, const float* __restrict__ temperature,
                                                              const __nv_bfloat16* __restrict__ dw, const float* __restrict__ dTt,
                                                              const float* __restrict__ ds, __nv_bfloat16* __restrict__ dXF16,
                                                              float* __restrict__ dWs_part, float* __restrict__ dtau_part, int N, int H,
                                                              int nchunk, int clamp) {
  using S = StBwdSmem<G>;
  constexpr int KP = S::KP;
  // L [G] | dwv [G] | dWs^T [G]; dF [32] and dX [32] reuse the L / dwv columns (read into registers before the second MMA
  // group is issued): 96 columns -> 128 allocated for G = 32, so three CTAs share the SM's 512 columns
  constexpr int TMEM_COLS = G == 32 ? 128 : 256;
  constexpr uint32_t C_L = 0, C_DW = G, C_DF = 0, C_DX = G, C_WS = 2 * G;
  extern __shared__ __align__(1024) uint8_t smem_al[];   // SW128 tiles need 1024-byte aligned bases
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
  for (int idx = tid; idx < G * 8; idx += ST_TOK) {
    const int g = idx >> 3, c = idx & 7;
    float4 a = *reinterpret_cast<const float4*>(Ws + g * ST_D + c * 4);
    float4 t = *reinterpret_cast<const float4*>(dTh + g * ST_D + c * 4);
    a.x = st_rna(a.x); a.y = st_rna(a.y); a.z = st_rna(a.z); a.w = st_rna(a.w);
    t.x = st_rna(t.x); t.y = st_rna(t.y); t.z = st_rna(t.z); t.w = st_rna(t.w);
    *reinterpret_cast<float4*>(gen + S::WS + sw128_off(g, c)) = a;
    *reinterpret_cast<float4*>(gen + S::DT + sw128_off(g, c)) = t;
  }
  for (int idx = tid; idx < G * ST_D; idx += ST_TOK) {
    const int g = idx / ST_D, d = idx - g * ST_D;   // element (d, g) of the transposes: k-panel g/32, row d, column g%32
    const uint32_t off = (g >> 5) * 4096 + sw128_off(d, (g & 31) >> 2) + (g & 3) * 4;
    *reinterpret_cast<float*>(gen + S::WST + off) = st_rna(Ws[idx]);
    *reinterpret_cast<float*>(gen + S::DTT + off) = st_rna(dTh[idx]);
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
  constexpr uint32_t idesc_g = umma_idesc(2, G, 0, 0);    // tf32, N = G
  constexpr uint32_t idesc_d = umma_idesc(2, 32, 0, 0);   // tf32, N = dim_head
  constexpr uint32_t idesc_t = umma_idesc(1, G, 0, 0);    // bf16, N = G (token contraction)

  uint32_t ph_mma = 0;
  float dbs_acc = 0.f, dtau_acc = 0.f;
  int iter = 0;
  // one X / F stage per CTA (three CTAs per SM cover each other's latencies); the next chunk's tiles are requested as soon as
  // the second MMA group - the last reader of the w / dL tiles written over them - has completed
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
    const uint32_t dl = KP == 1 ? fs : (uint32_t)S::DL;
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
      for (int k = 0; k < 4; ++k)   // L = X Ws^T
        umma_tf32(tmem + C_L, umma_desc_kmajor_sw128(base + xs + k * 32), umma_desc_kmajor_sw128(base + S::WS + k * 32), idesc_g, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // F dTt^T
        umma_tf32(tmem + C_DW, umma_desc_kmajor_sw128(base + fs + k * 32), umma_desc_kmajor_sw128(base + S::DT + k * 32), idesc_g, k != 0);
      umma_commit(bar_mma);
    }
    {
      // meanwhile: column `tid` of X^T (A operand of dWs^T += X^T dL)
      float x[ST_D];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(gen + xs + sw128_off(tid, c));
        x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
      }
      st_scatter_col16<ST_D>(gen + S::XT, S::XT_P, tid, x);
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
      // softmax backward, temperature gradient; L <- w, dwv <- dL (both pre-rounded to tf32: MMA operands)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float wv = __expf(L[g] * inv_tau - mx) * inv;
        const float dLp = wv * (dwv[g] - dot);
        dtau_acc = fmaf(dLp, L[g], dtau_acc);
        L[g] = st_rna(wv);
        dwv[g] = st_rna(dLp * inv_tau);
      }
      // w row -> K-major A operand of dF = w dTt (overwrites this thread's own rows of the X / F tiles);
      // dL row -> K-major A operand of dX = dL Ws; dL column -> dL^T (B operand of dWs^T += X^T dL)
#pragma unroll
      for (int c = 0; c < G / 4; ++c) {
        *reinterpret_cast<float4*>(gen + xs + (c >> 3) * ST_TILE + sw128_off(tid, c & 7)) =
            make_float4(L[4 * c], L[4 * c + 1], L[4 * c + 2], L[4 * c + 3]);
        *reinterpret_cast<float4*>(gen + dl + (c >> 3) * ST_TILE + sw128_off(tid, c & 7)) =
            make_float4(dwv[4 * c], dwv[4 * c + 1], dwv[4 * c + 2], dwv[4 * c + 3]);
      }
      st_scatter_col16<G>(gen + S::LT, G * 128, tid, dwv);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < G / 8; ++k) {   // dF = w dTt ; dX = dL Ws   (K = G)
        const uint32_t ko = (k >> 2) * ST_TILE + (k & 3) * 32, kb = (k >> 2) * 4096 + (k & 3) * 32;
        umma_tf32(tmem + C_DF, umma_desc_kmajor_sw128(base + xs + ko), umma_desc_kmajor_sw128(base + S::DTT + kb), idesc_d, k != 0);
        umma_tf32(tmem + C_DX, umma_desc_kmajor_sw128(base + dl + ko), umma_desc_kmajor_sw128(base + S::WST + kb), idesc_d, k != 0);
      }
#pragma unroll
      for (int kp = 0; kp < 2; ++kp)     // dWs^T[d][g] += sum_t X^T[d][t] dL^T[g][t]   (bf16 operands, 16 tokens per MMA)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + C_WS, umma_desc_kmajor_sw128(base + S::XT + kp * S::XT_P + k * 32),
                    umma_desc_kmajor_sw128(base + S::LT + kp * (G * 128) + k * 32), idesc_t, (iter | kp | k) != 0);
      umma_commit(bar_mma);
    }
    if (tid < G) {   // dbs[g] += sum_t dL[t][g]: row g of dL^T (bf16, 2 panels of 64 tokens)
      float a = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(gen + S::LT + kp * (G * 128) + sw128_off(tid, c));
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            a += f.x + f.y;
          }
        }
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
          for (int c = 0; c < 4; ++c) {
            __nv_bfloat162 q[4] = {__floats2bfloat162_rn(o[8 * c], o[8 * c + 1]), __floats2bfloat162_rn(o[8 * c + 2], o[8 * c + 3]),
                                   __floats2bfloat162_rn(o[8 * c + 4], o[8 * c + 5]), __floats2bfloat162_rn(o[8 * c + 6], o[8 * c + 7])};
            *reinterpret_cast<uint4*>(dst + 8 * c) = *reinterpret_cast<uint4*>(q);
          }
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
static int launch_slice_bwd_tc(const float* XF, const float* Ws, const float* bs, const float* temperature, const __nv_bfloat16* dw,
                               const float* dTt, const float* ds, __nv_bfloat16* dXF16, float* dWs_part, float* dtau_part, int B, int N,
                               int H, int groups, int clamp, cudaStream_t st) {
  CUtensorMap tm;
  int rc = encode_xf(&tm, XF, B, N, 2 * H * ST_D);
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

extern "C" int tbns_pa_slice_fwd_tc(const float* XF, const float* Ws, const float* bs, const float* temperature, void* w16, float* part,
                                    int B, int N, int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF && Ws && bs && temperature && w16 && part, "tbns_pa_slice_fwd_tc: null pointer");
  TBNS_REQUIRE(tbns_pa_slice_tc_supported(D, G), "tbns_pa_slice_fwd_tc: needs dim_head 32 and slice_num 32 or 64 (got %d, %d)", D, G);
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "tbns_pa_slice_fwd_tc: bad dims");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(XF) & 15) == 0 && (reinterpret_cast<uintptr_t>(w16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(Ws) & 15) == 0,
               "tbns_pa_slice_fwd_tc: operands must be 16-byte aligned");
  const int groups = tbns_slice_groups(B, N, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (G == 32) return launch_slice_fwd_tc<32>(XF, Ws, bs, temperature, reinterpret_cast<__nv_bfloat16*>(w16), part, B, N, H, groups, clamp, st);
  return launch_slice_fwd_tc<64>(XF, Ws, bs, temperature, reinterpret_cast<__nv_bfloat16*>(w16), part, B, N, H, groups, clamp, st);
}

extern "C" int tbns_pa_slice_bwd_tc(const float* XF, const float* Ws, const float* bs, const float* temperature, const void* dw16,
                                    const float* dTt, const float* ds, void* dXF16, float* dWs_part, float* dtau_part, int B, int N,
                                    int H, int D, int G, int clamp, void* stream) {
  TBNS_REQUIRE(XF && Ws && bs && temperature && dw16 && dTt && ds && dXF16 && dWs_part && dtau_part, "tbns_pa_slice_bwd_tc: null pointer");
  const __nv_bfloat16* dw = reinterpret_cast<const __nv_bfloat16*>(dw16);
  TBNS_REQUIRE(tbns_pa_slice_tc_supported(D, G), "tbns_pa_slice_bwd_tc: needs dim_head 32 and slice_num 32 or 64 (got %d, %d)", D, G);
  TBNS_REQUIRE(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "tbns_pa_slice_bwd_tc: bad dims");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(XF) & 15) == 0 && (reinterpret_cast<uintptr_t>(dXF16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(Ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dTt) & 15) == 0,
               "tbns_pa_slice_bwd_tc: operands must be 16-byte aligned");
  const int groups = tbns_slice_groups(B, N, H);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dXF16);
  if (G == 32) return launch_slice_bwd_tc<32>(XF, Ws, bs, temperature, dw, dTt, ds, o, dWs_part, dtau_part, B, N, H, groups, clamp, st);
  return launch_slice_bwd_tc<64>(XF, Ws, bs, temperature, dw, dTt, ds, o, dWs_part, dtau_part, B, N, H, groups, clamp, st);
}
