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
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   : TMA producer  — cp.async.bulk.tensor 4D (A, one shifted box per tap) + 2D (W) into a STAGES-deep
//              128B-swizzled shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1   : allocates TMEM, one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//              tcgen05.commit releases ring slots / publishes the accumulator
//   warps 2-5: epilogue — tcgen05.ld (32 lanes x 32 columns per warp), + bias, fp32 stores
#include <cuda.h>

#include "common.cuh"

namespace tbns {

constexpr int TC_BM = 128;      // UMMA M
constexpr int TC_BK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_UK = 16;       // UMMA K for 16-bit inputs
constexpr int TC_THREADS = 192;

struct TcParams {
  float* C;
  long long ldc;
  const float* bias;
  int Bimg, Hg, Wg, Cin, taps, flip;
  int BW, BH;          // the 128-token M tile is a BH x BW patch of the grid (BW*BH == 128)
  int tiles_w, tiles_h;
  int N;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset = 1024 B between 8-row groups [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version 1 [46,48)
  d |= (uint64_t)2 << 61;                           // layout type SWIZZLE_128B [61,64)
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using S = TcSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar_full = base + S::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_acc = bar_empty + STAGES * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + S::BAR_OFF + (2 * STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  const int mt = blockIdx.x;
  const int tw = mt % p.tiles_w;
  const int th = (mt / p.tiles_w) % p.tiles_h;
  const int bimg = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.BW, h0 = th * p.BH;
  const int n0 = blockIdx.y * BN;
  const int kc_per_tap = p.Cin / TC_BK;
  const int nkb = p.taps * kc_per_tap;

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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
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
        tma_load_2d(sb, &tmB, bar_full + s * 8, kb * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UK; ++k) {
          const uint64_t ad = umma_desc_kmajor_sw128(sa + k * TC_UK * 2);
          const uint64_t bd = umma_desc_kmajor_sw128(sb + k * TC_UK * 2);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_empty + s * 8);  // slot free once these MMAs have read it
      }
      umma_commit(bar_acc);              // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // tile row == TMEM lane
    const int hh = r / p.BW, ww = r - hh * p.BW;
    const int h = h0 + hh, w = w0 + ww;
    const bool valid = (h < p.Hg) && (w < p.Wg);
    const long long grow = ((long long)bimg * p.Hg + h) * p.Wg + w;
    float* crow = p.C + grow * p.ldc + n0;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.bias) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + n0 + c0 + j);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          *reinterpret_cast<float4*>(crow + c0 + j) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
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

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int encode_bf16(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return TBNS_ERR_CUDA;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TBNS_ERR_CUDA;
  }
  return TBNS_OK;
}

template <int BN, int STAGES>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int m_tiles, cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  TBNS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  dim3 grid(m_tiles, p.N / BN);
  gemm_tc_kernel<BN, STAGES><<<grid, TC_THREADS, S::TOTAL, st>>>(tmA, tmB, p);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_cast_bf16(const float* in, void* out, long long n, void* stream) {
  TBNS_REQUIRE(in && out && n >= 0, "tbns_cast_bf16: bad args");
  if (n == 0) return TBNS_OK;
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "tbns_cast_bf16: unaligned");
  const long long threads = (n + 7) / 8;
  cast_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out), n);
  TBNS_LAUNCH_CHECK();
  return TBNS_OK;
}

extern "C" int tbns_gemm_tc_supported(int Cin, int N, int taps) {
  return (Cin > 0 && Cin % TC_BK == 0 && N >= 64 && N % 64 == 0 && (taps == 1 || taps == 9)) ? 1 : 0;
}

extern "C" int tbns_gemm_tc(const void* A_bf16, const void* W_bf16, float* C, long long ldc, const float* bias, int Bimg, int Hg,
                            int Wg, int Cin, int N, int taps, int flip, void* stream) {
  TBNS_REQUIRE(A_bf16 && W_bf16 && C, "tbns_gemm_tc: null pointer");
  TBNS_REQUIRE(tbns_gemm_tc_supported(Cin, N, taps), "tbns_gemm_tc: unsupported shape Cin=%d N=%d taps=%d (need Cin%%64==0, N%%64==0)", Cin, N, taps);
  TBNS_REQUIRE(Bimg > 0 && Hg > 0 && Wg > 0, "tbns_gemm_tc: bad dims");
  TBNS_REQUIRE((reinterpret_cast<uintptr_t>(A_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(W_bf16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 4 == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
               "tbns_gemm_tc: operands must be 16-byte aligned");
  // pick the BH x BW token patch (BW*BH == 128) that wastes the fewest MMA rows
  int bestBW = 128;
  double bestU = -1.0;
  for (int bw = 8; bw <= 128; bw *= 2) {
    const int bh = TC_BM / bw;
    const double u = (double)Wg * Hg / ((double)cdiv(Wg, bw) * bw * (double)cdiv(Hg, bh) * bh);
    if (u > bestU + 1e-9) { bestU = u; bestBW = bw; }
  }
  TcParams p;
  p.C = C; p.ldc = ldc; p.bias = bias;
  p.Bimg = Bimg; p.Hg = Hg; p.Wg = Wg; p.Cin = Cin; p.taps = taps; p.flip = flip;
  p.BW = bestBW; p.BH = TC_BM / bestBW;
  p.tiles_w = cdiv(Wg, p.BW); p.tiles_h = cdiv(Hg, p.BH);
  p.N = N;
  const long long m_tiles = (long long)Bimg * p.tiles_w * p.tiles_h;
  TBNS_REQUIRE(m_tiles <= 0x7fffffffLL, "tbns_gemm_tc: too many tiles");

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Wg, (cuuint64_t)Hg, (cuuint64_t)Bimg};
    cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)Wg * Cin * 2, (cuuint64_t)Hg * Wg * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
    int rc = encode_bf16(&tmA, A_bf16, 4, dims, str, box);
    if (rc) return rc;
  }
  const int BN = (N % 256 == 0 && m_tiles * (N / 256) >= 148) ? 256 : (N % 128 == 0 ? 128 : 64);
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * Cin, (cuuint64_t)N};
    cuuint64_t str[1] = {(cuuint64_t)taps * Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
    int rc = encode_bf16(&tmB, W_bf16, 2, dims, str, box);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (BN == 256) return launch_tc<256, 4>(tmA, tmB, p, (int)m_tiles, st);
  if (BN == 128) return launch_tc<128, 6>(tmA, tmB, p, (int)m_tiles, st);
  return launch_tc<64, 8>(tmA, tmB, p, (int)m_tiles, st);
}
