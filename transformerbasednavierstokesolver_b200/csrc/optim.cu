// AdamW over FLAT parameter / gradient / moment buffers (exp_ns.py:172-173 builds torch.optim.AdamW; its step runs once per
// batch at exp_ns.py:208-209).  The model's ~180 parameter tensors are views into one fp32 buffer (train.FlatAdamW), the
// gradients already live in one flat buffer (train.FlatGradients), so the whole optimizer step is ONE elementwise pass at HBM
// speed (7 streams of 4 bytes per parameter) instead of a dozen multi-tensor launches.  The hyper-parameters of the step come
// from a small device array written by the host before the launch, so the launch can be replayed from a CUDA graph while a
// scheduler (OneCycleLR: lr AND beta1 change every step) keeps working.
//   hp[0] = lr, hp[1] = beta1, hp[2] = beta2, hp[3] = eps, hp[4] = weight_decay, hp[5] = 1 - beta1^t, hp[6] = 1 - beta2^t
// Update (torch.optim.AdamW, amsgrad = False, maximize = False):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
#include "common.cuh"

namespace tbns {

__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, const float* __restrict__ hp, long long n4, long long n) {
  pdl_sync();
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], bc1 = hp[5], bc2 = hp[6];
  const float decay = 1.0f - lr * wd, step = lr / bc1, rs2 = rsqrtf(bc2), c1 = 1.0f - b1, c2 = 1.0f - b2;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    pp *= decay;
    mm = fmaf(b1, mm, c1 * gg);
    vv = fmaf(b2, vv, c2 * gg * gg);
    pp -= step * mm / fmaf(sqrtf(vv), rs2, eps);
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n not a multiple of 4)
  for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

}  // namespace tbns

using namespace tbns;

extern "C" int tbns_adamw_flat(float* p, const float* g, float* m, float* v, const float* hp, long long n, void* stream) {
  TBNS_REQUIRE(p && g && m && v && hp && n >= 0, "tbns_adamw_flat: bad args");
  if (n == 0) return TBNS_OK;
  TBNS_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0,
               "tbns_adamw_flat: buffers must be 16-byte aligned");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  TBNS_CUDA(launch_pdl(adamw_flat_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, hp, n4, n));
  return TBNS_OK;
}
