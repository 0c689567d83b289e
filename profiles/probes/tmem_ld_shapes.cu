// Probe: which (TMEM lane, column) does each register of tcgen05.ld.sync.aligned.16x256b.x2 return?
// Writes value = lane*1000 + column with tcgen05.st.32x32b, reads back with 16x256b at lane offsets 0 and 16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_ld_shapes.cu ; run on a B200
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(float* out) {
  __shared__ uint32_t slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  // each warp writes its 32 lanes x 16 columns
  uint32_t v[16];
  for (int j = 0; j < 16; ++j) v[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + j));
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16)),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32 + half * 16) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[((warp * 2 + half) * 32 + lane) * 8 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}
int main() {
  float* d; cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  static float h[4 * 2 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int warp = 0; warp < 2; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int lane = 0; lane < 32; ++lane) {
        printf("w%d h%d t%02d:", warp, half, lane);
        for (int j = 0; j < 8; ++j) printf(" (%3d,%2d)", (int)h[((warp * 2 + half) * 32 + lane) * 8 + j] / 1000, (int)h[((warp * 2 + half) * 32 + lane) * 8 + j] % 1000);
        printf("\n");
      }
  return 0;
}
