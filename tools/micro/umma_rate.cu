// Microbenchmark: cycles per tcgen05.mma (kind::f16, cta_group::1, A and B from shared memory) as a function of the tile
// shape, one CTA, one issuing thread, R back-to-back accumulating MMAs on the same operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I asr_transformer_b200/csrc tools/micro/umma_rate.cu -o /tmp/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace asr;

__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int a_mn, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;   // 1.0h
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (uint32_t(a_mn) << 15) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
    const uint64_t ad = umma_smem_desc_sw128(smem_u32(base), a_mn ? 1024 : 16, 1024);
    const uint64_t bd = umma_smem_desc_sw128(smem_u32(base + 32768), 16, 1024);
    for (int w = 0; w < 2; ++w) {   // second round is the measured one
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
        umma_f16_ss(tmem_ptr, ad + uint64_t((r & 3) * 2), bd + uint64_t((r & 3) * 2), idesc, r != 0);
      umma_commit(&bar);
      mbar_wait(&bar, w & 1);
      const long long t1 = clock64();
      out[w] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_ptr, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 512;
  const int shapes[][3] = {{128, 16, 0}, {128, 32, 0}, {128, 64, 0}, {128, 128, 0}, {128, 256, 0}, {64, 16, 0}, {64, 64, 0},
                           {64, 256, 0}, {64, 16, 1}, {128, 16, 1}};
  for (auto& s : shapes) {
    rate_kernel<<<1, 128, 100 * 1024>>>(s[0], s[1], s[2], reps, out);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d a_mn=%d: %.1f cycles per MMA (K=16)  [%s]\n", s[0], s[1], s[2], double(h[1]) / reps,
           cudaGetErrorString(e));
  }
  return 0;
}
