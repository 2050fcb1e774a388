// Micro-probe: latency / throughput of legacy mma.sync.m16n8k16 (bf16, fp32 acc) on sm_100a, per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void probe(long long* out, float* sink, int iters) {
  float acc[NACC][4];
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma(acc[j], a0, a1, a2, a3, b0, b1);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
template <int NACC>
void run(int warps, const char* what) {
  long long* d; float* sink; cudaMalloc(&d, 8 * 148); cudaMalloc(&sink, 4 * 148 * 1024);
  const int iters = 2000;
  probe<NACC><<<1, warps * 32>>>(d, sink, iters);
  probe<NACC><<<1, warps * 32>>>(d, sink, iters);
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double per = double(h) / iters;
  printf("%-28s warps %2d  indep acc %d: %.1f cycles per loop, %.2f cycles per mma per warp, %.2f cycles per mma per SM\n", what, warps, NACC,
         per, per / NACC, per / NACC / warps);
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<1>(1, "dependent chain latency");
  run<2>(1, "2 chains");
  run<4>(1, "4 chains");
  run<8>(1, "8 chains");
  run<8>(4, "4 warps (1 per SMSP)");
  run<8>(8, "8 warps");
  run<4>(8, "8 warps 4 chains");
  run<2>(8, "8 warps 2 chains");
  run<8>(16, "16 warps");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
