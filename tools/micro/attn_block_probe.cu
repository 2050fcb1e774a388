// Micro-probe: cycles per 32-key attention block (the decoder's q K^T and P V inner blocks) for 8 warps of one CTA with the
// tiles already in shared memory: the floor a perfect ring could reach.  nvcc -O3 -arch=sm_100a attn_block_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm(uint32_t addr, uint4& r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsmt(uint32_t addr, uint4& r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
}
// MODE 0: QK block (8 ldsm + 16 mma chained + adds + max)   1: QK without the loads (constant fragments)
// MODE 2: PV block (8 ldsm.trans + 8 mma)                   3: loads only (8 ldsm, consumed by an xor)
template <int MODE, int WAIT = 0, int SKEW = 0>
__global__ void __launch_bounds__(256) probe(long long* out, float* sink, int iters) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, mat = lane >> 3, r = lane & 7;
  for (int i = threadIdx.x; i < 8 * 4096 * 4 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u & 0x3c003c00u;
  __syncthreads();
  uint4 qa[4];
  for (int k = 0; k < 4; ++k) qa[k] = make_uint4(lane + k, lane * 3 + k, lane * 5 + k, lane * 7 + k);
  float o[4][4] = {}, mx = -1e30f;
  uint32_t acc = 0;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + warp * 4096;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");   // phase 0 complete
  }
  __syncthreads();
  if (SKEW && warp >= 4) { long long t = clock64(); while (clock64() - t < SKEW) {} }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (WAIT) {   // the ring's acquire: a try_wait on a phase that is already complete
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(0u) : "memory");
      if (!ok) __trap();
    }
    const uint32_t kbase = base + (it & 3) * 32768;
    uint4 kf[4][2];
    if (MODE == 0 || MODE == 3) {
      const uint32_t row = kbase + r * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ldsm(row + j * 1024 + ((mat ^ r) << 4), kf[j][0]);
        ldsm(row + j * 1024 + (((mat + 4) ^ r) << 4), kf[j][1]);
      }
    } else if (MODE == 2) {
      const uint32_t row = kbase + ((mat >> 1) * 8 + r) * 128;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const uint32_t ch = (((2 * mt + (mat & 1)) ^ r) & 7) << 4;
        ldsmt(row + ch, kf[mt][0]);
        ldsmt(row + 2048 + ch, kf[mt][1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { kf[j][0] = make_uint4(it + j, it * 3 + j, lane + j, lane * 9 + j); kf[j][1] = make_uint4(it + j + 7, it * 5 + j, lane + j + 3, lane * 11 + j); }
    }
    if (MODE == 0 || MODE == 1) {
      float cc[4][4] = {};
#pragma unroll
      for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4& k = kf[j][kt >> 1];
          mma(cc[j], qa[kt], (kt & 1) ? k.z : k.x, (kt & 1) ? k.w : k.y);
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) { mx = fmaxf(mx, cc[j][0] + cc[j][2]); mx = fmaxf(mx, cc[j][1] + cc[j][3]); }
    } else if (MODE == 2) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) { mma(o[mt], kf[mt][0], qa[0].x, qa[0].y); mma(o[mt], kf[mt][1], qa[0].z, qa[0].w); }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc ^= kf[j][0].x ^ kf[j][0].w ^ kf[j][1].y ^ kf[j][1].z;
    }
  }
  long long t1 = clock64();
  float s = mx + float(acc);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += o[i][j];
  sink[blockIdx.x * 256 + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
template <int MODE, int WAIT = 0, int SKEW = 0>
void run(int warps, const char* what) {
  long long* d; float* sink; cudaMalloc(&d, 8 * 148); cudaMalloc(&sink, 4 * 148 * 256);
  cudaFuncSetAttribute(probe<MODE, WAIT, SKEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) probe<MODE, WAIT, SKEW><<<1, warps * 32, 131072>>>(d, sink, iters);
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps %d: %.1f cycles per block\n", what, warps, double(h) / iters);
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<0>(8, "QK block: 8 ldsm + 16 mma + adds");
  run<0>(4, "QK block: 8 ldsm + 16 mma + adds");
  run<1>(8, "QK block without loads");
  run<2>(8, "PV block: 8 ldsm.trans + 8 mma");
  run<2>(4, "PV block: 8 ldsm.trans + 8 mma");
  run<3>(8, "loads only (8 ldsm)");
  run<3>(4, "loads only (8 ldsm)");
  run<0, 1, 0>(8, "QK block + try_wait, warps in step");
  run<0, 1, 170>(8, "QK block + try_wait, warps 4-7 skewed 170");
  run<0, 1, 250>(8, "QK block + try_wait, warps 4-7 skewed 250");
  run<2, 1, 0>(8, "PV block + try_wait, warps in step");
  run<2, 1, 130>(8, "PV block + try_wait, warps 4-7 skewed 130");
  run<0, 1, 0>(4, "QK block + try_wait, 4 warps");
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
