// Y[M,N] = X[M,K] * W[N,K]^T with fused bias / ReLU / positional-encoding / residual epilogue.
//
// Replaces every nn.Linear call site of the reference forward path (model.py:47,123; layers.py:16-18,40,54-57)
// for the batched (encoder / teacher-forced decoder / cross-K/V) case.
//
// sm_100a design: one 128 x BN output tile per CTA.
//   warp 4  : TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B, 64-wide K blocks, STAGES-deep mbarrier ring)
//   warp 5  : UMMA issuer   (one elected thread, tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warps 0-3: epilogue     (tcgen05.ld 32 lanes x 32 columns per warp -> registers -> fused epilogue -> HBM)
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace asr {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;

template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
  return size_t(STAGES) * (A_STAGE_BYTES + BN * BK * 2) + (2 * STAGES + 4) * 8 + 32 + 2 * BN * 4 + 1024;
}

// Compile-time specialised epilogue of one full, vector-aligned 32-column chunk (the common case): only the loads,
// math and stores of this GEMM's epilogue are emitted.  FLAGS: 1 bias, 2 relu, 4 positional rowvec, 8 residual,
// 16 fp32 out, 32 f16 out, 64 f16 out as a hi | lo pair (lo = f16(v - hi) stored f16_lo_off elements further: the
// A operand of a following a_split GEMM, fp32-accurate to 2^-22).
enum { EF_BIAS = 1, EF_RELU = 2, EF_PE = 4, EF_RES = 8, EF_F32 = 16, EF_H16 = 32, EF_SPLIT = 64 };
template <int FLAGS>
__device__ __forceinline__ void epilogue_chunk_fast(const GemmEpilogue& ep, const uint32_t (&r)[32], int row, int col0,
                                                    const float* sbias_chunk) {
  float4 res[8], pe[8];
  if (FLAGS & EF_RES) {
    const float4* rp = reinterpret_cast<const float4*>(ep.residual + size_t(row) * ep.ld_res + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) res[j] = rp[j];
  }
  if (FLAGS & EF_PE) {
    const float4* pp = reinterpret_cast<const float4*>(ep.rowvec + size_t(row % ep.rowvec_period) * ep.ld_rowvec + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) pe[j] = __ldg(pp + j);
  }
  float* of = (FLAGS & EF_F32) ? ep.out_f32 + size_t(row) * ep.ld_f32 + col0 : nullptr;
  f16* ob = (FLAGS & EF_H16) ? ep.out_f16 + size_t(row) * ep.ld_f16 + col0 : nullptr;
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[4 * j + i]);
    if (FLAGS & EF_BIAS) {
      const float4 b0 = *reinterpret_cast<const float4*>(sbias_chunk + 4 * j);
      const float4 b1 = *reinterpret_cast<const float4*>(sbias_chunk + 4 * j + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (FLAGS & EF_RELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (FLAGS & EF_PE) {
      v[0] += pe[j].x; v[1] += pe[j].y; v[2] += pe[j].z; v[3] += pe[j].w;
      v[4] += pe[j + 1].x; v[5] += pe[j + 1].y; v[6] += pe[j + 1].z; v[7] += pe[j + 1].w;
    }
    if (FLAGS & EF_RES) {
      v[0] += res[j].x; v[1] += res[j].y; v[2] += res[j].z; v[3] += res[j].w;
      v[4] += res[j + 1].x; v[5] += res[j + 1].y; v[6] += res[j + 1].z; v[7] += res[j + 1].w;
    }
    if (FLAGS & EF_F32) {
      *reinterpret_cast<float4*>(of + 4 * j) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(of + 4 * j + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (FLAGS & EF_H16) {
      uint4 hi = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(ob + 4 * j) = hi;
      if (FLAGS & EF_SPLIT)
        *reinterpret_cast<uint4*>(ob + ep.f16_lo_off + 4 * j) =
            make_uint4(f16x2_residual(v[0], v[1], hi.x), f16x2_residual(v[2], v[3], hi.y),
                       f16x2_residual(v[4], v[5], hi.z), f16x2_residual(v[6], v[7], hi.w));
    }
  }
}

// sbias: this CTA's bias slice in shared memory (nullptr = no bias), indexed by column - n0
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& ep, const uint32_t (&r)[32], int row, int col0,
                                               int n_store, const float* sbias, int n0) {
  // 32 consecutive columns of one row, processed 4 at a time.
  const bool res_vec = ep.residual && (ep.ld_res % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
  const bool f32_vec = ep.out_f32 && (ep.ld_f32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0);
  const bool h16_vec = ep.out_f16 && (ep.ld_f16 % 4 == 0) && (ep.f16_lo_off % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(ep.out_f16) & 7) == 0);
  const float* pe_row = ep.rowvec ? ep.rowvec + size_t(row % ep.rowvec_period) * ep.ld_rowvec : nullptr;
  if (col0 + 32 <= n_store && (!ep.residual || res_vec) && (!ep.out_f32 || f32_vec) && (!ep.out_f16 || h16_vec) &&
      (!pe_row || ((ep.ld_rowvec % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.rowvec) & 15) == 0)))) {
    // fast path: every global load of the chunk is in flight before the first use (the epilogue is latency-bound)
    float4 res[8], pe[8];
    if (ep.residual) {
      const float4* rp = reinterpret_cast<const float4*>(ep.residual + size_t(row) * ep.ld_res + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) res[j] = rp[j];
    }
    if (pe_row) {
      const float4* pp = reinterpret_cast<const float4*>(pe_row + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) pe[j] = __ldg(pp + j);
    }
    const bool h16_wide = ep.out_f16 && (ep.ld_f16 % 8 == 0) && ((reinterpret_cast<uintptr_t>(ep.out_f16) & 15) == 0);
    uint32_t pk[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[4 * j + i]);
      if (sbias) {
        const float4 b = *reinterpret_cast<const float4*>(sbias + col0 - n0 + 4 * j);
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      if (ep.relu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (pe_row) { v[0] += pe[j].x; v[1] += pe[j].y; v[2] += pe[j].z; v[3] += pe[j].w; }
      if (ep.residual) { v[0] += res[j].x; v[1] += res[j].y; v[2] += res[j].z; v[3] += res[j].w; }
      if (ep.out_f32)
        *reinterpret_cast<float4*>(ep.out_f32 + size_t(row) * ep.ld_f32 + col0 + 4 * j) = make_float4(v[0], v[1], v[2], v[3]);
      if (ep.out_f16) {
        uint2 t;
        t.x = pack_f16x2(v[0], v[1]);
        t.y = pack_f16x2(v[2], v[3]);
        if (ep.f16_lo_off)
          *reinterpret_cast<uint2*>(ep.out_f16 + size_t(row) * ep.ld_f16 + ep.f16_lo_off + col0 + 4 * j) =
              make_uint2(f16x2_residual(v[0], v[1], t.x), f16x2_residual(v[2], v[3], t.y));
        if (!h16_wide) {
          *reinterpret_cast<uint2*>(ep.out_f16 + size_t(row) * ep.ld_f16 + col0 + 4 * j) = t;
        } else if (j & 1) {      // 16-byte stores: half the L2 write transactions of this row-per-thread layout
          *reinterpret_cast<uint4*>(ep.out_f16 + size_t(row) * ep.ld_f16 + col0 + 4 * (j - 1)) =
              make_uint4(pk[0], pk[1], t.x, t.y);
        } else {
          pk[0] = t.x;
          pk[1] = t.y;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const int col = col0 + j;
    if (col >= n_store) break;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[j + i]);
    const bool full = (col + 4 <= n_store);
    if (ep.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (full || col + i < n_store) v[i] += __ldg(ep.bias + col + i);
    }
    if (ep.relu) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (pe_row) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (full || col + i < n_store) v[i] += __ldg(pe_row + col + i);
    }
    if (ep.residual) {
      const float* rp = ep.residual + size_t(row) * ep.ld_res + col;
      if (full && res_vec) {
        const float4 t = *reinterpret_cast<const float4*>(rp);
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (col + i < n_store) v[i] += rp[i];
      }
    }
    if (ep.out_f32) {
      float* op = ep.out_f32 + size_t(row) * ep.ld_f32 + col;
      if (full && f32_vec) {
        *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (col + i < n_store) op[i] = v[i];
      }
    }
    if (ep.out_f16) {
      f16* op = ep.out_f16 + size_t(row) * ep.ld_f16 + col;
      if (full && h16_vec) {
        uint2 t;
        t.x = pack_f16x2(v[0], v[1]);
        t.y = pack_f16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(op) = t;
        if (ep.f16_lo_off)
          *reinterpret_cast<uint2*>(op + ep.f16_lo_off) =
              make_uint2(f16x2_residual(v[0], v[1], t.x), f16x2_residual(v[2], v[3], t.y));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (col + i < n_store) {
            const f16 hh = f16_sat(v[i]);
            op[i] = hh;
            if (ep.f16_lo_off) op[ep.f16_lo_off + i] = f16_sat(v[i] - __half2float(hh));
          }
      }
    }
  }
}

// Persistent kernel: one CTA per SM walks the output tiles (n fastest, so concurrently processed tiles share their A
// rows in L2); the fp32 accumulator is double-buffered in TMEM, so the epilogue of tile i (8 warps: 4 lane quadrants x
// 2 column halves) overlaps the TMA loads and MMAs of tile i+1.  These GEMMs have short K (256 .. 1216): what bounds
// them is the epilogue and per-tile fixed latency, not the tensor pipe - hence persistent + overlapped, not bigger tiles.
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = (EPI_WARPS + 2) * 32;

// TMAS (bias + f16 output, BN = 128 only: the packed QKV and cross-K/V GEMMs): the 32 x 64 f16 block of every epilogue warp
// leaves as ONE tensor store from a swizzled 4 KB staging tile instead of 8 row-per-thread 16-byte stores per thread
// (32 sectors per instruction through the LSU).
template <int BN, int STAGES, int FLAGS, bool TMAS = false>   // FLAGS >= 0: the specialised epilogue this launch uses; -1: generic
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, GemmEpilogue ep, int M, int n_store, int K, int nkw, int tiles_n,
               int n_tiles) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  static_assert(!TMAS || (BN == 128 && FLAGS == (EF_BIAS | EF_H16)), "tensor-store epilogue: bias + f16 output, BN 128");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* stiles = sB + STAGES * B_STAGE_BYTES;                     // (TMAS) [EPI_WARPS][4 KB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stiles + (TMAS ? EPI_WARPS * 4096 : 0));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 1) + 15) & ~uintptr_t(15));   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = K / BK;

  if (warp == EPI_WARPS + 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == EPI_WARPS) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kb * BK, m0);
          tma_load_2d(sB + s * B_STAGE_BYTES, &tmB, &full_bar[s], (kb < nkw ? kb : kb - nkw) * BK, n0);   // a_split: W twice
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ------------------------------------------------ UMMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0, 0);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1, use = lt >> 1;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(sB + s * B_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)   // +32 B per 16-element K step inside the 128 B swizzle row
            umma_f16_ss(acc, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc, (kb | k) != 0);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue: warp w -> TMEM lanes 32 (w % 4).., columns half w / 4
    const int quad = warp & 3, half = warp >> 2;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      float* sb = sbias + buf * BN;
      if (ep.bias)
        for (int i = threadIdx.x; i < BN; i += EPI_WARPS * 32) sb[i] = (n0 + i < n_store) ? __ldg(ep.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // bias slice visible to the epilogue warps
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      if (TMAS) {
        uint4* tl = reinterpret_cast<uint4*>(stiles + warp * 4096);
        if (lane == 0) tma_store_wait_read<0>();                        // the previous tile's store (issued a tile ago)
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = half * 64 + c * 32;
          uint32_t r[32];
          tmem_ld32(tmem_base + buf * BN + (uint32_t(quad * 32) << 16) + uint32_t(col), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]) + sb[col + 8 * j + i];
            tl[lane * 8 + ((c * 4 + j) ^ (lane & 7))] =
                make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO, tl, n0 + half * 64, m0 + quad * 32);      // rows >= M / columns >= n_store: clipped by the map
          tma_store_commit();
        }
      } else
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col = half * (BN / 2) + c * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + buf * BN + (uint32_t(quad * 32) << 16) + uint32_t(col), r);
        tmem_ld_wait();
        if (row < M) {
          const int c0 = n0 + col;
          if (FLAGS >= 0 && c0 + 32 <= n_store) epilogue_chunk_fast<(FLAGS >= 0 ? FLAGS : 0)>(ep, r, row, c0, sb + col);
          else epilogue_chunk(ep, r, row, c0, n_store, ep.bias ? sb : nullptr, n0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    if (TMAS && lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 2 * BN);
}

template <int BN, int STAGES, int FLAGS>
int launch_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int n_store, int n_pad,
                int K, int nkw, cudaStream_t s) {
  int n_slots = 0;   // persistent grid size: one CTA per SM (98 KB+ of shared memory each)
  if (int rc = device_props(&n_slots, nullptr)) return rc;
  const int tiles_n = n_pad / BN, n_tiles = tiles_n * ((M + BM - 1) / BM);
  const int grid = n_tiles < n_slots ? n_tiles : n_slots;
  if constexpr (BN == 128 && FLAGS == (EF_BIAS | EF_H16)) {   // f16 output by tensor stores (ASR_B200_LN_TMA=0: off)
    static const bool tma_off = [] {
      const char* e = std::getenv("ASR_B200_LN_TMA");
      return e && e[0] == '0';
    }();
    CUtensorMap tmO;
    const uint64_t dims[2] = {(uint64_t)n_store, (uint64_t)M}, str[2] = {2, (uint64_t)ep.ld_f16 * 2};
    const uint32_t box[2] = {64, 32};
    if (!tma_off && n_store % 8 == 0 && make_tmap_f16(&tmO, ep.out_f16, 2, dims, str, box, nullptr) == 0) {
      auto kern = gemm_tc_kernel<BN, STAGES, FLAGS, true>;
      constexpr size_t smem = gemm_smem_bytes<BN, STAGES>() + EPI_WARPS * 4096;
      if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
      kern<<<grid, GEMM_THREADS, smem, s>>>(tmA, tmB, tmO, ep, M, n_store, K, nkw, tiles_n, n_tiles);
      ASR_CUDA_OK(cudaGetLastError());
      ASR_LAUNCHED(1);
      return 0;
    }
  }
  auto kern = gemm_tc_kernel<BN, STAGES, FLAGS, false>;
  constexpr size_t smem = gemm_smem_bytes<BN, STAGES>();
  if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
  kern<<<grid, GEMM_THREADS, smem, s>>>(tmA, tmB, tmA, ep, M, n_store, K, nkw, tiles_n, n_tiles);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

// pick the compile-time epilogue: the flag combinations of this path's GEMMs when every pointer / leading dimension
// allows 16-byte accesses, the generic (fully predicated) epilogue otherwise
template <int BN, int STAGES>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int n_store, int n_pad,
               int K, int nkw, cudaStream_t s) {
  const int flags = (ep.bias ? EF_BIAS : 0) | (ep.relu ? EF_RELU : 0) | (ep.rowvec ? EF_PE : 0) |
                    (ep.residual ? EF_RES : 0) | (ep.out_f32 ? EF_F32 : 0) | (ep.out_f16 ? EF_H16 : 0) |
                    ((ep.out_f16 && ep.f16_lo_off) ? EF_SPLIT : 0);
  const bool aligned =
      (!ep.residual || ((ep.ld_res % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0))) &&
      (!ep.out_f32 || ((ep.ld_f32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0))) &&
      (!ep.out_f16 || ((ep.ld_f16 % 8 == 0) && (ep.f16_lo_off % 8 == 0) &&
                       ((reinterpret_cast<uintptr_t>(ep.out_f16) & 15) == 0))) &&
      (!ep.rowvec || ((ep.ld_rowvec % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.rowvec) & 15) == 0)));
  if (aligned) {
    switch (flags) {
      case EF_BIAS | EF_H16: return launch_inst<BN, STAGES, EF_BIAS | EF_H16>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
      case EF_BIAS | EF_RELU | EF_H16:
        return launch_inst<BN, STAGES, EF_BIAS | EF_RELU | EF_H16>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
      case EF_BIAS | EF_RELU | EF_H16 | EF_SPLIT:
        return launch_inst<BN, STAGES, EF_BIAS | EF_RELU | EF_H16 | EF_SPLIT>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
      case EF_BIAS | EF_RES | EF_F32:
        return launch_inst<BN, STAGES, EF_BIAS | EF_RES | EF_F32>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
      case EF_BIAS | EF_PE | EF_F32:
        return launch_inst<BN, STAGES, EF_BIAS | EF_PE | EF_F32>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
      case EF_F32: return launch_inst<BN, STAGES, EF_F32>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
      default: break;
    }
  }
  return launch_inst<BN, STAGES, -1>(tmA, tmB, ep, M, n_store, n_pad, K, nkw, s);
}

// ------------------------------------------------------------------------------------------------
// Full-row GEMM + LayerNorm: one 128 x 256 output tile per CTA step holds WHOLE rows of the model dimension, so the
// LayerNorm that follows every out projection / FFN / input projection of the pre-LN layers (model.py:20,23,52,67,70,73)
// runs in the epilogue instead of in its own launch: pass 1 reads the accumulator, adds bias / positional encoding /
// residual, stores the new residual row (fp32) and writes the value back to TMEM (tcgen05.st) while accumulating shifted
// row statistics; the two column halves of a row (two warps) combine their (mean, M2) through shared memory (Chan's
// formula); pass 2 re-reads TMEM, normalises and stores the next GEMM's A operand (fp16, hi | lo pair when split).
// 48 KB per pipeline stage (A 128 x 64 + B 256 x 64), accumulator double-buffered: all 512 TMEM columns.
constexpr int LN_BN = 256;
constexpr int LN_B_STAGE = LN_BN * BK * 2;
template <int STAGES, bool TMAE = false>
constexpr size_t gemm_ln_smem() {
  return size_t(STAGES) * (A_STAGE_BYTES + LN_B_STAGE) + (TMAE ? EPI_WARPS * (3 * 4096 + 2 * 8) : 0) + (2 * STAGES + 4) * 8 +
         32 + 3 * LN_BN * 4 + 2 * BM * 2 * 8 + 1024;
}

// Epilogue of one full-row (256-column) tile for the thread that owns `row` and the column half `half` (128 columns at
// TMEM address tacc): see gemm_ln_kernel.  st: this row's two (mean, M2) slots in shared memory (double-buffered by the
// caller); all EPI_WARPS * 32 epilogue threads call this together (one named barrier inside).
__device__ __forceinline__ void ln_epilogue_tile(const GemmEpilogue& ep, const LnEpilogue& ln, uint32_t tacc, int row,
                                                 bool live, int half, const float* sbias, const float* sgamma,
                                                 const float* sbeta, float2* st) {
      const float* pe_row = ep.rowvec ? ep.rowvec + size_t(row % ep.rowvec_period) * ep.ld_rowvec : nullptr;
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col = half * 128 + c * 32;
        uint32_t r[32];
        tmem_ld32(tacc + uint32_t(c * 32), r);
        tmem_ld_wait();
        float4 res[8], pe[8];
        if (ep.residual && live) {
          const float4* rp = reinterpret_cast<const float4*>(ep.residual + size_t(row) * ep.ld_res + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) res[j] = rp[j];
        }
        if (pe_row && live) {
          const float4* pp = reinterpret_cast<const float4*>(pe_row + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) pe[j] = __ldg(pp + j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[4 * j + i]) + sbias[col + 4 * j + i];
          if (ep.relu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (pe_row && live) { v[0] += pe[j].x; v[1] += pe[j].y; v[2] += pe[j].z; v[3] += pe[j].w; }
          if (ep.residual && live) { v[0] += res[j].x; v[1] += res[j].y; v[2] += res[j].z; v[3] += res[j].w; }
          if (ep.out_f32 && live)
            *reinterpret_cast<float4*>(ep.out_f32 + size_t(row) * ep.ld_f32 + col + 4 * j) = make_float4(v[0], v[1], v[2], v[3]);
          if (c == 0 && j == 0) shift = v[0];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float d = v[i] - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
            r[4 * j + i] = __float_as_uint(v[i]);
          }
        }
        tmem_st32(tacc + uint32_t(c * 32), r);      // the finished row values go back to TMEM for the second pass
      }
      tmem_st_wait();
      const float mean_h = shift + s1 * (1.0f / 128.0f);
      const float m2_h = fmaxf(s2 - s1 * s1 * (1.0f / 128.0f), 0.f);
      st[half] = make_float2(mean_h, m2_h);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      const float2 other = st[half ^ 1];
      const float mean = 0.5f * (mean_h + other.x);
      const float dm = mean_h - other.x;
      const float var = (m2_h + other.y + dm * dm * 64.0f) * (1.0f / 256.0f);
      const float rstd = 1.0f / sqrtf(var + ln.eps);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col = half * 128 + c * 32;
        uint32_t r[32];
        tmem_ld32(tacc + uint32_t(c * 32), r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              y[i] = (__uint_as_float(r[8 * j + i]) - mean) * rstd * sgamma[col + 8 * j + i] + sbeta[col + 8 * j + i];
            if (ln.out_f32) {
              float* op = ln.out_f32 + size_t(row) * LN_BN + col + 8 * j;
              *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
            if (ln.out_f16) {
              f16* op = ln.out_f16 + size_t(row) * (ln.split ? 2 * LN_BN : LN_BN) + col + 8 * j;
              const uint4 hi = make_uint4(pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]), pack_f16x2(y[4], y[5]),
                                          pack_f16x2(y[6], y[7]));
              *reinterpret_cast<uint4*>(op) = hi;
              if (ln.split)
                *reinterpret_cast<uint4*>(op + LN_BN) =
                    make_uint4(f16x2_residual(y[0], y[1], hi.x), f16x2_residual(y[2], y[3], hi.y),
                               f16x2_residual(y[4], y[5], hi.z), f16x2_residual(y[6], y[7], hi.w));
            }
          }
        }
      }
}

// The same epilogue with every row stream carried by TMA (gemm_ln_kernel<.., true>): one row per thread means a direct
// global access sends 32 separate sectors through the LSU per instruction, and the three row streams of this epilogue
// (residual in, fp32 residual out, f16 hi | lo out: 65 MB each at 256 utterances) then cost 20-30 us each against 10 us at
// the HBM roofline (ablation in DESIGN.md section 8).  Here a warp owns three 4 KB staging tiles ([32 rows][128 B],
// 128-byte swizzle: chunk j of row r at j ^ (r & 7), conflict free for the row-per-thread view): lane 0 requests the
// residual box of chunk c + 1 (32 rows x 32 fp32) while chunk c is processed, the finished fp32 chunk and the f16 hi / lo
// halves are written to a tile and leave as tensor stores.  Rows beyond M are clipped / zero-filled by the tensor maps.
struct LnTma {
  const CUtensorMap* res;     // fp32 [M][ld_res], box {32, 32}
  const CUtensorMap* out32;   // fp32 [M][ld_f32], box {32, 32}
  const CUtensorMap* out16;   // f16  [M][ldo],    box {64, 32}
  uint8_t* tile[3];           // this warp's three 4 KB staging tiles (1024-byte aligned)
  uint64_t* bars;             // this warp's 2 mbarriers (residual double buffer)
  uint32_t* n_loads;          // this warp's count of residual boxes requested so far (buffer = n & 1, parity = n >> 1 & 1)
};
__device__ __forceinline__ void ln_epilogue_tile_tma(const GemmEpilogue& ep, const LnEpilogue& ln, uint32_t tacc, int row0,
                                                     int M, int half, const float* sbias, const float* sgamma,
                                                     const float* sbeta, float2* st, const LnTma& t, uint32_t& n_loads) {
      const int lane = threadIdx.x & 31, row = row0 + lane;
      const bool live = row < M;
      const float* pe_row = ep.rowvec ? ep.rowvec + size_t(row % ep.rowvec_period) * ep.ld_rowvec : nullptr;
      auto tile = [&](int i) { return reinterpret_cast<uint4*>(t.tile[i]); };
      auto sw = [&](int j) { return lane * 8 + (j ^ (lane & 7)); };
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      if (ep.residual && lane == 0) {
        tma_store_wait_read<0>();                   // the previous tile's last stores have left tiles 0 / 1
        const uint32_t b = n_loads & 1u;
        mbar_expect_tx(&t.bars[b], 4096);
        tma_load_2d(t.tile[b], t.res, &t.bars[b], half * 128, row0);
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col = half * 128 + c * 32;
        uint32_t r[32];
        tmem_ld32(tacc + uint32_t(c * 32), r);
        float4 res[8], pe[8];
        if (ep.residual) {
          const uint32_t k = n_loads + c, b = k & 1u;
          __syncwarp();                             // every lane is done with the tile the next request overwrites
          if (lane == 0 && c + 1 < 4) {
            mbar_expect_tx(&t.bars[b ^ 1u], 4096);
            tma_load_2d(t.tile[b ^ 1u], t.res, &t.bars[b ^ 1u], col + 32, row0);
          }
          mbar_wait(&t.bars[b], (k >> 1) & 1u);
          const uint4* rt = tile(b);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 v = rt[sw(j)];
            res[j] = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
          }
        }
        if (pe_row && live) {
          const float4* pp = reinterpret_cast<const float4*>(pe_row + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) pe[j] = __ldg(pp + j);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[4 * j + i]) + sbias[col + 4 * j + i];
          if (ep.relu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (pe_row && live) { v[0] += pe[j].x; v[1] += pe[j].y; v[2] += pe[j].z; v[3] += pe[j].w; }
          if (ep.residual) { v[0] += res[j].x; v[1] += res[j].y; v[2] += res[j].z; v[3] += res[j].w; }
          if (c == 0 && j == 0) shift = v[0];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float d = v[i] - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
            r[4 * j + i] = __float_as_uint(v[i]);
          }
        }
        if (ep.out_f32) {
          if (lane == 0) tma_store_wait_read<0>();  // the previous chunk's store has read tile 2 (issued a chunk ago)
          __syncwarp();
          uint4* ot = tile(2);
#pragma unroll
          for (int j = 0; j < 8; ++j) ot[sw(j)] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(t.out32, ot, col, row0);
            tma_store_commit();
          }
        }
        tmem_st32(tacc + uint32_t(c * 32), r);      // the finished row values go back to TMEM for the second pass
      }
      if (ep.residual) n_loads += 4;
      tmem_st_wait();
      const float mean_h = shift + s1 * (1.0f / 128.0f);
      const float m2_h = fmaxf(s2 - s1 * s1 * (1.0f / 128.0f), 0.f);
      st[half] = make_float2(mean_h, m2_h);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      const float2 other = st[half ^ 1];
      const float mean = 0.5f * (mean_h + other.x);
      const float dm = mean_h - other.x;
      const float var = (m2_h + other.y + dm * dm * 64.0f) * (1.0f / 256.0f);
      const float rstd = 1.0f / sqrtf(var + ln.eps);
      int n_st = 0;                                  // stores of this pass: tile n_st % 3, at most 2 still reading
#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {               // 64 columns per round: one 128-byte f16 row per tile
        const int col = half * 128 + cp * 64;
        uint4* th = tile(n_st % 3);
        uint4* tl = tile((n_st + 1) % 3);
        if (ln.out_f16) {
          if (lane == 0) tma_store_wait_read<1>();   // both tiles of this round are free (their stores: >= 2 groups back)
          __syncwarp();
        }
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
          const int c = cp * 2 + h2;
          uint32_t r[32];
          tmem_ld32(tacc + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cc = half * 128 + c * 32 + 8 * j;
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              y[i] = (__uint_as_float(r[8 * j + i]) - mean) * rstd * sgamma[cc + i] + sbeta[cc + i];
            if (ln.out_f32 && live) {                                // (the model's last LayerNorm only: row per thread)
              float* op = ln.out_f32 + size_t(row) * LN_BN + cc;
              *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
            if (ln.out_f16) {
              const uint4 hi = make_uint4(pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]), pack_f16x2(y[4], y[5]),
                                          pack_f16x2(y[6], y[7]));
              th[sw(h2 * 4 + j)] = hi;
              if (ln.split)
                tl[sw(h2 * 4 + j)] = make_uint4(f16x2_residual(y[0], y[1], hi.x), f16x2_residual(y[2], y[3], hi.y),
                                                f16x2_residual(y[4], y[5], hi.z), f16x2_residual(y[6], y[7], hi.w));
            }
          }
        }
        if (ln.out_f16) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(t.out16, th, col, row0);
            tma_store_commit();
            if (ln.split) {
              tma_store_2d(t.out16, tl, LN_BN + col, row0);
              tma_store_commit();
            }
          }
          n_st += 2;
        }
      }
}

template <int STAGES, bool TMAE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmO32,
               const __grid_constant__ CUtensorMap tmO16, GemmEpilogue ep, LnEpilogue ln, int M, int K, int nkw, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* stiles = sB + STAGES * LN_B_STAGE;                          // (TMAE) [EPI_WARPS][3][4 KB] epilogue staging tiles
  uint64_t* tbars = reinterpret_cast<uint64_t*>(stiles + (TMAE ? EPI_WARPS * 3 * 4096 : 0));   // (TMAE) [EPI_WARPS][2]
  uint64_t* full_bar = tbars + (TMAE ? EPI_WARPS * 2 : 0);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 1) + 15) & ~uintptr_t(15));   // [256]
  float* sgamma = sbias + LN_BN;
  float* sbeta = sgamma + LN_BN;
  float2* sstat = reinterpret_cast<float2*>(sbeta + LN_BN);     // [tile parity][row][column half] (mean, M2)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = K / BK;

  if (warp == EPI_WARPS + 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], EPI_WARPS);
    }
    if (TMAE)
      for (int b = 0; b < EPI_WARPS * 2; ++b) mbar_init(&tbars[b], 1);
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (TMAE) {
        tma_prefetch_desc(&tmRes);
        tma_prefetch_desc(&tmO32);
        tma_prefetch_desc(&tmO16);
      }
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 2 * LN_BN);
    tmem_relinquish();
  }
  if (threadIdx.x < LN_BN) {
    sbias[threadIdx.x] = ep.bias ? __ldg(ep.bias + threadIdx.x) : 0.f;
    sgamma[threadIdx.x] = __ldg(ln.gamma + threadIdx.x);
    sbeta[threadIdx.x] = __ldg(ln.beta + threadIdx.x);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == EPI_WARPS) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = tile * BM;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + LN_B_STAGE);
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kb * BK, m0);
          tma_load_2d(sB + s * LN_B_STAGE, &tmB, &full_bar[s], (kb < nkw ? kb : kb - nkw) * BK, 0);   // a_split: W twice
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, LN_BN, 0, 0);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1, use = lt >> 1;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * LN_BN;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(sB + s * LN_B_STAGE), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss(acc, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc, (kb | k) != 0);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ---- epilogue: warp w -> TMEM lanes 32 (w % 4).., column half w / 4 (128 columns of the row)
    const int quad = warp & 3, half = warp >> 2;
    const int rl = quad * 32 + lane;
    uint32_t lt = 0, n_loads = 0;
    LnTma tm;
    tm.res = &tmRes; tm.out32 = &tmO32; tm.out16 = &tmO16;
    for (int i = 0; i < 3; ++i) tm.tile[i] = stiles + (warp * 3 + i) * 4096;
    tm.bars = tbars + warp * 2; tm.n_loads = nullptr;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      const int row = tile * BM + rl;
      const bool live = row < M;
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * LN_BN + (uint32_t(quad * 32) << 16) + uint32_t(half * 128);
      if (TMAE)
        ln_epilogue_tile_tma(ep, ln, tacc, tile * BM + quad * 32, M, half, sbias, sgamma, sbeta,
                             sstat + (size_t(buf) * BM + rl) * 2, tm, n_loads);
      else
        ln_epilogue_tile(ep, ln, tacc, row, live, half, sbias, sgamma, sbeta, sstat + (size_t(buf) * BM + rl) * 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    if (TMAE && lane == 0) tma_store_wait_all<0>();   // the last tensor stores have left shared memory and are complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 2 * LN_BN);
}

// ------------------------------------------------------------------------------------------------
// Fused position-wise FFN (reference layers.py:53-58 + the residual add of model.py:24,74 + the LayerNorm that follows):
//   h += W2 relu(W1 x + b1) + b2 ;  y = LayerNorm(h)
// for one 128-row tile per CTA step, the (rows x FF) hidden activation never leaving the SM.  FF is walked in chunks of
// 128 columns: GEMM 1 (x tile [128 x K1] . W1 chunk^T) accumulates a 128 x 128 chunk in TMEM (double-buffered), the
// epilogue warps add bias, apply ReLU and write the chunk as an fp16 hi | lo A operand (four swizzled 128 x 64 k-blocks)
// into shared memory, GEMM 2 (hidden chunk . W2[:, chunk]^T, the lo k-blocks against the same W2 columns) accumulates the
// full 128 x 256 output row tile in the other 256 TMEM columns; the MMA thread issues GEMM 1 of chunk c+1 before GEMM 2
// of chunk c, so the tensor pipe works while the epilogue converts.  After the last chunk the full-row epilogue of
// gemm_ln_kernel (bias + residual -> h, LayerNorm -> next operand) runs on the output tile.
// Ring stages of 32 KB: (x k-block 16 KB + W1 k-block 16 KB) for GEMM 1, one W2 k-block (256 x 64) for GEMM 2.
constexpr int FFN_STAGE = 32768;
constexpr int FFN_HID_BYTES = 4 * A_STAGE_BYTES;      // hi k-blocks 0, 1 | lo k-blocks 2, 3 of a 128-column chunk
template <int STAGES, bool TMAE = false>
constexpr size_t ffn_fused_smem(int FF) {
  return size_t(STAGES) * FFN_STAGE + FFN_HID_BYTES + (TMAE ? EPI_WARPS * (4096 + 2 * 8) : 0) + (2 * STAGES + 16) * 8 + 32 +
         (size_t(FF) + 3 * LN_BN) * 4 + 2 * BM * 2 * 8 + 1024;
}

template <int STAGES, bool TMAE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16, GemmEpilogue ep,
                 LnEpilogue ln, const float* __restrict__ b1, int M, int FF, int nk1, int nkw1, int split, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* hid = smem + STAGES * FFN_STAGE;
  uint8_t* stiles = hid + FFN_HID_BYTES;                               // (TMAE) [EPI_WARPS][4 KB]: third staging tile per warp
  uint64_t* tbars = reinterpret_cast<uint64_t*>(stiles + (TMAE ? EPI_WARPS * 4096 : 0));   // (TMAE) [EPI_WARPS][2]
  uint64_t* full_bar = tbars + (TMAE ? EPI_WARPS * 2 : 0);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc1_full = empty_bar + STAGES;     // [2]
  uint64_t* acc1_empty = acc1_full + 2;         // [2]
  uint64_t* hid_full = acc1_empty + 2;          // epilogue -> MMA: the hidden chunk is in shared memory
  uint64_t* hid_empty = hid_full + 1;           // MMA -> epilogue: GEMM 2 of the chunk has read it
  uint64_t* acc2_full = hid_empty + 1;
  uint64_t* acc2_empty = acc2_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc2_empty + 1);
  float* sb1 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 1) + 15) & ~uintptr_t(15));   // [FF]
  float* sbias = sb1 + FF;                      // [256] b2
  float* sgamma = sbias + LN_BN;
  float* sbeta = sgamma + LN_BN;
  float2* sstat = reinterpret_cast<float2*>(sbeta + LN_BN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NC = FF / 128;                      // hidden chunks per tile
  const int nk2 = split ? 4 : 2;                // k-blocks of GEMM 2 per chunk (hi | lo)

  if (warp == EPI_WARPS + 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], EPI_WARPS);
    }
    mbar_init(hid_full, EPI_WARPS);
    mbar_init(hid_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, EPI_WARPS);
    if (TMAE)
      for (int b = 0; b < EPI_WARPS * 2; ++b) mbar_init(&tbars[b], 1);
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmW1);
      tma_prefetch_desc(&tmW2);
      if (TMAE) {
        tma_prefetch_desc(&tmRes);
        tma_prefetch_desc(&tmO32);
        tma_prefetch_desc(&tmO16);
      }
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < FF; i += GEMM_THREADS) sb1[i] = __ldg(b1 + i);
  if (threadIdx.x < LN_BN) {
    sbias[threadIdx.x] = ep.bias ? __ldg(ep.bias + threadIdx.x) : 0.f;
    sgamma[threadIdx.x] = ln.gamma ? __ldg(ln.gamma + threadIdx.x) : 1.f;
    sbeta[threadIdx.x] = ln.beta ? __ldg(ln.beta + threadIdx.x) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tm_acc2 = tmem_base + 256;
  const int my_tiles = (n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int total_chunks = my_tiles * NC;       // flat chunk index g = (local tile) * NC + c

  if (warp == EPI_WARPS) {
    // ------------------------------------------------ TMA producer (same order as the MMA thread consumes)
    if (lane == 0) {
      uint32_t it = 0;
      auto g1 = [&](int g) {
        const int tile = blockIdx.x + (g / NC) * gridDim.x, c = g % NC;
        for (int kb = 0; kb < nk1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], 2 * A_STAGE_BYTES);
          tma_load_2d(ring + s * FFN_STAGE, &tmX, &full_bar[s], kb * BK, tile * BM);
          tma_load_2d(ring + s * FFN_STAGE + A_STAGE_BYTES, &tmW1, &full_bar[s], (kb < nkw1 ? kb : kb - nkw1) * BK, c * 128);
        }
      };
      auto g2 = [&](int g) {
        const int c = g % NC;
        for (int kb = 0; kb < nk2; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], FFN_STAGE);
          tma_load_2d(ring + s * FFN_STAGE, &tmW2, &full_bar[s], c * 128 + (kb & 1) * BK, 0);   // lo pass: same columns
        }
      };
      if (total_chunks > 0) g1(0);
      for (int g = 0; g < total_chunks; ++g) {
        if (g + 1 < total_chunks) g1(g + 1);
        g2(g);
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_f16(BM, 128, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(BM, 256, 0, 0);
      uint32_t it = 0;
      auto g1 = [&](int g) {
        const uint32_t b = g & 1;
        mbar_wait(&acc1_empty[b], ((g >> 1) & 1) ^ 1);          // the epilogue has drained this chunk accumulator
        tc_fence_after();
        for (int kb = 0; kb < nk1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(ring + s * FFN_STAGE), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(ring + s * FFN_STAGE + A_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss(tmem_base + b * 128, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc1, (kb | k) != 0);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc1_full[b]);
      };
      auto g2 = [&](int g) {
        const int c = g % NC, lt = g / NC;
        if (c == 0) {
          mbar_wait(acc2_empty, (lt & 1) ^ 1);                   // the previous tile's output epilogue is done
          tc_fence_after();
        }
        mbar_wait(hid_full, g & 1);                              // hidden chunk g converted and in shared memory
        tc_fence_after();
        for (int kb = 0; kb < nk2; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(hid + kb * A_STAGE_BYTES), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(ring + s * FFN_STAGE), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss(tm_acc2, a_desc + uint64_t(k * 2), b_desc + uint64_t(k * 2), idesc2, (c | kb | k) != 0);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(hid_empty);
        if (c == NC - 1) umma_commit(acc2_full);
      };
      if (total_chunks > 0) g1(0);
      for (int g = 0; g < total_chunks; ++g) {
        if (g + 1 < total_chunks) g1(g + 1);
        g2(g);
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps: warp w -> TMEM lanes 32 (w % 4).., column half w / 4
    const int quad = warp & 3, half = warp >> 2;
    const int rl = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    // (TMAE) staging tiles of the output epilogue: the two 4 KB pieces of the hidden-chunk buffer that only THIS warp
    // writes when it converts a chunk (k-blocks `half` and 2 + half, rows 32 quad .. + 31: free once the tile's last
    // GEMM 2 has completed) and one private tile
    uint32_t n_loads = 0;
    LnTma tm;
    tm.res = &tmRes; tm.out32 = &tmO32; tm.out16 = &tmO16;
    tm.tile[0] = hid + half * A_STAGE_BYTES + quad * 4096;
    tm.tile[1] = hid + (2 + half) * A_STAGE_BYTES + quad * 4096;
    tm.tile[2] = stiles + warp * 4096;
    tm.bars = tbars + warp * 2; tm.n_loads = nullptr;
    for (int g = 0; g < total_chunks; ++g) {
      const int lt = g / NC, c = g % NC;
      const uint32_t b = g & 1;
      mbar_wait(&acc1_full[b], (g >> 1) & 1);
      tc_fence_after();
      mbar_wait(hid_empty, (g & 1) ^ 1);                         // GEMM 2 of the previous chunk has read `hid`
      // 64 columns of the chunk per thread: bias + ReLU -> fp16 hi (k-block `half`) | lo (k-block 2 + half), one
      // 16-byte swizzled chunk per 8 columns (K-major SWIZZLE_128B A operand, row = this thread's tile row)
      uint8_t* hrow_hi = hid + half * A_STAGE_BYTES + rl * 128;
      uint8_t* hrow_lo = hid + (2 + half) * A_STAGE_BYTES + rl * 128;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + b * 128 + lane_addr + uint32_t(half * 64 + cc * 32), r);
        tmem_ld_wait();
        const float* bb = sb1 + c * 128 + half * 64 + cc * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(r[8 * j + i]) + bb[8 * j + i], 0.f);
          const uint4 hi = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]),
                                      pack_f16x2(v[6], v[7]));
          const int chunk = ((cc * 4 + j) ^ (rl & 7)) << 4;
          *reinterpret_cast<uint4*>(hrow_hi + chunk) = hi;
          if (split)
            *reinterpret_cast<uint4*>(hrow_lo + chunk) =
                make_uint4(f16x2_residual(v[0], v[1], hi.x), f16x2_residual(v[2], v[3], hi.y),
                           f16x2_residual(v[4], v[5], hi.z), f16x2_residual(v[6], v[7], hi.w));
        }
      }
      fence_proxy_async();        // hid: generic writes -> UMMA operand reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(hid_full);
        mbar_arrive(&acc1_empty[b]);
      }
      if (c == NC - 1) {          // the output row tile is complete: bias + residual -> h, LayerNorm -> next operand
        const int tile = blockIdx.x + lt * gridDim.x;
        const int row = tile * BM + rl;
        mbar_wait(acc2_full, lt & 1);
        tc_fence_after();
        if (TMAE) {
          ln_epilogue_tile_tma(ep, ln, tm_acc2 + lane_addr + uint32_t(half * 128), tile * BM + quad * 32, M, half, sbias,
                               sgamma, sbeta, sstat + (size_t(lt & 1) * BM + rl) * 2, tm, n_loads);
          if (lane == 0) tma_store_wait_read<0>();   // the stores have left tiles 0 / 1 before this warp converts into them
          __syncwarp();
        } else {
          ln_epilogue_tile(ep, ln, tm_acc2 + lane_addr + uint32_t(half * 128), row, row < M, half, sbias, sgamma, sbeta,
                           sstat + (size_t(lt & 1) * BM + rl) * 2);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty);
      }
    }
  }
  if (TMAE && warp < EPI_WARPS && lane == 0) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

__global__ void gemm_naive_kernel(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K,
                                  GemmEpilogue ep) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= N || row >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc = fmaf(__half2float(X[size_t(row) * ldx + k]), __half2float(W[size_t(col) * ldw + k]), acc);
  if (ep.bias) acc += ep.bias[col];
  if (ep.relu) acc = fmaxf(acc, 0.f);
  if (ep.rowvec) acc += ep.rowvec[size_t(row % ep.rowvec_period) * ep.ld_rowvec + col];
  if (ep.residual) acc += ep.residual[size_t(row) * ep.ld_res + col];
  if (ep.out_f32) ep.out_f32[size_t(row) * ep.ld_f32 + col] = acc;
  if (ep.out_f16) ep.out_f16[size_t(row) * ep.ld_f16 + col] = __float2half_rn(acc);
}

}  // namespace

int launch_gemm_tc(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep_in,
                   cudaStream_t s, int a_split) {
  if (M <= 0) return 0;
  const int KA = a_split ? 2 * K : K;   // a_split: X = [hi | lo] halves of K columns each, both multiplied by W
  if (K % BK != 0 || K <= 0) return set_error(-2, "gemm: K=%d must be a positive multiple of 64", K);
  const int n_pad = (N + 63) / 64 * 64;   // W must hold n_pad rows (zero rows past N)
  GemmEpilogue ep = ep_in;
  const int n_store = ep.n_store > 0 ? ep.n_store : N;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)KA, (uint64_t)M};
    uint64_t str[2] = {2, (uint64_t)ldx * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tmap_f16(&tmA, X, 2, dims, str, box, nullptr);
    if (rc) return rc;
  }
  // Tile choice: 128-wide tiles (2 x 128 TMEM columns), 3-stage TMA ring (98 KB; ASR_B200_GEMM_STAGES=6 selects the
  // 192 KB variant); 64-wide when N is not a multiple of 128.  ASR_B200_GEMM_TILE (128 / 64)
  // forces a width for experiments.
  // (256-wide tiles, ASR_B200_GEMM_TILE=256, measured identical at C2: these GEMMs are bound by their epilogue's global
  // stores, not by the L2 -> shared-memory operand stream.)
  int bn = (n_pad % 128 == 0) ? 128 : 64;
  if (const char* e = std::getenv("ASR_B200_GEMM_TILE")) {
    const int f = std::atoi(e);
    if ((f == 256 || f == 128 || f == 64) && n_pad % f == 0) bn = f;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)n_pad};
    uint64_t str[2] = {2, (uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    int rc = make_tmap_f16(&tmB, W, 2, dims, str, box, nullptr);
    if (rc) return rc;
  }
  static const int stages = [] {
    const char* e = std::getenv("ASR_B200_GEMM_STAGES");
    return e && e[0] ? std::atoi(e) : 3;
  }();
  switch (bn) {
    case 256: return launch_one<256, 4>(tmA, tmB, ep, M, n_store, n_pad, KA, K / BK, s);
    case 128:
      // 3-stage ring (98 KB) by default.  Alone the 6-stage (192 KB) variant is 4 % faster, but the encoder runs beside
      // the cluster decoder on the ~20 SMs it leaves free, and there the shallower ring measured better for both
      // (serving loop: 12.22 -> 12.06 ms per 128 utterances).  The grid follows the occupancy the runtime reports
      // (1 CTA per SM on the B200 driver used here; forcing 2 per SM changed nothing in the serving loop).
      if (stages == 6) return launch_one<128, 6>(tmA, tmB, ep, M, n_store, n_pad, KA, K / BK, s);
      return launch_one<128, 3>(tmA, tmB, ep, M, n_store, n_pad, KA, K / BK, s);
    default: return launch_one<64, 8>(tmA, tmB, ep, M, n_store, n_pad, KA, K / BK, s);
  }
}

int launch_gemm_ln(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                   const LnEpilogue& ln, cudaStream_t s, int a_split) {
  if (M <= 0) return 0;
  static const bool disabled = [] {
    const char* e = std::getenv("ASR_B200_FUSE_LN");
    return e && e[0] == '0';
  }();
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (disabled || N != LN_BN || K % BK != 0 || K <= 0 || !ln.gamma || !ln.beta || (!ln.out_f16 && !ln.out_f32) ||
      ep.out_f16 || (ep.n_store && ep.n_store != N) ||
      (ep.residual && (ep.ld_res % 4 != 0 || !al16(ep.residual))) || (ep.out_f32 && (ep.ld_f32 % 4 != 0 || !al16(ep.out_f32))) ||
      (ep.rowvec && (ep.ld_rowvec % 4 != 0 || !al16(ep.rowvec))) || !al16(ln.out_f16) || !al16(ln.out_f32))
    return 1;
  const int KA = a_split ? 2 * K : K;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)KA, (uint64_t)M};
    uint64_t str[2] = {2, (uint64_t)ldx * 2};
    uint32_t box[2] = {BK, BM};
    if (int rc = make_tmap_f16(&tmA, X, 2, dims, str, box, nullptr)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)LN_BN};
    uint64_t str[2] = {2, (uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)LN_BN};
    if (int rc = make_tmap_f16(&tmB, W, 2, dims, str, box, nullptr)) return rc;
  }
  int n_sm = 0;
  if (int rc = device_props(&n_sm, nullptr)) return rc;
  const int n_tiles = (M + BM - 1) / BM;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  // Row streams of the epilogue by TMA (ASR_B200_LN_TMA=0: the row-per-thread epilogue): needs dense-enough strides
  static const bool tma_off = [] {
    const char* e = std::getenv("ASR_B200_LN_TMA");
    return e && e[0] == '0';
  }();
  const int ldo = ln.split ? 2 * LN_BN : LN_BN;
  CUtensorMap tmRes = tmA, tmO32 = tmA, tmO16 = tmA;
  bool tmae = !tma_off;
  if (tmae) {
    const uint32_t box32[2] = {32, 32}, box16[2] = {64, 32};
    if (ep.residual) {
      uint64_t dims[2] = {(uint64_t)LN_BN, (uint64_t)M}, str[2] = {4, (uint64_t)ep.ld_res * 4};
      tmae = tmae && make_tmap_f32(&tmRes, ep.residual, 2, dims, str, box32, nullptr) == 0;
    }
    if (ep.out_f32) {
      uint64_t dims[2] = {(uint64_t)LN_BN, (uint64_t)M}, str[2] = {4, (uint64_t)ep.ld_f32 * 4};
      tmae = tmae && make_tmap_f32(&tmO32, ep.out_f32, 2, dims, str, box32, nullptr) == 0;
    }
    if (ln.out_f16) {
      uint64_t dims[2] = {(uint64_t)ldo, (uint64_t)M}, str[2] = {2, (uint64_t)ldo * 2};
      tmae = tmae && make_tmap_f16(&tmO16, ln.out_f16, 2, dims, str, box16, nullptr) == 0;
    }
  }
  if (tmae) {
    constexpr int STAGES = 2;   // 2 x 48 KB ring + 96 KB of epilogue staging tiles
    auto kern = gemm_ln_kernel<STAGES, true>;
    constexpr size_t smem = gemm_ln_smem<STAGES, true>();
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    kern<<<grid, GEMM_THREADS, smem, s>>>(tmA, tmB, tmRes, tmO32, tmO16, ep, ln, M, KA, K / BK, n_tiles);
  } else {
    constexpr int STAGES = 4;
    auto kern = gemm_ln_kernel<STAGES, false>;
    constexpr size_t smem = gemm_ln_smem<STAGES, false>();
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    kern<<<grid, GEMM_THREADS, smem, s>>>(tmA, tmB, tmRes, tmO32, tmO16, ep, ln, M, KA, K / BK, n_tiles);
  }
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_ffn_fused(const f16* X, int ldx, const f16* W1, const float* b1, const f16* W2, int M, int D, int FF,
                     const GemmEpilogue& ep, const LnEpilogue& ln, cudaStream_t s, int split) {
  if (M <= 0) return 0;
  static const bool disabled = [] {
    const char* e = std::getenv("ASR_B200_FUSE_FFN");
    return e && e[0] == '0';
  }();
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  constexpr int STAGES = 4;
  if (disabled || D != LN_BN || FF % 128 != 0 || FF <= 0 || ffn_fused_smem<STAGES>(FF) > 227 * 1024 || !b1 ||
      ep.out_f16 || ep.rowvec || (ep.residual && (ep.ld_res % 4 != 0 || !al16(ep.residual))) ||
      (ep.out_f32 && (ep.ld_f32 % 4 != 0 || !al16(ep.out_f32))) || !al16(ln.out_f16) || !al16(ln.out_f32))
    return 1;
  const int K1 = split ? 2 * D : D;
  CUtensorMap tmX, tmW1, tmW2;
  {
    uint64_t dims[2] = {(uint64_t)K1, (uint64_t)M};
    uint64_t str[2] = {2, (uint64_t)ldx * 2};
    uint32_t box[2] = {BK, BM};
    if (int rc = make_tmap_f16(&tmX, X, 2, dims, str, box, nullptr)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)FF};
    uint64_t str[2] = {2, (uint64_t)D * 2};
    uint32_t box[2] = {BK, 128};
    if (int rc = make_tmap_f16(&tmW1, W1, 2, dims, str, box, nullptr)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)FF, (uint64_t)D};
    uint64_t str[2] = {2, (uint64_t)FF * 2};
    uint32_t box[2] = {BK, 256};
    if (int rc = make_tmap_f16(&tmW2, W2, 2, dims, str, box, nullptr)) return rc;
  }
  int n_sm = 0;
  if (int rc = device_props(&n_sm, nullptr)) return rc;
  LnEpilogue lnn = ln;
  const int n_tiles = (M + BM - 1) / BM;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  static const bool tma_off = [] {
    const char* e = std::getenv("ASR_B200_LN_TMA");
    return e && e[0] == '0';
  }();
  const int ldo = ln.split ? 2 * LN_BN : LN_BN;
  CUtensorMap tmRes = tmX, tmO32 = tmX, tmO16 = tmX;
  bool tmae = !tma_off && ffn_fused_smem<3, true>(FF) <= 227 * 1024;
  if (tmae) {
    const uint32_t box32[2] = {32, 32}, box16[2] = {64, 32};
    if (ep.residual) {
      uint64_t dims[2] = {(uint64_t)LN_BN, (uint64_t)M}, str[2] = {4, (uint64_t)ep.ld_res * 4};
      tmae = tmae && make_tmap_f32(&tmRes, ep.residual, 2, dims, str, box32, nullptr) == 0;
    }
    if (ep.out_f32) {
      uint64_t dims[2] = {(uint64_t)LN_BN, (uint64_t)M}, str[2] = {4, (uint64_t)ep.ld_f32 * 4};
      tmae = tmae && make_tmap_f32(&tmO32, ep.out_f32, 2, dims, str, box32, nullptr) == 0;
    }
    if (ln.out_f16) {
      uint64_t dims[2] = {(uint64_t)ldo, (uint64_t)M}, str[2] = {2, (uint64_t)ldo * 2};
      tmae = tmae && make_tmap_f16(&tmO16, ln.out_f16, 2, dims, str, box16, nullptr) == 0;
    }
  }
  if (tmae) {   // 3-stage ring (measured: as fast as 4 here) + one private staging tile per epilogue warp
    auto kern = ffn_fused_kernel<3, true>;
    const size_t smem = ffn_fused_smem<3, true>(FF);
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    kern<<<grid, GEMM_THREADS, smem, s>>>(tmX, tmW1, tmW2, tmRes, tmO32, tmO16, ep, lnn, b1, M, FF, K1 / BK, D / BK, split,
                                          n_tiles);
  } else {
    auto kern = ffn_fused_kernel<STAGES, false>;
    const size_t smem = ffn_fused_smem<STAGES, false>(FF);
    if (int rc = ensure_dyn_smem((const void*)kern, smem)) return rc;
    kern<<<grid, GEMM_THREADS, smem, s>>>(tmX, tmW1, tmW2, tmRes, tmO32, tmO16, ep, lnn, b1, M, FF, K1 / BK, D / BK, split,
                                          n_tiles);
  }
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_gemm_naive(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                      cudaStream_t s) {
  if (M <= 0) return 0;
  const int n = ep.n_store > 0 ? ep.n_store : N;
  dim3 grid((n + 127) / 128, M);
  gemm_naive_kernel<<<grid, 128, 0, s>>>(X, ldx, W, ldw, M, n, K, ep);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// UMMA probe: D[128, N] = A[128, 64] * B, one CTA, one K block of 64.
//   b_mn_major == 0: B given as [N, 64] (K contiguous)   -> D = A * B^T
//   b_mn_major == 1: B given as [64 (k), N=64] (N contiguous, what a V tile looks like) -> D = A * B
// Used by tests/test_ops_gpu.py to pin the descriptor encodings the GEMM and attention kernels rely on.
namespace {
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int N,
                  int b_mn_major) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 16384;         // up to 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t b_bytes = b_mn_major ? 64 * 128 : N * 128;
    mbar_expect_tx(&bar[0], 16384 + b_bytes);
    tma_load_2d(sA, &tmA, &bar[0], 0, 0);
    tma_load_2d(sB, &tmB, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, N, 0, b_mn_major);
    const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(sB), b_mn_major ? 1024 : 16, 1024);
    for (int k = 0; k < 4; ++k) {
      // K-major: +32 B per K step; MN-major: 16 K rows of 128 B = +2048 B per K step
      const uint64_t b_adv = b_mn_major ? uint64_t(k * (2048 >> 4)) : uint64_t(k * 2);
      umma_f16_ss(tmem_base, a_desc + uint64_t(k * 2), b_desc + b_adv, idesc, k != 0);
    }
    umma_commit(&bar[1]);
  }
  __syncwarp();
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(c * 32), r);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
    for (int j = 0; j < 32; ++j) D[size_t(row) * N + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}
}  // namespace

int launch_umma_probe(const f16* A, const f16* Bm, float* D, int N, int b_mn_major, cudaStream_t s) {
  if (b_mn_major ? (N != 64) : (N % 32 != 0 || N < 32 || N > 128))
    return set_error(-2, "umma_probe: unsupported N=%d for b_mn_major=%d", N, b_mn_major);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {64, 128};
    uint64_t str[2] = {2, 128};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_f16(&tmA, A, 2, dims, str, box, nullptr);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)(b_mn_major ? 64 : N)};
    uint64_t str[2] = {2, 128};
    uint32_t box[2] = {64, (uint32_t)(b_mn_major ? 64 : N)};
    int rc = make_tmap_f16(&tmB, Bm, 2, dims, str, box, nullptr);
    if (rc) return rc;
  }
  const size_t smem = 32768 + 64 + 1024;
  if (int rc = ensure_dyn_smem((const void*)umma_probe_kernel, smem)) return rc;
  umma_probe_kernel<<<1, 128, smem, s>>>(tmA, tmB, D, N, b_mn_major);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
