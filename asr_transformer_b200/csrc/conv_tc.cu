// Conv front-end on the 5th-generation tensor cores: Transformer.input_layer (reference model.py:168-171),
// Conv2d(1, 64, 3, stride 2) + ReLU + Conv2d(64, 64, 3, stride 2) + ReLU, one kernel, conv2 as a tcgen05 implicit GEMM.
//
// A persistent CTA owns a tile of 6 output frames x all 19 output bins of one utterance (input_dim 80):
//   1. warps 0-15: spectrogram patch -> shared memory (cp.async), conv1 + ReLU in fp32 on the CUDA cores (C_in = 1: nine
//      FMAs per channel, packed FFMA2), result split into f16 hi | lo parts and written to shared memory as the A operand
//      of conv2.  The patch is stored as FOUR PARITY PLANES (time parity x frequency parity of the conv1 pixel), each a
//      no-swizzle K-major UMMA operand [8 channel chunks][plane rows][16 B]: conv2 has stride 2, so for tap (kh, kw) the
//      128 input pixels of the 128 output pixels of the tile are 128 CONSECUTIVE rows of plane (kw & 1, kh & 1), starting
//      (kw >> 1) * 20 + (kh >> 1) rows in - an im2col tile is just a descriptor start address.
//      GEMM row m = 20 * frame + bin (bin 19 is a dummy pixel: rows of a plane are 20 apart per frame), M = 128 >= 120.
//   2. warp 16, one thread: 9 taps x 4 k-steps x (hi, lo) = 72 tcgen05.mma (M128 N64 K16, fp16 in, fp32 accumulate in
//      TMEM) against the conv2 weights, resident in shared memory as [tap][channel chunk][64 out][8 in] (unpacked once
//      per CTA from the fragment-major packing the legacy kernels use - no ABI change);
//   3. warps 0-15: accumulator from TMEM, + bias, ReLU, f16 hi | lo, stores into z (B*T', [hi (F'*64) | lo (F'*64)]) -
//      deferred by one tile: the epilogue of tile i runs under the MMAs of tile i + 1 (two accumulators in TMEM).
// The spectrogram patch of the next tile is requested under the MMAs too.  The mma.sync edition (simple_ops.cu,
// conv_fused_kernel) stays as the fallback for other input dimensions and as the cross-check
// (tests/test_ops_gpu.py::test_conv_tc_matches_fused).
#include "kernels.h"
#include "ptx.cuh"
#ifdef ASR_CONV_DBG
#include <cstdio>
#endif

namespace asr {
namespace {

constexpr int CT_WARPS = 16;               // conv1 / epilogue warps
constexpr int CT_PROD = CT_WARPS * 32;
constexpr int CT_THREADS = CT_PROD + 32;   // + the MMA warp
constexpr int CT_TT2 = 6;                  // output frames per tile
constexpr int CT_F = 80, CT_F1 = 39, CT_F2 = 19, CT_FH = 20;
constexpr int CT_W = 4 * CT_TT2 + 3;       // input frames per tile
constexpr int ROWS_E = (CT_TT2 + 1) * CT_FH, ROWS_O = CT_TT2 * CT_FH;   // plane rows: even / odd conv1 frame index
// bytes between the 16-byte channel chunks of a plane: the rows, padded so that LBO = 16 (mod 128): the eight chunks of one
// pixel then fall into eight different 16-byte bank groups (without the padding the conv1 stores of a pixel hit one or
// two groups: 4- to 8-way bank conflicts, which made conv1 the bottleneck of the kernel)
constexpr uint32_t lbo_pad(uint32_t b) { return b + ((16 + 128 - (b % 128)) % 128); }
constexpr uint32_t LBO_E = lbo_pad(ROWS_E * 16), LBO_O = lbo_pad(ROWS_O * 16);
static_assert(LBO_E % 128 == 16 && LBO_O % 128 == 16, "chunk stride");
constexpr uint32_t PLANES_BYTES = 2 * 8 * (LBO_E + LBO_O);              // 4 planes of one precision (66,560)
constexpr uint32_t W_BYTES = 9 * 8 * 64 * 16;                           // 73,728
constexpr uint32_t OFF_W = 0;
constexpr uint32_t OFF_Y = OFF_W + W_BYTES;                             // hi planes, then lo planes
constexpr uint32_t OFF_SLACK = OFF_Y + 2 * PLANES_BYTES;                // dummy rows of the last plane read past it
constexpr uint32_t OFF_X = OFF_SLACK + 512;
constexpr uint32_t OFF_W1 = OFF_X + ((CT_F * CT_W * 4 + 15) & ~15);
constexpr uint32_t OFF_BAR = OFF_W1 + (9 * 64 + 64 + 64) * 4;
constexpr uint32_t CT_SMEM = OFF_BAR + 64;

// K-major operand without swizzle: 8-row x 16-byte core matrices, rows 16 B apart; sbo = bytes between 8-row groups,
// lbo = bytes between the two 16-byte K chunks of one MMA (K = 16 f16)
__device__ __forceinline__ uint64_t umma_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell); layout type (bits 61-63) 0 = no swizzle
  return d;
}
// byte offset of plane (time parity pt, frequency parity pf) inside one precision's block
__device__ __forceinline__ uint32_t plane_off(int pt, int pf) {
  return pt ? 2 * 8 * LBO_E + uint32_t(pf) * 8 * LBO_O : uint32_t(pf) * 8 * LBO_E;
}

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const float* __restrict__ spec, const float* __restrict__ w1, const float* __restrict__ b1,
               const uint2* __restrict__ wfrag, const float* __restrict__ b2, int B, int T, int T2, int n_tt,
               f16* __restrict__ z, int split) {
  extern __shared__ __align__(128) uint8_t sm[];
  float* sx = reinterpret_cast<float*>(sm + OFF_X);
  float* sw1 = reinterpret_cast<float*>(sm + OFF_W1);
  float* sb1 = sw1 + 9 * 64;
  float* sb2 = sb1 + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t* y_full = bars;        // conv1 patch written (CT_PROD arrivals)
  uint64_t* d_full = bars + 1;    // MMAs of the tile complete (accumulator ready, patch free)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int F = CT_F, W = CT_W, F1 = CT_F1, F2 = CT_F2;

  // conv2 weights: fragment-major [36 k-steps][8 n-tiles][32 lanes][4 f16] (engine.pack_conv2_fragments: k-step = (tap *
  // 2 + half) * 2 + sub, lane = g * 4 + c, out channel nt * 8 + g, in channels half * 32 + c * 8 + sub * 4 + 0..3) ->
  // [tap][chunk = half * 4 + c][out channel][8 in channels]
  for (int i = tid; i < 36 * 8 * 32; i += CT_THREADS) {
    const int ks = i >> 8, nt = (i >> 5) & 7, ln = i & 31, g = ln >> 2, c = ln & 3;
    const int tap = ks >> 2, half = (ks >> 1) & 1, sub = ks & 1;
    *reinterpret_cast<uint2*>(sm + OFF_W + ((tap * 8 + half * 4 + c) * 64 + nt * 8 + g) * 16 + sub * 8) = wfrag[i];
  }
  for (int i = tid; i < 9 * 64; i += CT_THREADS) sw1[i] = w1[i];
  if (tid < 64) {
    sb1[tid] = b1[tid];
    sb2[tid] = b2[tid];
  }
  if (tid == 0) {
    mbar_init(y_full, CT_PROD);
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == CT_WARPS) {
    tmem_alloc(tmem_ptr, 128);    // two accumulators of 64 columns
    tmem_relinquish();
  }
  fence_proxy_async();            // weights: generic writes -> UMMA operand reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr;
  const int n_tiles = B * n_tt;

  if (warp == CT_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0, 0);
      const uint32_t wb = smem_u32(sm + OFF_W), yb = smem_u32(sm + OFF_Y);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(y_full, it & 1);
        tc_fence_after();
        const uint32_t d = tmem_d + (it & 1) * 64;
        uint32_t acc = 0;
        for (int prec = 0; prec < (split ? 2 : 1); ++prec)
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int kh = tap / 3, kw = tap - kh * 3;             // kh: frequency offset, kw: time offset
            const uint32_t lbo = (kw & 1) ? LBO_O : LBO_E;
            const uint32_t a0 = yb + prec * PLANES_BYTES + plane_off(kw & 1, kh & 1) + ((kw >> 1) * CT_FH + (kh >> 1)) * 16;
            const uint32_t b0 = wb + tap * 8192;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_f16_ss(d, umma_smem_desc_nosw(a0 + 2 * ks * lbo, lbo, 128),
                          umma_smem_desc_nosw(b0 + 2 * ks * 1024, 1024, 128), idesc, acc);
              acc = 1;
            }
          }
        umma_commit(d_full);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ conv1 producers / epilogue (CT_PROD threads)
    // Order per tile i: [MMAs of tile i-1 done: the patch is free] conv1(i) -> y_full -> epilogue(i-1), which therefore
    // runs UNDER the MMAs of tile i (the accumulator is double-buffered in TMEM).
    auto load_patch = [&](int tl) {
      const int b = tl / n_tt, t20 = (tl - b * n_tt) * CT_TT2, ntt = min(CT_TT2, T2 - t20);
      const int c0 = 4 * t20, wcols = 4 * ntt + 3;
      const float* xin = spec + size_t(b) * F * T + c0;
      for (int i = tid; i < F * W; i += CT_PROD) {
        const int f = i / W, cc = i - f * W;
        if (cc < wcols && c0 + cc < T) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sx + i)), "l"(xin + size_t(f) * T + cc)
                       : "memory");
        } else {
          sx[i] = 0.f;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // accumulator of tile (bb, tt0, nt) (buffer buf) -> + bias, ReLU, f16 hi | lo -> z: rows 32 (warp % 4) .. + 31,
    // columns 16 (warp / 4) .. + 15 per warp
    auto epilogue = [&](int buf, int bb, int tt0, int nt) {
      const int m = (warp & 3) * 32 + lane, tt = m / CT_FH, f2 = m - tt * CT_FH, col0 = (warp >> 2) * 16;
      uint32_t rr[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
            "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
          : "r"(tmem_d + uint32_t(buf * 64) + (uint32_t((warp & 3) * 32) << 16) + uint32_t(col0))
          : "memory");
      tmem_ld_wait();
      if (tt < nt && f2 < F2) {
        const size_t zrow = size_t(split ? 2 : 1) * F2 * 64;
        f16* dst = z + (size_t(bb) * T2 + tt0 + tt) * zrow + size_t(f2) * 64 + col0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(rr[8 * q + i]) + sb2[col0 + 8 * q + i], 0.f);
          uint4 o;
          o.x = pack_f16x2(v[0], v[1]);
          o.y = pack_f16x2(v[2], v[3]);
          o.z = pack_f16x2(v[4], v[5]);
          o.w = pack_f16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(dst + 8 * q) = o;
          if (split)
            *reinterpret_cast<uint4*>(dst + size_t(F2) * 64 + 8 * q) =
                make_uint4(f16x2_residual(v[0], v[1], o.x), f16x2_residual(v[2], v[3], o.y),
                           f16x2_residual(v[4], v[5], o.z), f16x2_residual(v[6], v[7], o.w));
        }
      }
      tc_fence_before();                                     // these TMEM reads before the MMAs that reuse the buffer
    };
    if (blockIdx.x < n_tiles) load_patch(blockIdx.x);
    // conv1: 16 threads per pixel, 4 channels each (two packed FMA chains per pixel, two pixels per pass).  17 warps cap
    // the kernel at 96 registers per thread: 8 channels per thread (72 weight registers) spill and run slower (measured).
    const int cg = tid & 15;
    float2 wr[9][2], bias[2];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int k = 0; k < 2; ++k) wr[tap][k] = *reinterpret_cast<const float2*>(sw1 + tap * 64 + cg * 4 + 2 * k);
#pragma unroll
    for (int k = 0; k < 2; ++k) bias[k] = *reinterpret_cast<const float2*>(sb1 + cg * 4 + 2 * k);
    uint32_t it = 0;
    int pb = 0, pt0 = 0, pnt = 0;                            // the previous tile (its epilogue is still due)
#ifdef ASR_CONV_DBG
    long long tph[6] = {0, 0, 0, 0, 0, 0}, tl = clock64();
#define CONV_MARK(i) { const long long n_ = clock64(); tph[i] += n_ - tl; tl = n_; }
#else
#define CONV_MARK(i)
#endif
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int b = tile / n_tt, t20 = (tile - b * n_tt) * CT_TT2, ntt = min(CT_TT2, T2 - t20);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      CONV_MARK(0)
      if (it > 0) {                                          // MMAs of the previous tile done: the conv1 patch is free
        mbar_wait(d_full, (it - 1) & 1);
        tc_fence_after();
      }
      CONV_MARK(1)
      asm volatile("bar.sync 1, %0;" ::"n"(CT_PROD) : "memory");   // spectrogram patch complete in shared memory
      CONV_MARK(2)
      // ---- conv1 + ReLU -> f16 hi | lo parity planes (one FMA per tap and channel, taps in the reference order, as in
      // conv_fused_kernel)
      const int npix = (2 * ntt + 1) * F1;
      for (int pix0 = tid >> 4; pix0 < npix; pix0 += 2 * (CT_PROD / 16)) {
        const int pix1 = pix0 + CT_PROD / 16;
        const bool two = pix1 < npix;
        const int ra = pix0 / F1, fa = pix0 - ra * F1;
        const int pxb = two ? pix1 : pix0, rb = pxb / F1, fb = pxb - rb * F1;
        const float* xa = sx + (2 * fa) * W + 2 * ra;
        const float* xb = sx + (2 * fb) * W + 2 * rb;
        float2 acc[2][2];
#pragma unroll
        for (int k = 0; k < 2; ++k) acc[0][k] = acc[1][k] = bias[k];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float x0 = xa[kh * W + kw], x1 = xb[kh * W + kw];
            const float2 xx0 = make_float2(x0, x0), xx1 = make_float2(x1, x1);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              acc[0][k] = __ffma2_rn(xx0, wr[kh * 3 + kw][k], acc[0][k]);
              acc[1][k] = __ffma2_rn(xx1, wr[kh * 3 + kw][k], acc[1][k]);
            }
          }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !two) break;
          const int r = q ? rb : ra, f1 = q ? fb : fa;
          const float v0 = fmaxf(acc[q][0].x, 0.f), v1 = fmaxf(acc[q][0].y, 0.f), v2 = fmaxf(acc[q][1].x, 0.f),
                      v3 = fmaxf(acc[q][1].y, 0.f);
          uint2 o;
          o.x = pack_f16x2(v0, v1);
          o.y = pack_f16x2(v2, v3);
          const uint32_t off = plane_off(r & 1, f1 & 1) + uint32_t(cg >> 1) * ((r & 1) ? LBO_O : LBO_E) +
                               uint32_t((r >> 1) * CT_FH + (f1 >> 1)) * 16 + uint32_t(cg & 1) * 8;
          *reinterpret_cast<uint2*>(sm + OFF_Y + off) = o;
          if (split)
            *reinterpret_cast<uint2*>(sm + OFF_Y + PLANES_BYTES + off) =
                make_uint2(f16x2_residual(v0, v1, o.x), f16x2_residual(v2, v3, o.y));
        }
      }
      CONV_MARK(3)
      fence_proxy_async();                                   // patch: generic writes -> UMMA operand reads
      mbar_arrive(y_full);
      asm volatile("bar.sync 1, %0;" ::"n"(CT_PROD) : "memory");   // every thread is done with sx
      if (tile + int(gridDim.x) < n_tiles) load_patch(tile + gridDim.x);   // next patch: lands under the MMAs
      CONV_MARK(4)
      if (it > 0) epilogue((it - 1) & 1, pb, pt0, pnt);      // previous tile's accumulator, under this tile's MMAs
      CONV_MARK(5)
      pb = b; pt0 = t20; pnt = ntt;
    }
    if (it > 0) {
      mbar_wait(d_full, (it - 1) & 1);
      tc_fence_after();
      epilogue((it - 1) & 1, pb, pt0, pnt);
    }
#ifdef ASR_CONV_DBG
    if (blockIdx.x == 0 && tid == 0)
      printf("conv_tc tiles %u: cp.async wait %lld, mma wait %lld, bar %lld, conv1 %lld, arrive+bar+prefetch %lld, epilogue %lld (cycles per tile)\n",
             it, tph[0] / it, tph[1] / it, tph[2] / it, tph[3] / it, tph[4] / it, tph[5] / it);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CT_WARPS) tmem_dealloc(tmem_d, 128);
}

}  // namespace

// returns 1 (nothing launched) when the shape is not the one this kernel is built for (input_dim 80): the caller falls
// back to conv_fused_kernel
int launch_conv_tc(const float* spec, const float* w1, const float* b1, const f16* w2frag, const float* b2, int B, int F,
                   int T, f16* z, cudaStream_t s, int split) {
  if (F != CT_F) return 1;
  const int T1 = (T - 3) / 2 + 1, T2 = (T1 - 3) / 2 + 1;
  if (T2 <= 0) return set_error(-2, "conv: input %dx%d too small", F, T);
  if (B <= 0) return 0;
  int n_sm = 0, max_smem = 0;
  if (int rc = device_props(&n_sm, &max_smem)) return rc;
  if (size_t(max_smem) < CT_SMEM) return 1;
  if (int rc = ensure_dyn_smem((const void*)conv_tc_kernel, CT_SMEM)) return rc;
  const int n_tt = (T2 + CT_TT2 - 1) / CT_TT2;
  const long long n_tiles = (long long)B * n_tt;
  const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);
  conv_tc_kernel<<<grid, CT_THREADS, CT_SMEM, s>>>(spec, w1, b1, reinterpret_cast<const uint2*>(w2frag), b2, B, T, T2, n_tt,
                                                   z, split);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
