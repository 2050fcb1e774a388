// LayerNorm, embedding + positional encoding, dtype casts and the conv subsampling front-end.
#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

// ---------------------------------------------------------------- LayerNorm (nn.LayerNorm, eps 1e-5, affine)
// One warp per row, 16-byte vectorised loads, fp32 statistics via warp shuffles; optional fp32 and f16 outputs.
constexpr int LN_MAX_CHUNKS = 8;   // D <= 1024

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int rows, int D, float eps, float* __restrict__ y_f32, f16* __restrict__ y_f16, int split) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + size_t(row) * D;
  float4 v[LN_MAX_CHUNKS];
  float sum = 0.f;
  const int nch = D / 128;   // full 128-float chunks; D % 128 == 0 enforced by the launcher
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    if (i < nch) {
      v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(sum) / float(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    if (i < nch) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / float(D) + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    if (i < nch) {
      const int col = i * 128 + lane * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + col);
      const float4 bt = *reinterpret_cast<const float4*>(beta + col);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bt.x;
      o.y = (v[i].y - mean) * rstd * g.y + bt.y;
      o.z = (v[i].z - mean) * rstd * g.z + bt.z;
      o.w = (v[i].w - mean) * rstd * g.w + bt.w;
      if (y_f32) *reinterpret_cast<float4*>(y_f32 + size_t(row) * D + col) = o;
      if (y_f16) {   // split: row = [hi (D) | lo (D)], the A operand of an a_split GEMM
        uint2 t;
        t.x = pack_f16x2(o.x, o.y);
        t.y = pack_f16x2(o.z, o.w);
        f16* yr = y_f16 + size_t(row) * (split ? 2 * D : D) + col;
        *reinterpret_cast<uint2*>(yr) = t;
        if (split) *reinterpret_cast<uint2*>(yr + D) = make_uint2(f16x2_residual(o.x, o.y, t.x), f16x2_residual(o.z, o.w, t.y));
      }
    }
  }
}

// ---------------------------------------------------------------- embedding + positional encoding
// out[b, t, :] = E[tokens[b, t]] + pe[t]   (reference model.py:117)
__global__ void embed_pe_kernel(const int32_t* tokens, int ld_tok, const float* emb, const float* pe, int L, int D,
                                int vocab, float* out) {
  const int t = blockIdx.x, b = blockIdx.y;
  int tok = tokens[size_t(b) * ld_tok + t];
  tok = min(max(tok, 0), vocab - 1);
  const float* e = emb + size_t(tok) * D;
  const float* pr = pe + size_t(t) * D;
  float* o = out + (size_t(b) * L + t) * D;
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(e + d);
    const float4 c = *reinterpret_cast<const float4*>(pr + d);
    *reinterpret_cast<float4*>(o + d) = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ x, f16* __restrict__ y, size_t n4) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 t;
    t.x = pack_f16x2(v.x, v.y);
    t.y = pack_f16x2(v.z, v.w);
    reinterpret_cast<uint2*>(y)[i] = t;
  }
}
// x fp32 [rows, D] -> y f16 [rows, 2D] = [hi | lo]
__global__ void f32_to_f16_split_kernel(const float* __restrict__ x, f16* __restrict__ y, size_t rows, int D) {
  const int d4 = D / 4;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x, n4 = rows * d4;
  for (; i < n4; i += stride) {
    const size_t r = i / d4;
    const int c = int(i - r * d4) * 4;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 t;
    t.x = pack_f16x2(v.x, v.y);
    t.y = pack_f16x2(v.z, v.w);
    f16* yr = y + r * 2 * D + c;
    *reinterpret_cast<uint2*>(yr) = t;
    *reinterpret_cast<uint2*>(yr + D) = make_uint2(f16x2_residual(v.x, v.y, t.x), f16x2_residual(v.z, v.w, t.y));
  }
}
__global__ void f32_to_f16_tail_kernel(const float* x, f16* y, size_t start, size_t n) {
  const size_t i = start + threadIdx.x;
  if (i < n) y[i] = __float2half_rn(x[i]);
}

// ---------------------------------------------------------------- conv1: Conv2d(1,64,3,stride 2) + ReLU
// spectrum fp32 (B,1,F,T) -> y1 f16 channels-last (B, T1, F1, 64).  C_in = 1, so this is CUDA-core work bound by
// the 128 B / pixel output stream.  One CTA per (utterance, CONV1_TT consecutive output frames): the F x (2 TT + 1)
// input patch is staged in shared memory with loads coalesced along T (the input's contiguous axis), then 8 threads
// per output pixel produce 8 channels each and one 16-byte store (pixels of a frame are contiguous: coalesced).
constexpr int CONV1_TT = 32;
__global__ void __launch_bounds__(256)
conv1_kernel(const float* __restrict__ spec, const float* __restrict__ w1, const float* __restrict__ b1, int B, int F,
             int T, int F1, int T1, f16* __restrict__ y1, size_t lo_plane) {
  extern __shared__ float conv1_smem[];
  float* sw = conv1_smem;            // [9][64]
  float* sb = sw + 9 * 64;           // [64]
  float* sx = sb + 64;               // [F][2 TT + 1]
  constexpr int W = 2 * CONV1_TT + 1;
  const int b = blockIdx.y, t10 = blockIdx.x * CONV1_TT;
  const int nt1 = min(CONV1_TT, T1 - t10), wcols = 2 * nt1 + 1;
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) sw[i] = w1[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = b1[threadIdx.x];
  const float* xin = spec + size_t(b) * F * T + 2 * t10;
  for (int i = threadIdx.x; i < F * W; i += blockDim.x) {
    const int f = i / W, c = i - f * W;
    sx[i] = (c < wcols) ? __ldg(xin + size_t(f) * T + c) : 0.f;
  }
  __syncthreads();
  const int cg = threadIdx.x & 7;
  float wr[9][8], bias[8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int c = 0; c < 8; ++c) wr[tap][c] = sw[tap * 64 + cg * 8 + c];
#pragma unroll
  for (int c = 0; c < 8; ++c) bias[c] = sb[cg * 8 + c];
  f16* yout = y1 + (size_t(b) * T1 + t10) * F1 * 64;
  for (int pix = threadIdx.x >> 3; pix < nt1 * F1; pix += 32) {
    const int tl = pix / F1, f1 = pix - tl * F1;
    const float* xp = sx + (2 * f1) * W + 2 * tl;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = bias[c];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float x = xp[kh * W + kw];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(x, wr[kh * 3 + kw][c], acc[c]);
      }
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fmaxf(acc[c], 0.f);
    uint4 o;
    o.x = pack_f16x2(acc[0], acc[1]);
    o.y = pack_f16x2(acc[2], acc[3]);
    o.z = pack_f16x2(acc[4], acc[5]);
    o.w = pack_f16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(yout + size_t(pix) * 64 + cg * 8) = o;
    if (lo_plane)   // hi | lo split of the conv2 operand (second plane)
      *reinterpret_cast<uint4*>(yout + lo_plane + size_t(pix) * 64 + cg * 8) =
          make_uint4(f16x2_residual(acc[0], acc[1], o.x), f16x2_residual(acc[2], acc[3], o.y),
                     f16x2_residual(acc[4], acc[5], o.z), f16x2_residual(acc[6], acc[7], o.w));
  }
}

// ---------------------------------------------------------------- conv2: Conv2d(64,64,3,stride 2) + ReLU
// Implicit GEMM on tensor cores: M = B*T2*F2 output pixels, N = 64, K = 9 taps * 64 channels.
// Weights live in shared memory pre-arranged as mma fragments (host packs them, see packing.py), so each
// B fragment is one conflict-free 8-byte LDS.  A fragments are 16-byte channel vectors read straight from
// the channels-last y1.  Output z[pixel, co] with pixel = (b, t2, f2): i.e. (B, T2, F2*64), column order (f, c);
// the matching column permutation of _lin_in.weight is done once at weight-pack time.
constexpr int CONV2_KSTEPS = 36;   // 9 taps * 2 halves of 32 channels * 2 sub-steps of 16
constexpr int CONV2_W_BYTES = CONV2_KSTEPS * 8 * 32 * 8;   // 73,728

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// two adjacent channels of one output pixel -> z (hi, and lo = f16(v - hi) lo_off elements further when lo_off != 0)
__device__ __forceinline__ void store_pair_split(f16* dst, bool ok, float a, float b, size_t lo_off) {
  if (!ok) return;
  const uint32_t hi = pack_f16x2(a, b);
  *reinterpret_cast<uint32_t*>(dst) = hi;
  if (lo_off) *reinterpret_cast<uint32_t*>(dst + lo_off) = f16x2_residual(a, b, hi);
}

__global__ void __launch_bounds__(128)
conv2_kernel(const f16* __restrict__ y1, const uint2* __restrict__ wfrag, const float* __restrict__ b2, int B,
             int F1, int T1, int F2, int T2, f16* __restrict__ z, size_t lo_plane) {
  extern __shared__ uint2 sw[];   // [36][8][32] fragments
  for (int i = threadIdx.x; i < CONV2_KSTEPS * 8 * 32; i += blockDim.x) sw[i] = wfrag[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const long long M = (long long)B * T2 * F2;
  const long long ntiles = (M + 63) / 64;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m_lo = tile * 64 + warp * 16 + g;
    const long long m_hi = m_lo + 8;
    const f16* pa[2];
    size_t zo[2];      // output offsets: row (b, t2) of [hi (F2*64) | lo (F2*64)] when split, column f2*64
    const size_t zrow = size_t(lo_plane ? 2 : 1) * F2 * 64;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      long long m = i ? m_hi : m_lo;
      if (m >= M) m = M - 1;   // clamp (stores are predicated)
      const int f2 = int(m % F2);
      const int t2 = int((m / F2) % T2);
      const int b = int(m / ((long long)F2 * T2));
      pa[i] = y1 + ((size_t(b) * T1 + 2 * t2) * F1 + 2 * f2) * 64 + c * 8;
      zo[i] = (size_t(b) * T2 + t2) * zrow + size_t(f2) * 64;
    }
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int kh = tap / 3, kw = tap % 3;
      const size_t toff = (size_t(kw) * F1 + kh) * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint4 vlo = __ldg(reinterpret_cast<const uint4*>(pa[0] + toff + half * 32));
        const uint4 vhi = __ldg(reinterpret_cast<const uint4*>(pa[1] + toff + half * 32));
        uint4 ulo = make_uint4(0, 0, 0, 0), uhi = make_uint4(0, 0, 0, 0);
        if (lo_plane) {
          ulo = __ldg(reinterpret_cast<const uint4*>(pa[0] + lo_plane + toff + half * 32));
          uhi = __ldg(reinterpret_cast<const uint4*>(pa[1] + lo_plane + toff + half * 32));
        }
        const uint2* wk = sw + size_t((tap * 2 + half) * 2) * 8 * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint2 w0 = wk[nt * 32];
          const uint2 w1 = wk[8 * 32 + nt * 32];
          mma_f16_16816(acc[nt], vlo.x, vhi.x, vlo.y, vhi.y, w0.x, w0.y);
          mma_f16_16816(acc[nt], vlo.z, vhi.z, vlo.w, vhi.w, w1.x, w1.y);
          if (lo_plane) {
            mma_f16_16816(acc[nt], ulo.x, uhi.x, ulo.y, uhi.y, w0.x, w0.y);
            mma_f16_16816(acc[nt], ulo.z, uhi.z, ulo.w, uhi.w, w1.x, w1.y);
          }
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int co = nt * 8 + 2 * c;
      const float bb0 = __ldg(b2 + co), bb1 = __ldg(b2 + co + 1);
      store_pair_split(z + zo[0] + co, m_lo < M, fmaxf(acc[nt][0] + bb0, 0.f), fmaxf(acc[nt][1] + bb1, 0.f),
                       lo_plane ? size_t(F2) * 64 : 0);
      store_pair_split(z + zo[1] + co, m_hi < M, fmaxf(acc[nt][2] + bb0, 0.f), fmaxf(acc[nt][3] + bb1, 0.f),
                       lo_plane ? size_t(F2) * 64 : 0);
    }
  }
}


// ---------------------------------------------------------------- fused front-end: conv1 + ReLU + conv2 + ReLU
// One kernel for Transformer.input_layer (model.py:168-171).  The (B, T1, F1, 64) f16 intermediate (159 MB at C2) never
// leaves the SM: a persistent CTA owns a tile of TT2 output frames x all F2 output bins of one utterance, computes the
// conv1 patch it needs ((2 TT2 + 1) x F1 pixels x 64 channels, fp32 FMAs from a shared-memory patch of the spectrogram)
// into shared memory as f16 and runs conv2 on it as an implicit GEMM with mma.sync (same fragment-packed weights and
// the same arithmetic as conv1_kernel + conv2_kernel, so the two paths give identical bits).  The patch stores the
// 16-byte channel chunks of a pixel XOR-swizzled by bit 1 of its frequency index, which makes the A-fragment loads of
// neighbouring output pixels (input pixels two apart) land in different bank halves.
constexpr int CONVF_THREADS = 256;
__global__ void __launch_bounds__(CONVF_THREADS, 1)
conv_fused_kernel(const float* __restrict__ spec, const float* __restrict__ w1, const float* __restrict__ b1,
                  const uint2* __restrict__ wfrag, const float* __restrict__ b2, int B, int F, int T, int F1, int F2,
                  int T2, int TT2, int n_tt, f16* __restrict__ z, int split) {
  extern __shared__ __align__(16) uint8_t convf_smem[];
  uint2* sw = reinterpret_cast<uint2*>(convf_smem);                        // [36][8][32] conv2 fragments
  float* sw1 = reinterpret_cast<float*>(convf_smem + CONV2_W_BYTES);       // [9][64] conv1 taps
  float* sb1 = sw1 + 9 * 64;                                               // [64]
  float* sx = sb1 + 64;                                                    // [F][W] spectrogram patch
  const int W = 4 * TT2 + 3;
  f16* sy = reinterpret_cast<f16*>(sx + ((size_t(F) * W + 3) & ~size_t(3)));   // [(2 TT2 + 1)][F1][64] conv1 patch
  const size_t sy_lo = size_t(2 * TT2 + 1) * F1 * 64;                          // + the lo parts of the patch when split
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  for (int i = tid; i < CONV2_KSTEPS * 8 * 32; i += CONVF_THREADS) sw[i] = wfrag[i];
  for (int i = tid; i < 9 * 64; i += CONVF_THREADS) sw1[i] = w1[i];
  if (tid < 64) sb1[tid] = b1[tid];
  const int n_tiles = B * n_tt;
  // spectrogram patch of a tile: all F bins x its 4 ntt + 3 input frames, fetched with 4-byte cp.async so that every
  // load of the patch is in flight at once; the patch of tile i + 1 is requested as soon as conv1 of tile i has consumed
  // sx and lands under the MMAs of tile i
  auto load_patch = [&](int tl) {
    const int b = tl / n_tt, t20 = (tl - b * n_tt) * TT2, ntt = min(TT2, T2 - t20);
    const int c0 = 4 * t20, wcols = 4 * ntt + 3;
    const float* xin = spec + size_t(b) * F * T + c0;
    for (int i = tid; i < F * W; i += CONVF_THREADS) {
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
  if (blockIdx.x < n_tiles) load_patch(blockIdx.x);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / n_tt, t20 = (tile - b * n_tt) * TT2, ntt = min(TT2, T2 - t20);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                     // patch complete; the previous tile's MMAs are done with sy
    // ---- conv1 + ReLU -> f16 patch; 8 threads per pixel, 8 channels each.  The arithmetic of conv1_kernel (one
    // fused multiply-add per tap and channel, taps in the same order), issued as packed fp32x2 FMAs (FFMA2, sm_100):
    // same bits, half the issue slots.  (The kernel is instruction-bound overall: 34 k warp instructions per tile at
    // 1.7 IPC, a third of them the m16n8k16 MMAs and their per-MMA weight-fragment loads.)
    {
      const int cg = tid & 7;
      float2 wr[9][4], bias[4];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int k = 0; k < 4; ++k) wr[tap][k] = *reinterpret_cast<const float2*>(sw1 + tap * 64 + cg * 8 + 2 * k);
#pragma unroll
      for (int k = 0; k < 4; ++k) bias[k] = *reinterpret_cast<const float2*>(sb1 + cg * 8 + 2 * k);
      const int npix = (2 * ntt + 1) * F1;
      for (int pix = tid >> 3; pix < npix; pix += CONVF_THREADS / 8) {
        const int r = pix / F1, f1 = pix - r * F1;
        const float* xp = sx + (2 * f1) * W + 2 * r;
        float2 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = bias[k];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float x = xp[kh * W + kw];
            const float2 xx = make_float2(x, x);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = __ffma2_rn(xx, wr[kh * 3 + kw][k], acc[k]);
          }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float2(fmaxf(acc[k].x, 0.f), fmaxf(acc[k].y, 0.f));
        uint4 o;
        o.x = pack_f16x2(acc[0].x, acc[0].y);
        o.y = pack_f16x2(acc[1].x, acc[1].y);
        o.z = pack_f16x2(acc[2].x, acc[2].y);
        o.w = pack_f16x2(acc[3].x, acc[3].y);
        f16* dst = sy + size_t(pix) * 64 + ((cg ^ (((f1 >> 1) & 1) << 2)) << 3);
        *reinterpret_cast<uint4*>(dst) = o;
        if (split)
          *reinterpret_cast<uint4*>(dst + sy_lo) =
              make_uint4(f16x2_residual(acc[0].x, acc[0].y, o.x), f16x2_residual(acc[1].x, acc[1].y, o.y),
                         f16x2_residual(acc[2].x, acc[2].y, o.z), f16x2_residual(acc[3].x, acc[3].y, o.w));
      }
    }
    __syncthreads();
    if (tile + int(gridDim.x) < n_tiles) load_patch(tile + gridDim.x);   // sx is free: prefetch under the MMAs
    // ---- conv2 + ReLU: implicit GEMM, 16 output pixels x 64 channels per warp pass
    const int npx = ntt * F2, n_mt = (npx + 15) >> 4;
    for (int mt = warp; mt < n_mt; mt += CONVF_THREADS / 32) {
      const int m_lo = mt * 16 + g, m_hi = m_lo + 8;
      int pbase[2], fpar[2];
      size_t zo[2];
      const size_t zrow = size_t(split ? 2 : 1) * F2 * 64;   // output row (b, t2): [hi (F2*64) | lo (F2*64)] when split
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = min(i ? m_hi : m_lo, npx - 1);       // clamp (stores are predicated)
        const int tt = m / F2, f2 = m - tt * F2;
        pbase[i] = (2 * tt) * F1 + 2 * f2;
        fpar[i] = f2 & 1;
        zo[i] = size_t(tt) * zrow + size_t(f2) * 64;
      }
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap - kh * 3;
        const int poff = kw * F1 + kh;
        const int s0 = ((fpar[0] + (kh >> 1)) & 1) << 2, s1 = ((fpar[1] + (kh >> 1)) & 1) << 2;   // swizzle of the input pixel
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const f16* a0 = sy + size_t(pbase[0] + poff) * 64 + (((half * 4 + c) ^ s0) << 3);
          const f16* a1 = sy + size_t(pbase[1] + poff) * 64 + (((half * 4 + c) ^ s1) << 3);
          const uint4 vlo = *reinterpret_cast<const uint4*>(a0);
          const uint4 vhi = *reinterpret_cast<const uint4*>(a1);
          uint4 ulo = make_uint4(0, 0, 0, 0), uhi = make_uint4(0, 0, 0, 0);
          if (split) {
            ulo = *reinterpret_cast<const uint4*>(a0 + sy_lo);
            uhi = *reinterpret_cast<const uint4*>(a1 + sy_lo);
          }
          const uint2* wk = sw + size_t((tap * 2 + half) * 2) * 8 * 32 + lane;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const uint2 w0 = wk[nt * 32];
            const uint2 w1v = wk[8 * 32 + nt * 32];
            mma_f16_16816(acc[nt], vlo.x, vhi.x, vlo.y, vhi.y, w0.x, w0.y);
            mma_f16_16816(acc[nt], vlo.z, vhi.z, vlo.w, vhi.w, w1v.x, w1v.y);
            if (split) {   // lo parts of the activations against the same weight fragments
              mma_f16_16816(acc[nt], ulo.x, uhi.x, ulo.y, uhi.y, w0.x, w0.y);
              mma_f16_16816(acc[nt], ulo.z, uhi.z, ulo.w, uhi.w, w1v.x, w1v.y);
            }
          }
        }
      }
      f16* zb = z + (size_t(b) * T2 + t20) * zrow;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int co = nt * 8 + 2 * c;
        const float bb0 = __ldg(b2 + co), bb1 = __ldg(b2 + co + 1);
        store_pair_split(zb + zo[0] + co, m_lo < npx, fmaxf(acc[nt][0] + bb0, 0.f), fmaxf(acc[nt][1] + bb1, 0.f),
                         split ? size_t(F2) * 64 : 0);
        store_pair_split(zb + zo[1] + co, m_hi < npx, fmaxf(acc[nt][2] + bb0, 0.f), fmaxf(acc[nt][3] + bb1, 0.f),
                         split ? size_t(F2) * 64 : 0);
      }
    }
  }
}

size_t conv_fused_smem(int F, int F1, int TT2, int split) {
  const size_t W = 4 * size_t(TT2) + 3;
  return CONV2_W_BYTES + (9 * 64 + 64) * 4 + ((size_t(F) * W + 3) & ~size_t(3)) * 4 +
         (2 * size_t(TT2) + 1) * F1 * 128 * (split ? 2 : 1);
}

}  // namespace

int launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int D, float eps, float* y_f32,
                     f16* y_f16, cudaStream_t s, int split) {
  if (rows <= 0) return 0;
  if (D % 128 != 0 || D > 128 * LN_MAX_CHUNKS) return set_error(-2, "layernorm: D=%d must be a multiple of 128, <= 1024", D);
  const int rows_per_block = 8;
  layernorm_kernel<<<(rows + rows_per_block - 1) / rows_per_block, rows_per_block * 32, 0, s>>>(x, gamma, beta, rows,
                                                                                                  D, eps, y_f32, y_f16, split);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_embed_pe(const int32_t* tokens, int ld_tok, const float* emb, const float* pe, int B, int L, int D,
                    int vocab, float* out, cudaStream_t s) {
  if (B <= 0 || L <= 0) return 0;
  embed_pe_kernel<<<dim3(L, B), 64, 0, s>>>(tokens, ld_tok, emb, pe, L, D, vocab, out);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_f32_to_f16(const float* x, f16* y, size_t n, cudaStream_t s) {
  if (n == 0) return 0;
  const size_t n4 = n / 4;
  if (n4) {
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    f32_to_f16_kernel<<<blocks, 256, 0, s>>>(x, y, n4);
  }
  if (n % 4) f32_to_f16_tail_kernel<<<1, 4, 0, s>>>(x, y, n4 * 4, n);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED((n4 ? 1 : 0) + (n % 4 ? 1 : 0));
  return 0;
}

int launch_f32_to_f16_split(const float* x, f16* y, size_t rows, int D, cudaStream_t s) {
  if (rows == 0) return 0;
  if (D % 8 != 0) return set_error(-2, "f32_to_f16_split: D=%d must be a multiple of 8", D);
  const size_t n4 = rows * (D / 4);
  const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  f32_to_f16_split_kernel<<<blocks, 256, 0, s>>>(x, y, rows, D);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_conv1(const float* spec, const float* w1, const float* b1, int B, int F, int T, f16* y1, cudaStream_t s,
                 int split) {
  const int F1 = (F - 3) / 2 + 1, T1 = (T - 3) / 2 + 1;
  if (F1 <= 0 || T1 <= 0) return set_error(-2, "conv1: input %dx%d too small", F, T);
  const size_t smem = (size_t(9 * 64 + 64) + size_t(F) * (2 * CONV1_TT + 1)) * sizeof(float);
  if (smem > 200 * 1024) return set_error(-2, "conv1: input_dim %d too large for the shared-memory patch", F);
  if (int rc = ensure_dyn_smem((const void*)conv1_kernel, smem)) return rc;
  dim3 grid((T1 + CONV1_TT - 1) / CONV1_TT, B);
  conv1_kernel<<<grid, 256, smem, s>>>(spec, w1, b1, B, F, T, F1, T1, y1, split ? size_t(B) * T1 * F1 * 64 : 0);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

// Fused front-end; returns 1 (and launches nothing) when no tile size fits the shared memory, so that the caller can
// fall back to conv1_kernel + conv2_kernel.
int launch_conv_fused(const float* spec, const float* w1, const float* b1, const f16* w2frag, const float* b2, int B,
                      int F, int T, f16* z, cudaStream_t s, int split) {
  const int F1 = (F - 3) / 2 + 1, T1 = (T - 3) / 2 + 1, F2 = (F1 - 3) / 2 + 1, T2 = (T1 - 3) / 2 + 1;
  if (F2 <= 0 || T2 <= 0) return set_error(-2, "conv: input %dx%d too small", F, T);
  int n_sm = 0, max_smem = 0;
  if (int rc = device_props(&n_sm, &max_smem)) return rc;
  // frames per tile: fill the 8 warps' 16-pixel passes (128 pixels) without exceeding the shared memory
  int TT2 = 0;
  for (int t = 1; t <= 8 && t <= T2; ++t)
    if (conv_fused_smem(F, F1, t, split) <= size_t(max_smem) && t * F2 <= 128) TT2 = t;
  if (TT2 == 0 && conv_fused_smem(F, F1, 1, split) <= size_t(max_smem)) TT2 = 1;
  if (TT2 == 0) return 1;
  const size_t smem = conv_fused_smem(F, F1, TT2, split);
  if (int rc = ensure_dyn_smem((const void*)conv_fused_kernel, smem)) return rc;
  const int n_tt = (T2 + TT2 - 1) / TT2;
  const long long n_tiles = (long long)B * n_tt;
  const int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);
  conv_fused_kernel<<<grid, CONVF_THREADS, smem, s>>>(spec, w1, b1, reinterpret_cast<const uint2*>(w2frag), b2, B, F, T,
                                                      F1, F2, T2, TT2, n_tt, z, split);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_conv2(const f16* y1, const f16* w2frag, const float* b2, int B, int F1, int T1, f16* z, cudaStream_t s,
                 int split) {
  const int F2 = (F1 - 3) / 2 + 1, T2 = (T1 - 3) / 2 + 1;
  if (F2 <= 0 || T2 <= 0) return set_error(-2, "conv2: input %dx%d too small", F1, T1);
  if (int rc = ensure_dyn_smem((const void*)conv2_kernel, CONV2_W_BYTES)) return rc;
  const long long ntiles = ((long long)B * T2 * F2 + 63) / 64;
  const int blocks = (int)(ntiles < 148 * 3 ? ntiles : 148 * 3);
  conv2_kernel<<<blocks, 128, CONV2_W_BYTES, s>>>(y1, reinterpret_cast<const uint2*>(w2frag), b2, B, F1, T1, F2, T2, z,
                                                  split ? size_t(B) * T1 * F1 * 64 : 0);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
