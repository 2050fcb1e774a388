// Persistent cooperative greedy decoder: ONE launch runs all L decode steps of the whole utterance batch.
//
// Restates reference model.py:125-151 (Decoder.evaluate) with a device-resident KV cache; no host sync, no kernel
// boundary per token.  One CTA per SM (cooperative launch), phases separated by grid-wide barriers:
//
//   per layer:  A  LN1 + packed QKV linear (grid-split over output columns, K/V rows appended to the cache)   | barrier
//               B  utterance-local chain in ONE CTA per utterance: causal self attention -> out projection + residual
//                  -> LN2 -> cross-attention query -> cross attention over the precomputed encoder K/V -> out
//                  projection + residual  (no global traffic between the sub-steps, everything stays in smem)     | barrier
//               C  LN3 + FFN squeeze + ReLU (grid-split)                                                          | barrier
//               D  FFN unsqueeze + residual (grid-split)                                                          | barrier
//   per step:   E  utterance-local: classifier (no final LayerNorm, model.py:142) -> argmax (lowest index wins)
//                  -> EOS bookkeeping -> embedding + PE of the next token                                         | barrier
//
// Memory system design (this path is bound by HBM/L2 traffic and latency, not by math):
//   * the encoder K/V of an (utterance, layer) - the dominant stream, 1.53 MB per utterance-step at the default
//     model - is pulled by TMA bulk copies (cp.async.bulk, L2 evict-first) into a 3-stage shared-memory ring; the
//     first stages of the NEXT layer are requested as soon as the current layer's cross attention is done, so they
//     land while the grid-split FFN / QKV phases run;
//   * weights are read with an L2 evict-last policy so the 11 MB weight set stays L2-resident under that stream;
//   * every loop keeps >= 8 independent 16-byte loads in flight per lane before the first use.
// Precision: Linear layers see fp32-accurate activations (bf16 hi + lo split, two tensor-core passes per weight
// tile, fp32 accumulate); K/V caches bf16; attention scores / softmax / residual / LayerNorm fp32.
// Data written by one CTA and read by another after a barrier is loaded with ld.global.cg (L2), never through L1.
#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

constexpr int NT = 256;              // threads per CTA
constexpr int NW = NT / 32;
constexpr int MROWS = 16;            // batch rows per grid-split work item (one m16 MMA tile)
constexpr int PAD = 32;              // smem row padding (elements): row stride == 64 B mod 128 B -> conflict-free LDS.128
constexpr int RING = 3;              // cross-K/V ring stages
constexpr int CHUNK_BYTES = 32768;   // bytes per ring stage (= 128 / H keys of all heads, K and V)

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// read-only 16-byte weight load that asks L2 to keep the line (weights are re-read every step by every CTA)
__device__ __forceinline__ uint4 ldg_keep(const bf16* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier, L2 evict-first (streamed once per step)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// Team-wide barrier on a monotonically increasing counter (zeroed by the launcher).  Polling uses relaxed gpu-scope
// loads (served by L2, no L1 invalidation per poll); one fence on each side orders the data.  Bounded spin: a
// protocol bug becomes a launch failure, never a hung GPU.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, int ncta) {
  target += ncta;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (ld_relaxed_u32(counter) < target) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

// split fp32 -> bf16 hi + lo and store 4 consecutive elements
__device__ __forceinline__ void store_hilo4(bf16* hi, bf16* lo, float4 v) {
  const bf16 h0 = __float2bfloat16(v.x), h1 = __float2bfloat16(v.y), h2 = __float2bfloat16(v.z), h3 = __float2bfloat16(v.w);
  uint2 a, b;
  a.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  a.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
  b.x = pack_bf16x2(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
  b.y = pack_bf16x2(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
  *reinterpret_cast<uint2*>(hi) = a;
  *reinterpret_cast<uint2*>(lo) = b;
}

struct Smem {
  bf16* hi;        // [MROWS][kmax + PAD]
  bf16* lo;
  float* red;      // [NW tiles][8 K chunks][32 lanes][4] partial sums
  float* vec;      // utterance-local fp32 vectors: h[D], x[D] (attention out / LN out), q[D], logits[V]
  float* part;     // [NW][64] partial attention outputs
  float* stat;     // [0,2NW) softmax stats, [32] token, [40,72) phase timers
  uint8_t* ring;   // [RING][CHUNK_BYTES] cross-K/V stages
  uint64_t* full;  // [RING] mbarriers
};

enum Epi { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_QKV = 2 };

struct LinArgs {
  const float* x; int ldx; int K;
  const float* ln_g; const float* ln_b;     // nullable: LayerNorm prologue over K
  const bf16* w; const float* bias; int N;  // W [N_pad][K] bf16
  int relu; int epi;
  float* out; int ldo;                      // EPI_RESIDUAL: out += (in place)
  bf16* cache; int cache_rows; int cache_col0; int step;   // EPI_QKV
};

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core part of the grid-split linear.  lpu = 16-byte weight loads per (tile, K chunk) unit (1, 2, 4 or 8);
// 8 / lpu units are kept in flight per warp so that 8 independent loads are outstanding before the first MMA.
__device__ __forceinline__ void mma_units(const LinArgs& a, const Smem& sm, int nunits, int nch, int kc, int slot, int S,
                                          int base, int ld, uint64_t pol_w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int lpu = kc >> 5;                       // power of two
  const int lsh = 31 - __clz(lpu);
  const int uf = 8 >> lsh;
  float4* red = reinterpret_cast<float4*>(sm.red);
  for (int j0 = 0; warp + j0 * NW < nunits; j0 += uf) {
    uint4 wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = j >> lsh, u = j & (lpu - 1);
      const int q = warp + (j0 + f) * NW;
      if (q < nunits) {
        const int ti = q / nch, ch = q % nch;
        const int tile = slot + (base + ti) * S;
        wv[j] = ldg_keep(a.w + size_t(tile * 8 + g) * a.K + ch * kc + c * 8 + u * 32, pol_w);
      }
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = j >> lsh, u = j & (lpu - 1);
      const int q = warp + (j0 + f) * NW;
      if (q < nunits) {
        const int ti = q / nch, ch = q % nch;
        if (u == 0) acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        const bf16* ah = sm.hi + g * ld + ch * kc + c * 8 + u * 32;
        const bf16* al = sm.lo + g * ld + ch * kc + c * 8 + u * 32;
        const uint4 h0 = *reinterpret_cast<const uint4*>(ah);
        const uint4 h1 = *reinterpret_cast<const uint4*>(ah + 8 * ld);
        const uint4 l0 = *reinterpret_cast<const uint4*>(al);
        const uint4 l1 = *reinterpret_cast<const uint4*>(al + 8 * ld);
        mma16816(acc, h0.x, h1.x, h0.y, h1.y, wv[j].x, wv[j].y);
        mma16816(acc, h0.z, h1.z, h0.w, h1.w, wv[j].z, wv[j].w);
        mma16816(acc, l0.x, l1.x, l0.y, l1.y, wv[j].x, wv[j].y);
        mma16816(acc, l0.z, l1.z, l0.w, l1.w, wv[j].z, wv[j].w);
        if (u == lpu - 1) red[(ti * 8 + ch) * 32 + lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      }
    }
  }
}

// Grid-split linear: out[B, N] = epi( LN?(x)[B, K] * W^T + bias ).  CTA -> (16-row block, set of 8-column tiles);
// the CTA's 8 warps are spread over (tile, K chunk) units; partial sums are reduced through smem in a fixed order.
__device__ __forceinline__ void linear_phase(const LinArgs& a, int B, const Smem& sm, uint64_t pol_w, int cta,
                                             int ncta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int MB = (B + MROWS - 1) / MROWS;
  const int S = ncta / MB;                  // CTAs ("slots") per row block
  if (S == 0 || cta >= MB * S) return;
  const int mb = cta % MB, slot = cta / MB;
  const int ntiles = (a.N + 7) / 8;
  if (slot >= ntiles) return;
  const int n_my = (ntiles - slot + S - 1) / S;        // tiles slot, slot+S, ...
  const int ld = a.K + PAD;
  const int row0 = mb * MROWS;

  // ---- stage activations (optional LayerNorm) as bf16 hi/lo: warp w owns rows w and w + 8, both loaded up front
  if (a.ln_g) {                              // K = D <= 512 on this path (launcher check)
    float4 v[2][4];
    float mean[2], rstd[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = row0 + warp + rr * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i * 128 < a.K)
          v[rr][i] = row < B ? ldcg4(a.x + size_t(row) * a.ldx + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i * 128 < a.K) sum += v[rr][i].x + v[rr][i].y + v[rr][i].z + v[rr][i].w;
      mean[rr] = warp_sum(sum) / float(a.K);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i * 128 < a.K) {
          const float d0 = v[rr][i].x - mean[rr], d1 = v[rr][i].y - mean[rr], d2 = v[rr][i].z - mean[rr],
                      d3 = v[rr][i].w - mean[rr];
          sq += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      rstd[rr] = 1.0f / sqrtf(warp_sum(sq) / float(a.K) + 1e-5f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i * 128 < a.K) {
        const int k = i * 128 + lane * 4;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(a.ln_g + k));
        const float4 bt = __ldg(reinterpret_cast<const float4*>(a.ln_b + k));
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          float4 o;
          o.x = (v[rr][i].x - mean[rr]) * rstd[rr] * gm.x + bt.x;
          o.y = (v[rr][i].y - mean[rr]) * rstd[rr] * gm.y + bt.y;
          o.z = (v[rr][i].z - mean[rr]) * rstd[rr] * gm.z + bt.z;
          o.w = (v[rr][i].w - mean[rr]) * rstd[rr] * gm.w + bt.w;
          const int r = warp + rr * 8;
          store_hilo4(sm.hi + r * ld + k, sm.lo + r * ld + k, o);
        }
      }
  } else {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp + rr * 8, row = row0 + r;
      const float* xr = a.x + size_t(row) * a.ldx;
#pragma unroll 4
      for (int k = lane * 4; k < a.K; k += 128)
        store_hilo4(sm.hi + r * ld + k, sm.lo + r * ld + k, row < B ? ldcg4(xr + k) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  __syncthreads();

  // ---- tensor-core passes.  K is cut into a FIXED number of chunks that depends on K only (never on the batch or
  // the grid), every (tile, chunk) partial sum starts from zero and the chunks are added in index order, so the
  // result of a row is bit-identical whatever batch it is decoded in (batch invariance).
  int nch = 8;
  while (nch > 1 && a.K % (32 * nch) != 0) nch >>= 1;
  const int kc = a.K / nch;                                  // multiple of 32, <= 256 for K <= 2048
  float4* red = reinterpret_cast<float4*>(sm.red);           // [NW tiles][8 chunks][32 lanes]
  for (int base = 0; base < n_my; base += NW) {
    const int nr = min(NW, n_my - base);
    mma_units(a, sm, nr * nch, nch, kc, slot, S, base, ld, pol_w);
    __syncthreads();
    if (warp < nr) {
      const int tile = slot + (base + warp) * S;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ch = 0; ch < nch; ++ch) {
        const float4 t = red[(warp * 8 + ch) * 32 + lane];
        acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
      }
      const int col = tile * 8 + 2 * c;
      if (col < a.N) {
        const bool two = col + 1 < a.N;
        const float b0 = a.bias ? __ldg(a.bias + col) : 0.f;
        const float b1 = (a.bias && two) ? __ldg(a.bias + col + 1) : 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = row0 + g + hh * 8;
          if (row >= B) continue;
          float v0 = acc[hh * 2] + b0, v1 = acc[hh * 2 + 1] + b1;
          if (a.relu) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          float* op = a.out + size_t(row) * a.ldo + col;
          if (a.epi == EPI_RESIDUAL) {
            v0 += __ldcg(op);
            if (two) v1 += __ldcg(op + 1);
          }
          op[0] = v0;
          if (two) op[1] = v1;
          if (a.epi == EPI_QKV && col >= a.cache_col0) {
            const int w = a.N - a.cache_col0;
            bf16* kp = a.cache + (size_t(row) * a.cache_rows + a.step) * w + (col - a.cache_col0);
            kp[0] = __float2bfloat16(v0);
            if (two) kp[1] = __float2bfloat16(v1);
          }
        }
      }
    }
    if (base + NW < n_my) __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Utterance-local helpers (one CTA = one utterance, M = 1)

// y[n] = W[n, :] . x + bias[n] for n < N; x: fp32 vector in smem (exact, no bf16 split needed on this path),
// W bf16 [N_pad][K], K % 128 == 0.  CUDA-core FMAs: at M = 1 the tensor cores would be 1/16 used and their dependent
// accumulate chain is the latency bottleneck.  Warp w handles 8-row tiles w, w+8, ...; lane (g, c) owns row g of the
// tile and the K slices [kb + 8c, kb + 8c + 8): 16-byte weight loads, 8 independent FMA chains, 4-lane reduction.
// Two tiles x 128 K elements (8 loads per lane) are requested before the first FMA.
__device__ __forceinline__ void fma8(float (&acc)[8], const uint4 w, const float4 x0, const float4 x1) {
  const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&w);
  const float2 a = __bfloat1622float2(w2[0]), b = __bfloat1622float2(w2[1]), c = __bfloat1622float2(w2[2]),
               d = __bfloat1622float2(w2[3]);
  acc[0] = fmaf(a.x, x0.x, acc[0]); acc[1] = fmaf(a.y, x0.y, acc[1]);
  acc[2] = fmaf(b.x, x0.z, acc[2]); acc[3] = fmaf(b.y, x0.w, acc[3]);
  acc[4] = fmaf(c.x, x1.x, acc[4]); acc[5] = fmaf(c.y, x1.y, acc[5]);
  acc[6] = fmaf(d.x, x1.z, acc[6]); acc[7] = fmaf(d.y, x1.w, acc[7]);
}
__device__ __forceinline__ void matvec_cta(const float* x, int K, const bf16* W, const float* bias, int N, float* y,
                                           uint64_t pol_w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int ntiles = (N + 7) / 8;
  for (int tile0 = warp; tile0 < ntiles; tile0 += 2 * NW) {
    const int tile1 = tile0 + NW;
    const bool has1 = tile1 < ntiles;
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    const bf16* w0 = W + size_t(tile0 * 8 + g) * K + c * 8;
    const bf16* w1 = W + size_t((has1 ? tile1 : tile0) * 8 + g) * K + c * 8;
#pragma unroll 1
    for (int kb = 0; kb < K; kb += 128) {
      uint4 wa[4], wb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        wa[u] = ldg_keep(w0 + kb + u * 32, pol_w);
        wb[u] = ldg_keep(w1 + kb + u * 32, pol_w);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* xp = x + kb + u * 32 + c * 8;
        const float4 x0 = *reinterpret_cast<const float4*>(xp);
        const float4 x1 = *reinterpret_cast<const float4*>(xp + 4);
        fma8(a0, wa[u], x0, x1);
        fma8(a1, wb[u], x0, x1);
      }
    }
    float r0 = ((a0[0] + a0[1]) + (a0[2] + a0[3])) + ((a0[4] + a0[5]) + (a0[6] + a0[7]));
    float r1 = ((a1[0] + a1[1]) + (a1[2] + a1[3])) + ((a1[4] + a1[5]) + (a1[6] + a1[7]));
    r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
    r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
    if (c == 0) {
      const int n0 = tile0 * 8 + g, n1 = tile1 * 8 + g;
      if (n0 < N) y[n0] = r0 + (bias ? __ldg(bias + n0) : 0.f);
      if (has1 && n1 < N) y[n1] = r1 + (bias ? __ldg(bias + n1) : 0.f);
    }
  }
}

// in-place LayerNorm of a smem fp32 vector (warp 0); other warps wait at the caller's __syncthreads
__device__ __forceinline__ void ln_vec(float* x, int D, const float* g, const float* b) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  float sum = 0.f;
  for (int k = lane; k < D; k += 32) sum += x[k];
  const float mean = warp_sum(sum) / float(D);
  float sq = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float d = x[k] - mean;
    sq += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / float(D) + 1e-5f);
  for (int k = lane; k < D; k += 32) x[k] = (x[k] - mean) * rstd * __ldg(g + k) + __ldg(b + k);
}

// ---- single-query attention for all H heads of one utterance, flash style.
// Warp w serves head w % H with key subset (w / H) of NW / H; 8 lanes per 128-byte head row, 4 key groups per warp.
// Every group keeps a running (max, sum, 8-dim accumulator) in log2 units; groups and warps are merged at the end.
struct AttnState {
  float qv[8];
  float m, l;
  float o[8];
};

__device__ __forceinline__ void attn_begin(AttnState& st, const float* q, int H, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c8 = lane & 7;
  const int h = warp % H;
  const float4 a = *reinterpret_cast<const float4*>(q + h * 64 + c8 * 8);
  const float4 b = *reinterpret_cast<const float4*>(q + h * 64 + c8 * 8 + 4);
  const float sc = scale * 1.4426950408889634f;    // p = exp2(s - m)
  st.qv[0] = a.x * sc; st.qv[1] = a.y * sc; st.qv[2] = a.z * sc; st.qv[3] = a.w * sc;
  st.qv[4] = b.x * sc; st.qv[5] = b.y * sc; st.qv[6] = b.z * sc; st.qv[7] = b.w * sc;
  st.m = -INFINITY;
  st.l = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) st.o[i] = 0.f;
}

// fold NB keys (rows already in registers) into the running state; valid[u] masks the tail
template <int NB>
__device__ __forceinline__ void attn_fold(AttnState& st, const uint4 (&kr)[NB], const uint4 (&vr)[NB],
                                          const bool (&valid)[NB], unsigned gmask) {
  float sc[NB];
  float bm = -INFINITY;
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    float sv = 0.f;
    if (valid[u]) {
      const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kr[u]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(k2[i]);
        sv = fmaf(st.qv[2 * i], f.x, sv);
        sv = fmaf(st.qv[2 * i + 1], f.y, sv);
      }
    }
    sv += __shfl_xor_sync(gmask, sv, 1);
    sv += __shfl_xor_sync(gmask, sv, 2);
    sv += __shfl_xor_sync(gmask, sv, 4);
    sc[u] = valid[u] ? sv : -INFINITY;
    bm = fmaxf(bm, sc[u]);
  }
  if (bm == -INFINITY) return;                    // no valid key in this batch (uniform inside the 8-lane group)
  const float m_new = fmaxf(st.m, bm);
  const float alpha = exp2f(st.m - m_new);        // st.m == -inf -> 0
  st.l *= alpha;
#pragma unroll
  for (int i = 0; i < 8; ++i) st.o[i] *= alpha;
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    if (valid[u]) {
      const float pw = exp2f(sc[u] - m_new);
      st.l += pw;
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vr[u]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(v2[i]);
        st.o[2 * i] = fmaf(pw, f.x, st.o[2 * i]);
        st.o[2 * i + 1] = fmaf(pw, f.y, st.o[2 * i + 1]);
      }
    }
  }
  st.m = m_new;
}

// keys straight from global memory (self-attention cache, written by other CTAs before the barrier -> ld.global.cg)
__device__ __forceinline__ void attn_global(AttnState& st, const bf16* kbase, const bf16* vbase, int ldkv, int n, int H) {
  constexpr int AU = 8;                              // keys in flight per group: 16 x 16 B per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c8 = lane & 7, sub = lane >> 3;
  const int wph = NW / H, h = warp % H, part = warp / H;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int stride = wph * 4;
  const bf16* kp = kbase + h * 64 + c8 * 8;
  const bf16* vp = vbase + h * 64 + c8 * 8;
  for (int k0 = part * 4 + sub; k0 < n; k0 += stride * AU) {   // trip count is uniform inside an 8-lane group
    uint4 kr[AU], vr[AU];
    bool valid[AU];
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      const int kj = k0 + u * stride;
      valid[u] = kj < n;
      if (valid[u]) {
        kr[u] = __ldcg(reinterpret_cast<const uint4*>(kp + size_t(kj) * ldkv));
        vr[u] = __ldcg(reinterpret_cast<const uint4*>(vp + size_t(kj) * ldkv));
      }
    }
    attn_fold<AU>(st, kr, vr, valid, gmask);
  }
}

// one ring stage: nk key rows of [K(H*64) | V(H*64)] bf16 in shared memory; each group owns 4 key slots per chunk
__device__ __forceinline__ void attn_chunk(AttnState& st, const uint8_t* stage, int nk, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c8 = lane & 7, sub = lane >> 3;
  const int wph = NW / H, h = warp % H, part = warp / H;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int stride = wph * 4;
  const int row_bytes = H * 256;                     // K and V of all heads
  const uint8_t* base = stage + h * 128 + c8 * 16;
  uint4 kr[4], vr[4];
  bool valid[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int kl = part * 4 + sub + u * stride;
    valid[u] = kl < nk;
    if (valid[u]) {
      kr[u] = *reinterpret_cast<const uint4*>(base + size_t(kl) * row_bytes);
      vr[u] = *reinterpret_cast<const uint4*>(base + size_t(kl) * row_bytes + H * 128);
    }
  }
  attn_fold<4>(st, kr, vr, valid, gmask);
}

// merge the key groups of each warp and the warps of each head; out: smem fp32 [H*64]
__device__ __forceinline__ void attn_finish(AttnState& st, int H, const Smem& sm, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c8 = lane & 7, sub = lane >> 3;
  const int wph = NW / H;
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, st.m, off);
    const float lo = __shfl_xor_sync(0xffffffffu, st.l, off);
    const float mn = fmaxf(st.m, mo);
    const float fa = (st.m == -INFINITY) ? 0.f : exp2f(st.m - mn);
    const float fb = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
    st.l = st.l * fa + lo * fb;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float oo = __shfl_xor_sync(0xffffffffu, st.o[i], off);
      st.o[i] = st.o[i] * fa + oo * fb;
    }
    st.m = mn;
  }
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm.part[warp * 64 + c8 * 8 + i] = st.o[i];
    if (c8 == 0) {
      sm.stat[warp] = st.m;
      sm.stat[NW + warp] = st.l;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < H * 64; d += NT) {
    const int hh = d >> 6, dd = d & 63;
    float mm = -INFINITY;
    for (int pI = 0; pI < wph; ++pI) mm = fmaxf(mm, sm.stat[pI * H + hh]);
    float t = 0.f, ls = 0.f;
    for (int pI = 0; pI < wph; ++pI) {
      const float mw = sm.stat[pI * H + hh];
      const float f = (mw == -INFINITY) ? 0.f : exp2f(mw - mm);
      t += sm.part[(pI * H + hh) * 64 + dd] * f;
      ls += sm.stat[NW + pI * H + hh] * f;
    }
    out[d] = ls > 0.f ? t / ls : 0.f;                // fully masked -> zeros (layers.py:25)
  }
  __syncthreads();
}

// request ring stages [c0, c1) of one (utterance, layer) encoder K/V block; thread 0 only
__device__ __forceinline__ void ring_issue(const Smem& sm, const bf16* ck, int Tp, int H, int ck_keys, int c0, int c1,
                                           uint64_t pol) {
  const int row_bytes = H * 256;
  for (int c = c0; c < c1; ++c) {
    const int s = c % RING;
    const int nk = min(ck_keys, Tp - c * ck_keys);
    const uint32_t bytes = uint32_t(nk) * row_bytes;
    mbar_expect_tx(&sm.full[s], bytes);
    bulk_g2s(sm.ring + size_t(s) * CHUNK_BYTES, reinterpret_cast<const uint8_t*>(ck) + size_t(c) * ck_keys * row_bytes,
             bytes, &sm.full[s], pol);
  }
}

__global__ void __launch_bounds__(NT, 1) dec_persistent_kernel(const __grid_constant__ PersistentParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem sm;
  {
    uint8_t* ptr = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int ld = p.kmax + PAD;
    sm.ring = ptr; ptr += size_t(RING) * CHUNK_BYTES;
    sm.hi = reinterpret_cast<bf16*>(ptr); ptr += size_t(MROWS) * ld * 2;
    sm.lo = reinterpret_cast<bf16*>(ptr); ptr += size_t(MROWS) * ld * 2;
    sm.red = reinterpret_cast<float*>(ptr); ptr += NW * 8 * 32 * 16;
    sm.vec = reinterpret_cast<float*>(ptr); ptr += size_t(4) * p.kmax * 4;
    sm.part = reinterpret_cast<float*>(ptr); ptr += NW * 64 * 4;
    sm.stat = reinterpret_cast<float*>(ptr); ptr += 80 * 4;
    sm.full = reinterpret_cast<uint64_t*>(ptr);
  }
  // Teams: the grid is cut into p.teams independent groups of CTAs; each decodes its own contiguous slice of the
  // batch with its own barrier, so the (latency-bound) phase chains of different slices overlap on the machine.
  const int ncta = gridDim.x / p.teams;
  const int team = blockIdx.x / ncta, cta = blockIdx.x % ncta;
  const int per_team = (p.B + p.teams - 1) / p.teams;
  const int b0 = team * per_team;
  const int B = min(per_team, p.B - b0);             // utterances of this team
  if (team >= p.teams || B <= 0) return;
  const int D = p.D, H = p.H;
  const uint64_t pol_w = l2_policy_evict_last();
  const uint64_t pol_kv = l2_policy_evict_first();
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) mbar_init(&sm.full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  unsigned* bar_counter = p.barrier + team * 32;      // 128 B apart
  unsigned* done_counter = p.done_count + team * 32;
  float* g_h = p.h + size_t(b0) * D;
  float* g_qkv = p.qkv + size_t(b0) * 3 * D;
  float* g_ff = p.ff + size_t(b0) * p.FF;
  int32_t* g_tokens = p.tokens + size_t(b0) * (p.L + 1);
  int32_t* g_ntok = p.n_tokens ? p.n_tokens + b0 : nullptr;
  int32_t* g_fin = p.finished + b0;
  float* g_logits = p.step_logits ? p.step_logits + size_t(b0) * p.L * p.V : nullptr;
  unsigned target = 0;
  unsigned ring_parity = 0;            // bit s = parity of the next completion to wait for on stage s
  int pre_u = -1, pre_l = -1;          // (utterance, layer) whose first ring stages have already been requested
  const int ck_keys = CHUNK_BYTES / (H * 256);                 // keys per ring stage
  const int nchunks = (p.Tp + ck_keys - 1) / ck_keys;
  // optional per-phase cycle accounting (asr_decode_profile): thread 0 of every CTA accumulates clock64 deltas
  long long* tacc = reinterpret_cast<long long*>(sm.stat + 40);
  long long t_prev = 0;
  if (p.timing && threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
    t_prev = clock64();
  }
#define PHASE_DONE(idx)                                  \
  if (p.timing && threadIdx.x == 0) {                    \
    const long long now_ = clock64();                    \
    tacc[idx] += now_ - t_prev;                          \
    t_prev = now_;                                       \
  }
  // sub-phases of B (slots 10..15) are measured with their own clock and do not disturb the phase clock
  long long t_sub = 0;
#define SUB_START() if (p.timing && threadIdx.x == 0) t_sub = clock64();
#define SUB_DONE(idx)                                    \
  if (p.timing && threadIdx.x == 0) {                    \
    const long long now_ = clock64();                    \
    tacc[idx] += now_ - t_sub;                           \
    t_sub = now_;                                        \
  }
  float* v_h = sm.vec;                 // [D] residual row
  float* v_x = sm.vec + p.kmax;        // [D] attention output / scratch
  float* v_q = sm.vec + 2 * p.kmax;    // [D] query
  float* v_l = sm.vec + 3 * p.kmax;    // logits (V <= kmax)

  // the encoder K/V of this CTA's first utterance, layer 0, can be requested before anything else runs
  if (cta < B) {
    if (threadIdx.x == 0)
      ring_issue(sm, p.ckv + size_t(b0 + cta) * p.Tp * 2 * D, p.Tp, H, ck_keys, 0, min(RING, nchunks), pol_kv);
    pre_u = cta;
    pre_l = 0;
  }

  for (int t = 0; t < p.L; ++t) {
    for (int l = 0; l < p.nd; ++l) {
      const PersistentLayer& w = p.layer[l];
      bf16* cache = p.cache + (size_t(l) * p.B + b0) * p.L * 2 * D;          // this team's slice of layer l
      const bf16* ckv = p.ckv + (size_t(l) * p.B + b0) * p.Tp * 2 * D;
      for (int ph = 0; ph < 4; ++ph) {
        if (ph != 1) {
          // ---- A: LN1 + QKV, append K/V (model.py:67-68, layers.py:16-18)
          // ---- C: LN3 + FFN squeeze + ReLU (model.py:73-74, layers.py:54-55)
          // ---- D: FFN unsqueeze + residual
          LinArgs a;
          a.x = ph == 3 ? g_ff : g_h;
          a.ldx = a.K = ph == 3 ? p.FF : D;
          a.ln_g = ph == 0 ? w.ln1_g : (ph == 2 ? w.ln3_g : nullptr);
          a.ln_b = ph == 0 ? w.ln1_b : (ph == 2 ? w.ln3_b : nullptr);
          a.w = ph == 0 ? w.w_qkv : (ph == 2 ? w.w1 : w.w2);
          a.bias = ph == 0 ? w.b_qkv : (ph == 2 ? w.b1 : w.b2);
          a.N = ph == 0 ? 3 * D : (ph == 2 ? p.FF : D);
          a.relu = ph == 2;
          a.epi = ph == 0 ? EPI_QKV : (ph == 2 ? EPI_STORE : EPI_RESIDUAL);
          a.out = ph == 0 ? g_qkv : (ph == 2 ? g_ff : g_h);
          a.ldo = a.N;
          a.cache = cache; a.cache_rows = p.L; a.cache_col0 = D; a.step = t;
          linear_phase(a, B, sm, pol_w, cta, ncta);
        } else {
          // ---- B: utterance-local attention chain
          for (int u = cta; u < B; u += ncta) {
            SUB_START()
            for (int d = threadIdx.x * 4; d < D; d += NT * 4) {
              *reinterpret_cast<float4*>(v_h + d) = ldcg4(g_h + size_t(u) * D + d);
              *reinterpret_cast<float4*>(v_q + d) = ldcg4(g_qkv + size_t(u) * 3 * D + d);
            }
            __syncthreads();
            SUB_DONE(10)
            AttnState st;
            const bf16* kc = cache + size_t(u) * p.L * 2 * D;
            attn_begin(st, v_q, H, p.scale);
            attn_global(st, kc, kc + D, 2 * D, t + 1, H);                  // causal self attention: keys 0..t
            attn_finish(st, H, sm, v_x);
            SUB_DONE(11)
            matvec_cta(v_x, D, w.w_o, w.b_o, D, v_q, pol_w);              // out projection (v_q reused as scratch)
            __syncthreads();
            for (int d = threadIdx.x; d < D; d += NT) {                    // residual (model.py:68); keep h, LN a copy
              const float hv = v_h[d] + v_q[d];
              v_h[d] = hv;
              v_x[d] = hv;
            }
            __syncthreads();
            SUB_DONE(12)
            ln_vec(v_x, D, w.ln2_g, w.ln2_b);                              // LN2 (model.py:70)
            __syncthreads();
            matvec_cta(v_x, D, w.w_qc, w.b_qc, D, v_q, pol_w);            // cross-attention query
            __syncthreads();
            SUB_DONE(13)
            // cross attention over the encoder K/V, streamed through the shared-memory ring (never masked)
            const bf16* ck = ckv + size_t(u) * p.Tp * 2 * D;
            if (!(pre_u == u && pre_l == l) && threadIdx.x == 0)
              ring_issue(sm, ck, p.Tp, H, ck_keys, 0, min(RING, nchunks), pol_kv);
            attn_begin(st, v_q, H, p.scale);
            for (int c = 0; c < nchunks; ++c) {
              const int s = c % RING;
              mbar_wait(&sm.full[s], (ring_parity >> s) & 1u);
              ring_parity ^= 1u << s;
              attn_chunk(st, sm.ring + size_t(s) * CHUNK_BYTES, min(ck_keys, p.Tp - c * ck_keys), H);
              __syncthreads();                                              // stage s drained by every warp
              if (c + RING < nchunks && threadIdx.x == 0)
                ring_issue(sm, ck, p.Tp, H, ck_keys, c + RING, c + RING + 1, pol_kv);
            }
            // request the first stages of the next (utterance, layer) this CTA will serve: they land while the
            // grid-split phases run.  Static data (computed once per utterance), so no ordering hazard.
            {
              int nu = u + ncta, nl = l;
              if (nu >= B) {
                nu = cta;
                nl = l + 1;
                if (nl >= p.nd) nl = (t + 1 < p.L) ? 0 : -1;
              }
              if (nl >= 0) {
                if (threadIdx.x == 0)
                  ring_issue(sm, p.ckv + (size_t(nl) * p.B + b0 + nu) * p.Tp * 2 * D, p.Tp, H, ck_keys, 0, min(RING, nchunks),
                             pol_kv);
                pre_u = nu;
                pre_l = nl;
              } else {
                pre_u = pre_l = -1;
              }
            }
            attn_finish(st, H, sm, v_x);
            SUB_DONE(14)
            matvec_cta(v_x, D, w.w_oc, w.b_oc, D, v_q, pol_w);
            __syncthreads();
            for (int d = threadIdx.x; d < D; d += NT) g_h[size_t(u) * D + d] = v_h[d] + v_q[d];   // residual (model.py:71)
            __syncthreads();
            SUB_DONE(15)
          }
        }
        PHASE_DONE(2 * ph)
        grid_barrier(bar_counter, target, ncta);
        PHASE_DONE(2 * ph + 1)
      }
    }
    // ---- E: classifier (no final LayerNorm, model.py:142) + argmax + EOS + next embedding
    for (int u = cta; u < B; u += ncta) {
      for (int d = threadIdx.x * 4; d < D; d += NT * 4)
        *reinterpret_cast<float4*>(v_h + d) = ldcg4(g_h + size_t(u) * D + d);
      __syncthreads();
      matvec_cta(v_h, D, p.classifier, nullptr, p.V, v_l, pol_w);
      __syncthreads();
      if (g_logits)
        for (int v = threadIdx.x; v < p.V; v += NT) g_logits[(size_t(u) * p.L + t) * p.V + v] = v_l[v];
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int v = lane; v < p.V; v += 32)
          if (v_l[v] > best) {
            best = v_l[v];
            bi = v;
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
          }
        }
        if (bi == 0x7fffffff) bi = 0;
        int tok = bi;
        if (p.stop_at_eos) {
          const int fin = g_fin[u];
          if (fin) tok = p.pad;
          else if (tok == p.eos && lane == 0) {
            g_fin[u] = 1;
            if (g_ntok) g_ntok[u] = t + 2;
            atomicAdd(done_counter, 1u);
          }
        }
        if (lane == 0) {
          g_tokens[size_t(u) * (p.L + 1) + t + 1] = tok;
          sm.stat[32] = __int_as_float(tok);
        }
      }
      __syncthreads();
      if (t + 1 < p.L) {   // embedding + positional encoding of the next input token (model.py:137)
        const int tok = __float_as_int(sm.stat[32]);
        for (int d = threadIdx.x * 4; d < D; d += NT * 4) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + size_t(tok) * D + d));
          const float4 q = __ldg(reinterpret_cast<const float4*>(p.pe + size_t(t + 1) * D + d));
          *reinterpret_cast<float4*>(g_h + size_t(u) * D + d) = make_float4(e.x + q.x, e.y + q.y, e.z + q.z, e.w + q.w);
        }
      }
      __syncthreads();
    }
    PHASE_DONE(8)
    grid_barrier(bar_counter, target, ncta);
    PHASE_DONE(9)
    if (p.stop_at_eos && ld_acquire_u32(done_counter) >= (unsigned)B) break;   // uniform: read after the barrier
  }
  // drain: a requested-but-unconsumed prefetch must land before the CTA (and its shared memory) goes away
  if (pre_u >= 0) {
    for (int c = 0; c < min(RING, nchunks); ++c) mbar_wait(&sm.full[c % RING], (ring_parity >> (c % RING)) & 1u);
  }
  if (p.timing && threadIdx.x == 0)
    for (int i = 0; i < 16; ++i) p.timing[size_t(blockIdx.x) * 16 + i] = tacc[i];
#undef PHASE_DONE
#undef SUB_START
#undef SUB_DONE
}

}  // namespace

size_t persistent_smem_bytes(int D, int FF, int V) {
  int kmax = D > FF ? D : FF;
  if (V > kmax) kmax = (V + 127) / 128 * 128;
  return 128 + size_t(RING) * CHUNK_BYTES + size_t(2) * MROWS * (kmax + PAD) * 2 + NW * 8 * 32 * 16 +
         size_t(4) * kmax * 4 + NW * 64 * 4 + 80 * 4 + RING * 8 + 16;
}

bool persistent_supported(int D, int FF, int V, int H, int nd) {
  return nd <= PERSIST_MAX_LAYERS && (H == 2 || H == 4 || H == 8) && D == 64 * H && D % 128 == 0 && D <= 512 &&
         FF % 64 == 0 && FF <= 2048 && persistent_smem_bytes(D, FF, V) <= 227 * 1024;
}

int launch_dec_persistent(PersistentParams& p, cudaStream_t s) {
  if (p.nd > PERSIST_MAX_LAYERS) return set_error(-2, "persistent decoder: more than %d layers", PERSIST_MAX_LAYERS);
  if (p.H > NW || NW % p.H != 0 || p.H < 2) return set_error(-2, "persistent decoder: num_heads %d must be 2, 4 or 8", p.H);
  if (p.D != 64 * p.H || p.D % 128 != 0 || p.D > 512 || p.FF % 64 != 0 || p.FF > 2048)
    return set_error(-2, "persistent decoder: unsupported D=%d / FF=%d", p.D, p.FF);
  p.kmax = p.D > p.FF ? p.D : p.FF;
  if (p.V > p.kmax) p.kmax = (p.V + 127) / 128 * 128;
  const size_t smem = persistent_smem_bytes(p.D, p.FF, p.V);
  if (smem > 227 * 1024) return set_error(-2, "persistent decoder: needs %zu B of shared memory", smem);
  static size_t configured = 0;
  if (smem > configured) {
    ASR_CUDA_OK(cudaFuncSetAttribute(dec_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int dev = 0, sms = 0, per_sm = 0;
  ASR_CUDA_OK(cudaGetDevice(&dev));
  ASR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ASR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dec_persistent_kernel, NT, smem));
  if (per_sm < 1) return set_error(-2, "persistent decoder: kernel does not fit on an SM");
  if (p.teams < 1) p.teams = 1;
  while (p.teams > 1 && (sms / p.teams) < 8) p.teams >>= 1;
  ASR_CUDA_OK(cudaMemsetAsync(p.barrier, 0, 2 * PERSIST_MAX_TEAMS * 32 * sizeof(unsigned), s));   // barrier + done arrays
  void* args[] = {&p};
  ASR_CUDA_OK(cudaLaunchCooperativeKernel((void*)dec_persistent_kernel, dim3(sms), dim3(NT), args, smem, s));
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
