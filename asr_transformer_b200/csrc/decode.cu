// Greedy decode step kernels (reference model.py:125-151, restated with a device-resident KV cache).
//
// Precision rule: the decode-step Linear layers must see fp32-accurate activations or the
// autoregressive token stream drifts from the fp32 reference.  Activations are therefore split into f16
// hi + lo parts (x = hi + lo up to 2^-17 relative) and each weight tile is multiplied by both on the tensor
// cores (mma.sync m16n8k16, fp32 accumulate); weights are stored as f16.  These kernels are bound by HBM/L2
// traffic (weights + K/V caches), not by the tensor pipe.
#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------- decode linear
// out[B, N] = epilogue( prologue(x)[B, K] * W[N, K]^T )
// grid (N_pad/32, ceil(B/32)), 4 warps; warp w owns output columns [n0 + 8w, n0 + 8w + 8) for 32 batch rows.
constexpr int DL_ROWS = 32;
constexpr int DL_KC = 256;                 // K chunk staged in smem
constexpr int DL_LD = DL_KC + 32;          // row stride (elements): 576 B == 64 mod 128 -> conflict-free LDS.128

__global__ void __launch_bounds__(128)
dec_linear_kernel(DecLinear p) {
  __shared__ __align__(16) f16 s_hi[DL_ROWS * DL_LD];
  __shared__ __align__(16) f16 s_lo[DL_ROWS * DL_LD];
  __shared__ float s_mean[DL_ROWS], s_rstd[DL_ROWS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const int m0 = blockIdx.y * DL_ROWS;
  const int n_base = blockIdx.x * 32 + warp * 8;
  const bool ln = p.ln_gamma != nullptr;

  if (ln) {   // LayerNorm statistics over the full row (K == D), one warp per row
    for (int r = warp; r < DL_ROWS; r += 4) {
      const int row = m0 + r;
      float mean = 0.f, rstd = 0.f;
      if (row < p.B) {
        const float* xr = p.x + size_t(row) * p.ldx;
        float sum = 0.f;
        for (int k = lane * 4; k < p.K; k += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + k);
          sum += v.x + v.y + v.z + v.w;
        }
        mean = warp_sum(sum) / float(p.K);
        float sq = 0.f;
        for (int k = lane * 4; k < p.K; k += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + k);
          const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
          sq += a * a + b * b + cc * cc + d * d;
        }
        rstd = 1.0f / sqrtf(warp_sum(sq) / float(p.K) + 1e-5f);
      }
      if (lane == 0) {
        s_mean[r] = mean;
        s_rstd[r] = rstd;
      }
    }
    __syncthreads();
  }

  float acc[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;

  const f16* wrow = p.w + size_t(n_base + g) * p.K + c * 8;

  for (int kc = 0; kc < p.K; kc += DL_KC) {
    const int kw = min(DL_KC, p.K - kc);   // multiple of 32
    if (kc > 0) __syncthreads();
    // stage activations: fp32 -> (LayerNorm) -> f16 hi / lo
    for (int idx = threadIdx.x; idx < DL_ROWS * (DL_KC / 4); idx += blockDim.x) {
      const int r = idx / (DL_KC / 4);
      const int k = (idx % (DL_KC / 4)) * 4;
      const int row = m0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.B && k < kw) {
        v = *reinterpret_cast<const float4*>(p.x + size_t(row) * p.ldx + kc + k);
        if (ln) {
          const float4 gm = *reinterpret_cast<const float4*>(p.ln_gamma + kc + k);
          const float4 bt = *reinterpret_cast<const float4*>(p.ln_beta + kc + k);
          const float mean = s_mean[r], rstd = s_rstd[r];
          v.x = (v.x - mean) * rstd * gm.x + bt.x;
          v.y = (v.y - mean) * rstd * gm.y + bt.y;
          v.z = (v.z - mean) * rstd * gm.z + bt.z;
          v.w = (v.w - mean) * rstd * gm.w + bt.w;
        }
      }
      const f16 h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z),
                 h3 = __float2half_rn(v.w);
      uint2 hi, lo;
      hi.x = pack_f16x2(__half2float(h0), __half2float(h1));
      hi.y = pack_f16x2(__half2float(h2), __half2float(h3));
      lo.x = pack_f16x2(v.x - __half2float(h0), v.y - __half2float(h1));
      lo.y = pack_f16x2(v.z - __half2float(h2), v.w - __half2float(h3));
      *reinterpret_cast<uint2*>(s_hi + r * DL_LD + k) = hi;
      *reinterpret_cast<uint2*>(s_lo + r * DL_LD + k) = lo;
    }
    __syncthreads();
    // Each 32-wide k block: thread (g,c) holds W[n_base+g][k0 + 8c .. 8c+7] (one 16-byte load) and uses it for two
    // k16 MMA steps; the A fragments use the same K permutation, so the dot products are unchanged.
    const int nkb = kw / 32;
    uint4 wv[DL_KC / 32];
#pragma unroll
    for (int kb = 0; kb < DL_KC / 32; ++kb)
      if (kb < nkb) wv[kb] = __ldg(reinterpret_cast<const uint4*>(wrow + kc + kb * 32));
#pragma unroll
    for (int kb = 0; kb < DL_KC / 32; ++kb) {
      if (kb < nkb) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r0 = mt * 16 + g;
          const uint4 h_a = *reinterpret_cast<const uint4*>(s_hi + r0 * DL_LD + kb * 32 + c * 8);
          const uint4 h_b = *reinterpret_cast<const uint4*>(s_hi + (r0 + 8) * DL_LD + kb * 32 + c * 8);
          const uint4 l_a = *reinterpret_cast<const uint4*>(s_lo + r0 * DL_LD + kb * 32 + c * 8);
          const uint4 l_b = *reinterpret_cast<const uint4*>(s_lo + (r0 + 8) * DL_LD + kb * 32 + c * 8);
          mma_f16_16816(acc[mt], h_a.x, h_b.x, h_a.y, h_b.y, wv[kb].x, wv[kb].y);
          mma_f16_16816(acc[mt], h_a.z, h_b.z, h_a.w, h_b.w, wv[kb].z, wv[kb].w);
          mma_f16_16816(acc[mt], l_a.x, l_b.x, l_a.y, l_b.y, wv[kb].x, wv[kb].y);
          mma_f16_16816(acc[mt], l_a.z, l_b.z, l_a.w, l_b.w, wv[kb].z, wv[kb].w);
        }
      }
    }
  }

  // epilogue: thread holds rows (g, g+8) of each m16 tile, columns n_base + 2c, 2c+1
  const int col = n_base + 2 * c;
  if (col >= p.N) return;
  const bool two = col + 1 < p.N;
  const float b0 = p.bias ? __ldg(p.bias + col) : 0.f;
  const float b1 = (p.bias && two) ? __ldg(p.bias + col + 1) : 0.f;
  const int step = p.step ? *p.step : 0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = m0 + mt * 16 + g + hh * 8;
      if (row >= p.B) continue;
      float v0 = acc[mt][hh * 2 + 0] + b0;
      float v1 = acc[mt][hh * 2 + 1] + b1;
      if (p.relu) {
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      if (p.residual) {
        v0 += p.residual[size_t(row) * p.ld_res + col];
        if (two) v1 += p.residual[size_t(row) * p.ld_res + col + 1];
      }
      float* op = p.out + size_t(row) * p.ldo + col;
      op[0] = v0;
      if (two) op[1] = v1;
      if (p.kv_cache && col >= p.kv_col0) {
        const int w = p.N - p.kv_col0;
        f16* kp = p.kv_cache + (size_t(row) * p.kv_rows + step) * w + (col - p.kv_col0);
        kp[0] = __float2half_rn(v0);
        if (two) kp[1] = __float2half_rn(v1);
      }
    }
  }
}

// ---------------------------------------------------------------- single-query attention over a f16 K/V cache
// One CTA per (head, utterance); 8 lanes per key row (8 x 16 B = the 128-byte head slice), fp32 math.
__global__ void __launch_bounds__(128)
dec_attn_kernel(DecAttn p) {
  extern __shared__ float s_sc[];          // [n_keys] scores / probabilities
  __shared__ float s_red[16 * 64];
  __shared__ float s_stat[8];
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c8 = tid & 7, grp = tid >> 3;  // 16 key groups
  const int n = p.step ? (*p.step + 1) : p.n_keys;

  float q[8];
  {
    const float* qp = p.q + size_t(b) * p.ldq + h * 64 + c8 * 8;
    const float4 a = *reinterpret_cast<const float4*>(qp);
    const float4 bq = *reinterpret_cast<const float4*>(qp + 4);
    q[0] = a.x * p.scale; q[1] = a.y * p.scale; q[2] = a.z * p.scale; q[3] = a.w * p.scale;
    q[4] = bq.x * p.scale; q[5] = bq.y * p.scale; q[6] = bq.z * p.scale; q[7] = bq.w * p.scale;
  }
  const f16* kb = p.k + size_t(b) * p.kv_batch_stride + h * 64 + c8 * 8;
  const f16* vb = p.v + size_t(b) * p.kv_batch_stride + h * 64 + c8 * 8;

  const unsigned gmask = 0xFFu << (lane & 24);
  float mx = -INFINITY;
  for (int kj = grp; kj < n; kj += 16) {
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(kb + size_t(kj) * p.ldkv));
    const __half2* k2 = reinterpret_cast<const __half2*>(&kv);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(k2[i]);
      s = fmaf(q[2 * i], f.x, s);
      s = fmaf(q[2 * i + 1], f.y, s);
    }
    s += __shfl_xor_sync(gmask, s, 1);   // reduce inside the 8-lane key group only: trip counts differ
    s += __shfl_xor_sync(gmask, s, 2);   // between the four groups of a warp when n % 16 != 0
    s += __shfl_xor_sync(gmask, s, 4);
    if (c8 == 0) s_sc[kj] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) s_stat[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
  float sum = 0.f;
  for (int kj = tid; kj < n; kj += 128) {
    const float e = __expf(s_sc[kj] - mx);
    s_sc[kj] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_stat[4 + warp] = sum;
  __syncthreads();
  const float l = s_stat[4] + s_stat[5] + s_stat[6] + s_stat[7];

  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  for (int kj = grp; kj < n; kj += 16) {
    const float pw = s_sc[kj];
    const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vb + size_t(kj) * p.ldkv));
    const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(v2[i]);
      o[2 * i] = fmaf(pw, f.x, o[2 * i]);
      o[2 * i + 1] = fmaf(pw, f.y, o[2 * i + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_red[grp * 64 + c8 * 8 + i] = o[i];
  __syncthreads();
  if (tid < 64) {
    float t = 0.f;
#pragma unroll
    for (int gI = 0; gI < 16; ++gI) t += s_red[gI * 64 + tid];
    p.out[size_t(b) * p.ldo + h * 64 + tid] = l > 0.f ? t / l : 0.f;
  }
}

// ---------------------------------------------------------------- argmax + EOS bookkeeping + next embedding
// Single CTA (so the step counter can be advanced race-free); one warp per utterance row, lowest index wins ties
// (torch.argmax semantics, reference model.py:143).
__global__ void __launch_bounds__(1024)
dec_select_kernel(DecSelect p, const float* emb, const float* pe, int D, float* h_next) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int step = *p.step;
  for (int b = warp; b < p.B; b += nwarps) {
    const float* lg = p.logits + size_t(b) * p.ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < p.V; v += 32) {
      const float x = lg[v];
      if (p.step_logits) p.step_logits[(size_t(b) * p.L + step) * p.V + v] = x;
      if (x > best) {   // strict: within a lane the lowest index is kept
        best = x;
        bi = v;
      }
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
      const int fin = p.finished[b];
      if (fin) tok = p.pad;
      else if (tok == p.eos && lane == 0) {
        p.finished[b] = 1;
        if (p.n_tokens) p.n_tokens[b] = step + 2;
      }
    }
    if (lane == 0) p.tokens[size_t(b) * p.ld_tok + step + 1] = tok;
    if (step + 1 < p.L && h_next) {   // embedding + PE of the next input token (reference model.py:137)
      const float* e = emb + size_t(tok) * D;
      const float* pr = pe + size_t(step + 1) * D;
      for (int d = lane * 4; d < D; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(e + d);
        const float4 c = *reinterpret_cast<const float4*>(pr + d);
        *reinterpret_cast<float4*>(h_next + size_t(b) * D + d) = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *p.step = step + 1;
}

__global__ void dec_embed_kernel(const int32_t* tokens, int ld_tok, const int32_t* step, const float* emb,
                                 const float* pe, int D, int vocab, float* h) {
  const int b = blockIdx.x;
  const int t = step ? *step : 0;
  int tok = tokens[size_t(b) * ld_tok + t];
  tok = min(max(tok, 0), vocab - 1);
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(emb + size_t(tok) * D + d);
    const float4 c = *reinterpret_cast<const float4*>(pe + size_t(t) * D + d);
    *reinterpret_cast<float4*>(h + size_t(b) * D + d) = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

}  // namespace

int launch_dec_linear(const DecLinear& p, cudaStream_t s) {
  if (p.B <= 0) return 0;
  if (p.K % 32 != 0 || p.K <= 0) return set_error(-2, "dec_linear: K=%d must be a multiple of 32", p.K);
  if (p.ln_gamma && p.K % 128 != 0) return set_error(-2, "dec_linear: LayerNorm prologue needs K %% 128 == 0");
  if (p.ldx % 4 != 0) return set_error(-2, "dec_linear: ldx must be a multiple of 4");
  const int n_pad = (p.N + 31) / 32 * 32;   // weight buffer must hold n_pad rows
  dim3 grid(n_pad / 32, (p.B + DL_ROWS - 1) / DL_ROWS);
  dec_linear_kernel<<<grid, 128, 0, s>>>(p);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_dec_attention(const DecAttn& p, cudaStream_t s) {
  if (p.B <= 0) return 0;
  const int max_keys = p.step ? p.n_keys : p.n_keys;   // n_keys = capacity when step-driven
  const size_t smem = size_t(max_keys) * sizeof(float);
  if (smem > 160 * 1024) return set_error(-2, "dec_attention: %d keys exceed the shared-memory score buffer", max_keys);
  if (int rc = ensure_dyn_smem((const void*)dec_attn_kernel, smem)) return rc;
  dec_attn_kernel<<<dim3(p.H, p.B), 128, smem, s>>>(p);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_dec_embed(const int32_t* tokens, int ld_tok, const int32_t* step, const float* emb, const float* pe, int B,
                     int D, int vocab, float* h, cudaStream_t s) {
  if (B <= 0) return 0;
  dec_embed_kernel<<<B, 64, 0, s>>>(tokens, ld_tok, step, emb, pe, D, vocab, h);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_dec_select_embed(const DecSelect& p, const float* emb, const float* pe, int D, float* h_next,
                            cudaStream_t s) {
  if (p.B <= 0) return 0;
  dec_select_kernel<<<1, 1024, 0, s>>>(p, emb, pe, D, h_next);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
