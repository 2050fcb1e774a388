// Fused multi-head attention core: softmax(mask(Q K^T * scale)) V, flash style, never materialising (B,S,S').
//
// Replaces reference layers.py:20-27 (bmm, masked_fill(-inf), softmax, nan_to_num, bmm) for all heads at once.
// Quirks reproduced: scale is a runtime float (= emb_dim**-0.5, layers.py:20), fully masked rows give zeros
// (nan_to_num, layers.py:25).
//
// Two sm_100a kernels, dh = 64.  attn_ts_kernel (below, persistent, probabilities in tensor memory) takes every launch
// with at most 256 keys and index-only masks (k_lens): the encoder's self attention and the cross attention of the
// teacher-forced decoder.  attn_tc_kernel takes the rest (causal, byte / dense masks, longer sequences):
// one CTA per (q-tile of 128 rows, head, utterance):
//   warp 4 (one elected thread): TMA loads of the Q tile and a 2-deep K/V tile ring (SWIZZLE_128B), and all
//                                tcgen05.mma issue: S = Q K^T (M128 N128 K64) and O_j = P V (M128 N64 K128,
//                                V consumed in place as an MN-major B operand, P from shared memory).
//   warps 0-3 (one query row per thread): tcgen05.ld S from TMEM, scale + mask + online softmax in fp32,
//                                P -> f16 into the swizzled smem A-operand tile, per-tile partial O_j read back
//                                from TMEM and folded into fp32 register accumulators (no TMEM rescale needed).
#include "kernels.h"
#include "ptx.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace asr {
namespace {

constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per tile
constexpr int DH = 64;
constexpr int TILE_BYTES = BKV * DH * 2;  // 16 KB (Q, K and V tiles alike)
constexpr int P_BYTES = BQ * BKV * 2;     // 32 KB
constexpr size_t ATTN_SMEM = TILE_BYTES * 5 + P_BYTES + 128;   // 112.1 KB: two CTAs per SM (228 KB, 1 KB reserved each)
constexpr uint32_t TMEM_COLS = 256;       // S: [0,128), O_j: [128,192)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnDev {
  f16* out; int ldo; long long o_batch_stride; int out_lo_off;
  int H, Sq, Sk;
  float scale_log2;
  int causal;
  const int32_t* k_lens;
  const uint8_t* q_valid;
  const uint8_t* k_valid;
  const uint8_t* dense_mask;
  int mask_B;
  int tma_out;    // the output goes by tensor stores (tmO valid)
};

__global__ void __launch_bounds__(160, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, AttnDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // no static shared memory in this kernel: the base is 1 KB aligned
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();   // swizzled TMA / UMMA tiles need it
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;          // 2 stages
  uint8_t* sV = smem + 3 * TILE_BYTES;      // 2 stages
  uint8_t* sP = smem + 5 * TILE_BYTES;      // two 64-key K blocks of 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;

  int k_end = p.Sk;
  if (p.k_lens) k_end = min(k_end, max(0, p.k_lens[b]));
  if (p.causal) k_end = min(k_end, q0 + BQ);
  const int nkv = (k_end + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0 && nkv > 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * DH, q0, b);
      for (int j = 0; j < 2 && j < nkv; ++j) {
        mbar_expect_tx(&kv_full[j], 2 * TILE_BYTES);
        tma_load_3d(sK + j * TILE_BYTES, &tmK, &kv_full[j], h * DH, j * BKV, b);
        tma_load_3d(sV + j * TILE_BYTES, &tmV, &kv_full[j], h * DH, j * BKV, b);
      }
      constexpr uint32_t idesc_S = umma_idesc_f16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_O = umma_idesc_f16(BQ, DH, 0, 1);   // B = V tile, MN-major
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(sK + s * TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_f16_ss(tmem_S, q_desc + uint64_t(k * 2), k_desc + uint64_t(k * 2), idesc_S, k != 0);
        umma_commit(s_full);
        if (j >= 1 && j + 1 < nkv) {   // refill the stage tile j-1 used, once P V_{j-1} has drained it
          const int sp = (j - 1) & 1;
          mbar_wait(&kv_empty[sp], ((j - 1) >> 1) & 1);
          mbar_expect_tx(&kv_full[sp], 2 * TILE_BYTES);
          tma_load_3d(sK + sp * TILE_BYTES, &tmK, &kv_full[sp], h * DH, (j + 1) * BKV, b);
          tma_load_3d(sV + sp * TILE_BYTES, &tmV, &kv_full[sp], h * DH, (j + 1) * BKV, b);
        }
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(sV + s * TILE_BYTES), 1024, 1024);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t p_desc = umma_smem_desc_sw128(smem_u32(sP + (k >> 2) * (BQ * 128)), 16, 1024) + uint64_t((k & 3) * 2);
          umma_f16_ss(tmem_O, p_desc, v_desc + uint64_t(k * (2048 >> 4)), idesc_O, k != 0);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ---------------- softmax warps: thread <-> query row
    const int r = warp * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    bool row_masked = false;
    if (p.q_valid && qi < p.Sq) row_masked = p.q_valid[size_t(b) * p.Sq + qi] == 0;
    const uint8_t* kvalid = p.k_valid ? p.k_valid + size_t(b) * p.Sk : nullptr;
    const uint8_t* dmask = nullptr;
    if (p.dense_mask && qi < p.Sq)
      dmask = p.dense_mask + (size_t(p.mask_B > 1 ? b : 0) * p.Sq + qi) * p.Sk;
    int k_lim = p.Sk;
    if (p.k_lens) k_lim = min(k_lim, max(0, p.k_lens[b]));
    if (p.causal) k_lim = min(k_lim, qi + 1);
    if (row_masked) k_lim = 0;
    const bool fast = !kvalid && !dmask;

    float m = -INFINITY, l = 0.f;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * BKV;
      // pass 1: row maximum of the masked, scaled scores
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(tmem_S + lane_addr + uint32_t(c * 32), rr);
        tmem_ld_wait();
        if (fast && kbase + c * 32 + 32 <= k_lim) {      // unmasked chunk: no per-element predicates
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(rr[i]));
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2);   // scale > 0
        } else if (fast) {                               // index-only mask (sequence end / causal edge): selects, no loads
          const int nv = k_lim - (kbase + c * 32);       // valid keys of this chunk (may be <= 0)
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], i < nv ? __uint_as_float(rr[i]) : -INFINITY);
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kj = kbase + c * 32 + i;
            bool ok = kj < k_lim;
            if (kvalid && ok) ok = kvalid[kj] != 0;
            if (dmask && ok) ok = dmask[kj] == 0;
            if (ok) mx = fmaxf(mx, __uint_as_float(rr[i]) * p.scale_log2);
          }
        }
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = (m_new == -INFINITY) ? 1.f : fast_exp2(m - m_new);
      // fold in the pending partial O_{j-1} (relative to the old max), then rescale to the new max
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < DH / 32; ++c) {
          uint32_t rr[32];
          tmem_ld32(tmem_O + lane_addr + uint32_t(c * 32), rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(rr[i]);
        }
      }
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] *= alpha;
      // pass 2: probabilities -> f16 P tile (swizzled K-major A operand), row sum in fp32
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(tmem_S + lane_addr + uint32_t(c * 32), rr);
        tmem_ld_wait();
        uint32_t packed[16];
        if (fast && kbase + c * 32 + 32 <= k_lim) {
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(rr[i]), p.scale_log2, -m_new));
            const float p1 = fast_exp2(fmaf(__uint_as_float(rr[i + 1]), p.scale_log2, -m_new));
            l4[(i >> 1) & 3] += p0 + p1;
            packed[i >> 1] = pack_f16x2(p0, p1);
          }
          lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        } else if (fast) {
          const int nv = k_lim - (kbase + c * 32);
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = i < nv ? fast_exp2(fmaf(__uint_as_float(rr[i]), p.scale_log2, -m_new)) : 0.f;
            const float p1 = i + 1 < nv ? fast_exp2(fmaf(__uint_as_float(rr[i + 1]), p.scale_log2, -m_new)) : 0.f;
            l4[(i >> 1) & 3] += p0 + p1;
            packed[i >> 1] = pack_f16x2(p0, p1);
          }
          lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float pv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int kj = kbase + c * 32 + i + u;
              bool ok = kj < k_lim;
              if (kvalid && ok) ok = kvalid[kj] != 0;
              if (dmask && ok) ok = dmask[kj] == 0;
              pv[u] = ok ? fast_exp2(__uint_as_float(rr[i + u]) * p.scale_log2 - m_new) : 0.f;
            }
            lsum += pv[0] + pv[1];
            packed[i >> 1] = pack_f16x2(pv[0], pv[1]);
          }
        }
        uint8_t* blk = sP + (c >> 1) * (BQ * 128) + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (c & 1) * 4 + q4;
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(packed[q4 * 4 + 0], packed[q4 * 4 + 1], packed[q4 * 4 + 2], packed[q4 * 4 + 3]);
        }
      }
      l = l * alpha + lsum;
      m = m_new;
      fence_proxy_async();   // P writes (generic proxy) -> UMMA operand reads (async proxy)
      tc_fence_before();     // order our TMEM loads before the MMAs that overwrite S / O_j
      mbar_arrive(p_ready);
    }
    if (nkv > 0) {
      mbar_wait(o_full, (nkv - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(tmem_O + lane_addr + uint32_t(c * 32), rr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(rr[i]);
      }
    }
    if (p.tma_out) {
      // Output by tensor stores: a thread owns a query row, so direct stores would send 32 separate sectors through the
      // LSU per instruction.  The P tile is free now (the last P V has completed): each warp stages its 32 rows x 64 dims
      // there as f16 hi (and lo) rows of 128 B (128-byte swizzle) and one lane stores the box; rows >= Sq are clipped.
      const float inv = l > 0.f ? 1.f / l : 0.f;   // fully masked row -> zeros (layers.py:25)
      uint4* th = reinterpret_cast<uint4*>(sP + warp * 4096);
      uint4* tl = reinterpret_cast<uint4*>(sP + 16384 + warp * 4096);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d = 8 * j;
        uint4 t;
        t.x = pack_f16x2(o[d + 0] * inv, o[d + 1] * inv);
        t.y = pack_f16x2(o[d + 2] * inv, o[d + 3] * inv);
        t.z = pack_f16x2(o[d + 4] * inv, o[d + 5] * inv);
        t.w = pack_f16x2(o[d + 6] * inv, o[d + 7] * inv);
        th[lane * 8 + (j ^ (lane & 7))] = t;
        if (p.out_lo_off)
          tl[lane * 8 + (j ^ (lane & 7))] =
              make_uint4(f16x2_residual(o[d + 0] * inv, o[d + 1] * inv, t.x), f16x2_residual(o[d + 2] * inv, o[d + 3] * inv, t.y),
                         f16x2_residual(o[d + 4] * inv, o[d + 5] * inv, t.z), f16x2_residual(o[d + 6] * inv, o[d + 7] * inv, t.w));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && q0 + warp * 32 < p.Sq) {
        tma_store_3d(&tmO, th, h * DH, q0 + warp * 32, b);
        if (p.out_lo_off) tma_store_3d(&tmO, tl, p.out_lo_off + h * DH, q0 + warp * 32, b);
        tma_store_commit();
        tma_store_wait_read<0>();                  // the tiles are read before the CTA (and its shared memory) goes away
      }
    } else if (qi < p.Sq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;   // fully masked row -> zeros (layers.py:25)
      f16* op = p.out + size_t(b) * p.o_batch_stride + size_t(qi) * p.ldo + h * DH;
#pragma unroll
      for (int d = 0; d < DH; d += 8) {
        uint4 t;
        t.x = pack_f16x2(o[d + 0] * inv, o[d + 1] * inv);
        t.y = pack_f16x2(o[d + 2] * inv, o[d + 3] * inv);
        t.z = pack_f16x2(o[d + 4] * inv, o[d + 5] * inv);
        t.w = pack_f16x2(o[d + 6] * inv, o[d + 7] * inv);
        *reinterpret_cast<uint4*>(op + d) = t;
        if (p.out_lo_off)
          *reinterpret_cast<uint4*>(op + p.out_lo_off + d) =
              make_uint4(f16x2_residual(o[d + 0] * inv, o[d + 1] * inv, t.x), f16x2_residual(o[d + 2] * inv, o[d + 3] * inv, t.y),
                         f16x2_residual(o[d + 4] * inv, o[d + 5] * inv, t.z), f16x2_residual(o[d + 6] * inv, o[d + 7] * inv, t.w));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ persistent ping-pong kernel (Sk <= 256)
// attn_ts_kernel.  The encoder's self attention (T' = 250 keys at 10 s) and the teacher-forced decoder's cross attention
// see at most 256 keys: a whole score row fits one TMEM accumulator (128 x 256 fp32), so there is no online rescaling and
// K / V are loaded once per (utterance, head).  One persistent CTA per SM (18 warps) walks a contiguous range of
// (utterance, head, q tile) items:
//   warp 16 (one thread): TMA producer.  Q tile per item (2 stages); the K and V tiles (256 rows, zero-filled past Sk,
//                         2 stages each) only when (utterance, head) changes: it runs two items / one head ahead.
//   warp 17 (one thread): tcgen05.mma issue.  S_i = Q K^T (M128 N256 K64) into TMEM buffer i & 1 (256 columns each), then
//                         O_{i-1} = P V (M128 N64, K = the used keys, TS form: P from TMEM) into columns [64,128) of
//                         buffer (i-1) & 1.
//   warps 0-7 / 8-15:     two softmax groups, even / odd items (ping-pong).  Within a group warps [0,4) take score columns
//                         [0,128) of the 128 rows and warps [4,8) columns [128,256) (TMEM lanes are tied to warp id % 4, so
//                         both warps of a row see the same lanes): pass 1 row maximum, pass 2 exp2 -> packed f16
//                         probabilities written by tcgen05.st over score columns the thread has consumed itself (half 0:
//                         columns [0,64), half 1: [128,192)) + fp32 row sum; the halves exchange maxima and sums through
//                         shared memory (256-thread named barriers); then each half reads 32 of the 64 O columns, scales
//                         by 1/l and stages f16 hi | lo rows in the group's tiles, stored by TMA (the store of item i
//                         drains under the softmax of item i+2).
// While group A waits for P V / the next S, group B's exponentials run and vice versa.  The probabilities never touch
// shared memory: earlier versions of this kernel measured, per layer at 256 utterances (legacy kernel: 70 us):
//   57 us  one thread per row (8 softmax warps), 64 KB f16 P tile per group in shared memory, single Q / K / V buffers
//          (every tensor-map load on the critical path);
//   51 us  P in TMEM, double-buffered Q / K / V, private staging;
//   47 us  16 softmax warps (this layout); 45-46 us with the item coordinates of the MMA thread tracked incrementally
//          (its integer divisions sat between a barrier wait and the tcgen05.mma issue) and the P V issue loop unrolled.
// What bounds it is the instruction stream of the softmax warps, not a pipe: ncu shows the XU pipe 50-70 % busy and 0.5
// issue slots per cycle used, but every variant that ADDS instructions to remove XU work is slower - exponentials scaled
// by 2^-112 and packed to f16 by integer ops instead of F2FP (LEA / IMAD + PRMT: +0 ... +2 us), a degree-4 polynomial
// exp2 on the FMA pipe for 4 / 5 / 7 / 8 / 10 of every 16 columns (+2 / +2 / +4 / +4 / +6 us).
// Timeline of a group in steady state (clock64 stamps, -DASR_ATTN_TIMELINE), cycles: pass 1 1.2 k, pass 2 3.2 k, wait for
// P V 1.5 k (the 16 TS-form MMAs of N = 64 take 1.1 k), epilogue 1.0 k, loop 0.4 k: 7.5 k per two items.
constexpr int PP_KEYS = 256;
constexpr int PP_Q_BYTES = BQ * DH * 2;            // 16 KB
constexpr int PP_KV_BYTES = PP_KEYS * DH * 2;      // 32 KB
constexpr int TS_STAGE_BYTES = 2 * BQ * DH * 2;    // 32 KB per group: f16 hi rows | lo rows of one 128-query tile
constexpr size_t TS_SMEM = 2 * PP_Q_BYTES + 4 * PP_KV_BYTES + 2 * TS_STAGE_BYTES + 256 + 2048;   // 226.25 KB (barriers, exchange)

#ifdef ASR_ATTN_TIMELINE   // diagnostic build: clock64 stamps of CTA 0 (softmax warp 0 of group 0, MMA thread), printed by the launcher
__device__ long long g_attn_tl[2 * 32 * 8];
#define TL(role, it, k) do { if (blockIdx.x == 0 && (it) < 32) g_attn_tl[((role) * 32 + (it)) * 8 + (k)] = clock64(); } while (0)
#else
#define TL(role, it, k) do { } while (0)
#endif
// packed fp32 pairs (fma.rn.f32x2 / add.rn.f32x2, sm_100): the softmax warps are bound by their instruction count, and a
// pair op is one issue slot for two score columns
__device__ __forceinline__ uint64_t f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_split(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void group_sync(int g) {   // the 256 softmax threads of group g (named barriers 1 and 2)
  if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}

__global__ void __launch_bounds__(576, 1)
attn_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, AttnDev p, int nq,
                 int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint8_t* sQ = smem;                               // [2] x 16 KB
  uint8_t* sK = sQ + 2 * PP_Q_BYTES;                // [2] x 32 KB
  uint8_t* sV = sK + 2 * PP_KV_BYTES;               // [2] x 32 KB
  uint8_t* sO = sV + 2 * PP_KV_BYTES;               // [2 groups] x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * TS_STAGE_BYTES);
  uint64_t* q_full = bars + 0;     // every barrier below: [2]
  uint64_t* q_free = bars + 2;
  uint64_t* k_full = bars + 4;
  uint64_t* k_free = bars + 6;
  uint64_t* v_full = bars + 8;
  uint64_t* v_free = bars + 10;
  uint64_t* s_full = bars + 12;
  uint64_t* p_ready = bars + 14;
  uint64_t* o_full = bars + 16;
  uint64_t* o_read = bars + 18;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);
  float* xch = reinterpret_cast<float*>(bars + 22);   // [2 groups][2 halves][128 rows] row maxima, then row sums (2 KB)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int f0 = int((long long)blockIdx.x * n_items / gridDim.x);
  const int f1 = int((long long)(blockIdx.x + 1) * n_items / gridDim.x);
  const int n = f1 - f0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_free[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_free[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_free[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_read[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto key_chunks = [&](int b, int& k_lim) {
    k_lim = p.Sk;
    if (p.k_lens) k_lim = min(k_lim, max(0, p.k_lens[b]));
    return max(1, (k_lim + 31) >> 5);
  };

  if (warp == 16) {
    if (lane == 0 && n > 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      int prev_bh = -1, kc = -1;
      for (int it = 0; it < n; ++it) {
        const int f = f0 + it;
        const int bh = f / nq, qt = f - bh * nq;
        const int b = bh / p.H, h = bh - b * p.H;
        const int qs = it & 1;
        if (it >= 2) mbar_wait(&q_free[qs], ((it >> 1) - 1) & 1);
        mbar_expect_tx(&q_full[qs], PP_Q_BYTES);
        tma_load_3d(sQ + qs * PP_Q_BYTES, &tmQ, &q_full[qs], h * DH, qt * BQ, b);
        if (bh != prev_bh) {
          prev_bh = bh;
          ++kc;
          const int st = kc & 1;
          if (kc >= 2) mbar_wait(&k_free[st], ((kc >> 1) - 1) & 1);
          mbar_expect_tx(&k_full[st], PP_KV_BYTES);
          tma_load_3d(sK + st * PP_KV_BYTES, &tmK, &k_full[st], h * DH, 0, b);
          if (kc >= 2) mbar_wait(&v_free[st], ((kc >> 1) - 1) & 1);
          mbar_expect_tx(&v_full[st], PP_KV_BYTES);
          tma_load_3d(sV + st * PP_KV_BYTES, &tmV, &v_full[st], h * DH, 0, b);
        }
      }
    }
  } else if (warp == 17) {
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_S = umma_idesc_f16(BQ, PP_KEYS, 0, 0);
      constexpr uint32_t idesc_O = umma_idesc_f16(BQ, DH, 0, 1);   // B = V tile, MN-major
      // This thread shares its scheduler with four softmax warps, so every instruction it executes between a barrier
      // wait and a tcgen05.mma costs tens of cycles of the chain softmax -> P V -> epilogue -> next S: the item
      // coordinates are tracked incrementally (no divisions) and the P V issue loop is fully unrolled.
      int s_kc = -1, pv_kc = -1;
      const int bh0 = f0 / nq;
      int s_qt = f0 - bh0 * nq, s_new = 1;                    // S side: q tile of item it, item starts a new (b, h)
      int pv_qt = s_qt, pv_new = 1, pv_b = bh0 / p.H, pv_h = bh0 - pv_b * p.H;   // P V side, one item behind
      auto issue_pv = [&](int j) {
        const int s = j & 1;
        int k_lim;
        const int nk = 2 * key_chunks(pv_b, k_lim);          // 16-key MMA steps
        if (pv_new) {
          ++pv_kc;
          mbar_wait(&v_full[pv_kc & 1], (pv_kc >> 1) & 1);
        }
        const bool last_of_head = j + 1 == n || pv_qt + 1 == nq;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(sV + (pv_kc & 1) * PP_KV_BYTES), 1024, 1024);
        const uint32_t tbuf = tmem_base + uint32_t(s * 256);
        TL(1, j, 3);
        mbar_wait(&p_ready[s], (j >> 1) & 1);
        TL(1, j, 4);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 16; ++k)   // P of keys [0,128) in columns [0,64), of keys [128,256) in columns [128,192)
          if (k < nk)
            umma_f16_ts(tbuf + 64u, tbuf + uint32_t(k < 8 ? k * 8 : 64 + k * 8), v_desc + uint64_t(k * (2048 >> 4)), idesc_O,
                        k != 0);
        umma_commit(&o_full[s]);
        TL(1, j, 5);
        if (last_of_head) umma_commit(&v_free[pv_kc & 1]);
        pv_new = 0;
        if (++pv_qt == nq) {
          pv_qt = 0;
          pv_new = 1;
          if (++pv_h == p.H) { pv_h = 0; ++pv_b; }
        }
      };
      for (int it = 0; it < n; ++it) {
        const int s = it & 1;
        TL(1, it, 0);
        if (it >= 2) mbar_wait(&o_read[s], ((it >> 1) - 1) & 1);   // P and O of item it-2 have left this TMEM buffer
        TL(1, it, 1);
        mbar_wait(&q_full[s], (it >> 1) & 1);
        if (s_new) {
          ++s_kc;
          mbar_wait(&k_full[s_kc & 1], (s_kc >> 1) & 1);
        }
        tc_fence_after();
        const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(sQ + s * PP_Q_BYTES), 16, 1024);
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(sK + (s_kc & 1) * PP_KV_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_f16_ss(tmem_base + s * 256, q_desc + uint64_t(k * 2), k_desc + uint64_t(k * 2), idesc_S, k != 0);
        umma_commit(&s_full[s]);
        umma_commit(&q_free[s]);
        TL(1, it, 2);
        if (it + 1 == n || s_qt + 1 == nq) umma_commit(&k_free[s_kc & 1]);
        s_new = 0;
        if (++s_qt == nq) { s_qt = 0; s_new = 1; }
        if (it >= 1) issue_pv(it - 1);
      }
      issue_pv(n - 1);
    }
  } else {
    // ---------------- softmax: group g (items it = g mod 2), column half hf, rows 32 * w4 + lane
    const int g = warp >> 3, hf = (warp >> 2) & 1, w4 = warp & 3;
    const int r = w4 * 32 + lane;
    const uint32_t tB = tmem_base + uint32_t(g * 256) + (uint32_t(w4 * 32) << 16);   // the group's TMEM buffer, own lanes
    const uint32_t tS = tB + uint32_t(hf * 128);                                     // own score columns
    float* xmine = xch + (g * 2 + hf) * 128 + r;
    const float* xother = xch + (g * 2 + (hf ^ 1)) * 128 + r;
    uint4* th = reinterpret_cast<uint4*>(sO + g * TS_STAGE_BYTES + w4 * 4096);
    uint4* tl = reinterpret_cast<uint4*>(sO + g * TS_STAGE_BYTES + 16384 + w4 * 4096);
    int qt, h, b;                                       // coordinates of item it, advanced by two items per iteration
    {
      const int f = f0 + g;
      const int bh = f / nq;
      qt = f - bh * nq;
      b = bh / p.H;
      h = bh - b * p.H;
    }
    auto advance = [&]() {
      for (int u = 0; u < 2; ++u)
        if (++qt == nq) {
          qt = 0;
          if (++h == p.H) { h = 0; ++b; }
        }
    };
    for (int it = g; it < n; it += 2, advance()) {
      int k_lim;
      const int nch = min(4, key_chunks(b, k_lim) - 4 * hf);   // own 32-key chunks in use (<= 0: none)
      const int k_off = hf * 128;
      const uint32_t par = (it >> 1) & 1;
      const bool tl_on = warp == 0 && lane == 0;
      if (tl_on) TL(0, it, 0);
      mbar_wait(&s_full[g], par);
      if (tl_on) TL(0, it, 1);
      tc_fence_after();
      // pass 1: row maximum over the own columns
      float mx = -INFINITY;
      if (nch > 0) {
        uint32_t ra[32];
        auto red = [&](const uint32_t (&rr)[32], int c) {
          const int nv = k_lim - k_off - c * 32;
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (nv >= 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(rr[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], i < nv ? __uint_as_float(rr[i]) : -INFINITY);
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        };
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {   // four warps per scheduler hide the TMEM latency: no register double buffer
          tmem_ld32(tS + uint32_t(c * 32), ra);
          tmem_ld_wait();
          red(ra, c);
        }
      }
      if (hf == 0 && lane == 0) tma_store_wait_read<0>();   // the boxes of item it-2 have left the staging tiles
      *xmine = mx;
      if (tl_on) TL(0, it, 2);
      group_sync(g);
      if (tl_on) TL(0, it, 3);
      mx = fmaxf(mx, *xother);
      const float m_s = mx * p.scale_log2;                  // scale > 0; -inf when no key is valid
      const float m_use = (m_s == -INFINITY) ? 0.f : m_s;
      float lsum = 0.f;
      const uint64_t sc2 = f32x2(p.scale_log2, p.scale_log2), nm2 = f32x2(-m_use, -m_use);
      // The chunk that straddles Sk (250 keys: 26 of 32 columns) would take the predicated path below, ~2.6x the
      // instructions of a full chunk, and the other half of the row waits for it at the exchange.  When nothing but the
      // end of the sequence masks it, the K and V rows past Sk are the tensor map's zero fill: those scores are exactly
      // 0, their probabilities meet zero V rows, and their known contribution 2^-m each is taken out of the row sum.
      const bool tail_zero = k_lim == p.Sk && m_use > -64.f;
      if (nch > 0) {
        uint32_t ra[32];
        auto emit = [&](const uint32_t (&rr)[32], int c) {
          const int nv = k_lim - k_off - c * 32;
          uint32_t packed[16];
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
          if (nv >= 32 || tail_zero) {
            uint64_t a2[2] = {0ull, 0ull};                   // two pairs of partial sums
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float x0, x1;
              f32x2_split(fma_f32x2(f32x2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), sc2, nm2), x0, x1);
              const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
              a2[(i >> 1) & 1] = add_f32x2(a2[(i >> 1) & 1], f32x2(p0, p1));
              packed[i >> 1] = pack_f16x2(p0, p1);
            }
            f32x2_split(add_f32x2(a2[0], a2[1]), l4[0], l4[1]);
            if (nv < 32) l4[2] = -float(32 - nv) * fast_exp2(-m_use);   // the columns past Sk: score 0, V row 0
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = i < nv ? fast_exp2(fmaf(__uint_as_float(rr[i]), p.scale_log2, -m_use)) : 0.f;
              const float p1 = i + 1 < nv ? fast_exp2(fmaf(__uint_as_float(rr[i + 1]), p.scale_log2, -m_use)) : 0.f;
              l4[(i >> 1) & 3] += p0 + p1;
              packed[i >> 1] = pack_f16x2(p0, p1);
            }
          }
          lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
          tmem_st16(tS + uint32_t(c * 16), packed);          // over own, consumed score columns
        };
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          tmem_ld32(tS + uint32_t(c * 32), ra);
          tmem_ld_wait();
          emit(ra, c);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[g]);
      if (tl_on) TL(0, it, 4);
      group_sync(g);             // everyone has read the row maxima: the exchange slots take the row sums
      *xmine = lsum;

      mbar_wait(&o_full[g], par);
      if (tl_on) TL(0, it, 5);
      tc_fence_after();
      uint32_t ro[32];
      tmem_ld32(tB + 64u + uint32_t(hf * 32), ro);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_read[g]);   // the TMEM buffer may take S of item it+2
      if (tl_on) TL(0, it, 6);
      group_sync(g);             // row sums written
      lsum += *xother;
      const float inv = lsum > 0.f ? 1.f / lsum : 0.f;   // fully masked row -> zeros (layers.py:25)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = 8 * j;
        const float o0 = __uint_as_float(ro[d + 0]) * inv, o1 = __uint_as_float(ro[d + 1]) * inv;
        const float o2 = __uint_as_float(ro[d + 2]) * inv, o3 = __uint_as_float(ro[d + 3]) * inv;
        const float o4 = __uint_as_float(ro[d + 4]) * inv, o5 = __uint_as_float(ro[d + 5]) * inv;
        const float o6 = __uint_as_float(ro[d + 6]) * inv, o7 = __uint_as_float(ro[d + 7]) * inv;
        uint4 t;
        t.x = pack_f16x2(o0, o1);
        t.y = pack_f16x2(o2, o3);
        t.z = pack_f16x2(o4, o5);
        t.w = pack_f16x2(o6, o7);
        const int ch = (hf * 4 + j) ^ (lane & 7);
        th[lane * 8 + ch] = t;
        if (p.out_lo_off)
          tl[lane * 8 + ch] = make_uint4(f16x2_residual(o0, o1, t.x), f16x2_residual(o2, o3, t.y),
                                         f16x2_residual(o4, o5, t.z), f16x2_residual(o6, o7, t.w));
      }
      fence_proxy_async();
      group_sync(g);             // both halves of every row are staged
      if (hf == 0 && lane == 0) {
        const int row0 = qt * BQ + w4 * 32;
        if (row0 < p.Sq) {
          tma_store_3d(&tmO, th, h * DH, row0, b);
          if (p.out_lo_off) tma_store_3d(&tmO, tl, p.out_lo_off + h * DH, row0, b);
          tma_store_commit();
        }
      }
      if (tl_on) TL(0, it, 7);
    }
    if (hf == 0 && lane == 0) tma_store_wait_read<0>();   // the staging tiles are read before the CTA goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ naive cross-check (CUDA cores, one warp per row)
__global__ void attn_naive_kernel(AttnParams p) {
  const int qi = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x;
  const f16* q = p.q + size_t(b) * p.q_batch_stride + size_t(qi) * p.ldq + h * DH;
  const float q0 = __half2float(q[lane * 2]), q1 = __half2float(q[lane * 2 + 1]);
  bool row_masked = p.q_valid && p.q_valid[size_t(b) * p.Sq + qi] == 0;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int kj = 0; kj < p.Sk; ++kj) {
    bool ok = !row_masked;
    if (p.k_lens && kj >= p.k_lens[b]) ok = false;
    if (p.causal && kj > qi) ok = false;
    if (p.k_valid && p.k_valid[size_t(b) * p.Sk + kj] == 0) ok = false;
    if (p.dense_mask && p.dense_mask[(size_t(p.mask_B > 1 ? b : 0) * p.Sq + qi) * p.Sk + kj] != 0) ok = false;
    if (!ok) continue;
    const f16* k = p.k + size_t(b) * p.k_batch_stride + size_t(kj) * p.ldk + h * DH;
    const f16* v = p.v + size_t(b) * p.v_batch_stride + size_t(kj) * p.ldv + h * DH;
    float s = q0 * __half2float(k[lane * 2]) + q1 * __half2float(k[lane * 2 + 1]);
    s = warp_sum(s) * p.scale;
    const float m_new = fmaxf(m, s);
    const float a = expf(m - m_new), pw = expf(s - m_new);
    l = l * a + pw;
    o0 = o0 * a + pw * __half2float(v[lane * 2]);
    o1 = o1 * a + pw * __half2float(v[lane * 2 + 1]);
    m = m_new;
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  f16* op = p.out + size_t(b) * p.o_batch_stride + size_t(qi) * p.ldo + h * DH;
  const f16 h0 = f16_sat(o0 * inv), h1 = f16_sat(o1 * inv);
  op[lane * 2] = h0;
  op[lane * 2 + 1] = h1;
  if (p.out_lo_off) {
    op[p.out_lo_off + lane * 2] = f16_sat(o0 * inv - __half2float(h0));
    op[p.out_lo_off + lane * 2 + 1] = f16_sat(o1 * inv - __half2float(h1));
  }
}

int check_params(const AttnParams& p) {
  if (p.B <= 0 || p.H <= 0 || p.Sq <= 0 || p.Sk <= 0) return set_error(-2, "attention: empty problem");
  if (!p.q || !p.k || !p.v || !p.out) return set_error(-2, "attention: null pointer");
  return 0;
}

}  // namespace

int launch_attention_tc(const AttnParams& p, cudaStream_t s) {
  if (int rc = check_params(p)) return rc;
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[3] = {DH, BQ, 1};
  {
    uint64_t dims[3] = {(uint64_t)p.H * DH, (uint64_t)p.Sq, (uint64_t)p.B};
    uint64_t str[3] = {2, (uint64_t)p.ldq * 2, (uint64_t)p.q_batch_stride * 2};
    if (int rc = make_tmap_f16(&tmQ, p.q, 3, dims, str, box, nullptr)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.H * DH, (uint64_t)p.Sk, (uint64_t)p.B};
    uint64_t str[3] = {2, (uint64_t)p.ldk * 2, (uint64_t)p.k_batch_stride * 2};
    if (int rc = make_tmap_f16(&tmK, p.k, 3, dims, str, box, nullptr)) return rc;
    str[1] = (uint64_t)p.ldv * 2;
    str[2] = (uint64_t)p.v_batch_stride * 2;
    if (int rc = make_tmap_f16(&tmV, p.v, 3, dims, str, box, nullptr)) return rc;
  }
  AttnDev d;
  d.out = p.out; d.ldo = p.ldo; d.o_batch_stride = p.o_batch_stride; d.out_lo_off = p.out_lo_off;
  d.H = p.H; d.Sq = p.Sq; d.Sk = p.Sk;
  d.scale_log2 = p.scale * 1.4426950408889634f;
  d.causal = p.causal; d.k_lens = p.k_lens; d.q_valid = p.q_valid; d.k_valid = p.k_valid;
  d.dense_mask = p.dense_mask; d.mask_B = p.mask_B;
  CUtensorMap tmO = tmQ;
  d.tma_out = 0;
  {
    static const bool tma_off = [] {
      const char* e = std::getenv("ASR_B200_LN_TMA");
      return e && e[0] == '0';
    }();
    const uint64_t cols = uint64_t(p.H) * DH + uint64_t(p.out_lo_off);
    if (!tma_off && cols <= uint64_t(p.ldo) && p.out_lo_off % 8 == 0) {
      uint64_t dims[3] = {cols, (uint64_t)p.Sq, (uint64_t)p.B};
      uint64_t str[3] = {2, (uint64_t)p.ldo * 2, (uint64_t)p.o_batch_stride * 2};
      const uint32_t obox[3] = {DH, 32, 1};
      d.tma_out = make_tmap_f16(&tmO, p.out, 3, dims, str, obox, nullptr) == 0;
    }
  }
  static const bool pp_off = [] {
    const char* e = std::getenv("ASR_B200_ATTN_PP");
    return e && e[0] == '0';
  }();
  if (!pp_off && d.tma_out && p.Sk <= PP_KEYS && !p.causal && !p.q_valid && !p.k_valid && !p.dense_mask) {
    // every key of an utterance fits one TMEM accumulator: persistent ping-pong kernel
    const uint32_t kvbox[3] = {DH, PP_KEYS, 1};
    uint64_t dims[3] = {(uint64_t)p.H * DH, (uint64_t)p.Sk, (uint64_t)p.B};
    uint64_t str[3] = {2, (uint64_t)p.ldk * 2, (uint64_t)p.k_batch_stride * 2};
    if (int rc = make_tmap_f16(&tmK, p.k, 3, dims, str, kvbox, nullptr)) return rc;
    str[1] = (uint64_t)p.ldv * 2;
    str[2] = (uint64_t)p.v_batch_stride * 2;
    if (int rc = make_tmap_f16(&tmV, p.v, 3, dims, str, kvbox, nullptr)) return rc;
    int n_sm = 0;
    if (int rc = device_props(&n_sm, nullptr)) return rc;
    const int nq = (p.Sq + BQ - 1) / BQ;
    const long long items = (long long)p.B * p.H * nq;
    if (items < (1ll << 30)) {
      const int grid = (int)std::min<long long>(n_sm, items);
      if (int rc = ensure_dyn_smem((const void*)attn_ts_kernel, TS_SMEM)) return rc;
      attn_ts_kernel<<<grid, 576, TS_SMEM, s>>>(tmQ, tmK, tmV, tmO, d, nq, (int)items);
#ifdef ASR_ATTN_TIMELINE
        {
          static long long tl[2 * 32 * 8];
          cudaStreamSynchronize(s);
          cudaMemcpyFromSymbol(tl, g_attn_tl, sizeof(tl));
          const long long t0 = tl[32 * 8];   // MMA thread, item 0, stamp 0
          for (int it = 0; it < 16; ++it) {
            fprintf(stderr, "it %2d  mma:", it);
            for (int k = 0; k < 6; ++k) fprintf(stderr, " %7lld", tl[(32 + it) * 8 + k] - t0);
            if (it % 2 == 0) {
              fprintf(stderr, "   wg0:");
              for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", tl[it * 8 + k] - t0);
            }
            fprintf(stderr, "\n");
          }
        }
#endif
      ASR_CUDA_OK(cudaGetLastError());
      ASR_LAUNCHED(1);
      return 0;
    }
  }
  if (int rc = ensure_dyn_smem((const void*)attn_tc_kernel, ATTN_SMEM)) return rc;
  dim3 grid((p.Sq + BQ - 1) / BQ, p.H, p.B);
  attn_tc_kernel<<<grid, 160, ATTN_SMEM, s>>>(tmQ, tmK, tmV, tmO, d);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

int launch_attention_naive(const AttnParams& p, cudaStream_t s) {
  if (int rc = check_params(p)) return rc;
  dim3 grid(p.Sq, p.H, p.B);
  attn_naive_kernel<<<grid, 32, 0, s>>>(p);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
