// Cluster greedy decoder: one thread-block CLUSTER owns a small group of utterances for the whole decode (all L steps,
// all layers); the CTAs of the cluster split every layer head-parallel ("tensor parallel" over distributed shared
// memory) and nothing but three tiny all-reduces per layer crosses CTA boundaries.
//
// Restates reference model.py:125-151 (Decoder.evaluate) with a device-resident KV cache.  Why this shape: at the
// BASELINE batch (64 utterances per GPU) a decode step is a chain of ~50 dependent micro-operations; grid-wide
// barriers (decode_persistent.cu) cost microseconds each, and one CTA per utterance (decode_stream.cu) has to pull
// the whole 11 MB weight set through ONE SM per step.  Here
//   * cluster size = num_heads; CTA r owns head r: its slice of the packed QKV / cross-Q weights, the attention of
//     that head over the self cache and the encoder K/V, the K-slice of both out projections, FF/H rows of the FFN
//     squeeze and the matching K-slice of the unsqueeze.  Out projections and the FFN unsqueeze produce partial sums
//     over the full model dimension, combined by an all-reduce written straight into the peers' shared memory
//     (st.async + mbarrier complete_tx); no cluster-wide barrier, no global memory round trip;
//   * every byte a CTA reads (its weight slices in consumption order, the K/V rows of its head) arrives through a
//     deep shared-memory ring filled by TMA (1-D bulk copies for the packed weight image and the self cache, 2-D
//     tensor-map copies for the encoder K/V) issued by a dedicated producer thread that runs a full ring ahead:
//     the access sequence is static, so memory latency is hidden and the step is bound by the L2 -> SM stream;
//   * Linear layers run on the tensor cores with the roles swapped (weights = A operand, the <= 8 utterances of the
//     cluster = N dimension of mma.m16n8k16) from a fragment-major packed image: one conflict-free LDS.128 IS the A
//     fragment of one MMA; activations are fp32-accurate (bf16 hi + lo split, two passes, fp32 accumulate - Q13).
// LayerNorm / softmax / residual / logits are fp32; K/V caches bf16; argmax lowest-index tie-break (model.py:143).
// The kernel is specialised at compile time for (heads, FFN rows per CTA, vocabulary rows per CTA, utterance slots).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

constexpr int NCW = 8;                    // consumer warps
constexpr int NCT = NCW * 32;             // consumer threads
constexpr int NTHREADS = NCT + 32;        // + producer warp
constexpr int STAGE_BYTES = 32768;
constexpr int MAX_STAGES = 6;
constexpr float LOG2E = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 8-byte store into a peer CTA's shared memory that also signals 8 bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_v2(uint32_t raddr, float a, float b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "f"(a), "f"(b), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must surface as a launch failure, not a hang
  }
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------ static shapes
template <int H_, int FFS_, int VS_, int GUP_>
struct Shape {
  static constexpr int H = H_, CS = H_, D = 64 * H_, FFS = FFS_, VS = VS_, GUP = GUP_;
  static constexpr int RPS = (STAGE_BYTES / 256) / GUP;   // key rows per utterance slot per ring stage
  static constexpr int WPU = NCW / GUP;                   // attention warps per utterance slot
  static constexpr int NGROUPS = 4 * WPU;                 // 8-lane key groups per utterance slot
  static constexpr int SMALL_FLOATS = 256 + FFS + 9 * D;
  static constexpr uint32_t SMALL_BYTES = (SMALL_FLOATS * 4 + 127) / 128 * 128;
  static constexpr int LDX = (D + 32) * 2, LDH = (FFS + 32) * 2, LDO = 96 * 2;   // activation row strides (bytes)
};
// One packed matrix: MT m-tiles of 16 rows, KB k-blocks of 32 columns; KBS k-blocks per ring stage; the units
// (m-tile, k-group) are dealt to the 8 warps; KG > 1 only where MT alone does not divide by 8.
template <int MT_, int KB_>
struct Mat {
  static constexpr int MT = MT_, KB = KB_;
  static constexpr int KBS = (32 / MT) < 1 ? 1 : ((32 / MT) > KB ? KB : (32 / MT));
  static constexpr int KG = (MT % 8 == 0) ? 1 : (MT % 4 == 0) ? 2 : (MT % 2 == 0) ? 4 : 8;
  static constexpr int UPW = MT * KG / 8, KPG = KBS / KG, NST = KB / KBS;
  static constexpr uint32_t ST_BYTES = KBS * MT * 1024u;
  static_assert(MT >= 1 && MT <= 32 && KBS % KG == 0 && KB % KBS == 0 && (MT * KG) % 8 == 0 && UPW <= 4, "tiling");
};

struct SmemMap {
  uint32_t ring, h, xn_hi, xn_lo, hid_hi, hid_lo, o_hi, o_lo, q, kvrow, scratch, prm, lg, recv, arg, part, stat, tok,
      ctrl, bars, total;
};
template <class S>
__host__ __device__ inline SmemMap smem_map(int nstages) {
  SmemMap m;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t o = off;
    off += (bytes + 127u) & ~127u;
    return o;
  };
  m.ring = take(uint32_t(nstages) * STAGE_BYTES);
  m.h = take(S::GUP * S::D * 4);
  m.xn_hi = take(S::GUP * S::LDX);
  m.xn_lo = take(S::GUP * S::LDX);
  m.hid_hi = take(S::GUP * S::LDH);
  m.hid_lo = take(S::GUP * S::LDH);
  m.o_hi = take(S::GUP * S::LDO);
  m.o_lo = take(S::GUP * S::LDO);
  m.q = take(S::GUP * 64 * 4);
  m.kvrow = take(S::GUP * 128 * 2);
  m.scratch = take(32 * 32 * 16);                      // [KG * MT <= 32][32 lanes] float4 partial tiles
  m.prm = take(S::SMALL_BYTES);
  m.lg = take(S::GUP * S::VS * 4);
  m.recv = take(2 * S::CS * S::D * S::GUP * 4);        // [parity][source rank][D][GUP] fp32 partial sums
  m.arg = take(2 * S::CS * S::GUP * 8);                // [parity][source rank][GUP] (value, index)
  m.part = take(NCW * 64 * 4);
  m.stat = take(2 * NCW * 4);
  m.tok = take(2 * 8 * 4);                             // [0..7] next tokens, [8..15] finished flags
  m.ctrl = take(16);                                   // [0] steps done, [1] stop, [2] cache rows written
  m.bars = take((2 * MAX_STAGES + 4) * 8);
  m.total = off;
  return m;
}

struct Ring {
  uint8_t* buf;
  uint64_t* full;
  uint64_t* empty;
  int nstages;
};

struct Producer {
  Ring r;
  int slot = 0;
  uint32_t round = 0;
  long long waited = 0;
  __device__ __forceinline__ uint8_t* begin(uint32_t bytes) {
    const long long w0 = clock64();
    mbar_wait(&r.empty[slot], (round & 1u) ^ 1u);
    waited += clock64() - w0;
    mbar_expect_tx(&r.full[slot], bytes);
    return r.buf + size_t(slot) * STAGE_BYTES;
  }
  __device__ __forceinline__ uint64_t* bar() { return &r.full[slot]; }
  __device__ __forceinline__ void end() {
    if (++slot == r.nstages) {
      slot = 0;
      ++round;
    }
  }
  template <class M>
  __device__ __forceinline__ void mat(const uint8_t* src, uint64_t pol) {   // consecutive stages of KBS k-blocks
#pragma unroll 1
    for (int s = 0; s < M::NST; ++s) {
      uint8_t* dst = begin(M::ST_BYTES);
      bulk_load(dst, src, M::ST_BYTES, bar(), pol);
      src += M::ST_BYTES;
      end();
    }
  }
};

struct Consumer {
  Ring r;
  int slot = 0;
  uint32_t round = 0;
  long long waited = 0;
  __device__ __forceinline__ const uint8_t* acquire() {
    const long long w0 = clock64();
    mbar_wait(&r.full[slot], round & 1u);
    waited += clock64() - w0;
    return r.buf + size_t(slot) * STAGE_BYTES;
  }
  __device__ __forceinline__ void release() {       // every consumer warp calls this once per stage
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&r.empty[slot]);
    if (++slot == r.nstages) {
      slot = 0;
      ++round;
    }
  }
};

// ------------------------------------------------------------------------------------------------ streamed matmul
// out[n, u] = sum_k W[n, k] x[u, k] for the rows n of one packed matrix and the GUP utterance rows of x, W streamed
// through the ring.  Image layout per k-block kb (32 columns) and m-tile mt (16 rows): 1 KB =
// [k-tile s (2)][lane = g*4 + tg (32)][16 B] with the 16 bytes = the mma.m16n8k16 A fragment {a0,a1,a2,a3} of that
// lane: {W[g][c..c+1], W[g+8][c..c+1], W[g][c+2..c+3], W[g+8][c+2..c+3]}, c = 32 kb + 8 tg + 4 s, i.e. the K
// permutation k_mma {2tg, 2tg+1, 2tg+8, 2tg+9} <-> columns {c..c+3}.  The B fragment applies the same permutation:
// lane (g, tg) reads the 16 bytes x[u = g][32 kb + 8 tg .. +7] (bf16 hi and lo copies; row stride == 64 mod 128 B:
// conflict free) and feeds halves s = 0 / 1 to the two MMAs.  Unit (m-tile mt, k-group kg) belongs to warp
// (mt * KG + kg) % 8.  The finished tile {(n = 16mt+g, u = 2tg), (n, u+1), (n+8, u), (n+8, u+1)} goes to
// epi(n, u0, v(u0), v(u0+1)); with KG > 1 the k-group partials are first combined through `scratch`.
template <class M, int GUP, class Epi>
__device__ __forceinline__ void mm_stream(Consumer& c, const uint8_t* xhi, const uint8_t* xlo, int ldx, float4* scratch,
                                          Epi&& epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int kg = warp % M::KG, mt0 = warp / M::KG;          // unit j of this warp: m-tile mt0 + j * (8 / KG)
  float acc[M::UPW][2][4];
#pragma unroll
  for (int j = 0; j < M::UPW; ++j)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][s][i] = 0.f;
  const bool xrow = g < GUP;
  const uint8_t* xh = xhi + g * ldx + tg * 16 + kg * M::KPG * 64;
  const uint8_t* xl = xlo + g * ldx + tg * 16 + kg * M::KPG * 64;
#pragma unroll 1
  for (int st_i = 0; st_i < M::NST; ++st_i) {
    const uint8_t* st = c.acquire() + lane * 16 + (size_t(kg) * M::KPG * M::MT + mt0) * 1024;
#pragma unroll
    for (int q = 0; q < M::KPG; ++q) {
      uint4 bh = make_uint4(0, 0, 0, 0), bl = make_uint4(0, 0, 0, 0);
      if (xrow) {
        bh = lds128(xh + (st_i * M::KBS + q) * 64);
        bl = lds128(xl + (st_i * M::KBS + q) * 64);
      }
#pragma unroll
      for (int j = 0; j < M::UPW; ++j) {
        const uint8_t* a = st + (size_t(q) * M::MT + j * (8 / M::KG)) * 1024;
        const uint4 a0 = lds128(a), a1 = lds128(a + 512);
        mma16816(acc[j][0], a0, bh.x, bh.y);
        mma16816(acc[j][1], a1, bh.z, bh.w);
        mma16816(acc[j][0], a0, bl.x, bl.y);
        mma16816(acc[j][1], a1, bl.z, bl.w);
      }
    }
    c.release();
  }
  if (M::KG == 1) {
#pragma unroll
    for (int j = 0; j < M::UPW; ++j) {
      const int n0 = 16 * (mt0 + j * 8) + g;
      if (2 * tg < GUP) {
        epi(n0, 2 * tg, acc[j][0][0] + acc[j][1][0], acc[j][0][1] + acc[j][1][1]);
        epi(n0 + 8, 2 * tg, acc[j][0][2] + acc[j][1][2], acc[j][0][3] + acc[j][1][3]);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < M::UPW; ++j)
      scratch[(kg * M::MT + mt0 + j * (8 / M::KG)) * 32 + lane] =
          make_float4(acc[j][0][0] + acc[j][1][0], acc[j][0][1] + acc[j][1][1], acc[j][0][2] + acc[j][1][2],
                      acc[j][0][3] + acc[j][1][3]);
    consumer_sync();
    for (int it = threadIdx.x; it < M::MT * 32; it += NCT) {
      const int ln = it & 31, g2 = ln >> 2, tg2 = ln & 3;
      if (2 * tg2 < GUP) {
        float4 s = scratch[it];
#pragma unroll
        for (int k = 1; k < M::KG; ++k) {
          const float4 v = scratch[k * M::MT * 32 + it];
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const int n0 = 16 * (it >> 5) + g2;
        epi(n0, 2 * tg2, s.x, s.y);
        epi(n0 + 8, 2 * tg2, s.z, s.w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm / split
// rows u < GU of h (fp32, stride D) -> bf16 hi + lo rows (stride ld elements); warp u handles row u.
template <int D>
__device__ __forceinline__ void rows_to_hilo(const float* h, int GU, const float* gam, const float* bet, bf16* hi,
                                             bf16* lo, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < GU) {
    const float* src = h + warp * D;
    float x[D / 32];
#pragma unroll
    for (int i = 0; i < D / 32; ++i) x[i] = src[lane + 32 * i];
    if (gam) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < D / 32; ++i) sum += x[i];
      const float mean = warp_sum(sum) * (1.0f / float(D));
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < D / 32; ++i) {
        const float d = x[i] - mean;
        sq += d * d;
      }
      const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / float(D)) + 1e-5f);
#pragma unroll
      for (int i = 0; i < D / 32; ++i) x[i] = (x[i] - mean) * rstd * gam[lane + 32 * i] + bet[lane + 32 * i];
    }
#pragma unroll
    for (int i = 0; i < D / 32; ++i) {
      const bf16 hh = __float2bfloat16(x[i]);
      hi[warp * ld + lane + 32 * i] = hh;
      lo[warp * ld + lane + 32 * i] = __float2bfloat16(x[i] - __bfloat162float(hh));
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// single-query attention of ONE head, flash style: one running (max, sum, acc[8]) per 8-lane key group, log2 units
struct Attn {
  float qv[8];
  float m, l;
  float o[8];
};
__device__ __forceinline__ void attn_begin(Attn& st, const float* q /* 64 floats, pre-scaled by scale*log2e */) {
  const int c8 = threadIdx.x & 7;
  const float4 a = *reinterpret_cast<const float4*>(q + c8 * 8);
  const float4 b = *reinterpret_cast<const float4*>(q + c8 * 8 + 4);
  st.qv[0] = a.x; st.qv[1] = a.y; st.qv[2] = a.z; st.qv[3] = a.w;
  st.qv[4] = b.x; st.qv[5] = b.y; st.qv[6] = b.z; st.qv[7] = b.w;
  st.m = -INFINITY;
  st.l = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) st.o[i] = 0.f;
}
template <int NB>
__device__ __forceinline__ void attn_fold(Attn& st, const uint4 (&kr)[NB], const uint4 (&vr)[NB], const bool (&valid)[NB],
                                          unsigned gmask) {
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
  if (bm == -INFINITY) return;
  const float m_new = fmaxf(st.m, bm);
  const float alpha = exp2f(st.m - m_new);
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
// one ring stage: per utterance slot [K: RPS rows x 128 B][V: RPS rows x 128 B]; this warp's slot holds n_rows keys.
// Group grp takes rows grp + j * NGROUPS, j = 0..3 (RPS == 4 * NGROUPS).
template <class S>
__device__ __forceinline__ void attn_stage(Attn& st, const uint8_t* slot, int n_rows, int grp, bool active) {
  const int lane = threadIdx.x & 31, c8 = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const uint8_t* kb = slot + c8 * 16 + grp * 128;
  uint4 kr[4], vr[4];
  bool valid[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    valid[j] = active && (grp + j * S::NGROUPS) < n_rows;
    if (valid[j]) {
      kr[j] = lds128(kb + j * S::NGROUPS * 128);
      vr[j] = lds128(kb + j * S::NGROUPS * 128 + S::RPS * 128);
    }
  }
  attn_fold<4>(st, kr, vr, valid, gmask);
}
// merge the key groups of the cluster's utterance slots and emit o (bf16 hi + lo rows, stride 96 elements)
template <class S>
__device__ __forceinline__ void attn_finish(Attn& st, int GU, float* part_buf, float* stat, bf16* o_hi, bf16* o_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c8 = lane & 7, sub = lane >> 3;
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
    *reinterpret_cast<float4*>(part_buf + warp * 64 + c8 * 8) = make_float4(st.o[0], st.o[1], st.o[2], st.o[3]);
    *reinterpret_cast<float4*>(part_buf + warp * 64 + c8 * 8 + 4) = make_float4(st.o[4], st.o[5], st.o[6], st.o[7]);
    if (c8 == 0) {
      stat[warp] = st.m;
      stat[NCW + warp] = st.l;
    }
  }
  consumer_sync();
  for (int d = threadIdx.x; d < GU * 64; d += NCT) {
    const int uu = d >> 6, dd = d & 63;
    float mm = -INFINITY;
#pragma unroll
    for (int pI = 0; pI < S::WPU; ++pI) mm = fmaxf(mm, stat[pI * S::GUP + uu]);
    float t = 0.f, ls = 0.f;
#pragma unroll
    for (int pI = 0; pI < S::WPU; ++pI) {
      const float mw = stat[pI * S::GUP + uu];
      const float f = (mw == -INFINITY) ? 0.f : exp2f(mw - mm);
      t += part_buf[(pI * S::GUP + uu) * 64 + dd] * f;
      ls += stat[NCW + pI * S::GUP + uu] * f;
    }
    const float y = ls > 0.f ? t / ls : 0.f;
    const bf16 hh = __float2bfloat16(y);
    o_hi[uu * 96 + dd] = hh;
    o_lo[uu * 96 + dd] = __float2bfloat16(y - __bfloat162float(hh));
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ the kernel
template <class S>
__global__ void __launch_bounds__(NTHREADS, 1)
dec_cluster_kernel(const __grid_constant__ ClusterParams p, const __grid_constant__ CUtensorMap ckv_map) {
  constexpr int D = S::D, H = S::H, CS = S::CS, GUP = S::GUP, FFS = S::FFS, VS = S::VS, RPS = S::RPS;
  using MQkv = Mat<12, D / 32>;
  using MWo = Mat<D / 16, 2>;
  using MWqc = Mat<4, D / 32>;
  using MW1 = Mat<FFS / 16, D / 32>;
  using MW2 = Mat<D / 16, FFS / 32>;
  using MCls = Mat<VS / 16, D / 32>;

  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemMap sm = smem_map<S>(p.nstages);
  Ring ring;
  ring.buf = smem + sm.ring;
  ring.nstages = p.nstages;
  float* s_h = reinterpret_cast<float*>(smem + sm.h);
  bf16* xn_hi = reinterpret_cast<bf16*>(smem + sm.xn_hi);
  bf16* xn_lo = reinterpret_cast<bf16*>(smem + sm.xn_lo);
  bf16* hid_hi = reinterpret_cast<bf16*>(smem + sm.hid_hi);
  bf16* hid_lo = reinterpret_cast<bf16*>(smem + sm.hid_lo);
  bf16* o_hi = reinterpret_cast<bf16*>(smem + sm.o_hi);
  bf16* o_lo = reinterpret_cast<bf16*>(smem + sm.o_lo);
  float* s_q = reinterpret_cast<float*>(smem + sm.q);
  bf16* kv_row = reinterpret_cast<bf16*>(smem + sm.kvrow);
  float4* scratch = reinterpret_cast<float4*>(smem + sm.scratch);
  float* prm = reinterpret_cast<float*>(smem + sm.prm);
  float* s_lg = reinterpret_cast<float*>(smem + sm.lg);
  float* recv = reinterpret_cast<float*>(smem + sm.recv);
  float2* arg = reinterpret_cast<float2*>(smem + sm.arg);
  float* part_buf = reinterpret_cast<float*>(smem + sm.part);
  float* stat = reinterpret_cast<float*>(smem + sm.stat);
  int* s_tok = reinterpret_cast<int*>(smem + sm.tok);
  volatile int* ctrl = reinterpret_cast<volatile int*>(smem + sm.ctrl);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  ring.full = bars;
  ring.empty = bars + MAX_STAGES;
  uint64_t* xbar = bars + 2 * MAX_STAGES;       // [2] all-reduce parity barriers
  uint64_t* abar = xbar + 2;                    // [2] argmax exchange parity barriers

  const int rank = int(cluster_ctarank());
  const int cluster_id = blockIdx.x / CS;
  const int ubase = cluster_id * p.GU;
  const int GU = min(p.GU, p.B - ubase);         // utterances of this cluster (>= 1 by construction of the grid)
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;

  if (tid == 0) {
    if (smem_u32(smem) & 127u) __trap();         // TMA destinations need 128-byte alignment
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(&ring.full[s], 1);
      mbar_init(&ring.empty[s], NCW);
    }
    mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1);
    mbar_init(&abar[0], 1); mbar_init(&abar[1], 1);
    ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0;
    fence_barrier_init();
  }
  // zero every activation buffer once: rows of absent utterances (u >= GU) stay zero for the whole decode
  for (uint32_t i = sm.h / 4 + tid; i < sm.scratch / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid < 16) s_tok[tid] = 0;
  __syncthreads();
  cluster_sync_all();   // peers' barriers are initialised before any remote store can arrive

  const size_t cache_head = size_t(2) * p.L * 64;                 // elements per (layer, utterance, head): K rows | V rows
  const uint8_t* my_image = p.image + size_t(rank) * p.rank_bytes;

  if (warp == NCW) {
    // =============================== producer: one thread walks the static access sequence
    if (tid == NCT) {
      const uint64_t pol_w = policy_evict_last();
      const uint64_t pol_kv = p.kv_evict_first ? policy_evict_first() : policy_evict_last();
      Producer pr;
      pr.r = ring;
#pragma unroll 1
      for (int t = 0; t < p.L; ++t) {
        if (p.stop_at_eos) {                       // strict gate: nothing of step t is requested before step t-1 ended
          while (ctrl[0] < t && !ctrl[1]) {
          }
          if (ctrl[1]) break;
        }
#pragma unroll 1
        for (int l = 0; l < p.nd; ++l) {
          const uint8_t* img = my_image + size_t(l) * p.layer_bytes;
          {
            uint8_t* dst = pr.begin(S::SMALL_BYTES);
            bulk_load(dst, img + p.off_small, S::SMALL_BYTES, pr.bar(), pol_w);
            pr.end();
          }
          pr.mat<MQkv>(img + p.off_qkv, pol_w);
          if (t > 0) {   // self cache rows 0..t-1 of this layer (written by this CTA in earlier steps)
            const int need = (t - 1) * p.nd + l + 1;
            while (ctrl[2] < need) {
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            for (int c0 = 0; c0 < t; c0 += RPS) {
              const int n = min(RPS, t - c0);
              uint8_t* dst = pr.begin(uint32_t(GU) * 2u * n * 128u);
              for (int u = 0; u < GU; ++u) {
                const bf16* kp = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head + size_t(c0) * 64;
                bulk_load(dst + u * RPS * 256, kp, n * 128, pr.bar(), pol_kv);
                bulk_load(dst + u * RPS * 256 + RPS * 128, kp + size_t(p.L) * 64, n * 128, pr.bar(), pol_kv);
              }
              pr.end();
            }
          }
          pr.mat<MWo>(img + p.off_wo, pol_w);
          pr.mat<MWqc>(img + p.off_wqc, pol_w);
          for (int c0 = 0; c0 < p.Tp; c0 += RPS) {   // encoder K/V of this head: 2-D boxes [RPS rows][64 columns]
            uint8_t* dst = pr.begin(uint32_t(GU) * 2u * RPS * 128u);
            for (int u = 0; u < GU; ++u) {
              const int row = (l * p.B + ubase + u) * p.Tp + c0;
              tma_load_2d_hint(dst + u * RPS * 256, &ckv_map, pr.bar(), rank * 64, row, pol_kv);
              tma_load_2d_hint(dst + u * RPS * 256 + RPS * 128, &ckv_map, pr.bar(), D + rank * 64, row, pol_kv);
            }
            pr.end();
          }
          pr.mat<MWo>(img + p.off_woc, pol_w);
          pr.mat<MW1>(img + p.off_w1, pol_w);
          pr.mat<MW2>(img + p.off_w2, pol_w);
        }
        pr.mat<MCls>(my_image + p.off_cls, pol_w);
      }
      if (p.timing) {
        p.timing[size_t(blockIdx.x) * 16 + 3] = pr.waited;
        p.timing[size_t(blockIdx.x) * 16 + 4] = (long long)pr.round * p.nstages + pr.slot;
      }
    }
    __syncwarp();
  } else {
    // ================================= consumers
    Consumer c;
    c.r = ring;
    const long long t_begin = clock64();
    long long t_xchg = 0;
    uint32_t n_xchg = 0, n_arg = 0;
    const int lane = tid & 31, sub = lane >> 3;
    const int au = warp % GUP, agrp = (warp / GUP) * 4 + sub;       // attention: utterance slot / key group
    const bool a_active = au < GU;
    const float qscale = p.scale * LOG2E;
    const uint8_t* xh = reinterpret_cast<const uint8_t*>(xn_hi);
    const uint8_t* xl = reinterpret_cast<const uint8_t*>(xn_lo);

    // embedding + PE of the first token (host-side init kernel)
    for (int i = tid * 4; i < GU * D; i += NCT * 4)
      *reinterpret_cast<float4*>(s_h + i) = *reinterpret_cast<const float4*>(p.h0 + size_t(ubase) * D + i);
    consumer_sync();

    // partial sums over the full model dimension (this CTA's K-slice) -> all-reduce across the cluster:
    // h[u][n] += sum over ranks + bias[n].  Called right after the mm_stream that sent the partial tiles.
    auto send_partial = [&](int n, int u0, float v0, float v1) {
      const uint32_t par = n_xchg & 1u;
      float* mine = recv + ((size_t(par) * CS + rank) * D + n) * GUP + u0;
      *reinterpret_cast<float2*>(mine) = make_float2(v0, v1);
      const uint32_t ma = smem_u32(mine), ba = smem_u32(&xbar[par]);
#pragma unroll
      for (int r = 0; r < CS; ++r)
        if (r != rank) st_async_v2(mapa_u32(ma, r), v0, v1, mapa_u32(ba, r));
    };
    auto all_reduce_finish = [&](const float* bias) {
      const uint32_t par = n_xchg & 1u, phase = (n_xchg >> 1) & 1u;
      if (tid == 0) mbar_expect_tx(&xbar[par], uint32_t(CS - 1) * D * GUP * 4u);
      consumer_sync();                                     // own slot written by every thread
      const long long w0 = clock64();
      mbar_wait_cluster(&xbar[par], phase);
      t_xchg += clock64() - w0;
      const float* rv = recv + size_t(par) * CS * D * GUP;
      for (int idx = tid; idx < D * GUP; idx += NCT) {
        const int n = idx / GUP, u = idx % GUP;
        if (u < GU) {
          float s = 0.f;
#pragma unroll
          for (int r = 0; r < CS; ++r) s += rv[r * D * GUP + idx];
          s_h[u * D + n] += s + bias[n];
        }
      }
      consumer_sync();
      ++n_xchg;
    };

    int t = 0;
#pragma unroll 1
    for (; t < p.L; ++t) {
#pragma unroll 1
      for (int l = 0; l < p.nd; ++l) {
        // ---- this layer's biases and LayerNorm parameters (one ring stage)
        {
          const float4* src = reinterpret_cast<const float4*>(c.acquire());
          float4* dst = reinterpret_cast<float4*>(prm);
          for (int i = tid; i < int(S::SMALL_BYTES / 16); i += NCT) dst[i] = src[i];
          c.release();
          consumer_sync();
        }
        const float* b_qkv = prm;                 // [192] q | k | v rows of this head
        const float* b_qc = prm + 192;            // [64]
        const float* b_1 = prm + 256;             // [FFS]
        const float* b_o = prm + 256 + FFS;       // [D]
        const float* b_oc = b_o + D;
        const float* b_2 = b_oc + D;
        const float* ln = b_2 + D;                // ln1 g,b | ln2 g,b | ln3 g,b

        // ---- LN1 -> q, k, v of this head (model.py:67-68, layers.py:16-18)
        rows_to_hilo<D>(s_h, GU, ln, ln + D, xn_hi, xn_lo, D + 32);
        consumer_sync();
        mm_stream<MQkv, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, float v0, float v1) {
          const float y0 = v0 + b_qkv[n], y1 = v1 + b_qkv[n];
          if (n < 64) {
            s_q[u0 * 64 + n] = y0 * qscale;
            s_q[(u0 + 1) * 64 + n] = y1 * qscale;
          } else {
            kv_row[u0 * 128 + (n - 64)] = __float2bfloat16(y0);
            kv_row[(u0 + 1) * 128 + (n - 64)] = __float2bfloat16(y1);
          }
        });
        consumer_sync();
        // append k_t, v_t (bf16) to the device-resident cache: [layer][utterance][head][K rows | V rows][64]
        if (tid < GU * 16) {
          const int u = tid >> 4, ch = tid & 15, kv = ch >> 3, c16 = ch & 7;
          bf16* dst = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head + size_t(kv) * p.L * 64 +
                      size_t(t) * 64 + c16 * 8;
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(kv_row + u * 128 + kv * 64 + c16 * 8);
          asm volatile("fence.proxy.async;" ::: "memory");   // later read back by the producer's bulk copies
        }
        // ---- causal self attention over keys 0..t (the current row comes from shared memory)
        Attn st;
        attn_begin(st, s_q + au * 64);
        {
          uint4 kr[1], vr[1];
          bool valid[1];
          valid[0] = a_active && agrp == 0;
          if (valid[0]) {
            kr[0] = lds128(kv_row + au * 128 + (lane & 7) * 8);
            vr[0] = lds128(kv_row + au * 128 + 64 + (lane & 7) * 8);
          }
          attn_fold<1>(st, kr, vr, valid, 0xFFu << (lane & 24));
        }
#pragma unroll 1
        for (int c0 = 0; c0 < t; c0 += RPS) {
          const uint8_t* stg = c.acquire();
          attn_stage<S>(st, stg + au * RPS * 256, min(RPS, t - c0), agrp, a_active);
          c.release();
        }
        attn_finish<S>(st, GU, part_buf, stat, o_hi, o_lo);   // (two consumer barriers inside)
        if (tid == 0) {                                       // cache row t of this layer is published
          __threadfence_block();
          ctrl[2] = t * p.nd + l + 1;
        }
        mm_stream<MWo, GUP>(c, reinterpret_cast<const uint8_t*>(o_hi), reinterpret_cast<const uint8_t*>(o_lo), S::LDO,
                            scratch, send_partial);
        all_reduce_finish(b_o);                               // out projection + residual (model.py:68)

        // ---- LN2 -> cross-attention query -> attention over the encoder K/V, never masked (model.py:70-71)
        rows_to_hilo<D>(s_h, GU, ln + 2 * D, ln + 3 * D, xn_hi, xn_lo, D + 32);
        consumer_sync();
        mm_stream<MWqc, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, float v0, float v1) {
          s_q[u0 * 64 + n] = (v0 + b_qc[n]) * qscale;
          s_q[(u0 + 1) * 64 + n] = (v1 + b_qc[n]) * qscale;
        });
        consumer_sync();
        attn_begin(st, s_q + au * 64);
#pragma unroll 1
        for (int c0 = 0; c0 < p.Tp; c0 += RPS) {
          const uint8_t* stg = c.acquire();
          attn_stage<S>(st, stg + au * RPS * 256, min(RPS, p.Tp - c0), agrp, a_active);
          c.release();
        }
        attn_finish<S>(st, GU, part_buf, stat, o_hi, o_lo);
        mm_stream<MWo, GUP>(c, reinterpret_cast<const uint8_t*>(o_hi), reinterpret_cast<const uint8_t*>(o_lo), S::LDO,
                            scratch, send_partial);
        all_reduce_finish(b_oc);

        // ---- LN3 -> FFN: squeeze rows of this CTA + ReLU, then the matching K-slice of unsqueeze (model.py:73-74)
        rows_to_hilo<D>(s_h, GU, ln + 4 * D, ln + 5 * D, xn_hi, xn_lo, D + 32);
        consumer_sync();
        mm_stream<MW1, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, float v0, float v1) {
          const float y0 = fmaxf(v0 + b_1[n], 0.f), y1 = fmaxf(v1 + b_1[n], 0.f);
          const bf16 h0 = __float2bfloat16(y0), h1 = __float2bfloat16(y1);
          hid_hi[u0 * (FFS + 32) + n] = h0;
          hid_lo[u0 * (FFS + 32) + n] = __float2bfloat16(y0 - __bfloat162float(h0));
          hid_hi[(u0 + 1) * (FFS + 32) + n] = h1;
          hid_lo[(u0 + 1) * (FFS + 32) + n] = __float2bfloat16(y1 - __bfloat162float(h1));
        });
        consumer_sync();
        mm_stream<MW2, GUP>(c, reinterpret_cast<const uint8_t*>(hid_hi), reinterpret_cast<const uint8_t*>(hid_lo),
                            S::LDH, scratch, send_partial);
        all_reduce_finish(b_2);
      }

      // ---- classifier WITHOUT the final LayerNorm (model.py:142): VS vocabulary rows per CTA
      rows_to_hilo<D>(s_h, GU, nullptr, nullptr, xn_hi, xn_lo, D + 32);
      consumer_sync();
      mm_stream<MCls, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, float v0, float v1) {
        s_lg[u0 * VS + n] = v0;
        s_lg[(u0 + 1) * VS + n] = v1;
      });
      consumer_sync();
      const int v_lo = rank * VS, v_n = max(0, min(VS, p.V - v_lo));   // this CTA's vocabulary range
      if (p.step_logits)
        for (int i = tid; i < GU * v_n; i += NCT) {
          const int u = i / v_n, v = i - u * v_n;
          p.step_logits[(size_t(ubase + u) * p.L + t) * p.V + v_lo + v] = s_lg[u * VS + v];
        }
      // ---- argmax: local (warp u), then across the cluster (lowest index wins ties, model.py:143)
      const uint32_t apar = n_arg & 1u, aphase = (n_arg >> 1) & 1u;
      if (tid == 0) mbar_expect_tx(&abar[apar], uint32_t(CS - 1) * GUP * 8u);
      if (warp < GUP) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        if (warp < GU)
          for (int v = lane; v < v_n; v += 32) {
            const float x = s_lg[warp * VS + v];
            if (x > best) {
              best = x;
              bi = v_lo + v;
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
        if (lane == 0) {
          float2* mine = arg + (size_t(apar) * CS + rank) * GUP + warp;
          *mine = make_float2(best, __int_as_float(bi));
          const uint32_t ma = smem_u32(mine), ba = smem_u32(&abar[apar]);
          for (int r = 0; r < CS; ++r)
            if (r != rank) st_async_v2(mapa_u32(ma, r), best, __int_as_float(bi), mapa_u32(ba, r));
        }
      }
      consumer_sync();
      {
        const long long w0 = clock64();
        mbar_wait_cluster(&abar[apar], aphase);
        t_xchg += clock64() - w0;
      }
      ++n_arg;
      if (tid < GU) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int r = 0; r < CS; ++r) {   // ranks own ascending vocabulary ranges: strict > keeps the lowest index
          const float2 a = arg[(size_t(apar) * CS + r) * GUP + tid];
          const int ai = __float_as_int(a.y);
          if (a.x > best || (a.x == best && ai < bi)) {
            best = a.x;
            bi = ai;
          }
        }
        if (bi == 0x7fffffff) bi = 0;
        int tok = bi;
        if (p.stop_at_eos) {
          if (s_tok[8 + tid]) tok = p.pad;
          else if (tok == p.eos) {
            s_tok[8 + tid] = 1;
            if (rank == 0 && p.n_tokens) p.n_tokens[ubase + tid] = t + 2;
          }
        }
        s_tok[tid] = tok;
        if (rank == 0) p.tokens[size_t(ubase + tid) * (p.L + 1) + t + 1] = tok;
      }
      consumer_sync();
      bool stop = false;
      if (p.stop_at_eos) {
        int fin = 0;
        for (int u = 0; u < GU; ++u) fin += s_tok[8 + u];
        stop = fin == GU;
        if (tid == 0) {
          if (stop) ctrl[1] = 1;
          __threadfence_block();
          ctrl[0] = t + 1;
        }
      }
      if (stop) {
        ++t;
        break;
      }
      if (t + 1 < p.L) {   // embedding + PE of the next input token (model.py:137)
        for (int i = tid * 4; i < GU * D; i += NCT * 4) {
          const int u = i / D, d = i % D;
          const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + size_t(s_tok[u]) * D + d));
          const float4 q = __ldg(reinterpret_cast<const float4*>(p.pe + size_t(t + 1) * D + d));
          *reinterpret_cast<float4*>(s_h + i) = make_float4(e.x + q.x, e.y + q.y, e.z + q.z, e.w + q.w);
        }
      }
      consumer_sync();
    }
    // early exit (every utterance of the cluster finished): the remaining positions are padding
    if (p.stop_at_eos && rank == 0)
      for (int i = tid; i < GU * (p.L - t); i += NCT) {
        const int u = i / (p.L - t), k = t + 1 + i % (p.L - t);
        p.tokens[size_t(ubase + u) * (p.L + 1) + k] = p.pad;
      }
    if (p.timing && tid == 0) {
      p.timing[size_t(blockIdx.x) * 16 + 0] = clock64() - t_begin;
      p.timing[size_t(blockIdx.x) * 16 + 1] = c.waited;
      p.timing[size_t(blockIdx.x) * 16 + 2] = t_xchg;
    }
  }
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------ instances
typedef void (*ClusterKernel)(const ClusterParams, const CUtensorMap);
struct Instance {
  int H, FFS, VS, GUP;
  ClusterKernel fn;
  SmemMap (*map)(int);
};
#define ASR_INST(H, FFS, VS, G) {H, FFS, VS, G, dec_cluster_kernel<Shape<H, FFS, VS, G>>, smem_map<Shape<H, FFS, VS, G>>}
const Instance kInstances[] = {
    ASR_INST(4, 256, 64, 2), ASR_INST(4, 256, 64, 4), ASR_INST(4, 256, 64, 8),     // C1-C4: d_model 256, FFN 1024
    ASR_INST(2, 128, 128, 2), ASR_INST(2, 128, 128, 4), ASR_INST(2, 128, 128, 8),  // T0: d_model 128, FFN 256
    ASR_INST(8, 256, 32, 2), ASR_INST(8, 256, 32, 4),                              // C5: d_model 512, FFN 2048
};
const Instance* find_instance(int H, int FFS, int VS, int GUP) {
  for (const Instance& i : kInstances)
    if (i.H == H && i.FFS == FFS && i.VS == VS && i.GUP == GUP) return &i;
  return nullptr;
}

}  // namespace

bool cluster_layout(int D, int H, int FF, int V, int nd, ClusterLayout* out) {
  ClusterLayout L{};
  if (H < 2 || H > 8 || (H & (H - 1)) || D != 64 * H || FF % (32 * H) != 0 || V < 1 || nd < 1) return false;
  L.CS = H;
  L.FFS = FF / H;
  L.VS = ((V + H - 1) / H + 15) / 16 * 16;
  if (!find_instance(H, L.FFS, L.VS, 2)) return false;   // only the compiled shapes
  L.small_floats = 256 + L.FFS + 9 * D;
  L.small_bytes = (size_t(L.small_floats) * 4 + 127) / 128 * 128;
  size_t off = 0;
  L.off_small = off; off += L.small_bytes;
  L.off_qkv = off;   off += size_t(192) * D * 2;
  L.off_wo = off;    off += size_t(D) * 64 * 2;
  L.off_wqc = off;   off += size_t(64) * D * 2;
  L.off_woc = off;   off += size_t(D) * 64 * 2;
  L.off_w1 = off;    off += size_t(L.FFS) * D * 2;
  L.off_w2 = off;    off += size_t(D) * L.FFS * 2;
  L.layer_bytes = off;
  L.off_cls = size_t(nd) * L.layer_bytes;
  L.rank_bytes = L.off_cls + size_t(L.VS) * D * 2;
  L.total_bytes = L.rank_bytes * H;
  if (out) *out = L;
  return true;
}

int launch_dec_cluster(ClusterParams& p, cudaStream_t s) {
  ClusterLayout lay;
  if (!cluster_layout(p.D, p.H, p.FF, p.V, p.nd, &lay))
    return set_error(-2, "cluster decoder: unsupported config D=%d H=%d FF=%d V=%d", p.D, p.H, p.FF, p.V);
  if (!p.image || p.image_bytes != lay.total_bytes)
    return set_error(-1, "cluster decoder: packed image missing or wrong size (%zu, expected %zu bytes)", p.image_bytes,
                     lay.total_bytes);
  p.FFS = lay.FFS; p.VS = lay.VS; p.small_bytes = (uint32_t)lay.small_bytes;
  p.rank_bytes = lay.rank_bytes; p.layer_bytes = lay.layer_bytes;
  p.off_small = lay.off_small; p.off_qkv = lay.off_qkv; p.off_wo = lay.off_wo; p.off_wqc = lay.off_wqc;
  p.off_woc = lay.off_woc; p.off_w1 = lay.off_w1; p.off_w2 = lay.off_w2; p.off_cls = lay.off_cls;

  int dev = 0, max_smem = 0;
  ASR_CUDA_OK(cudaGetDevice(&dev));
  ASR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // utterances per cluster: the smallest group that lets every cluster be resident at once (one wave)
  const char* ge = std::getenv("ASR_B200_CLUSTER_GU");
  const int gu_forced = ge && ge[0] ? std::atoi(ge) : 0;
  const Instance* chosen = nullptr;
  int chosen_gu = 0, chosen_stages = 0;
  for (int gu = 1; gu <= 8; gu *= 2) {
    if (gu_forced && gu != gu_forced) continue;
    const Instance* inst = find_instance(p.H, p.FFS, p.VS, gu < 2 ? 2 : gu);
    if (!inst) break;
    int nst = MAX_STAGES;
    while (nst >= 3 && inst->map(nst).total > (uint32_t)max_smem) --nst;
    if (nst < 3) break;   // larger groups need even more shared memory
    static const void* configured[16] = {};
    bool done = false;
    for (const void* q : configured) done |= (q == (const void*)inst->fn);
    if (!done) {
      ASR_CUDA_OK(cudaFuncSetAttribute((const void*)inst->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      for (auto& q : configured)
        if (!q) {
          q = (const void*)inst->fn;
          break;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.H, 1, 1);
    cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = inst->map(nst).total;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.H; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int max_clusters = 0;
    ASR_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)inst->fn, &cfg));
    chosen = inst;
    chosen_gu = gu;
    chosen_stages = nst;
    if (gu_forced || (p.B + gu - 1) / gu <= max_clusters) break;
  }
  if (!chosen) return set_error(-2, "cluster decoder: no instance fits (D=%d FF=%d GU=%d)", p.D, p.FF, gu_forced);
  p.GU = chosen_gu;
  p.GUP = chosen->GUP;
  p.nstages = chosen_stages;
  {
    const char* e = std::getenv("ASR_B200_KV_POLICY");   // "first" (default: stream K/V past the L2-resident weights) / "last"
    p.kv_evict_first = !(e && e[0] == 'l');
  }

  // encoder K/V as a 2-D tensor: [nd * B * Tp rows][2D columns] bf16; box = [RPS rows][64 columns] (one head)
  CUtensorMap map;
  const int rps = (STAGE_BYTES / 256) / p.GUP;
  const uint64_t dims[2] = {uint64_t(2 * p.D), uint64_t(p.nd) * p.B * p.Tp};
  const uint64_t strides[2] = {0, uint64_t(4 * p.D)};
  const uint32_t box[2] = {64u, uint32_t(rps)};
  if (int rc = make_tmap_bf16(&map, p.ckv, 2, dims, strides, box, nullptr, /*swizzle=*/0)) return rc;

  const int n_clusters = (p.B + p.GU - 1) / p.GU;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_clusters * p.H, 1, 1);
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = chosen->map(p.nstages).total;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.H; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  ASR_CUDA_OK(cudaLaunchKernelEx(&cfg, chosen->fn, p, map));
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
