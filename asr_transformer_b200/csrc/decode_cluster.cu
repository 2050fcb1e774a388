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
//     fragment of one MMA; activations are fp32-accurate (f16 hi + lo split, two passes, fp32 accumulate - Q13).
// LayerNorm / softmax / residual / logits are fp32; K/V caches f16; argmax lowest-index tie-break (model.py:143).
// The kernel is specialised at compile time for (heads, FFN rows per CTA, vocabulary rows per CTA, utterance slots).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

#ifndef ASR_PEEK_MODE
#define ASR_PEEK_MODE 0
#endif
#ifndef ASR_NCW
#define ASR_NCW 8
#endif
#ifndef ASR_STAGE_KB
#define ASR_STAGE_KB 32
#endif
constexpr int NCW = ASR_NCW;              // consumer warps (16 measured slower: per-warp fixed costs dominate, not divisible work)
constexpr int NCT = NCW * 32;             // consumer threads
constexpr int NTHREADS = NCT + 32;        // + producer warp
constexpr int STAGE_BYTES = ASR_STAGE_KB * 1024;
constexpr int MAX_STAGES = 6 * 32 / ASR_STAGE_KB;
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
__device__ __noinline__ void mbar_wait_cluster_slow(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must surface as a launch failure, not a hang
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  mbar_wait_cluster_slow(bar, parity);
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
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, "
      "%6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
// TMA store of a small contiguous block shared -> global (async proxy), tracked by the thread's bulk group
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------ static shapes
template <int H_, int FFS_, int VS_, int GUP_>
struct Shape {
  static constexpr int H = H_, CS = H_, D = 64 * H_, FFS = FFS_, VS = VS_, GUP = GUP_;
  static constexpr int RPS = (STAGE_BYTES / 128) / GUP;   // key rows per utterance slot per ring stage (K or V rows)
  static constexpr int WPU = NCW / GUP;                   // attention warps per utterance slot
  static constexpr int SCX = GUP >= 4 ? 4 : 2;            // K stages per super-chunk (softmax granularity): 256 keys (128 at GUP 8;
                                                          // 256 there: 15.4 -> 17.3 ms, the V stages fall too far behind their K stages)
  static constexpr int TPW = (RPS / 16) / WPU;            // 16-key tiles per warp per stage
  static constexpr int SMALL_FLOATS = 256 + FFS + 11 * D;   // ... | ln1 | ln2 | ln3 | ln1 of the NEXT layer
  static constexpr uint32_t SMALL_BYTES = (SMALL_FLOATS * 4 + 127) / 128 * 128;
  static constexpr int LDX = (D + 32) * 2, LDH = (FFS + 32) * 2, LDO = 96 * 2;   // activation row strides (bytes)
};
// One packed matrix: MT m-tiles of 16 rows, KB k-blocks of 32 columns; KBS k-blocks per ring stage.  Work units
// (m-tile mt, k-group kg of KBS / KG k-blocks) are dealt to the NCW warps: unit (mt, kg) -> warp (mt * KG + kg) % NCW.
// KG (a power of two dividing KBS and NCW) is the smallest split that deals the units evenly (no guards in the loop).
constexpr int pick_kg(int mt, int kbs) {
  int kg = 1;
  while ((mt % (NCW / kg)) != 0 && kg * 2 <= kbs && kg < NCW) kg *= 2;
  return kg;
}
template <int MT_, int KB_>
struct Mat {
  static constexpr int MT = MT_, KB = KB_;
  static constexpr int KBS = (ASR_STAGE_KB / MT) < 1 ? 1 : ((ASR_STAGE_KB / MT) > KB ? KB : (ASR_STAGE_KB / MT));
  static constexpr int KG = pick_kg(MT, KBS);
  static constexpr int MSTEP = NCW / KG;                               // m-tile stride between a warp's units
  static constexpr int UPW = (MT + MSTEP - 1) / MSTEP, KPG = KBS / KG, NST = KB / KBS;
  static constexpr bool EXACT = (MT % MSTEP) == 0;                     // every warp has exactly UPW units
  static constexpr uint32_t ST_BYTES = KBS * MT * 1024u;
  static_assert(MT >= 1 && MT <= 32 && KBS % KG == 0 && KB % KBS == 0 && NCW % KG == 0 && UPW <= 4 && MT * KG <= 32 &&
                    (KG == 1 || MT * KG <= 24) && EXACT, "tiling");
};

struct SmemMap {
  uint32_t ring, h, xn_hi, xn_lo, hid_hi, hid_lo, o_hi, o_lo, q, qf, kvrow, scratch, prm, lg, recv, arg, part, stat, red, tok,
      ctrl, ph, bars, total;
};
// Every buffer but the ring sits at a COMPILE-TIME offset (the ring, whose depth is chosen at launch, comes last): the
// kernel's shared-memory pointers are constants, not registers.
template <class S>
__host__ __device__ constexpr SmemMap smem_map(int nstages) {
  SmemMap m{};
  uint32_t off = 0;
  auto take = [&off](uint32_t bytes) {
    const uint32_t o = off;
    off += (bytes + 127u) & ~127u;
    return o;
  };
  m.h = take(S::GUP * S::D * 4);
  m.xn_hi = take(S::GUP * S::LDX);
  m.xn_lo = take(S::GUP * S::LDX);
  m.hid_hi = take(S::GUP * S::LDH);
  m.hid_lo = take(S::GUP * S::LDH);
  m.o_hi = take(S::GUP * S::LDO);
  m.o_lo = take(S::GUP * S::LDO);
  m.q = take(S::GUP * 64 * 4);                         // f16 hi [GUP][64] | f16 lo [GUP][64], fragment order
  m.qf = take(S::GUP * 64 * 4);                        // fp32 q of the self attention (score of the current key)
  m.kvrow = take(S::GUP * 128 * 2);
  // [KG * MT <= 24 when KG > 1][32 lanes] float4 partial tiles of the split-K matmuls (QKV, cross q, classifier): they
  // run while the FFN hidden rows and the attention output rows are dead, so the scratch shares their space when it fits
  // (8 utterance slots: exactly 12 KB) - shared memory not spent here is ring depth.
  constexpr uint32_t kScratch = 24 * 32 * 16;
  if (m.q - m.hid_hi >= kScratch) {
    m.scratch = m.hid_hi;
  } else {
    m.scratch = take(kScratch);
  }
  m.prm = take(S::SMALL_BYTES);
  m.lg = take(S::GUP * S::VS * 4);
  m.recv = take(S::CS * S::GUP * 64 * 4);              // [source rank][GUP][64] fp32 partial sums of this CTA's column slice
  m.arg = take(2 * S::CS * S::GUP * 8);                // [parity][source rank][GUP] (value, index)
  m.part = take(NCW * 64 * 4);
  m.stat = take(2 * NCW * 4);
  m.red = take(NCW * 8 * 8);                           // LayerNorm partials [warp][utterance slot] (sum, sum of squares)
  m.tok = take(2 * 8 * 4);                             // [0..7] next tokens, [8..15] finished flags
  m.ctrl = take(16);                                   // [0] steps done, [1] stop, [2] cache rows written
  m.ph = take(12 * 8);                                 // phase clocks (timed launches)
  m.bars = take((2 * MAX_STAGES + 4) * 8);
  off = (off + 1023u) & ~1023u;                        // swizzled TMA destinations: 1024-byte aligned stages
  m.ring = off;
  m.total = off + uint32_t(nstages) * STAGE_BYTES;
  return m;
}

struct Ring {
  uint8_t* buf;
  uint64_t* full;
  uint64_t* empty;
  int nstages;
};

// The producer is a whole WARP walking the static access sequence in lock-step: lane 0 waits for the free slot and posts
// the byte count, then the lanes issue the stage's copies IN PARALLEL (lane u = utterance slot u of a K/V stage).  One
// thread issuing the eight copies of a K/V stage one after the other (address arithmetic + TMA issue, ~100 cycles each
// at single-thread latency) took as long as the consumers need to eat the stage: the ring ran dry in every attention.
struct Producer {
  Ring r;
  int slot = 0;
  uint32_t round = 0;
  long long waited = 0;
#ifdef ASR_TRACE
  long long* trace = nullptr;   // diagnostic build: per stage (wait begin, armed) clocks of one step of CTA 0
  int ntrace = 0;
  int tag = 0;                  // what the stage holds (low 4 bits of the recorded wait-begin clock)
#endif
  __device__ __forceinline__ uint8_t* begin(uint32_t bytes) {
    if ((threadIdx.x & 31) == 0) {
      const long long w0 = clock64();
      mbar_wait(&r.empty[slot], (round & 1u) ^ 1u);
      waited += clock64() - w0;
      mbar_expect_tx(&r.full[slot], bytes);
#ifdef ASR_TRACE
      if (trace && ntrace < 256) {
        trace[2 * ntrace] = (w0 & ~15LL) | tag;
        trace[2 * ntrace + 1] = clock64();
        ++ntrace;
      }
#endif
    }
    __syncwarp();                            // the slot is free and armed before any lane's copy can land in it
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
      if ((threadIdx.x & 31) == 0) bulk_load(dst, src, M::ST_BYTES, bar(), pol);
      src += M::ST_BYTES;
      end();
    }
  }
};

struct Consumer {
  Ring r;
  int slot = 0;
  uint32_t round = 0;
  uint32_t okm = 0;        // bit i: the stage i slots ahead of the current one is known to be full
#ifdef ASR_TRACE
  long long* trace = nullptr;   // diagnostic build: per acquire (begin, end) clocks of one step of CTA 0 / warp 0
  int ntrace = 0;
#endif
#ifdef ASR_COUNT_LATE
  long long late = 0;      // acquires that found the stage not yet full
#endif
  __device__ __forceinline__ uint64_t* bar_ahead(int i, uint32_t& parity) const {
    int sl = slot + i;
    uint32_t rd = round;
    if (sl >= r.nstages) {
      sl -= r.nstages;
      ++rd;
    }
    parity = rd & 1u;
    return &r.full[sl];
  }
  // Non-blocking look at the stage i slots ahead (mbarrier.test_wait never suspends the warp).  A completed-phase query
  // still takes ~100 cycles to come back, and with two warps per scheduler nothing hides it: every use site issues
  // its peeks EARLY (before the math of the stage in hand) and tests the cached bit when it gets there.
  __device__ __forceinline__ void peek(int i) {
#if ASR_PEEK_MODE
    if (i < r.nstages && !((okm >> i) & 1u)) {
      uint32_t parity;
      uint64_t* b = bar_ahead(i, parity);
#if ASR_PEEK_MODE == 1
      okm |= uint32_t(mbar_test_wait(b, parity)) << i;
#else
      okm |= uint32_t(mbar_try_wait(b, parity)) << i;
#endif
    }
#endif
  }
  // wait for the stage i slots ahead of the current one (0 = current) without consuming it
  __device__ __forceinline__ const uint8_t* acquire_ahead(int i) {
    uint32_t parity;
    uint64_t* b = bar_ahead(i, parity);
    if (!((okm >> i) & 1u)) {
#ifdef ASR_TRACE
      const long long tr0 = clock64();
      if (!mbar_try_wait(b, parity)) mbar_wait_slow(smem_u32(b), parity);
      if (trace && ntrace < 256 && threadIdx.x == 0) {
        trace[2 * ntrace] = tr0;
        trace[2 * ntrace + 1] = clock64();
        ++ntrace;
      }
#elif defined(ASR_COUNT_LATE)
      if (!mbar_test_wait(b, parity)) {      // diagnostic build: how long do the consumers wait for DATA?
        const long long w0 = clock64();
        if (!mbar_try_wait(b, parity)) mbar_wait_slow(smem_u32(b), parity);
        late += clock64() - w0;
      }
#else
      if (!mbar_try_wait(b, parity)) mbar_wait_slow(smem_u32(b), parity);   // stage not there yet (rare): out of line
#endif
      okm |= 1u << i;
    }
    return r.buf + size_t(b - r.full) * STAGE_BYTES;
  }
  __device__ __forceinline__ const uint8_t* acquire() { return acquire_ahead(0); }
  __device__ __forceinline__ void release() {       // every consumer warp calls this once per stage
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&r.empty[slot]);
    okm >>= 1;
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
// lane: {W[2g][c..c+1], W[2g+1][c..c+1], W[2g][c+2..c+3], W[2g+1][c+2..c+3]} (rows within the m-tile), c = 32 kb + 8 tg
// + 4 s, i.e. the K permutation k_mma {2tg, 2tg+1, 2tg+8, 2tg+9} <-> columns {c..c+3} and the row permutation MMA row
// g -> 2g, g + 8 -> 2g + 1: a lane's result tile holds two ADJACENT output rows (paired stores in every epilogue).  The B fragment applies the same permutation:
// lane (g, tg) reads the 16 bytes x[u = g][32 kb + 8 tg .. +7] (f16 hi and lo copies; row stride == 64 mod 128 B:
// conflict free) and feeds halves s = 0 / 1 to the two MMAs.  Unit (m-tile mt, k-group kg) belongs to warp
// (mt * KG + kg) % 8.  The finished tile v = {(n = 16mt+2g, u = 2tg), (n, u+1), (n+1, u), (n+1, u+1)} goes to
// epi(n, u, v); with KG > 1 the k-group partials are first combined through `scratch`.
template <class M, int GUP, class Epi>
__device__ __forceinline__ void mm_stream(Consumer& c, const uint8_t* xhi, const uint8_t* xlo, int ldx, float4* scratch,
                                          Epi&& epi) {
  constexpr bool COLS = 2 * GUP <= 8;   // hi and lo parts of x ride in different N columns of ONE mma (u | GUP + u)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int kg = warp % M::KG, mt0 = warp / M::KG;          // unit j of this warp: m-tile mt0 + j * MSTEP
  float acc[M::UPW][2][4];
#pragma unroll
  for (int j = 0; j < M::UPW; ++j)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][s][i] = 0.f;
  const bool xrow = g < (COLS ? 2 * GUP : GUP);
  const uint8_t* xh = ((COLS && g >= GUP) ? xlo + (g - GUP) * ldx : xhi + g * ldx) + tg * 16 + kg * M::KPG * 64;
  const uint8_t* xl = xlo + g * ldx + tg * 16 + kg * M::KPG * 64;
  // Software pipeline over the stages of the matrix (fully unrolled, two register buffers): the fragments of stage
  // i + 1 are loaded BEFORE the MMAs of stage i are issued, so the shared-memory reads of the eight warps (the LSU pipe
  // needs 4 cycles per LDS.128) overlap the tensor-core work instead of alternating with it; a stage is released as
  // soon as its MMAs have been issued (their operands are in registers by then).
  uint4 af[2][M::KPG][M::UPW][2], bh[2][M::KPG], bl[2][M::KPG];
  auto load_stage = [&](int buf, int st_i, const uint8_t* base) {
    const uint8_t* stp = base + lane * 16 + (size_t(kg) * M::KPG * M::MT + mt0) * 1024;
#pragma unroll
    for (int q = 0; q < M::KPG; ++q) {
      bh[buf][q] = make_uint4(0, 0, 0, 0);
      bl[buf][q] = make_uint4(0, 0, 0, 0);
      if (xrow) {
        bh[buf][q] = lds128(xh + (st_i * M::KBS + q) * 64);
        if (!COLS) bl[buf][q] = lds128(xl + (st_i * M::KBS + q) * 64);
      }
#pragma unroll
      for (int j = 0; j < M::UPW; ++j) {
        const uint8_t* a = stp + (size_t(q) * M::MT + j * M::MSTEP) * 1024;
        af[buf][q][j][0] = lds128(a);
        af[buf][q][j][1] = lds128(a + 512);
      }
    }
  };
  load_stage(0, 0, c.acquire());
  c.peek(1);
#pragma unroll
  for (int st_i = 0; st_i < M::NST; ++st_i) {
    const int cur = st_i & 1;
    if (st_i + 1 < M::NST) load_stage(cur ^ 1, st_i + 1, c.acquire_ahead(1));
    c.peek(st_i + 1 < M::NST ? 2 : 1);   // the stage after next (or the next phase's first one), under this stage's MMAs
#pragma unroll
    for (int q = 0; q < M::KPG; ++q)
#pragma unroll
      for (int j = 0; j < M::UPW; ++j) {
        mma16816(acc[j][0], af[cur][q][j][0], bh[cur][q].x, bh[cur][q].y);
        mma16816(acc[j][1], af[cur][q][j][1], bh[cur][q].z, bh[cur][q].w);
        if (!COLS) {
          mma16816(acc[j][0], af[cur][q][j][0], bl[cur][q].x, bl[cur][q].y);
          mma16816(acc[j][1], af[cur][q][j][1], bl[cur][q].z, bl[cur][q].w);
        }
      }
    c.release();
  }
  float4 v[M::UPW];
#pragma unroll
  for (int j = 0; j < M::UPW; ++j) {
    v[j] = make_float4(acc[j][0][0] + acc[j][1][0], acc[j][0][1] + acc[j][1][1], acc[j][0][2] + acc[j][1][2],
                       acc[j][0][3] + acc[j][1][3]);
    if (COLS) {   // column u (hi) + column GUP + u (lo): lanes tg and tg ^ (GUP / 2)
      v[j].x += __shfl_xor_sync(0xffffffffu, v[j].x, GUP / 2);
      v[j].y += __shfl_xor_sync(0xffffffffu, v[j].y, GUP / 2);
      v[j].z += __shfl_xor_sync(0xffffffffu, v[j].z, GUP / 2);
      v[j].w += __shfl_xor_sync(0xffffffffu, v[j].w, GUP / 2);
    }
  }
  if (M::KG == 1) {
#pragma unroll
    for (int j = 0; j < M::UPW; ++j)
      if (2 * tg < GUP) epi(16 * (mt0 + j * M::MSTEP) + 2 * g, 2 * tg, v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < M::UPW; ++j) scratch[(kg * M::MT + mt0 + j * M::MSTEP) * 32 + lane] = v[j];
    consumer_sync();
    for (int it = threadIdx.x; it < M::MT * 32; it += NCT) {
      const int ln = it & 31, g2 = ln >> 2, tg2 = ln & 3;
      if (2 * tg2 < GUP) {
        float4 s = scratch[it];
#pragma unroll
        for (int k = 1; k < M::KG; ++k) {
          const float4 w = scratch[k * M::MT * 32 + it];
          s.x += w.x; s.y += w.y; s.z += w.z; s.w += w.w;
        }
        epi(16 * (it >> 5) + 2 * g2, 2 * tg2, s);
      }
    }
  }
}

// Column skew of the [utterance][D] fp32 rows (residual stream, all-reduce receive slots): row u holds column n at
// n ^ hswz(u).  Bit 4 separates the two utterances a warp of the all-reduce tail reads together; bits 3-4 spread the
// four even (and the four odd) utterance rows the matmul epilogue writes together.  Multiples of 8 keep float4 groups.
__device__ __forceinline__ int hswz(int u) { return ((u & 1) << 4) ^ (((u >> 1) & 3) << 3); }

// ------------------------------------------------------------------------------------------------ LayerNorm / split
// rows u < GU of h (fp32, stride D) -> f16 hi + lo rows (stride ld elements); warp u handles row u.
template <int D>
__device__ __forceinline__ void rows_to_hilo(const float* h, int GU, const float* gam, const float* bet, f16* hi,
                                             f16* lo, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < GU) {
    const float* src = h + warp * D;
    float x[D / 32];
#pragma unroll
    for (int i = 0; i < D / 32; ++i) x[i] = src[(lane + 32 * i) ^ hswz(warp)];
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
      const f16 hh = __float2half_rn(x[i]);
      hi[warp * ld + lane + 32 * i] = hh;
      lo[warp * ld + lane + 32 * i] = __float2half_rn(x[i] - __half2float(hh));
    }
  }
}

// LayerNorm of the rows u < GU by ALL consumer warps: warp (part, u) owns D / WPU elements of row u; one pass over
// x - x[0] (shifted sums: no cancellation), partial (sum, sum of squares) per warp combined through shared memory.
template <class S>
__device__ __forceinline__ void ln_rows(const float* h, int GU, const float* gam, const float* bet, f16* hi, f16* lo,
                                        float* red /* [NCW][2] */) {
  constexpr int D = S::D, WPU = S::WPU, EPL = D / (32 * WPU);
  static_assert(D % (32 * WPU) == 0, "LayerNorm split");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = warp % S::GUP, part = warp / S::GUP;
  const float* src = h + u * D;
  const int k0 = part * (D / WPU) + lane, hsw = hswz(u);
  const float c0 = src[0];
  float x[EPL], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    x[i] = src[(k0 + 32 * i) ^ hsw] - c0;
    s1 += x[i];
    s2 = fmaf(x[i], x[i], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) *reinterpret_cast<float2*>(red + 2 * warp) = make_float2(s1, s2);
  consumer_sync();
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int pI = 0; pI < WPU; ++pI) {
    const float2 r = *reinterpret_cast<const float2*>(red + 2 * (pI * S::GUP + u));
    t1 += r.x;
    t2 += r.y;
  }
  const float mean = t1 * (1.0f / float(D));
  const float rstd = 1.0f / sqrtf(fmaxf(t2 * (1.0f / float(D)) - mean * mean, 0.f) + 1e-5f);
  if (u < GU) {
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int k = k0 + 32 * i;
      const float y = (x[i] - mean) * rstd * gam[k] + bet[k];
      const f16 hh = __float2half_rn(y);
      hi[u * (D + 32) + k] = hh;
      lo[u * (D + 32) + k] = __float2half_rn(y - __half2float(hh));
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// Single-query attention of ONE head on the tensor cores in blocks of 32 keys, flash style (log2 units):
//   S = K q   : A = K [16 keys x 16 dims] by ldmatrix.x4 (non-transposed) from the 128-byte-swizzled K rows TMA wrote,
//               B = q with even MMA columns = f16 hi part, odd columns = lo part (replicated): 2 key tiles x 4 dim tiles
//               = 8 MMAs per block.  Lane (g, tg) ends up with the scores of keys 16 mi + g and 16 mi + 8 + g (hi
//               column + lo column added in the lane), replicated over tg: 4 scores per lane and block.  (Round 2 first
//               ran the transposed form, A = q as two distinct rows of 16 and B = K^T: 16 MMAs and 8 scores per lane,
//               no lane exchange before P V; the legacy MMA issue rate of ~10 cycles per scheduler made that the bound
//               of the K stages - 15.7 -> 15.4 ms per 256-utterance launch for this form, same tokens.)
//   o += V^T p: A = V^T [16 dims x 16 keys] by ldmatrix.x4.trans of the row-major V rows, B = p with even columns = f16
//               hi parts, odd columns = lo parts: for k-tile kt lane (g, tg) must supply keys 16kt + 2tg (+1) and
//               16kt + 8 + 2tg (+1), which live in lanes (2tg, *) and (2tg + 1, *): 2 shuffles + 2 byte permutes per
//               k-tile (p_pack_ka).  o(dim) = c(even col) + c(odd col), replicated over tg.
// Running max m is warp-uniform; the running sum l is a per-lane partial over the lane's own keys (replicated over tg).
struct AttnT {
  float m, l;
  float o[4][4];
};
// q of one utterance: 64 half2 words [tg 4][k-tile 4][hi pair 0, lo pair 0, hi pair 1, lo pair 1] with pair 0 = dims
// 16 kt + 2 tg (+1), pair 1 = dims 16 kt + 8 + 2 tg (+1): lane (g, tg) reads the word group (tg, kt) with one LDS.128 (all g
// read the same words: broadcast) and keeps the hi or the lo halves (attn_q_bfrags).
__device__ __forceinline__ int q_word_index(int d) {
  const int r = d & 15;
  return ((((r & 7) >> 1) * 4 + (d >> 4)) * 4) + (r >> 3) * 2;
}
// f16 hi | lo split of a pair of values
__device__ __forceinline__ void hilo2(float a, float b, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(a, b);
  const float2 f = __half22float2(hi);
  lo = __floats2half2_rn(a - f.x, b - f.y);
}
__device__ __forceinline__ void q_store2(__half2* q, float* qf, int u, int d, float y0, float y1) {   // dims d (even), d+1
  __half2 hi, lo;
  hilo2(y0, y1, hi, lo);
  const int i = u * 64 + q_word_index(d);
  q[i] = hi;
  q[i + 1] = lo;
  if (qf) *reinterpret_cast<float2*>(qf + u * 64 + d) = make_float2(y0, y1);   // fp32 copy: score of the current key
}
// ldmatrix with a "memory" clobber but NOT volatile: ordered against the ring's acquire / release (which clobber
// memory) and among themselves, while the register-only MMAs are free to move between them.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint4& r) {
  asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint4& r) {
  asm("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
}
// B fragments of q: lane (g, tg) holds, for every k-tile, dims 16 kt + 2 tg (+1) and 16 kt + 8 + 2 tg (+1) of column g =
// the hi part (even g) or the lo part (odd g)
__device__ __forceinline__ void attn_q_bfrags(const __half2* q, int u, uint2 (&qb)[4]) {
  const int lane = threadIdx.x & 31, tg = lane & 3;
  const bool odd = (lane >> 2) & 1;
  const uint4* src = reinterpret_cast<const uint4*>(q + u * 64 + tg * 16);   // {hi pair 0, lo pair 0, hi pair 1, lo pair 1}
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const uint4 v = src[kt];
    qb[kt] = make_uint2(odd ? v.y : v.x, odd ? v.w : v.z);
  }
}
// A fragments of one K tile (32 rows of 128 B, chunk c of row r stored at chunk c ^ (r & 7)): matrices of an ldmatrix.x4 =
// keys 0-7 | 8-15 of dims 0-7, then of dims 8-15 of the (key tile mi, dim tile kt)
__device__ __forceinline__ void ka_load(uint32_t kbase, uint4 (&ka)[2][4]) {
  const int lane = threadIdx.x & 31, mat = lane >> 3, r = lane & 7;
  const uint32_t row = kbase + ((mat & 1) * 8 + r) * 128;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) ldsm_x4(row + mi * 2048 + (((2 * kt + (mat >> 1)) ^ r) << 4), ka[mi][kt]);
}
// s[2 mi + e] <-> key 16 mi + 8 e + g; keys >= n_valid get -inf (select, not arithmetic: rows past the valid keys may
// hold anything).  One accumulator per key tile, the four dim tiles chained through it.
__device__ __forceinline__ void ka_math(const uint4 (&ka)[2][4], int n_valid, const uint2 (&qb)[4], float (&s)[4]) {
  const int g = (threadIdx.x & 31) >> 2;
  float cc[2][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int i = 0; i < 4; ++i) cc[mi][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) mma16816(cc[mi], ka[mi][kt], qb[kt].x, qb[kt].y);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    s[2 * mi] = cc[mi][0] + cc[mi][1];                           // key 16 mi + g      (hi column + lo column)
    s[2 * mi + 1] = cc[mi][2] + cc[mi][3];                       // key 16 mi + 8 + g
  }
  if (n_valid < 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (16 * (i >> 1) + 8 * (i & 1) + g >= n_valid) s[i] = -INFINITY;
  }
}
// B fragments of the probabilities for o += V^T p: k-tile kt needs, in lane (g, tg), {keys 16 kt + 2 tg (+1), keys
// 16 kt + 8 + 2 tg (+1)} as f16 hi parts (even g) or lo parts (odd g); the lane owns keys 16 kt + g and 16 kt + 8 + g.
// Every lane publishes ONE word per k-tile, (part of key g, part of key g + 8) with part = hi in the replicas tg even, lo
// in the replicas tg odd, and the reader takes it from the replica of its own parity: 2 shuffles + 2 byte permutes per
// k-tile.
__device__ __forceinline__ void p_pack_ka(const float (&p)[4], uint32_t (&pb)[4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int src = 8 * tg + (g & 1);                              // lane (2 tg, g & 1); + 4: lane (2 tg + 1, g & 1)
#pragma unroll
  for (int kt = 0; kt < 2; ++kt) {
    __half2 hi, lo;
    hilo2(p[2 * kt], p[2 * kt + 1], hi, lo);
    const __half2 mine = (tg & 1) ? lo : hi;
    const uint32_t r = *reinterpret_cast<const uint32_t*>(&mine);
    const uint32_t x = __shfl_sync(0xffffffffu, r, src), y = __shfl_sync(0xffffffffu, r, src + 4);
    pb[2 * kt] = __byte_perm(x, y, 0x5410);                      // keys 16 kt + 2 tg, + 1
    pb[2 * kt + 1] = __byte_perm(x, y, 0x7632);                  // keys 16 kt + 8 + 2 tg, + 1
  }
}
// exp2 on the SFU (ex2.approx.ftz: 2 ulp, -inf -> +0); the library exp2f costs three more instructions per value
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// o += V^T p for one value tile (32 rows of 128 B, swizzled like K)
__device__ __forceinline__ void pv_load(uint32_t vbase, uint4 (&vf)[4][2]) {
  const int lane = threadIdx.x & 31, mat = lane >> 3, r = lane & 7;
  const uint32_t row = vbase + ((mat >> 1) * 8 + r) * 128;       // matrices 0, 1: keys 16kt + r; 2, 3: keys 16kt + 8 + r
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const uint32_t ch = (((2 * mt + (mat & 1)) ^ r) & 7) << 4;   // matrices 0, 2: dims 16mt .. +7; 1, 3: dims 16mt + 8 ..
    ldsm_x4_trans(row + ch, vf[mt][0]);
    ldsm_x4_trans(row + 2048 + ch, vf[mt][1]);
  }
}
__device__ __forceinline__ void pv_math(const uint4 (&vf)[4][2], AttnT& st, const uint32_t (&pb)[4]) {
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    mma16816(st.o[mt], vf[mt][0], pb[0], pb[1]);
    mma16816(st.o[mt], vf[mt][1], pb[2], pb[3]);
  }
}
// State before the first streamed key: empty, or (self attention, the warp that owns it) the CURRENT key k_t / v_t taken
// from shared memory: score = q . k_t in fp32 (q: fp32 copy; k_t | v_t: the f16 rows about to be appended, their 16-byte
// chunks swizzled by t & 7), p = 1, o = v_t.
__device__ __forceinline__ void attn_init(AttnT& st, const float* qf, const f16* kv_row_u, int swz) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  st.m = -INFINITY;
  st.l = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) st.o[i][j] = 0.f;
  if (qf) {   // warp-uniform
    const float2 q2 = *reinterpret_cast<const float2*>(qf + 2 * lane);
    const int d = 2 * lane, pos = (((d >> 3) ^ swz) << 3) + (d & 7);
    const float2 k2 = __half22float2(*reinterpret_cast<const __half2*>(kv_row_u + pos));
    st.m = warp_sum(fmaf(q2.x, k2.x, q2.y * k2.y));
    st.l = g == 0 ? 1.f : 0.f;     // per-lane partial sums are replicated over tg, distinct over g
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {                             // dims 16mt + g, 16mt + g + 8: chunks 2mt, 2mt + 1
      st.o[mt][0] = __half2float(kv_row_u[64 + (((2 * mt) ^ swz) << 3) + g]);
      st.o[mt][2] = __half2float(kv_row_u[64 + (((2 * mt + 1) ^ swz) << 3) + g]);
    }
  }
}
typedef uint2 QFrag;     // B fragment of q for one k-tile
constexpr int SPB = 4;   // scores per lane and 32-key block
// Attention of this warp's (utterance slot au, key partition apart) over n_keys rows streamed through the ring as
// super-chunks of up to SC K stages followed by the matching V stages; a stage holds, per utterance slot, WPU blocks of
// 32 keys ([slot][RPS rows][128 B]) and warp (au, apart) owns block apart of every stage.  All scores of a super-chunk are computed first, then ONE max / exp / sum, then all P V products.  n_keys rows
// are streamed (uniform over the CTA); only the first n_mine are valid keys of THIS warp's utterance (key padding), the
// rest are masked.  Stage presence (s < ns) is CTA-uniform; blocks without a valid key skip their math only.
template <class S, int SC>
__device__ __forceinline__ void attention(Consumer& c, AttnT& st, const QFrag (&qa)[4], int n_keys, int n_mine,
                                          bool active, int au, int apart) {
  constexpr int RPS = S::RPS;
  const uint32_t blk_off = (au * S::WPU + apart) * 4096;
  const int lane = threadIdx.x & 31, tg = lane & 3;
  int c0 = 0;
#pragma unroll 1
  while (c0 < n_keys) {
    const int nk = min(SC * RPS, n_keys - c0);                  // CTA-uniform
    const int ns = (nk + RPS - 1) / RPS;
    const int lim = active ? n_mine - c0 - 32 * apart : 0;      // my valid keys counted from my block of stage 0
    float sc[SC][SPB];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < SC; ++s) {
#pragma unroll
      for (int i = 0; i < SPB; ++i) sc[s][i] = -INFINITY;
      if (s < ns) {
        const uint32_t stg = smem_u32(c.acquire()) + blk_off;
        const int n = lim - s * RPS;
        if (n > 0) {
          // The slot goes back to the producer as soon as its rows are requested: the arrive is a release, so the
          // ldmatrix reads before it are ordered before the producer's next copy into the slot; the math then runs
          // from registers while the refill is under way (64 / 128-utterance launches -1.8 %, 256 unchanged).
          uint4 ka[2][4];
          ka_load(stg, ka);
          c.release();
          ka_math(ka, n, qa, sc[s]);
#pragma unroll
          for (int i = 0; i < SPB; ++i) mx = fmaxf(mx, sc[s][i]);
        } else {
          c.release();
        }
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));        // the eight g lanes own different keys
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    const float m_new = fmaxf(st.m, mx);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;     // a warp without any key: every p = ex2(-inf) = 0
    const float alpha = ex2(st.m - m_use);
    float lsum = 0.f;
    uint32_t pb[SC][4];
#pragma unroll
    for (int s = 0; s < SC; ++s)
      if (s < ns) {
        float pr[SPB];
#pragma unroll
        for (int i = 0; i < SPB; ++i) {
          pr[i] = ex2(sc[s][i] - m_use);
          lsum += pr[i];
        }
        p_pack_ka(pr, pb[s]);
      }
    st.l = st.l * alpha + lsum;
    st.m = m_new;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) st.o[i][j] *= alpha;
#pragma unroll
    for (int s = 0; s < SC; ++s)
      if (s < ns) {
        const uint32_t stg = smem_u32(c.acquire()) + blk_off;
        if (lim - s * RPS > 0) {
          uint4 vf[4][2];
          pv_load(stg, vf);
          c.release();
          pv_math(vf, st, pb[s]);
        } else {
          c.release();
        }
      }
    c0 += SC * RPS;
  }
  (void)tg;
}
// merge the key partitions of every utterance slot and emit o (f16 hi + lo rows, stride 96 elements)
template <class S>
__device__ __forceinline__ void attn_finish(AttnT& st, int GU, float* part_buf, float* stat, f16* o_hi, f16* o_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  float l = st.l;                                               // per-lane partial over the lane's keys
  l += __shfl_xor_sync(0xffffffffu, l, 4);                      // (replicated over tg, summed over g)
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
  if (S::WPU == 1) {   // one warp holds the whole utterance: no merge; lane (g, tg) emits dims 16 tg + g, 16 tg + g + 8
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
      if (tg == mt) {
        a = st.o[mt][0] + st.o[mt][1];
        b = st.o[mt][2] + st.o[mt][3];
      }
    if (warp < GU) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      const float y0 = a * inv, y1 = b * inv;
      const f16 h0 = __float2half_rn(y0), h1 = __float2half_rn(y1);
      const int i = warp * 96 + 16 * tg + g;
      o_hi[i] = h0;
      o_lo[i] = __float2half_rn(y0 - __half2float(h0));
      o_hi[i + 8] = h1;
      o_lo[i + 8] = __float2half_rn(y1 - __half2float(h1));
    }
    consumer_sync();
    return;
  }
  if (tg == 0) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      part_buf[warp * 64 + 16 * mt + g] = st.o[mt][0] + st.o[mt][1];
      part_buf[warp * 64 + 16 * mt + g + 8] = st.o[mt][2] + st.o[mt][3];
    }
  }
  if (lane == 0) {
    stat[warp] = st.m;
    stat[NCW + warp] = l;
  }
  consumer_sync();
  for (int d = threadIdx.x; d < GU * 64; d += NCT) {
    const int uu = d >> 6, dd = d & 63;
    float mm = -INFINITY;
#pragma unroll
    for (int pI = 0; pI < S::WPU; ++pI) mm = fmaxf(mm, stat[uu * S::WPU + pI]);
    float t = 0.f, ls = 0.f;
#pragma unroll
    for (int pI = 0; pI < S::WPU; ++pI) {
      const float mw = stat[uu * S::WPU + pI];
      const float f = (mw == -INFINITY) ? 0.f : ex2(mw - mm);
      t += part_buf[(uu * S::WPU + pI) * 64 + dd] * f;
      ls += stat[NCW + uu * S::WPU + pI] * f;
    }
    const float y = ls > 0.f ? t / ls : 0.f;
    const f16 hh = __float2half_rn(y);
    o_hi[uu * 96 + dd] = hh;
    o_lo[uu * 96 + dd] = __float2half_rn(y - __half2float(hh));
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ the kernel
// Registers: 9 warps put three on one scheduler, so ptxas stops at 168 per thread; the bodies need 151 (no spills).
// (setmaxnreg with a donor warpgroup was tried: ptxas drops the pair unless the producer branch fits 40 registers.)
template <class S>
__global__ void __launch_bounds__(NTHREADS, 1)
dec_cluster_kernel(const __grid_constant__ ClusterParams p, const __grid_constant__ CUtensorMap ckv_map,
                   const __grid_constant__ CUtensorMap cache_map) {
  constexpr int D = S::D, H = S::H, CS = S::CS, GUP = S::GUP, FFS = S::FFS, VS = S::VS, RPS = S::RPS;
  using MQkv = Mat<12, D / 32>;
  using MWo = Mat<D / 16, 2>;
  using MWqc = Mat<4, D / 32>;
  using MW1 = Mat<FFS / 16, D / 32>;
  using MW2 = Mat<D / 16, FFS / 32>;
  using MCls = Mat<VS / 16, D / 32>;

  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr SmemMap sm = smem_map<S>(0);       // offsets do not depend on the ring depth
  Ring ring;
  ring.buf = smem + sm.ring;
  ring.nstages = p.nstages;
  float* s_h = reinterpret_cast<float*>(smem + sm.h);
  f16* xn_hi = reinterpret_cast<f16*>(smem + sm.xn_hi);
  f16* xn_lo = reinterpret_cast<f16*>(smem + sm.xn_lo);
  f16* hid_hi = reinterpret_cast<f16*>(smem + sm.hid_hi);
  f16* hid_lo = reinterpret_cast<f16*>(smem + sm.hid_lo);
  f16* o_hi = reinterpret_cast<f16*>(smem + sm.o_hi);
  f16* o_lo = reinterpret_cast<f16*>(smem + sm.o_lo);
  __half2* q_frag = reinterpret_cast<__half2*>(smem + sm.q);   // [GUP][64] hi | lo words in A-fragment order
  f16* kv_row = reinterpret_cast<f16*>(smem + sm.kvrow);
  float4* scratch = reinterpret_cast<float4*>(smem + sm.scratch);
  float* prm = reinterpret_cast<float*>(smem + sm.prm);
  float* s_lg = reinterpret_cast<float*>(smem + sm.lg);
  float* recv = reinterpret_cast<float*>(smem + sm.recv);
  float2* arg = reinterpret_cast<float2*>(smem + sm.arg);
  float* part_buf = reinterpret_cast<float*>(smem + sm.part);
  float* stat = reinterpret_cast<float*>(smem + sm.stat);
  float2* red = reinterpret_cast<float2*>(smem + sm.red);
  int* s_tok = reinterpret_cast<int*>(smem + sm.tok);
  volatile int* ctrl = reinterpret_cast<volatile int*>(smem + sm.ctrl);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  ring.full = bars;
  ring.empty = bars + MAX_STAGES;
  uint64_t* xbar = bars + 2 * MAX_STAGES;       // all-reduce: [0] reduce-scatter, [1] all-gather
  uint64_t* abar = xbar + 2;                    // [2] argmax exchange parity barriers

  const int rank = int(cluster_ctarank());
  const int cluster_id = blockIdx.x / CS;
  const int ubase = cluster_id * p.GU;
  const int GU = min(p.GU, p.B - ubase);         // utterances of this cluster (>= 1 by construction of the grid)
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();        // swizzled TMA destinations need 1024-byte alignment
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
  for (uint32_t i = sm.h / 4 + tid; i < sm.prm / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid < 16) s_tok[tid] = 0;
  __syncthreads();
  cluster_sync_all();   // peers' barriers are initialised before any remote store can arrive

  const int Lc = (p.L + 31) & ~31;                                // cache rows per (layer, utterance, head, K|V): 32-key blocks
  const size_t cache_head = size_t(2) * Lc * 64;                  // elements per (layer, utterance, head): K rows | V rows
  const uint8_t* my_image = p.image + size_t(rank) * p.rank_bytes;

  if (warp >= NCW) {
    // =============================== producer warp (see struct Producer)
    {
      const int plane = tid & 31;
      const uint64_t pol_w = policy_evict_last();
      const uint64_t pol_kv = p.kv_evict_first ? policy_evict_first() : policy_evict_last();
      Producer pr;
      pr.r = ring;
      // Encoder K/V comes from HBM (392 MB per step over the GPU: it cannot stay in L2), and a stage requested when its
      // ring slot frees needs ~3 k cycles to land - more than the 4 stages of lead the ring gives in the attention
      // phases.  So every super-chunk is prefetched into L2 one super-chunk ahead; the ring's own copies then hit L2.  The
      // rows of an utterance are contiguous over all heads ([Tp][K | V, 2D]), so the cluster's CTAs split the rows of the
      // super-chunk and each lane u issues ONE bulk prefetch (rows x 4D bytes) for utterance slot u.
      auto kv_prefetch = [&](int l, int c0) {
        if (c0 >= p.Tp || plane >= GU) return;
        const int nk = min(S::SCX * RPS, p.Tp - c0), per = (nk + CS - 1) / CS;
        const int r0 = rank * per, n = min(per, nk - r0);
        if (n > 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.ckv + (size_t(l * p.B + ubase + plane) * p.Tp +
                                                                               c0 + r0) * 2 * D),
                       "r"(uint32_t(n) * 4u * D)
                       : "memory");
      };
      // The self K/V cache (201 MB at 256 utterances) does not stay in L2 either: the first self-attention stage of a layer
      // used to arrive ~1.7 k cycles late.  Its rows are prefetched one layer ahead (lane = utterance slot x (K | V), one
      // contiguous block of round32(t) rows each), issued where the producer is normally waiting for a free slot.
      auto self_prefetch = [&](int l2, int t2) {
        const int u = plane >> 1, n = (t2 + 31) & ~31;
        if (t2 > 0 && plane < 2 * GU)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.cache + ((size_t(l2) * p.B + ubase + u) * H + rank) *
                                                                                  cache_head + size_t(plane & 1) * Lc * 64),
                       "r"(uint32_t(n) * 128u)
                       : "memory");
      };
#pragma unroll 1
      for (int t = 0; t < p.L; ++t) {
#ifdef ASR_TRACE
        pr.trace = (p.timing && blockIdx.x == 0 && t == 70) ? p.timing + 148 * 16 + 512 : nullptr;
#endif
        if (p.stop_at_eos) {                       // strict gate: nothing of step t is requested before step t-1 ended
          while (ctrl[0] < t && !ctrl[1]) {
          }
          if (ctrl[1]) break;
        }
#pragma unroll 1
        for (int l = 0; l < p.nd; ++l) {
          const uint8_t* img = my_image + size_t(l) * p.layer_bytes;
#ifdef ASR_TRACE
          pr.tag = 1;
#endif
          {
            uint8_t* dst = pr.begin(S::SMALL_BYTES);
            if (plane == 0) bulk_load(dst, img + p.off_small, S::SMALL_BYTES, pr.bar(), pol_w);
            pr.end();
          }
          if (p.kv_prefetch) kv_prefetch(l, 0);
#ifdef ASR_TRACE
          pr.tag = 2;
#endif
          pr.mat<MQkv>(img + p.off_qkv, pol_w);
          if (t > 0) {   // self cache rows 0..t-1 of this layer (written by this CTA in earlier steps)
            const int need = (t - 1) * p.nd + l + 1;
            while (ctrl[2] < need) {
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            // lane u: K rows | V rows of utterance slot u, head `rank`
            const f16* cbase = p.cache + ((size_t(l) * p.B + ubase + (plane < GU ? plane : 0)) * H + rank) * cache_head;
            for (int c0 = 0; c0 < t; c0 += S::SCX * RPS) {
              const int nk = min(S::SCX * RPS, t - c0);
              for (int kv = 0; kv < 2; ++kv)               // K stages of the super-chunk, then its V stages
                for (int r0 = 0; r0 < nk; r0 += RPS) {
                  const int n = (min(RPS, nk - r0) + 31) & ~31;   // whole 32-key blocks (rows past t are zero)
#ifdef ASR_TRACE
                  pr.tag = 3 + kv;
#endif
                  if constexpr (RPS == 32) {   // 8 utterance slots of 32 rows: ONE 4-D box [slot][1][32 rows][64] per stage
                    uint8_t* dst = pr.begin(uint32_t(GUP) * RPS * 128u);
                    if (plane == 0)
                      tma_load_4d_hint(dst, &cache_map, pr.bar(), 0, c0 + r0, rank * 2 + kv, l * p.B + ubase, pol_kv);
                    pr.end();
                    (void)n;
                  } else {                     // (fewer, longer slots: a fixed box would copy rows nobody needs)
                    uint8_t* dst = pr.begin(uint32_t(GU) * n * 128u);
                    if (plane < GU)
                      bulk_load(dst + plane * RPS * 128, cbase + size_t(kv) * Lc * 64 + size_t(c0 + r0) * 64, n * 128,
                                pr.bar(), pol_kv);
                    pr.end();
                  }
                }
            }
          }
#ifdef ASR_TRACE
          pr.tag = 5;
#endif
          pr.mat<MWo>(img + p.off_wo, pol_w);
#ifdef ASR_TRACE
          pr.tag = 6;
#endif
          pr.mat<MWqc>(img + p.off_wqc, pol_w);
          {
            // ONE 3-D box per stage: [GUP utterance slots][RPS rows][64 columns of this head] (eight per-utterance 2-D
            // copies issued by eight lanes cost the producer ~700 cycles per stage - TMA issue is serial within a warp -
            // against the ~500 the consumers need: the ring ran dry in the second half of every cross attention).  Rows
            // past Tp and utterance slots past the batch are out of bounds: zero-filled, counted in the byte count.
            for (int c0 = 0; c0 < p.Tp; c0 += S::SCX * RPS) {
              const int nk = min(S::SCX * RPS, p.Tp - c0);
              if (p.kv_prefetch) kv_prefetch(l, c0 + S::SCX * RPS);   // the NEXT super-chunk: on its way to L2 meanwhile
              for (int kv = 0; kv < 2; ++kv)
                for (int r0 = 0; r0 < nk; r0 += RPS) {
#ifdef ASR_TRACE
                  pr.tag = 7 + kv;
#endif
                  uint8_t* dst = pr.begin(uint32_t(GUP) * RPS * 128u);
                  if (plane == 0)
                    tma_load_3d_hint(dst, &ckv_map, pr.bar(), kv * D + rank * 64, c0 + r0, l * p.B + ubase, pol_kv);
                  pr.end();
                }
            }
          }
#ifdef ASR_TRACE
          pr.tag = 9;
#endif
          if (p.kv_prefetch) {
            if (l + 1 < p.nd) self_prefetch(l + 1, t);
            else if (t + 1 < p.L) self_prefetch(0, t + 1);
          }
          pr.mat<MWo>(img + p.off_woc, pol_w);
#ifdef ASR_TRACE
          pr.tag = 10;
#endif
          pr.mat<MW1>(img + p.off_w1, pol_w);
#ifdef ASR_TRACE
          pr.tag = 11;
#endif
          pr.mat<MW2>(img + p.off_w2, pol_w);
        }
#ifdef ASR_TRACE
          pr.tag = 12;
#endif
        pr.mat<MCls>(my_image + p.off_cls, pol_w);
      }
      if (p.timing && plane == 0) {
        p.timing[size_t(blockIdx.x) * 16 + 3] = pr.waited;
        p.timing[size_t(blockIdx.x) * 16 + 4] = (long long)pr.round * p.nstages + pr.slot;
      }
    }
    __syncwarp();
  } else {
    // ================================= consumers
    Consumer c;
    c.r = ring;
    const bool timed = p.timing != nullptr;
    const long long t_begin = clock64();
    long long t_xchg = 0;
    uint32_t n_xchg = 0, n_arg = 0;
    const int lane = tid & 31;
    const int au = warp / S::WPU, apart = warp % S::WPU;            // attention: utterance slot / 32-key block of a stage
    float* q_f32 = reinterpret_cast<float*>(smem + sm.qf);
    const bool a_active = au < GU;
    // key-padding mask of the cross attention: encoder frames >= enc_lens[utterance] are not attended (nullable)
    const int n_cross = (p.enc_lens && a_active) ? max(0, min(p.Tp, p.enc_lens[ubase + au])) : p.Tp;
    const float qscale = p.scale * LOG2E;
    const uint8_t* xh = reinterpret_cast<const uint8_t*>(xn_hi);
    const uint8_t* xl = reinterpret_cast<const uint8_t*>(xn_lo);

    // embedding + PE of the first token (host-side init kernel)
    for (int i = tid * 4; i < GU * D; i += NCT * 4)         // (row u holds column n at n ^ hswz(u))
      *reinterpret_cast<float4*>(s_h + (i ^ hswz(i / D))) =
          *reinterpret_cast<const float4*>(p.h0 + size_t(ubase) * D + i);
    consumer_sync();

    // Partial sums over the full model dimension (this CTA's K-slice) -> all-reduce across the cluster, as a REDUCE-SCATTER
    // followed by an ALL-GATHER: h[u][n] += sum over ranks + bias[n].
    //   1. (matmul epilogue, send_partial) every partial tile goes straight to the CTA that owns its 64 columns (rank
    //      n / 64), into that CTA's receive slot of the sender: recv[src][u][64] fp32, 8-byte st.async + complete_tx on the
    //      owner's barrier xbar[0];
    //   2. the owner adds the CS slots in rank order (deterministic), the residual and the bias for its GUP x 64 slice and
    //      writes the new residual values into EVERY CTA's copy of the residual stream s_h (st.async, barrier xbar[1]);
    //   3. every CTA normalises the full rows locally (all_reduce_finish tail).
    // Against all-gathering the partials (every CTA summing everything) this moves 2 x 6 KB instead of 24 KB into each
    // CTA at 8 utterances, sums a quarter of the elements, and needs an 8 KB receive buffer instead of 64 KB - shared
    // memory that is ring depth now (3 -> 5 stages at 8 utterances per cluster).  Neither buffer is double-buffered: a CTA
    // can only send step n + 1 after it has completed step n, i.e. after it holds the step-n slice of every peer, and a
    // peer sends its slice only after it has finished reading its receive slots (and every CTA's s_h readers of step n
    // are done before that CTA sends its partials of step n + 1, without which nobody can produce a step n + 1 slice).
    // Both barriers simply alternate phases.  Windows of every rank (own included: the data path is uniform):
    // (the peers' windows are mapped where they are used: `mapa` is one instruction, twelve mapped addresses held for the
    // whole kernel are twelve registers of a kernel that sits at its register cap)
    // Residual stream s_h ([utterance][D]) and the receive slots ([source rank][utterance][64]) are fp32 rows with skewed
    // columns (hswz).
    auto send_partial = [&](int n, int u0, const float4& v) {     // v = {(n,u0), (n,u0+1), (n+1,u0), (n+1,u0+1)}
      const int dst = n >> 6, nn = n & 63;
      const uint32_t off0 = uint32_t((rank * GUP + u0) * 64 + (nn ^ hswz(u0))) * 4u;
      const uint32_t off1 = uint32_t((rank * GUP + u0 + 1) * 64 + (nn ^ hswz(u0 + 1))) * 4u;
      const uint32_t base = mapa_u32(smem_u32(recv), uint32_t(dst)), bar = mapa_u32(smem_u32(&xbar[0]), uint32_t(dst));
      st_async_v2(base + off0, v.x, v.z, bar);
      st_async_v2(base + off1, v.y, v.w, bar);
    };
    // ... and what follows every all-reduce: LayerNorm of the new rows (gam != nullptr) or the plain f16 hi | lo split
    // (classifier input).  Elements are dealt to the warps as blocks of 2 utterances x 16 column pairs: lane -> utterance
    // 2 up + (lane & 1), columns n, n + 1 with n = 32 j + 2 (lane >> 1); a warp always serves the same utterance pair up =
    // warp % (GUP / 2).  Every shared-memory access is then a conflict-free 8-byte (fp32 pair) or 4-byte (f16 pair) word
    // per lane.  Variance from one pass over (x - c), c = an element of the row.
    auto all_reduce_finish = [&](const float* bias, const float* gam, const float* bet) {
      constexpr int UP = GUP / 2, NB = UP * (D / 32), BPW = (NB + NCW - 1) / NCW;
      constexpr int NBS = UP * 2;                               // blocks of this CTA's 64-column slice (<= NCW)
      static_assert(NCW % UP == 0 && (NB % NCW == 0 || NB < NCW) && NBS <= NCW, "all-reduce block split");
      const uint32_t phase = n_xchg & 1u;
      const int up = warp % UP, u = 2 * up + (lane & 1), sw = hswz(u);
      float* hrow = s_h + u * D;
      if (tid == 0) {
        mbar_expect_tx(&xbar[0], uint32_t(CS) * 64 * GUP * 4u);
        mbar_expect_tx(&xbar[1], uint32_t(CS) * 64 * GUP * 4u);
      }
      if (timed) {
        const long long w0 = clock64();
        mbar_wait_cluster(&xbar[0], phase);
        t_xchg += clock64() - w0;
      } else {
        mbar_wait_cluster(&xbar[0], phase);
      }
      // ---- reduce: this CTA's slice, one column pair per thread of the first NBS warps
      if (warp < NBS) {
        const int nn = (warp / UP) * 32 + (lane >> 1) * 2, pos = nn ^ sw, n = rank * 64 + nn;
        float2 sum = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < CS; ++r) {
          const float2 w = *reinterpret_cast<const float2*>(recv + (r * GUP + u) * 64 + pos);
          sum.x += w.x;
          sum.y += w.y;
        }
        const float2 b = *reinterpret_cast<const float2*>(bias + n);
        const float2 h = *reinterpret_cast<const float2*>(hrow + (n ^ sw));
        const float hx = h.x + (sum.x + b.x), hy = h.y + (sum.y + b.y);
        const uint32_t off = uint32_t(u * D + (n ^ sw)) * 4u;
#pragma unroll
        for (int r = 0; r < CS; ++r)
          st_async_v2(mapa_u32(smem_u32(s_h) + off, uint32_t(r)), hx, hy, mapa_u32(smem_u32(&xbar[1]), uint32_t(r)));
      }
      if (timed) {
        const long long w0 = clock64();
        mbar_wait_cluster(&xbar[1], phase);
        t_xchg += clock64() - w0;
      } else {
        mbar_wait_cluster(&xbar[1], phase);
      }
      // ---- the new rows are complete in s_h: normalise
      const float c0 = hrow[0];
      float2 hv[BPW];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < BPW; ++k) {
        const int blk = warp + NCW * k;
        hv[k] = make_float2(c0, c0);
        if (NB % NCW == 0 || blk < NB) {
          const int n = (blk / UP) * 32 + (lane >> 1) * 2;
          hv[k] = *reinterpret_cast<const float2*>(hrow + (n ^ sw));
          const float d0 = hv[k].x - c0, d1 = hv[k].y - c0;
          s1 += d0 + d1;
          s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
        }
      }
      float mean = 0.f, rstd = 1.f;
      if (gam) {
#pragma unroll
        for (int o = 16; o >= 2; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane < 2) red[warp * 2 + lane] = make_float2(s1, s2);
        consumer_sync();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < NCW / UP; ++w) {
          const float2 r = red[(w * UP + up) * 2 + (lane & 1)];
          t1 += r.x;
          t2 += r.y;
        }
        mean = t1 * (1.0f / float(D));
        rstd = 1.0f / sqrtf(fmaxf(t2 * (1.0f / float(D)) - mean * mean, 0.f) + 1e-5f);
        mean += c0;
      }
      if (u < GU) {
#pragma unroll
        for (int k = 0; k < BPW; ++k) {
          const int blk = warp + NCW * k;
          if (NB % NCW == 0 || blk < NB) {
            const int n = (blk / UP) * 32 + (lane >> 1) * 2;
            float y0 = hv[k].x, y1 = hv[k].y;
            if (gam) {
              const float2 gg = *reinterpret_cast<const float2*>(gam + n), bb = *reinterpret_cast<const float2*>(bet + n);
              y0 = (y0 - mean) * rstd * gg.x + bb.x;
              y1 = (y1 - mean) * rstd * gg.y + bb.y;
            }
            __half2 hi, lo;
            hilo2(y0, y1, hi, lo);
            *reinterpret_cast<__half2*>(xn_hi + u * (D + 32) + n) = hi;
            *reinterpret_cast<__half2*>(xn_lo + u * (D + 32) + n) = lo;
          }
        }
      }
      consumer_sync();
      ++n_xchg;
    };

    // per-phase clock totals (if p.timing): thread 0's, kept in SHARED memory - as registers the 12 counters cost 24 of
    // the 168 registers per thread whether or not the launch is timed
    long long* ph = reinterpret_cast<long long*>(smem + sm.ph);
    if (tid == 0) {
      for (int i = 0; i < 11; ++i) ph[i] = 0;
      ph[11] = clock64();
    }
    auto mark = [&](int i) {
      if (timed && tid == 0) {
        const long long now = clock64();
        ph[i] += now - ph[11];
        ph[11] = now;
      }
    };
    int t = 0;
#pragma unroll 1
    for (; t < p.L; ++t) {
#ifdef ASR_TRACE
      c.trace = (p.timing && blockIdx.x == 0 && t == 70) ? p.timing + 148 * 16 : nullptr;
#endif
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

        mark(0);
        // ---- LN1 -> q, k, v of this head (model.py:67-68, layers.py:16-18); layers > 0 got LN1 from the all-reduce
        if (l == 0) {
          ln_rows<S>(s_h, GU, ln, ln + D, xn_hi, xn_lo, stat);
          consumer_sync();
        }
        // The layer is three residual sub-blocks.  They run as three passes of ONE loop body so that the attention
        // core, its merge, the out projection and the all-reduce tail exist once in the instruction stream (the two
        // attentions differ only in runtime arguments): the kernel's hot loop has to stay inside the SM's instruction
        // cache, which it shares with nothing when alone but loses to co-running kernels.
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const float* ar_bias;
          const float* ar_gam;
          if (pass < 2) {
            int n_keys, n_mine;
            bool cur = false;
            if (pass == 0) {
              // ---- masked self attention (model.py:67-68): q, k, v of this head, cache append, keys 0..t
              mark(1);
              mm_stream<MQkv, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, const float4& v) {
                const float2 b = *reinterpret_cast<const float2*>(b_qkv + n);
                if (n < 64) {
                  q_store2(q_frag, q_f32, u0, n, (v.x + b.x) * qscale, (v.z + b.y) * qscale);
                  q_store2(q_frag, q_f32, u0 + 1, n, (v.y + b.x) * qscale, (v.w + b.y) * qscale);
                } else {   // k_t | v_t rows in the cache's swizzled chunk order (chunk ^ (t & 7)): TMA-stored as they are
                  const int e = n - 64, pos = (e & 64) + ((((e & 63) >> 3) ^ (t & 7)) << 3) + (e & 7);
                  *reinterpret_cast<__half2*>(kv_row + u0 * 128 + pos) = __floats2half2_rn(v.x + b.x, v.z + b.y);
                  *reinterpret_cast<__half2*>(kv_row + (u0 + 1) * 128 + pos) = __floats2half2_rn(v.y + b.x, v.w + b.y);
                }
              });
              fence_proxy_async();                                   // kv_row: generic writes -> the TMA store below
              consumer_sync();
              mark(2);
              // append k_t, v_t (f16) to the device-resident cache: [layer][utterance][head][K rows | V rows][Lc][64], the
              // 16-byte chunks of row t stored swizzled (chunk ^ (t & 7)) so that the bulk copy lands ldmatrix-ready.  The
              // first row of every 32-row block also zeroes the block's other rows: whole 32-key blocks are always finite.
              if (tid < GU * 2) {                                    // one 128-byte TMA store per (utterance, K | V)
                const int u = tid >> 1, kv = tid & 1;
                f16* dst = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head + size_t(kv) * Lc * 64 +
                            size_t(t) * 64;
                bulk_store(dst, kv_row + u * 128 + kv * 64, 128);
              }
              if ((t & 31) == 0) {
                for (int i = tid; i < GU * 2 * 31 * 8; i += NCT) {
                  const int u = i / 496, rem = i - u * 496, kv = rem / 248, w = rem - kv * 248;   // w: 16-byte word in 31 rows
                  f16* dst = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head + size_t(kv) * Lc * 64 +
                              size_t(t + 1) * 64 + w * 8;
                  *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                }
                asm volatile("fence.proxy.async;" ::: "memory");   // generic zero fill -> later TMA reads (1 step in 32)
              }
              n_keys = t;
              n_mine = t;
              cur = a_active && apart == S::WPU - 1;                // the current key: from shared memory (attn_init)
              ar_bias = b_o;
              ar_gam = ln + 2 * D;                              // out projection + residual (model.py:68) -> LN2 (:70)
            } else {
              // ---- cross-attention query -> attention over the encoder K/V (model.py:70-71)
              mm_stream<MWqc, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, const float4& v) {
                const float2 b = *reinterpret_cast<const float2*>(b_qc + n);
                q_store2(q_frag, nullptr, u0, n, (v.x + b.x) * qscale, (v.z + b.y) * qscale);
                q_store2(q_frag, nullptr, u0 + 1, n, (v.y + b.x) * qscale, (v.w + b.y) * qscale);
              });
              consumer_sync();
              mark(7);
              n_keys = p.Tp;
              n_mine = n_cross;
              ar_bias = b_oc;
              ar_gam = ln + 4 * D;                              // -> LN3 (model.py:73)
            }
            AttnT st;
            QFrag qa[4];
            attn_init(st, cur ? q_f32 + au * 64 : nullptr, kv_row + au * 128, t & 7);
            attn_q_bfrags(q_frag, au, qa);
            attention<S, S::SCX>(c, st, qa, n_keys, n_mine, a_active, au, apart);
            if (pass == 0) {
              if (tid < GU * 2) bulk_store_wait();              // this step's cache rows are written (published below)
              mark(3);
            } else {
              mark(8);
            }
            attn_finish<S>(st, GU, part_buf, stat, o_hi, o_lo);   // (two consumer barriers inside)
            if (pass == 0 && tid == 0) {                          // cache row t of this layer is published
              __threadfence_block();
              ctrl[2] = t * p.nd + l + 1;
            }
            mark(4);
            mm_stream<MWo, GUP>(c, reinterpret_cast<const uint8_t*>(o_hi), reinterpret_cast<const uint8_t*>(o_lo),
                                S::LDO, scratch, send_partial);
            mark(5);
          } else {
            // ---- FFN: squeeze rows of this CTA + ReLU, then the matching K-slice of unsqueeze (model.py:73-74)
            mm_stream<MW1, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, const float4& v) {
              const float2 b = *reinterpret_cast<const float2*>(b_1 + n);
              __half2 hi, lo;
              hilo2(fmaxf(v.x + b.x, 0.f), fmaxf(v.z + b.y, 0.f), hi, lo);
              *reinterpret_cast<__half2*>(hid_hi + u0 * (FFS + 32) + n) = hi;
              *reinterpret_cast<__half2*>(hid_lo + u0 * (FFS + 32) + n) = lo;
              hilo2(fmaxf(v.y + b.x, 0.f), fmaxf(v.w + b.y, 0.f), hi, lo);
              *reinterpret_cast<__half2*>(hid_hi + (u0 + 1) * (FFS + 32) + n) = hi;
              *reinterpret_cast<__half2*>(hid_lo + (u0 + 1) * (FFS + 32) + n) = lo;
            });
            consumer_sync();
            mark(9);
            mm_stream<MW2, GUP>(c, reinterpret_cast<const uint8_t*>(hid_hi), reinterpret_cast<const uint8_t*>(hid_lo),
                                S::LDH, scratch, send_partial);
            mark(10);
            // -> LN1 of the next layer (its parameters travel in this layer's block), or the classifier's plain split
            ar_bias = b_2;
            ar_gam = (l + 1 < p.nd) ? ln + 6 * D : nullptr;
          }
          all_reduce_finish(ar_bias, ar_gam, ar_gam + D);         // beta follows gamma in the parameter block
          mark(6);
        }
      }

      // ---- classifier WITHOUT the final LayerNorm (model.py:142): VS vocabulary rows per CTA
      if (p.nd == 0) {
        rows_to_hilo<D>(s_h, GU, nullptr, nullptr, xn_hi, xn_lo, D + 32);
        consumer_sync();
      }
      mm_stream<MCls, GUP>(c, xh, xl, S::LDX, scratch, [&](int n, int u0, const float4& v) {
        *reinterpret_cast<float2*>(s_lg + u0 * VS + n) = make_float2(v.x, v.z);
        *reinterpret_cast<float2*>(s_lg + (u0 + 1) * VS + n) = make_float2(v.y, v.w);
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
          *reinterpret_cast<float4*>(s_h + (i ^ hswz(u))) = make_float4(e.x + q.x, e.y + q.y, e.z + q.z, e.w + q.w);
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
#ifdef ASR_COUNT_LATE
      p.timing[size_t(blockIdx.x) * 16 + 1] = c.late;   // stages this warp found not yet full (diagnostic build)
#else
      p.timing[size_t(blockIdx.x) * 16 + 1] = 0;
#endif
      p.timing[size_t(blockIdx.x) * 16 + 2] = t_xchg;
      mark(0);
      for (int i = 0; i < 11; ++i) p.timing[size_t(blockIdx.x) * 16 + 5 + i] = ph[i];
    }
  }
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------ instances
typedef void (*ClusterKernel)(const ClusterParams, const CUtensorMap, const CUtensorMap);
struct Instance {
  int H, FFS, VS, GUP;
  ClusterKernel fn;
  SmemMap (*map)(int);
};
#define ASR_INST(H, FFS, VS, G) {H, FFS, VS, G, dec_cluster_kernel<Shape<H, FFS, VS, G>>, smem_map<Shape<H, FFS, VS, G>>}
const Instance kInstances[] = {
#if ASR_NCW == 8
    ASR_INST(4, 256, 64, 2), ASR_INST(4, 256, 64, 4), ASR_INST(4, 256, 64, 8),     // C1-C4: d_model 256, FFN 1024
    ASR_INST(2, 128, 128, 2), ASR_INST(2, 128, 128, 4), ASR_INST(2, 128, 128, 8),  // T0: d_model 128, FFN 256
    ASR_INST(8, 256, 32, 2), ASR_INST(8, 256, 32, 4), ASR_INST(8, 256, 32, 8),     // C5: d_model 512, FFN 2048
#else
    ASR_INST(4, 256, 64, 2), ASR_INST(4, 256, 64, 4),                              // (experiment: 4 consumer warps)
#endif
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
  L.small_floats = 256 + L.FFS + 11 * D;
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

  int max_smem = 0;
  if (int rc = device_props(nullptr, &max_smem)) return rc;
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
    if (const char* se = std::getenv("ASR_B200_CLUSTER_STAGES"))
      if (std::atoi(se) >= 3 && std::atoi(se) <= MAX_STAGES) nst = std::atoi(se);
    while (nst >= 3 && inst->map(nst).total > (uint32_t)max_smem) --nst;
    if (nst < 3) break;   // larger groups need even more shared memory
    if (int rc = ensure_dyn_smem((const void*)inst->fn, (size_t)max_smem)) return rc;
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
    // L2 prefetch of the encoder K/V one super-chunk ahead: pays at 8 utterances per cluster (16.85 -> 16.44 ms per
    // 256-utterance decode), measured slightly negative at 4 (11.17 -> 11.50 ms); "0" / "1" force it off / on
    const char* e = std::getenv("ASR_B200_KV_PREFETCH");
    p.kv_prefetch = (e && e[0]) ? (e[0] != '0') : (p.GUP == 8);
  }
  {
    const char* e = std::getenv("ASR_B200_KV_POLICY");   // "first" (default: stream K/V past the L2-resident weights) / "last"
    p.kv_evict_first = !(e && e[0] == 'l');
  }

  // encoder K/V as a 3-D tensor: [nd * B utterances][Tp rows][2D columns] f16; box = [GUP utterances][RPS rows][64 columns
  // (one head)] = one ring stage
  CUtensorMap map;
  const int rps = (STAGE_BYTES / 128) / p.GUP;
  const uint64_t dims[3] = {uint64_t(2 * p.D), uint64_t(p.Tp), uint64_t(p.nd) * p.B};
  const uint64_t strides[3] = {0, uint64_t(4 * p.D), uint64_t(p.Tp) * 4 * p.D};
  const uint32_t box[3] = {64u, uint32_t(rps), uint32_t(p.GUP)};
  if (int rc = make_tmap_f16(&map, p.ckv, 3, dims, strides, box, nullptr, /*swizzle=*/128)) return rc;
  // self K/V cache as a 4-D tensor: [nd * B utterances][H heads x (K | V)][Lc rows][64] f16, rows stored pre-swizzled
  // (no TMA swizzle); box = [GUP utterances][1][RPS rows][64]
  CUtensorMap cmap;
  {
    const uint64_t Lc = uint64_t((p.L + 31) & ~31);
    const uint64_t cdims[4] = {64, Lc, uint64_t(2 * p.H), uint64_t(p.nd) * p.B};
    const uint64_t cstr[4] = {0, 128, Lc * 128, uint64_t(2 * p.H) * Lc * 128};
    const uint32_t cbox[4] = {64u, uint32_t(rps), 1u, uint32_t(p.GUP)};
    if (int rc = make_tmap_f16(&cmap, p.cache, 4, cdims, cstr, cbox, nullptr, /*swizzle=*/0)) return rc;
  }

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
  ASR_CUDA_OK(cudaLaunchKernelEx(&cfg, chosen->fn, p, map, cmap));
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
