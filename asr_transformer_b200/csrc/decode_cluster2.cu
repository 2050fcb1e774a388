// Cluster greedy decoder, tcgen05 edition: the whole decode (all L steps, all layers) in ONE launch, every operand
// stream going TMA -> shared memory -> 5th-generation tensor core, accumulators in TMEM.
//
// Restates reference model.py:125-151 (Decoder.evaluate) with a device-resident KV cache.  As in the first cluster
// decoder a thread-block CLUSTER of num_heads CTAs owns a group of <= 8 utterances for the whole decode and splits
// every layer head-parallel; what changed is who does the arithmetic:
//   * every Linear layer is a swap-AB UMMA: the weight tile (128 output rows x 64 k, fp16, written by the host in the
//     canonical K-major SWIZZLE_128B layout so that a 1-D bulk copy lands it ready to use) is the A operand straight
//     from the TMA ring, the activations of the cluster's 8 utterances are the B operand with N = 16 columns
//     (utterance u -> column 2u = fp16 hi part, 2u+1 = lo part: fp32-accurate activations), the
//     accumulator [128 rows x 16] lives in TMEM.  No weight byte passes through a register;
//   * attention runs on the tensor cores too: S = K q^T with the K tile (128 keys x 64 dims, exactly what TMA wrote) as
//     A and the q rows of all utterances as B (the columns of the tile's own utterance are used), softmax by 128 threads
//     with ONE key each (tcgen05.ld), P (fp16 hi | lo) back to shared memory as the B operand of O = V^T P (V tile
//     consumed in place as an MN-major A operand, M = 64 head dims);
//   * the eight consumer warps only run epilogues (TMEM -> registers -> bias / ReLU / softmax / LayerNorm -> next B
//     operand), one thread issues every MMA, one thread issues every TMA;
//   * partial sums of the K-split projections are REDUCE-SCATTERED over distributed shared memory (each CTA sums only
//     the D / H rows it owns, 32-byte st.async rows), the new residual rows are all-gathered, LayerNorm runs on the
//     gathered row: half the DSMEM bytes and a quarter of the additions of the all-gather-then-sum of the first version.
// LayerNorm / softmax / residual / logits are fp32; K/V caches fp16; argmax lowest-index tie-break (model.py:143).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

constexpr int NCW = 8;                     // consumer (epilogue) warps
constexpr int NCT = NCW * 32;
constexpr int NTHREADS = NCT + 64;         // + TMA producer warp + MMA issuer warp (one working thread each)
constexpr int STAGE = 16384;               // ring stage: one 128 x 64 fp16 operand tile
constexpr int MAX_STAGES = 10;
constexpr int NU = 8;                      // utterance slots per cluster: N = 2 NU = 16 MMA columns (hi | lo)
constexpr int DR = 64;                     // rows of the residual stream owned by one CTA (= D / H = head dim)
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t TM_ACC = 0, TM_S = 64, TM_O = 320, TM_COLS = 512;   // TMEM column map

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
__device__ __forceinline__ void st_async_v2(uint32_t raddr, float a, float b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "f"(a), "f"(b), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar)
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
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {   // 32 lanes x 16 columns
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld2_issue(uint32_t taddr, uint32_t& a, uint32_t& b) {   // 32 lanes x 2 columns
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
// fp16 hi | lo split of x (x = hi + lo to 2^-22); both saturate instead of overflowing
__device__ __forceinline__ void split16(float x, f16& hi, f16& lo) {
  hi = f16_sat(x);
  lo = f16_sat(x - __half2float(hi));
}
// Byte offset of element (row n < 16, k) in a K-major SWIZZLE_128B B operand: k-blocks of 64 elements, each
// [2 atoms of 8 rows][128 B], 16-byte chunk c of row r stored at chunk c ^ (r & 7)  (what TMA / UMMA expect)
__device__ __forceinline__ uint32_t b_off(int n, int k) {
  return uint32_t(((k >> 6) << 11) + ((n >> 3) << 10) + ((n & 7) << 7) + (((((k & 63) >> 3) ^ (n & 7))) << 4) + ((k & 7) << 1));
}
__device__ __forceinline__ void b_store(uint8_t* base, int u, int k, float x) {   // rows 2u (hi), 2u+1 (lo)
  f16 hi, lo;
  split16(x, hi, lo);
  *reinterpret_cast<f16*>(base + b_off(2 * u, k)) = hi;
  *reinterpret_cast<f16*>(base + b_off(2 * u + 1, k)) = lo;
}

// ------------------------------------------------------------------------------------------------ static shapes
template <int H_, int FFS_, int VS_>
struct Shape {
  static constexpr int H = H_, CS = H_, D = 64 * H_, FFS = FFS_, VS = VS_;
  static constexpr int KBD = D / 64, KBF = FFS / 64;                 // k-blocks of the model / FFN-slice dimension
  static constexpr int TQKV = 2, TWO = D / 128, TW1 = FFS / 128, TW2 = D / 128, TCLS = (VS + 127) / 128;
  static constexpr int SMALL_FLOATS = 256 + FFS + 3 * DR + 8 * D;   // b_qkv | b_qc | b_1 | b_o, b_oc, b_2 (own rows) | 4 LN (g, b)
  static constexpr uint32_t SMALL_BYTES = (SMALL_FLOATS * 4 + 127) / 128 * 128;
  static_assert(D % 128 == 0 && FFS % 128 == 0 && D / H == DR && TWO <= 4 && TW1 <= 4 && TCLS <= 4, "shape");
};

struct SmemMap {
  uint32_t ring, bx, bh, bq, bo, pp, h, recv, prm, kvrow, qkf, lg, red, arg, tok, ctrl, bars, tmem, total;
};
template <class S>
__host__ __device__ inline SmemMap smem_map(int nstages) {
  SmemMap m;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes, uint32_t align) {
    off = (off + align - 1) & ~(align - 1);
    const uint32_t o = off;
    off += bytes;
    return o;
  };
  m.ring = take(uint32_t(nstages) * STAGE, 1024);
  m.bx = take(S::KBD * 2048, 1024);                 // LayerNorm output: B operand [16 rows][D]
  m.bh = take(S::KBF * 2048, 1024);                 // FFN hidden slice: B operand [16 rows][FFS]
  m.bq = take(2048, 1024);                          // q of this head: B operand [16 rows][64]
  m.bo = take(2048, 1024);                          // attention output of this head: B operand [16 rows][64]
  m.pp = take(2 * 4096, 1024);                      // probabilities of a chunk: 2 key tiles x B operand [16 rows][128 keys]
  m.h = take(S::D * NU * 4, 128);                   // residual stream [D][NU] fp32 (full copy in every CTA)
  m.recv = take(2 * S::CS * DR * NU * 4, 128);      // [parity][source rank][own row][NU] fp32 partial sums
  m.prm = take(2 * S::SMALL_BYTES, 128);            // [layer parity] biases and LayerNorm parameters
  m.kvrow = take(NU * 128 * 2, 128);                // k_t | v_t of the step in the cache's (swizzled) row order
  m.qkf = take(3 * NU * 64 * 4, 128);               // q, k_t, v_t fp32 [NU][64]: the current-row terms of the self attention
  m.lg = take(NU * S::VS * 4, 128);                 // logits of this CTA's vocabulary rows [NU][VS]
  m.red = take(512, 128);                           // small reductions
  m.arg = take(2 * S::CS * NU * 8, 128);            // [parity][source rank][NU] (value, index)
  m.tok = take(2 * NU * 4, 16);                     // next tokens | finished flags
  m.ctrl = take(16, 16);                            // [0] steps done, [1] stop, [2] cache rows written
  m.bars = take((2 * MAX_STAGES + 16) * 8, 8);
  m.tmem = take(16, 16);
  m.total = off;
  return m;
}

// ------------------------------------------------------------------------------------------------ the kernel
template <class S>
__global__ void __launch_bounds__(NTHREADS, 1)
dec_cluster2_kernel(const __grid_constant__ ClusterParams p, const __grid_constant__ CUtensorMap ckv_map) {
  constexpr int D = S::D, H = S::H, CS = S::CS, FFS = S::FFS, VS = S::VS, KBD = S::KBD, KBF = S::KBF;

  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemMap sm = smem_map<S>(p.nstages);
  uint8_t* ring = smem + sm.ring;
  uint8_t* Bx = smem + sm.bx;
  uint8_t* Bh = smem + sm.bh;
  uint8_t* Bq = smem + sm.bq;
  uint8_t* Bo = smem + sm.bo;
  uint8_t* Pp = smem + sm.pp;
  float* s_h = reinterpret_cast<float*>(smem + sm.h);
  float* recv = reinterpret_cast<float*>(smem + sm.recv);
  f16* kv_row = reinterpret_cast<f16*>(smem + sm.kvrow);
  float* qf = reinterpret_cast<float*>(smem + sm.qkf);      // [NU][64] q (scaled, log2 units)
  float* kf = qf + NU * 64;                                  // [NU][64] k_t
  float* vf = kf + NU * 64;                                  // [NU][64] v_t
  float* s_lg = reinterpret_cast<float*>(smem + sm.lg);
  float* red = reinterpret_cast<float*>(smem + sm.red);
  float2* arg = reinterpret_cast<float2*>(smem + sm.arg);
  int* s_tok = reinterpret_cast<int*>(smem + sm.tok);
  volatile int* ctrl = reinterpret_cast<volatile int*>(smem + sm.ctrl);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* go_bar = bars + 2 * MAX_STAGES;       // consumers -> MMA thread: the next B operand is in shared memory
  uint64_t* done_bar = go_bar + 1;                // MMA thread -> consumers: the MMA group has completed
  uint64_t* pfull = done_bar + 1;                 // [2] parameter block of layer parity b has landed
  uint64_t* pempty = pfull + 2;                   // [2] ... has been consumed
  uint64_t* xbar = pempty + 2;                    // [2] reduce-scatter parity barriers
  uint64_t* hbar = xbar + 2;                      // [2] all-gather parity barriers
  uint64_t* abar = hbar + 2;                      // [2] argmax exchange parity barriers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + sm.tmem);

  const int rank = int(cluster_ctarank());
  const int cluster_id = blockIdx.x / CS;
  const int ubase = cluster_id * p.GU;
  const int GU = min(p.GU, p.B - ubase);         // utterances of this cluster (>= 1 by construction of the grid)
  const int warp = threadIdx.x >> 5, tid = threadIdx.x, lane = tid & 31;
  const int NS = p.nstages;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();        // swizzled operand tiles need 1024-byte alignment
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(go_bar, NCW);
    mbar_init(done_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pfull[i], 1);
      mbar_init(&pempty[i], 1);
      mbar_init(&xbar[i], 1);
      mbar_init(&hbar[i], 1);
      mbar_init(&abar[i], 1);
    }
    ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0;
    fence_barrier_init();
  }
  // zero every operand / activation buffer once: rows of absent utterances (u >= GU) stay zero for the whole decode
  for (uint32_t i = sm.bx / 4 + tid; i < sm.bars / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (warp == NCW) {
    tmem_alloc(tmem_ptr, TM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // peers' barriers are initialised before any remote store can arrive
  const uint32_t tmem = *tmem_ptr;

  const int Lc = (p.L + 15) & ~15;                                // cache rows per (layer, utterance, head, K|V)
  const size_t cache_head = size_t(2) * Lc * 64;                  // elements per (layer, utterance, head): K rows | V rows
  const uint8_t* my_image = p.image + size_t(rank) * p.rank_bytes;
  auto n_tiles_of = [](int keys) { return (keys + 127) >> 7; };

  if (warp == NCW) {
    // =============================== TMA producer: one thread walks the static access sequence
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_last();
      const uint64_t pol_kv = p.kv_evict_first ? policy_evict_first() : policy_evict_last();
      int slot = 0;
      uint32_t round = 0;
      long long waited = 0;
      auto begin = [&](uint32_t bytes) -> uint8_t* {
        const long long w0 = clock64();
        mbar_wait(&empty[slot], (round & 1u) ^ 1u);
        waited += clock64() - w0;
        mbar_expect_tx(&full[slot], bytes);
        return ring + size_t(slot) * STAGE;
      };
      auto end = [&]() {
        if (++slot == NS) {
          slot = 0;
          ++round;
        }
      };
      auto mat = [&](const uint8_t* src, int n) {   // n consecutive 16 KB operand tiles of the packed image
#pragma unroll 1
        for (int s = 0; s < n; ++s) {
          uint8_t* dst = begin(STAGE);
          bulk_load(dst, src, STAGE, &full[slot], pol_w);
          src += STAGE;
          end();
        }
      };
      uint32_t li = 0;   // running layer index (parameter buffer parity)
#pragma unroll 1
      for (int t = 0; t < p.L; ++t) {
        if (p.stop_at_eos) {                       // strict gate: nothing of step t is requested before step t-1 ended
          while (ctrl[0] < t && !ctrl[1]) {
          }
          if (ctrl[1]) break;
        }
#pragma unroll 1
        for (int l = 0; l < p.nd; ++l, ++li) {
          const uint8_t* img = my_image + size_t(l) * p.layer_bytes;
          {   // biases and LayerNorm parameters of the layer: their own double buffer
            const uint32_t b = li & 1u;
            mbar_wait(&pempty[b], ((li >> 1) & 1u) ^ 1u);
            mbar_expect_tx(&pfull[b], S::SMALL_BYTES);
            bulk_load(smem + sm.prm + b * S::SMALL_BYTES, img + p.off_small, S::SMALL_BYTES, &pfull[b], pol_w);
          }
          mat(img + p.off_qkv, S::TQKV * KBD);
          if (t > 0) {   // self cache rows 0..t-1 of this layer (written by this CTA in earlier steps)
            const int need = (t - 1) * p.nd + l + 1;
            while (ctrl[2] < need) {
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            for (int c0 = 0; c0 < t; c0 += 256) {
              const int nt = n_tiles_of(min(256, t - c0));
              for (int kv = 0; kv < 2; ++kv)               // K tiles of the chunk, then its V tiles
                for (int ti = 0; ti < nt; ++ti) {
                  const int rows = (min(128, t - c0 - ti * 128) + 15) & ~15;   // whole 16-key blocks (rows past t are zero)
                  for (int u = 0; u < GU; ++u) {
                    uint8_t* dst = begin(uint32_t(rows) * 128u);
                    const f16* src = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head +
                                     size_t(kv) * Lc * 64 + size_t(c0 + ti * 128) * 64;
                    bulk_load(dst, src, rows * 128, &full[slot], pol_kv);
                    end();
                  }
                }
            }
          }
          mat(img + p.off_wo, S::TWO);
          mat(img + p.off_wqc, KBD);
          for (int c0 = 0; c0 < p.Tp; c0 += 256) {   // encoder K/V of this head: 2-D boxes [128 rows][64 columns]
            const int nt = n_tiles_of(min(256, p.Tp - c0));
            for (int kv = 0; kv < 2; ++kv)
              for (int ti = 0; ti < nt; ++ti)
                for (int u = 0; u < GU; ++u) {
                  uint8_t* dst = begin(STAGE);
                  tma_load_2d_hint(dst, &ckv_map, &full[slot], kv * D + rank * 64,
                                   (l * p.B + ubase + u) * p.Tp + c0 + ti * 128, pol_kv);
                  end();
                }
          }
          mat(img + p.off_woc, S::TWO);
          mat(img + p.off_w1, S::TW1 * KBD);
          mat(img + p.off_w2, S::TW2 * KBF);
        }
        mat(my_image + p.off_cls, S::TCLS * KBD);
      }
      if (p.timing) {
        p.timing[size_t(blockIdx.x) * 16 + 3] = waited;
        p.timing[size_t(blockIdx.x) * 16 + 4] = (long long)round * NS + slot;
      }
    }
    __syncwarp();
  } else if (warp == NCW + 1) {
    // =============================== MMA issuer: one thread issues every tcgen05.mma of the decode, in program order
    if (lane == 0) {
      constexpr uint32_t idesc_w = umma_idesc_f16(128, 16, 0, 0);    // weights / K tile (A, K-major) x activations (B)
      constexpr uint32_t idesc_pv = umma_idesc_f16(64, 16, 1, 0);    // V^T (A, MN-major, 64 head dims) x P (B)
      int slot = 0;
      uint32_t round = 0, go_n = 0;
      auto wait_go = [&]() {
        mbar_wait(go_bar, go_n & 1u);
        ++go_n;
        tc_fence_after();
      };
      auto stage_wait = [&]() -> uint32_t {
        mbar_wait(&full[slot], round & 1u);
        tc_fence_after();
        return smem_u32(ring + size_t(slot) * STAGE);
      };
      auto stage_release = [&]() {
        umma_commit(&empty[slot]);
        if (++slot == NS) {
          slot = 0;
          ++round;
        }
      };
      // acc[mt] (+)= W[mt] (tiles x kblocks ring stages, tile-major) * B
      auto mm = [&](int tiles, int kblocks, const uint8_t* bop) {
        const uint32_t b0 = smem_u32(bop);
#pragma unroll 1
        for (int mt = 0; mt < tiles; ++mt)
#pragma unroll 1
          for (int kb = 0; kb < kblocks; ++kb) {
            const uint64_t ad = umma_smem_desc_sw128(stage_wait(), 16, 1024);
            const uint64_t bd = umma_smem_desc_sw128(b0 + kb * 2048, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem + TM_ACC + mt * 16, ad + uint64_t(k * 2), bd + uint64_t(k * 2), idesc_w, (kb | k) != 0);
            stage_release();
          }
        umma_commit(done_bar);
      };
      // single-query attention of every utterance over n_keys streamed rows, in chunks of <= 2 key tiles
      auto attn = [&](int n_keys) {
        const uint64_t qd = umma_smem_desc_sw128(smem_u32(Bq), 16, 1024);
#pragma unroll 1
        for (int c0 = 0; c0 < n_keys; c0 += 256) {
          const int nk = min(256, n_keys - c0), nt = n_tiles_of(nk);
          wait_go();                                               // q in place / S and P of the previous chunk drained
#pragma unroll 1
          for (int ti = 0; ti < nt; ++ti)
#pragma unroll 1
            for (int u = 0; u < GU; ++u) {                         // S(u, tile) = K tile (128 keys x 64) . q^T
              const uint64_t kd = umma_smem_desc_sw128(stage_wait(), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss(tmem + TM_S + (u * 2 + ti) * 16, kd + uint64_t(k * 2), qd + uint64_t(k * 2), idesc_w, k != 0);
              stage_release();
            }
          umma_commit(done_bar);
          wait_go();                                               // P in place
#pragma unroll 1
          for (int ti = 0; ti < nt; ++ti) {
            const int nks = (min(128, nk - ti * 128) + 15) >> 4;   // 16-key steps that hold real (or zero-filled) rows
            const uint32_t pb = smem_u32(Pp + ti * 4096);
#pragma unroll 1
            for (int u = 0; u < GU; ++u) {                         // O(u) (+)= V tile^T (64 dims x keys) . P^T
              const uint64_t vd = umma_smem_desc_sw128(stage_wait(), 1024, 1024);
#pragma unroll 1
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t pd = umma_smem_desc_sw128(pb + (ks >> 2) * 2048, 16, 1024) + uint64_t((ks & 3) * 2);
                umma_f16_ss(tmem + TM_O + u * 16, vd + uint64_t(ks * (2048 >> 4)), pd, idesc_pv, (ti | ks) != 0);
              }
              stage_release();
            }
          }
          umma_commit(done_bar);
        }
      };
      bool stop = false;
#pragma unroll 1
      for (int t = 0; t < p.L && !stop; ++t) {
#pragma unroll 1
        for (int l = 0; l < p.nd; ++l) {
          wait_go();                                               // LN1 output in place
          if (l == 0 && ctrl[1]) {                                 // every utterance of the cluster has finished
            stop = true;
            break;
          }
          mm(S::TQKV, KBD, Bx);
          attn(t);
          wait_go();                                               // attention output in place
          mm(S::TWO, 1, Bo);
          wait_go();                                               // LN2 output
          mm(1, KBD, Bx);
          attn(p.Tp);
          wait_go();
          mm(S::TWO, 1, Bo);
          wait_go();                                               // LN3 output
          mm(S::TW1, KBD, Bx);
          wait_go();                                               // hidden slice
          mm(S::TW2, KBF, Bh);
        }
        if (stop) break;
        wait_go();                                                 // classifier input (or, with nd == 0, the stop check)
        if (p.nd == 0 && ctrl[1]) break;
        mm(S::TCLS, KBD, Bx);
      }
    }
    __syncwarp();
  } else {
    // ================================= consumers: epilogues only
    const bool timed = p.timing != nullptr;
    const long long t_begin = clock64();
    long long t_xchg = 0;
    uint32_t n_xchg = 0, n_arg = 0, done_n = 0, li = 0;
    const int wq = warp & 3, wg = warp >> 2;                        // TMEM lane quadrant / epilogue group
    const uint32_t tm_lane = tmem + (uint32_t(wq * 32) << 16);
    const float qscale = p.scale * LOG2E;
    auto wait_done = [&]() {
      mbar_wait(done_bar, done_n & 1u);
      ++done_n;
      tc_fence_after();
    };
    auto signal_go = [&]() {                 // this warp's operand writes are visible to the tensor core; its TMEM reads done
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(go_bar);
    };
    // key-padding mask of the cross attention: encoder frames >= enc_lens[utterance] are not attended (nullable)
    int n_cross[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int u = wg * 4 + j;
      n_cross[j] = (p.enc_lens && u < GU) ? max(0, min(p.Tp, p.enc_lens[ubase + u])) : p.Tp;
    }

    // embedding + PE of the first token (host-side init kernel): s_h[n][u]
    for (int i = tid; i < GU * D; i += NCT) {
      const int u = i / D, n = i - u * D;
      s_h[n * NU + u] = p.h0[size_t(ubase + u) * D + n];
    }
    consumer_sync();

    // peer windows of the receive buffers / barriers (mapa is linear inside a CTA's window); peer k = rank (rank+1+k) % CS
    uint32_t peer_recv[CS - 1], peer_h[CS - 1], peer_xbar[CS - 1], peer_hbar[CS - 1];
#pragma unroll
    for (int k = 0; k < CS - 1; ++k) {
      const uint32_t r = uint32_t(rank + 1 + k) % CS;
      peer_recv[k] = mapa_u32(smem_u32(recv), r);
      peer_h[k] = mapa_u32(smem_u32(s_h), r);
      peer_xbar[k] = mapa_u32(smem_u32(&xbar[0]), r);
      peer_hbar[k] = mapa_u32(smem_u32(&hbar[0]), r);
    }

    // LayerNorm of the (complete) residual rows -> fp16 hi | lo B operand Bx; gam == nullptr: plain split (classifier
    // input).  Thread tid owns utterance u = tid % NU and the rows n = tid / NU + 32 k.
    auto ln_to_bx = [&](const float* gam, const float* bet) {
      constexpr int EP = D / 32;
      const int u = tid % NU, nb = tid / NU;
      const float c0 = s_h[u];
      float hv[EP], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        hv[k] = s_h[(nb + 32 * k) * NU + u];
        const float d = hv[k] - c0;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
      float mean = 0.f, rstd = 1.f;
      if (gam) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        if (lane < NU) *reinterpret_cast<float2*>(red + 2 * (warp * NU + lane)) = make_float2(s1, s2);
        consumer_sync();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) {
          const float2 r = *reinterpret_cast<const float2*>(red + 2 * (w * NU + u));
          t1 += r.x;
          t2 += r.y;
        }
        mean = t1 * (1.0f / float(D));
        rstd = 1.0f / sqrtf(fmaxf(t2 * (1.0f / float(D)) - mean * mean, 0.f) + 1e-5f);
        mean += c0;
      }
      if (u < GU) {
#pragma unroll
        for (int k = 0; k < EP; ++k) {
          const int n = nb + 32 * k;
          b_store(Bx, u, n, gam ? (hv[k] - mean) * rstd * gam[n] + bet[n] : hv[k]);
        }
      }
    };

    // K-split projection -> reduce-scatter of the partial sums (each CTA owns DR rows) -> + bias + residual -> all-gather
    // of the new rows -> LayerNorm (or plain split) into Bx.  Ends by releasing the MMA thread.
    auto project_exchange = [&](int tiles, const float* bias_own, const float* gam, const float* bet) {
      const uint32_t par = n_xchg & 1u, phase = (n_xchg >> 1) & 1u;
      wait_done();
      if (tid == 0) {
        mbar_expect_tx(&xbar[par], uint32_t(CS - 1) * DR * NU * 4u);
        mbar_expect_tx(&hbar[par], uint32_t(CS - 1) * DR * NU * 4u);
      }
#pragma unroll 1
      for (int mt = wg; mt < tiles; mt += 2) {
        float v[16];
        tmem_ld16(tm_lane + TM_ACC + mt * 16, v);
        const int r = mt * 128 + wq * 32 + lane;           // output row of the full model dimension
        const int owner = r / DR, rr = r - owner * DR;
        float s[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) s[u] = v[2 * u] + v[2 * u + 1];
        const uint32_t off = uint32_t(((par * CS + rank) * DR + rr) * NU) * 4u;
        if (owner == rank) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(recv) + off);
          dst[0] = make_float4(s[0], s[1], s[2], s[3]);
          dst[1] = make_float4(s[4], s[5], s[6], s[7]);
        } else {
          const int k = (owner - rank - 1 + CS) % CS;
          st_async_v4(peer_recv[k] + off, s[0], s[1], s[2], s[3], peer_xbar[k] + par * 8u);
          st_async_v4(peer_recv[k] + off + 16u, s[4], s[5], s[6], s[7], peer_xbar[k] + par * 8u);
        }
      }
      consumer_sync();                                      // own rows written
      {
        const long long w0 = timed ? clock64() : 0;
        mbar_wait_cluster(&xbar[par], phase);
        if (timed) t_xchg += clock64() - w0;
      }
      {   // sum the CS partial rows this CTA owns: DR x NU values, two per thread
        const int rr = tid >> 2, u0 = (tid & 3) * 2;
        const float* rv = recv + size_t(par) * CS * DR * NU + rr * NU + u0;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
          const float2 x = *reinterpret_cast<const float2*>(rv + r * DR * NU);
          a += x.x;
          b += x.y;
        }
        const int n = rank * DR + rr;
        float2* hp = reinterpret_cast<float2*>(s_h + n * NU + u0);
        const float2 old = *hp;
        a = old.x + (a + bias_own[rr]);
        b = old.y + (b + bias_own[rr]);
        *hp = make_float2(a, b);
        const uint32_t off = uint32_t(n * NU + u0) * 4u;
#pragma unroll
        for (int k = 0; k < CS - 1; ++k) st_async_v2(peer_h[k] + off, a, b, peer_hbar[k] + par * 8u);
      }
      consumer_sync();
      {
        const long long w0 = timed ? clock64() : 0;
        mbar_wait_cluster(&hbar[par], phase);
        if (timed) t_xchg += clock64() - w0;
      }
      ++n_xchg;
      ln_to_bx(gam, bet);
      signal_go();
    };

    // Softmax / output side of the single-query attention of this head (the MMA thread runs attn(n_keys) in lock step).
    // Group wg owns utterances 4 wg .. 4 wg + 3; thread (wq, lane) owns key wq * 32 + lane of every tile and, for the
    // output, head dim wq * 16 + lane (lanes < 16: the M = 64 accumulator sits in lanes 0..15 of every TMEM quadrant).
    // with_cur: the step's own k_t / v_t (fp32, shared memory) enter as the initial state of the online softmax.
    auto attention = [&](int n_keys, const int (&n_valid)[4], bool with_cur) {
      float m[4], lcur[4], lsum[4], o[4];
      const int dim = wq * 16 + (lane & 15);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = wg * 4 + j;
        m[j] = -INFINITY; lcur[j] = 0.f; lsum[j] = 0.f; o[j] = 0.f;
        if (with_cur && u < GU) {
          m[j] = red[64 + u];                                  // q . k_t (log2 units), computed after the QKV epilogue
          lcur[j] = 1.f;
          o[j] = vf[u * 64 + dim];
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < n_keys; c0 += 256) {
        const int nk = min(256, n_keys - c0), nt = (nk + 127) >> 7;
        wait_done();                                            // S of the chunk is in TMEM
        float sc[4][2];
        {
          uint32_t ra[4][2], rb[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ti = 0; ti < 2; ++ti) {
              ra[j][ti] = rb[j][ti] = 0u;
              const int u = wg * 4 + j;
              if (ti < nt && u < GU) tmem_ld2_issue(tm_lane + TM_S + (u * 2 + ti) * 16 + 2 * u, ra[j][ti], rb[j][ti]);
            }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ti = 0; ti < 2; ++ti) {
              const int key = c0 + ti * 128 + wq * 32 + lane;
              const bool ok = ti < nt && (wg * 4 + j) < GU && key < n_valid[j];
              sc[j][ti] = ok ? __uint_as_float(ra[j][ti]) + __uint_as_float(rb[j][ti]) : -INFINITY;
            }
        }
        float mx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mx[j] = fmaxf(sc[j][0], sc[j][1]);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], off));
        }
        if (lane == 0) *reinterpret_cast<float4*>(red + (wg * 4 + wq) * 4) = make_float4(mx[0], mx[1], mx[2], mx[3]);
        group_sync(wg);
        float alpha[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float mn = m[j];
#pragma unroll
          for (int w = 0; w < 4; ++w) mn = fmaxf(mn, red[(wg * 4 + w) * 4 + j]);
          const float m_use = (mn == -INFINITY) ? 0.f : mn;     // no key at all: every p = ex2(-inf) = 0
          alpha[j] = ex2(m[j] - m_use);                          // m = -inf -> 0
          m[j] = mn;
          const int u = wg * 4 + j;
          float ps = 0.f;
#pragma unroll
          for (int ti = 0; ti < 2; ++ti) {
            if (ti < nt && u < GU) {
              const float pr = ex2(sc[j][ti] - m_use);
              ps += pr;
              b_store(Pp + ti * 4096, u, wq * 32 + lane, pr);
            }
          }
          lsum[j] = lsum[j] * alpha[j] + ps;
          lcur[j] *= alpha[j];
        }
        signal_go();                                            // P in place (and S drained)
        wait_done();                                            // O of the chunk is in TMEM
        {
          uint32_t ra[4], rb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ra[j] = rb[j] = 0u;
            const int u = wg * 4 + j;
            if (u < GU) tmem_ld2_issue(tm_lane + TM_O + u * 16 + 2 * u, ra[j], rb[j]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = o[j] * alpha[j] + (__uint_as_float(ra[j]) + __uint_as_float(rb[j]));
        }
        if (c0 + 256 < n_keys) signal_go();                    // the next chunk may overwrite S, P and O
        else group_sync(wg);                                    // (red[] is reused below)
      }
      // row sums over the 128 key threads of the group, then o / l -> Bo (rows 2u, 2u+1; k = head dim)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lsum[j] += __shfl_xor_sync(0xffffffffu, lsum[j], off);
      if (lane == 0) *reinterpret_cast<float4*>(red + (wg * 4 + wq) * 4) = make_float4(lsum[0], lsum[1], lsum[2], lsum[3]);
      group_sync(wg);
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int u = wg * 4 + j;
          if (u < GU) {
            float l = lcur[j];
#pragma unroll
            for (int w = 0; w < 4; ++w) l += red[(wg * 4 + w) * 4 + j];
            b_store(Bo, u, dim, l > 0.f ? o[j] / l : 0.f);
          }
        }
      }
      group_sync(wg);                                           // red[] free again
      signal_go();                                              // attention output in place
    };

    long long ph[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // per-phase clock totals (thread 0; written if p.timing)
    long long ph_t = clock64();
    auto mark = [&](int i) {
      if (timed) {
        const long long now = clock64();
        ph[i] += now - ph_t;
        ph_t = now;
      }
    };
    int t = 0;
#pragma unroll 1
    for (; t < p.L; ++t) {
      const int self_valid[4] = {t, t, t, t};
#pragma unroll 1
      for (int l = 0; l < p.nd; ++l, ++li) {
        // ---- this layer's biases and LayerNorm parameters (double-buffered block, used in place)
        const uint32_t pb = li & 1u;
        mbar_wait(&pfull[pb], (li >> 1) & 1u);
        const float* prm = reinterpret_cast<const float*>(smem + sm.prm + pb * S::SMALL_BYTES);
        const float* b_qkv = prm;                 // [192] q | k | v rows of this head
        const float* b_qc = prm + 192;            // [64]
        const float* b_1 = prm + 256;             // [FFS]
        const float* b_o = prm + 256 + FFS;       // [DR] rows owned by this CTA
        const float* b_oc = b_o + DR;
        const float* b_2 = b_oc + DR;
        const float* ln = b_2 + DR;               // ln1 g,b | ln2 g,b | ln3 g,b | ln1 of the next layer g,b
        mark(0);
        if (l == 0) {   // LN1 of the first layer (later layers get it from the previous layer's last exchange)
          ln_to_bx(ln, ln + D);
          signal_go();
        }
        mark(1);
        // ---- q, k, v of this head (model.py:67-68, layers.py:16-18): tile 0 = q | k rows, tile 1 = v rows (+ padding)
        wait_done();
        {
          float v[16];
          tmem_ld16(tm_lane + TM_ACC + wg * 16, v);
          const int r = wg * 128 + wq * 32 + lane;
          if (r < 192) {
            const float bias = b_qkv[r];
            if (r < 64) {
#pragma unroll
              for (int u = 0; u < NU; ++u) {
                const float y = (v[2 * u] + v[2 * u + 1] + bias) * qscale;
                if (u < GU) {
                  b_store(Bq, u, r, y);
                  qf[u * 64 + r] = y;
                }
              }
            } else {   // k_t | v_t: fp32 copies for this step, fp16 row in the cache's swizzled chunk order for the append
              const int e = r - 64, d = e & 63;
              const int pos = (e & 64) + ((((e & 63) >> 3) ^ (t & 7)) << 3) + (e & 7);
              float* dstf = (e < 64 ? kf : vf);
#pragma unroll
              for (int u = 0; u < NU; ++u) {
                const float y = v[2 * u] + v[2 * u + 1] + bias;
                if (u < GU) {
                  dstf[u * 64 + d] = y;
                  kv_row[u * 128 + pos] = f16_sat(y);
                }
              }
            }
          }
        }
        fence_proxy_async();                                   // kv_row: generic writes -> the TMA store below
        consumer_sync();
        mark(2);
        // append k_t, v_t (fp16) to the device-resident cache: [layer][utterance][head][K rows | V rows][Lc][64], the
        // 16-byte chunks of row t stored swizzled (chunk ^ (t & 7)) so that the bulk copy lands as a UMMA operand tile.
        // The first row of every 16-row block also zeroes the block's other rows: whole 16-key blocks are always finite.
        if (tid < GU * 2) {                                    // one 128-byte TMA store per (utterance, K | V)
          const int u = tid >> 1, kv = tid & 1;
          f16* dst = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head + size_t(kv) * Lc * 64 +
                     size_t(t) * 64;
          bulk_store(dst, kv_row + u * 128 + kv * 64, 128);
        }
        if ((t & 15) == 0) {
          for (int i = tid; i < GU * 2 * 15 * 8; i += NCT) {
            const int u = i / 240, rem = i - u * 240, kv = rem / 120, w = rem - kv * 120;   // w: 16-byte word in 15 rows
            f16* dst = p.cache + ((size_t(l) * p.B + ubase + u) * H + rank) * cache_head + size_t(kv) * Lc * 64 +
                       size_t(t + 1) * 64 + w * 8;
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // generic zero fill -> later TMA reads (1 step in 16)
        }
        if (warp < GU) {   // score of the step's own key: q . k_t (fp32; q carries scale * log2 e)
          float s = qf[warp * 64 + lane] * kf[warp * 64 + lane] + qf[warp * 64 + 32 + lane] * kf[warp * 64 + 32 + lane];
          s = warp_sum(s);
          if (lane == 0) red[64 + warp] = s;
        }
        consumer_sync();
        if (t > 0) signal_go();                                 // q in place: the MMA thread may start S = K q^T
        // ---- masked self attention over the cache rows 0..t-1 + the current row (model.py:67-68)
        attention(t, self_valid, true);
        if (tid < GU * 2) bulk_store_wait();                    // this step's cache rows are written (published below)
        if (tid == 0) {                                         // cache row t of this layer is published
          __threadfence_block();
          ctrl[2] = t * p.nd + l + 1;
        }
        mark(3);
        project_exchange(S::TWO, b_o, ln + 2 * D, ln + 3 * D);  // out projection + residual (model.py:68) -> LN2 (:70)
        mark(6);
        // ---- cross-attention query -> attention over the encoder K/V (model.py:70-71)
        wait_done();
        if (wg == 0) {
          float v[16];
          tmem_ld16(tm_lane + TM_ACC, v);
          const int r = wq * 32 + lane;
          if (r < 64) {
            const float bias = b_qc[r];
#pragma unroll
            for (int u = 0; u < NU; ++u)
              if (u < GU) b_store(Bq, u, r, (v[2 * u] + v[2 * u + 1] + bias) * qscale);
          }
        }
        signal_go();
        mark(7);
        attention(p.Tp, n_cross, false);
        mark(8);
        project_exchange(S::TWO, b_oc, ln + 4 * D, ln + 5 * D);   // -> LN3 (model.py:73)
        mark(6);
        // ---- FFN: squeeze rows of this CTA + ReLU, then the matching K-slice of unsqueeze (model.py:73-74)
        wait_done();
#pragma unroll 1
        for (int mt = wg; mt < S::TW1; mt += 2) {
          float v[16];
          tmem_ld16(tm_lane + TM_ACC + mt * 16, v);
          const int r = mt * 128 + wq * 32 + lane;
          const float bias = b_1[r];
#pragma unroll
          for (int u = 0; u < NU; ++u)
            if (u < GU) b_store(Bh, u, r, fmaxf(v[2 * u] + v[2 * u + 1] + bias, 0.f));
        }
        signal_go();
        mark(9);
        // -> LN1 of the next layer (its parameters travel in this layer's block), or the classifier's plain split
        const bool last = l + 1 == p.nd;
        project_exchange(S::TW2, b_2, last ? nullptr : ln + 6 * D, last ? nullptr : ln + 7 * D);
        mark(10);
        consumer_sync();                                        // every thread is done with this layer's parameters
        if (tid == 0) mbar_arrive(&pempty[pb]);
      }

      // ---- classifier WITHOUT the final LayerNorm (model.py:142): VS vocabulary rows per CTA
      if (p.nd == 0) {
        ln_to_bx(nullptr, nullptr);
        signal_go();
      }
      wait_done();
#pragma unroll 1
      for (int mt = wg; mt < S::TCLS; mt += 2) {
        float v[16];
        tmem_ld16(tm_lane + TM_ACC + mt * 16, v);
        const int r = mt * 128 + wq * 32 + lane;
        if (r < VS) {
#pragma unroll
          for (int u = 0; u < NU; ++u) s_lg[u * VS + r] = v[2 * u] + v[2 * u + 1];
        }
      }
      tc_fence_before();
      consumer_sync();
      const int v_lo = rank * VS, v_n = max(0, min(VS, p.V - v_lo));   // this CTA's vocabulary range
      if (p.step_logits)
        for (int i = tid; i < GU * v_n; i += NCT) {
          const int u = i / v_n, v = i - u * v_n;
          p.step_logits[(size_t(ubase + u) * p.L + t) * p.V + v_lo + v] = s_lg[u * VS + v];
        }
      // ---- argmax: local (warp u), then across the cluster (lowest index wins ties, model.py:143)
      const uint32_t apar = n_arg & 1u, aphase = (n_arg >> 1) & 1u;
      if (tid == 0) mbar_expect_tx(&abar[apar], uint32_t(CS - 1) * NU * 8u);
      if (warp < NU) {
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
          float2* mine = arg + (size_t(apar) * CS + rank) * NU + warp;
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
          const float2 a = arg[(size_t(apar) * CS + r) * NU + tid];
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
        consumer_sync();                                        // ctrl[1] is visible before the MMA thread is woken
        if (t < p.L) signal_go();                               // wake the MMA thread: it leaves on the stop flag
        break;
      }
      if (t + 1 < p.L) {   // embedding + PE of the next input token (model.py:137)
        for (int i = tid; i < GU * D; i += NCT) {
          const int u = i / D, n = i - u * D;
          s_h[n * NU + u] = __ldg(p.emb + size_t(s_tok[u]) * D + n) + __ldg(p.pe + size_t(t + 1) * D + n);
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
      p.timing[size_t(blockIdx.x) * 16 + 1] = 0;
      p.timing[size_t(blockIdx.x) * 16 + 2] = t_xchg;
      mark(0);
      for (int i = 0; i < 11; ++i) p.timing[size_t(blockIdx.x) * 16 + 5 + i] = ph[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW) tmem_dealloc(tmem, TM_COLS);
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------ instances
typedef void (*Cluster2Kernel)(const ClusterParams, const CUtensorMap);
struct Instance {
  int H, FFS, VS;
  Cluster2Kernel fn;
  SmemMap (*map)(int);
};
#define ASR_INST2(H, FFS, VS) {H, FFS, VS, dec_cluster2_kernel<Shape<H, FFS, VS>>, smem_map<Shape<H, FFS, VS>>}
const Instance kInstances[] = {
    ASR_INST2(4, 256, 64),     // C1-C4: d_model 256, FFN 1024
    ASR_INST2(2, 128, 128),    // T0: d_model 128, FFN 256
    ASR_INST2(8, 256, 32),     // C5: d_model 512, FFN 2048
};
const Instance* find_instance(int H, int FFS, int VS) {
  for (const Instance& i : kInstances)
    if (i.H == H && i.FFS == FFS && i.VS == VS) return &i;
  return nullptr;
}

}  // namespace

bool cluster2_layout(int D, int H, int FF, int V, int nd, ClusterLayout* out) {
  ClusterLayout L{};
  if (H < 2 || H > 8 || (H & (H - 1)) || D != 64 * H || FF % (128 * H) != 0 || V < 1 || nd < 1) return false;
  L.CS = H;
  L.FFS = FF / H;
  L.VS = ((V + H - 1) / H + 15) / 16 * 16;
  if (!find_instance(H, L.FFS, L.VS)) return false;   // only the compiled shapes
  L.small_floats = 256 + L.FFS + 3 * DR + 8 * D;
  L.small_bytes = (size_t(L.small_floats) * 4 + 127) / 128 * 128;
  auto tiles = [](int rows) { return size_t((rows + 127) / 128); };
  size_t off = 0;
  L.off_small = off; off += (L.small_bytes + 1023) / 1024 * 1024;
  L.off_qkv = off;   off += 2 * size_t(D / 64) * STAGE;
  L.off_wo = off;    off += tiles(D) * STAGE;
  L.off_wqc = off;   off += size_t(D / 64) * STAGE;
  L.off_woc = off;   off += tiles(D) * STAGE;
  L.off_w1 = off;    off += tiles(L.FFS) * (D / 64) * STAGE;
  L.off_w2 = off;    off += tiles(D) * (L.FFS / 64) * STAGE;
  L.layer_bytes = off;
  L.off_cls = size_t(nd) * L.layer_bytes;
  L.rank_bytes = L.off_cls + tiles(L.VS) * (D / 64) * STAGE;
  L.total_bytes = L.rank_bytes * H;
  if (out) *out = L;
  return true;
}

int launch_dec_cluster2(ClusterParams& p, cudaStream_t s) {
  ClusterLayout lay;
  if (!cluster2_layout(p.D, p.H, p.FF, p.V, p.nd, &lay))
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
  const Instance* inst = find_instance(p.H, p.FFS, p.VS);
  int nst = MAX_STAGES;
  if (const char* se = std::getenv("ASR_B200_CLUSTER_STAGES"))
    if (std::atoi(se) >= 3 && std::atoi(se) <= MAX_STAGES) nst = std::atoi(se);
  while (nst >= 3 && inst->map(nst).total > (uint32_t)max_smem) --nst;
  if (nst < 3) return set_error(-2, "cluster decoder: shared memory too small (D=%d FF=%d)", p.D, p.FF);
  {
    static const void* configured[8][16] = {};    // per device ordinal: the attribute is per device
    bool done = false;
    const int di = dev & 7;
    for (const void* q : configured[di]) done |= (q == (const void*)inst->fn);
    if (!done) {
      ASR_CUDA_OK(cudaFuncSetAttribute((const void*)inst->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      for (auto& q : configured[di])
        if (!q) {
          q = (const void*)inst->fn;
          break;
        }
    }
  }
  p.GUP = NU;
  p.nstages = nst;
  {
    const char* e = std::getenv("ASR_B200_KV_POLICY");   // "first" (default: stream K/V past the L2-resident weights) / "last"
    p.kv_evict_first = !(e && e[0] == 'l');
  }
  // utterances per cluster: spread the batch over as many clusters as can be resident at once (one wave), at most NU
  // each.  The arithmetic of an utterance does not depend on how many share its cluster (the MMA has 16 columns either
  // way, the softmax is always split over the same 128 key threads), so tokens are invariant to this choice.
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = inst->map(nst).total;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.H; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(p.H, 1, 1);
  int max_clusters = 0;
  ASR_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)inst->fn, &cfg));
  if (max_clusters < 1) return set_error(-2, "cluster decoder: no cluster of %d CTAs can be resident", p.H);
  int gu = (p.B + max_clusters - 1) / max_clusters;
  if (const char* ge = std::getenv("ASR_B200_CLUSTER_GU"))
    if (ge[0] && std::atoi(ge) >= 1) gu = std::atoi(ge);
  p.GU = gu < 1 ? 1 : (gu > NU ? NU : gu);
  const int n_clusters = (p.B + p.GU - 1) / p.GU;

  // encoder K/V as a 2-D tensor: [nd * B * Tp rows][2D columns] f16; box = [128 rows][64 columns] (one head, one key tile)
  CUtensorMap map;
  const uint64_t dims[2] = {uint64_t(2 * p.D), uint64_t(p.nd) * p.B * p.Tp};
  const uint64_t strides[2] = {0, uint64_t(4 * p.D)};
  const uint32_t box[2] = {64u, 128u};
  if (int rc = make_tmap_f16(&map, p.ckv, 2, dims, strides, box, nullptr, /*swizzle=*/128)) return rc;

  cfg.gridDim = dim3(n_clusters * p.H, 1, 1);
  ASR_CUDA_OK(cudaLaunchKernelEx(&cfg, inst->fn, p, map));
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
