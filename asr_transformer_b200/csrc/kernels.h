// Internal launcher interface between the C-ABI layer (api.cu) and the sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace asr {

typedef __half f16;

// ---- error plumbing (thread-local message, negative return codes; never throws / aborts)
int set_error(int code, const char* fmt, ...);
const char* last_error();
#define ASR_CUDA_OK(expr)                                                                           \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::asr::set_error(-100, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---- launch accounting (bench.py reports the number of our kernels launched in the timed region)
extern unsigned long long g_kernel_launches;
#define ASR_LAUNCHED(n) (::asr::g_kernel_launches += (n))

// ---- per-device kernel configuration.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are properties
// of the CURRENT device: a process that drives several GPUs must configure every kernel on each of them, so the
// once-only guards are keyed by (device ordinal, kernel).  ensure_dyn_smem raises the opt-in limit of `fn` on the current
// device to at least `bytes` (no-op when already done); device_props reports the current device's SM count and
// opt-in shared-memory limit (cached per device).
int ensure_dyn_smem(const void* fn, size_t bytes);
int device_props(int* n_sm, int* max_smem_optin);

// ---- TMA tensor maps (driver entry point resolved at run time, no link-time libcuda dependency)
// f16 tensor, innermost dim contiguous. dims/strides innermost first; strides in BYTES for dims >= 1.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides /*nullable*/, int swizzle_bytes = 128 /* 0 = none */);
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides /*nullable*/, int swizzle_bytes = 128 /* 0 = none */);

// ---- GEMM  Y[M,N] = X[M,K] * W[N,K]^T (+bias)(relu)(+rowvec)(+residual)  (tcgen05 / TMEM / TMA)
struct GemmEpilogue {
  const float* bias = nullptr;       // [N]
  const float* residual = nullptr;   // fp32 [M, ld_res], added last
  int ld_res = 0;
  const float* rowvec = nullptr;     // fp32 [period, ld_rowvec] (positional encoding): + rowvec[row % period, col]
  int rowvec_period = 1;
  int ld_rowvec = 0;
  float* out_f32 = nullptr;          // [M, ld_f32]
  int ld_f32 = 0;
  f16* out_f16 = nullptr;          // [M, ld_f16]
  int ld_f16 = 0;
  int f16_lo_off = 0;                // != 0: also store lo = f16(v - f16(v)) at column + f16_lo_off (hi | lo A operand)
  int relu = 0;
  int n_store = 0;                   // columns >= n_store are not stored (0 => N)
};
// X: f16 [M, K] (row stride ldx elements), W: f16 [N_pad, K] (row stride ldw), K % 64 == 0, N_pad % 64 == 0.
// a_split != 0: X is [M, 2K] = [hi | lo] (x = hi + lo to 2^-22): Y = (hi + lo) W^T, i.e. fp32-accurate activations on the
// fp16 tensor cores for twice the MMAs (the K loop runs over both halves, W is streamed twice).
int launch_gemm_tc(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                   cudaStream_t s, int a_split = 0);
// Full-row GEMM with the LayerNorm that follows it fused into the epilogue (N == 256 only: one 128 x 256 tile holds whole
// rows): h = X W^T + bias (+rowvec) (+residual) -> ep.out_f32; y = LayerNorm(h) * gamma + beta -> ln.out_f16 (hi | lo pair
// when ln.split) and / or ln.out_f32.  Returns 1 (nothing launched) when the shape is unsupported: the caller then runs
// launch_gemm_tc + launch_layernorm.  Replaces the nn.LayerNorm launches at model.py:20,23,52,67,70,73,122.
struct LnEpilogue {
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  f16* out_f16 = nullptr;      // [M, D] or, with split, [M, 2D] = [hi | lo]
  int split = 0;
  float* out_f32 = nullptr;    // [M, D]
};
int launch_gemm_ln(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                   const LnEpilogue& ln, cudaStream_t s, int a_split = 0);
// Fused FFN (layers.py:53-58 + residual + following LayerNorm), D == 256 only (returns 1 otherwise -> two GEMMs):
// h = ep.residual + relu(X W1^T + b1) W2^T + ep.bias -> ep.out_f32; LayerNorm(h) -> ln outputs (ln.gamma == nullptr:
// no LayerNorm outputs wanted is not supported; pass the norm that follows).  X: [M, D] or, with split, [M, 2D] hi | lo;
// the hidden activation stays on the SM as an fp16 hi | lo (split) A operand.
int launch_ffn_fused(const f16* X, int ldx, const f16* W1, const float* b1, const f16* W2, int M, int D, int FF,
                     const GemmEpilogue& ep, const LnEpilogue& ln, cudaStream_t s, int split);
// Debug / cross-check: same contract, one thread per output element on CUDA cores.
int launch_gemm_naive(const f16* X, int ldx, const f16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                      cudaStream_t s);

// ---- multi-head attention core (flash style, tcgen05).  dh = 64 only.
struct AttnParams {
  const f16* q = nullptr;  int ldq = 0;  long long q_batch_stride = 0;   // element strides; head h at column h*64
  const f16* k = nullptr;  int ldk = 0;  long long k_batch_stride = 0;
  const f16* v = nullptr;  int ldv = 0;  long long v_batch_stride = 0;
  f16* out = nullptr;      int ldo = 0;  long long o_batch_stride = 0;   // [B, Sq, H*64]
  int out_lo_off = 0;                  // != 0: also store lo = f16(o - f16(o)) at column + out_lo_off (hi | lo rows)
  int B = 0, H = 0, Sq = 0, Sk = 0;
  float scale = 1.f;                   // emb_dim ** -0.5 (reference layers.py:20), NOT head_dim ** -0.5
  int causal = 0;                      // mask key j > query i
  const int32_t* k_lens = nullptr;     // [B]: keys >= k_lens[b] masked
  const uint8_t* q_valid = nullptr;    // [B, Sq]: 0 => whole query row masked
  const uint8_t* k_valid = nullptr;    // [B, Sk]: 0 => key masked
  const uint8_t* dense_mask = nullptr; // [mask_B (1 or B), Sq, Sk]: nonzero => masked
  int mask_B = 1;
};
int launch_attention_tc(const AttnParams& p, cudaStream_t s);
int launch_attention_naive(const AttnParams& p, cudaStream_t s);

// ---- UMMA probe (bring-up / regression check of descriptor encodings)
int launch_umma_probe(const f16* A, const f16* Bm, float* D, int N, int b_mn_major, cudaStream_t s);

// ---- elementwise / normalisation
// split != 0: y_f16 rows are [hi (D) | lo (D)] (row stride 2D): the A operand of an a_split GEMM
int launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int D, float eps, float* y_f32,
                     f16* y_f16, cudaStream_t s, int split = 0);
int launch_f32_to_f16_split(const float* x, f16* y, size_t rows, int D, cudaStream_t s);   // y [rows, 2D] = [hi | lo]
int launch_embed_pe(const int32_t* tokens, int ld_tok, const float* emb, const float* pe, int B, int L, int D,
                    int vocab, float* out, cudaStream_t s);
int launch_f32_to_f16(const float* x, f16* y, size_t n, cudaStream_t s);

// ---- conv front-end (reference model.py:168-171)
// conv1: spectrum fp32 (B,1,F,T) -> y1 f16 channels-last (B, T1, F1, 64)
// split != 0 (all three): activations travel as fp16 hi | lo pairs (x = hi + lo to 2^-22) and are multiplied twice on
// the tensor cores: y1 gets a second plane of lo parts (B*T1*F1*64 elements further), z rows become
// [hi (F2*64) | lo (F2*64)] (the A operand of the a_split _lin_in GEMM).
int launch_conv1(const float* spec, const float* w1 /*[9][64]*/, const float* b1, int B, int F, int T, f16* y1,
                 cudaStream_t s, int split = 0);
// conv2: y1 -> z f16 (B, T2, F2*64) with column order (f, c)
// conv1 + conv2 in one kernel (the f16 intermediate stays in shared memory); returns 1 if it does not fit -> use the
// two kernels below.  w2frag: pack_conv2_fragments order.
// tcgen05 edition of the fused front-end (conv_tc.cu; input_dim 80 only, returns 1 = not launched otherwise)
int launch_conv_tc(const float* spec, const float* w1, const float* b1, const f16* w2frag, const float* b2, int B, int F,
                   int T, f16* z, cudaStream_t s, int split);
int launch_conv_fused(const float* spec, const float* w1, const float* b1, const f16* w2frag, const float* b2, int B,
                      int F, int T, f16* z, cudaStream_t s, int split = 0);
int launch_conv2(const f16* y1, const f16* w2 /*[64 co][9][64 ci]*/, const float* b2, int B, int F1, int T1,
                 f16* z, cudaStream_t s, int split = 0);

// ---- power spectrogram front-end (reference dataset.py:34-35): audio (B, N) -> spec (B,1,n_fft/2+1,T), frames >= the
// signal's frame count zero-filled
int launch_spectrogram(const float* audio, int B, int n_samples, int n_fft, int hop, int T, float* spec, cudaStream_t s);

// ---- greedy decode step kernels
struct DecLinear {
  const float* x = nullptr;      // fp32 [B, K]
  int ldx = 0;
  const float* ln_gamma = nullptr, *ln_beta = nullptr;   // optional LayerNorm prologue over K (K == D)
  const f16* w = nullptr;       // f16 [N_pad, K]
  const float* bias = nullptr;   // [N] nullable
  int B = 0, N = 0, K = 0;
  int relu = 0;
  float* out = nullptr;          // fp32 [B, ldo]; with residual: out = residual + acc + bias (may alias residual)
  int ldo = 0;
  const float* residual = nullptr;
  int ld_res = 0;
  // QKV-append mode: columns [kv_col0, N) are also written as f16 into the cache row of step t
  f16* kv_cache = nullptr;      // [B, kv_rows, N - kv_col0]
  int kv_col0 = 0, kv_rows = 0;
  const int32_t* step = nullptr; // device step counter (row index into the cache)
};
int launch_dec_linear(const DecLinear& p, cudaStream_t s);

struct DecAttn {
  const float* q = nullptr; int ldq = 0;          // fp32 [B, ldq], head h at col h*64
  const f16* k = nullptr;  const f16* v = nullptr;
  int ldkv = 0; long long kv_batch_stride = 0;    // element strides
  int n_keys = 0;                                 // used when step == nullptr
  const int32_t* step = nullptr;                  // n_keys = *step + 1 (self attention over the cache)
  float* out = nullptr; int ldo = 0;
  int B = 0, H = 0;
  float scale = 1.f;
};
int launch_dec_attention(const DecAttn& p, cudaStream_t s);

struct DecSelect {
  const float* logits = nullptr; int ld = 0; int V = 0; int B = 0;
  int32_t* tokens = nullptr; int ld_tok = 0;      // tokens[b, step+1] = argmax
  int32_t* n_tokens = nullptr;                    // first EOS position (count of tokens incl. BOS/EOS), nullable
  int32_t* finished = nullptr;                    // [B]
  int32_t* step = nullptr;                        // incremented by block 0 after use
  float* step_logits = nullptr;                   // [B, L, V] nullable: copy of this step's logits
  int L = 0;
  int eos = 0, pad = 0, stop_at_eos = 0;
};
// argmax + EOS bookkeeping + embedding of the next token into h_next (single CTA; advances *step)
int launch_dec_select_embed(const DecSelect& p, const float* emb, const float* pe, int D, float* h_next,
                            cudaStream_t s);
int launch_dec_embed(const int32_t* tokens, int ld_tok, const int32_t* step, const float* emb, const float* pe, int B,
                     int D, int vocab, float* h, cudaStream_t s);

// ---- cluster greedy decoder (decode_cluster.cu): one thread-block cluster (num_heads CTAs, one head each) per group
// of <= 8 utterances; weights streamed from a fragment-major packed image (AsrWeights.dec_image)
struct ClusterMat { int MT, KB, KBS, KG; };   // m-tiles (16 rows), k-blocks (32 cols), k-blocks per ring stage, k-groups
struct ClusterLayout {                        // byte layout of the packed decoder image (see include/asr_b200.h)
  int CS, FFS, VS, small_floats;
  size_t small_bytes, off_small, off_qkv, off_wo, off_wqc, off_woc, off_w1, off_w2, layer_bytes, off_cls, rank_bytes,
      total_bytes;
};
bool cluster_layout(int D, int H, int FF, int V, int nd, ClusterLayout* out);
struct ClusterParams {
  int B, D, H, FF, V, L, Tp, nd;
  int GU, GUP, FFS, VS, nstages, kv_evict_first, kv_prefetch;
  ClusterMat m_qkv, m_wo, m_wqc, m_w1, m_w2, m_cls;
  const uint8_t* image; size_t image_bytes, rank_bytes, layer_bytes;
  size_t off_small, off_qkv, off_wo, off_wqc, off_woc, off_w1, off_w2, off_cls;
  uint32_t small_bytes;
  const float* emb; const float* pe; const float* h0;   // h0: fp32 [B][D] embedding + PE of the first token
  f16* cache;            // [nd][B][H][K rows | V rows][L][64]  (this kernel's own layout of the self-attention cache)
  const f16* ckv;        // [nd][B*Tp][2D]
  const int32_t* enc_lens;   // nullable [B]: valid encoder frames per utterance (cross-attention key-padding mask)
  int32_t* tokens; int32_t* n_tokens; float* step_logits;
  int eos, pad, stop_at_eos;
  float scale;
  long long* timing;      // nullable: [gridDim][16] {total, ring wait, exchange wait, producer wait, stages}
};
int launch_dec_cluster(ClusterParams& p, cudaStream_t s);

}  // namespace asr
