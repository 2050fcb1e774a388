// C-ABI layer (include/asr_b200.h): argument checking, workspace carving and kernel orchestration.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "../../include/asr_b200.h"
#include "kernels.h"
#include "ptx.cuh"

using namespace asr;

namespace {

inline int conv_len(int n) { return (n - 3) / 2 + 1; }

// Activations that feed a Linear / conv2 as the tensor-core A operand travel as fp16 hi | lo pairs (x = hi + lo to
// 2^-22) and are multiplied twice (a_split GEMMs): measured necessary for the 99 %-identical-tokens bar (a single fp16
// rounding of every linear input flips ~0.3 % of the utterances).  ASR_B200_SPLIT=0 switches to
// single fp16 operands (half the MMAs; for A/B measurements only).
int g_split = [] {
  const char* e = std::getenv("ASR_B200_SPLIT");
  return (e && e[0] == '0') ? 0 : 1;
}();
inline int SP() { return g_split ? 2 : 1; }

// Bump allocator over the caller's workspace (base == nullptr: size-only dry run).
struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct EncodeWs {
  f16 *y1, *z, *xn, *qkv, *att, *ff, *enc_f16;
  float* h;
  void carve(Bump& b, const AsrConfig& c, int B, int T) {
    const int F1 = conv_len(c.input_dim), T1 = conv_len(T), F2 = conv_len(F1), T2 = conv_len(T1);
    const size_t M = size_t(B) * T2, D = c.embedding_dim;
    y1 = b.take<f16>(size_t(B) * T1 * F1 * 64 * 2);   // (hi | lo planes / halves: see g_split)
    z = b.take<f16>(M * F2 * 64 * 2);
    h = b.take<float>(M * D);
    xn = b.take<f16>(M * D * 2);
    qkv = b.take<f16>(M * 3 * D);
    att = b.take<f16>(M * D * 2);
    ff = b.take<f16>(M * c.ff_dim * 2);
    enc_f16 = b.take<f16>(M * D * 2);
  }
};

struct DecFwdWs {
  f16 *enc_f16, *ckv, *xn, *qkv, *qc, *att, *ff;
  float* h;
  void carve(Bump& b, const AsrConfig& c, int B, int Tp, int L) {
    const size_t M = size_t(B) * Tp, R = size_t(B) * L, D = c.embedding_dim;
    enc_f16 = b.take<f16>(M * D * 2);
    ckv = b.take<f16>(size_t(c.decoder_num_layers) * M * 2 * D);
    h = b.take<float>(R * D);
    xn = b.take<f16>(R * D * 2);
    qkv = b.take<f16>(R * 3 * D);
    qc = b.take<f16>(R * D);
    att = b.take<f16>(R * D * 2);
    ff = b.take<f16>(R * c.ff_dim * 2);
  }
};

struct GreedyWs {
  f16 *enc_f16, *ckv, *cache;
  float *h, *qkv, *att, *qc, *ff, *logits;
  int32_t *step, *finished;
  void carve(Bump& b, const AsrConfig& c, int B, int Tp, int L) {
    const size_t M = size_t(B) * Tp, D = c.embedding_dim;
    const size_t vpad = (size_t(c.vocab_size) + 63) / 64 * 64;
    enc_f16 = b.take<f16>(M * D * 2);
    ckv = b.take<f16>(size_t(c.decoder_num_layers) * M * 2 * D);
    cache = b.take<f16>(size_t(c.decoder_num_layers) * B * ((L + 31) / 32 * 32) * 2 * D);   // rows padded to 32-key blocks
    h = b.take<float>(size_t(B) * D);
    qkv = b.take<float>(size_t(B) * 3 * D);
    att = b.take<float>(size_t(B) * D);
    qc = b.take<float>(size_t(B) * D);
    ff = b.take<float>(size_t(B) * c.ff_dim);
    logits = b.take<float>(size_t(B) * vpad);
    step = b.take<int32_t>(1);
    finished = b.take<int32_t>(B);
  }
};

struct GraphKey {
  void* ws = nullptr; const void* enc = nullptr; void* tokens = nullptr; void* n_tokens = nullptr;
  void* step_logits = nullptr; const void* first = nullptr; int B = 0, Tp = 0, L = 0, stop = 0;
  bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) == 0; }
};

// Optional per-kernel-class timing of the decode step (asr_decode_profile): CUDA event pairs around every launch.
enum DecClass { DC_QKV = 0, DC_SELF_ATTN, DC_OUT_PROJ, DC_CROSS_Q, DC_CROSS_ATTN, DC_FFN1, DC_FFN2, DC_CLASSIFIER,
                DC_SELECT, DC_COUNT };
struct StepProf {
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  size_t used = 0;
  cudaStream_t s = nullptr;
  void mark(int c) {
    if (used + 2 > ev.size()) {
      ev.resize(used + 2);
      cudaEventCreate(&ev[used]);
      cudaEventCreate(&ev[used + 1]);
    }
    cls.push_back(c);
    cudaEventRecord(ev[used], s);
  }
  void done() {
    cudaEventRecord(ev[used + 1], s);
    used += 2;
  }
  ~StepProf() {
    for (auto e : ev) cudaEventDestroy(e);
  }
};
#define PROF(prof, c, call)            \
  do {                                 \
    if (prof) (prof)->mark(c);         \
    if (int rc_ = (call)) return rc_;  \
    if (prof) (prof)->done();          \
  } while (0)

__global__ void dec_init_kernel(int32_t* tokens, int ld_tok, int32_t* n_tokens, int32_t* finished, int32_t* step,
                                int B, int L, int bos, const int32_t* first_tokens) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *step = 0;
  if (b < B) {
    tokens[size_t(b) * ld_tok] = first_tokens ? first_tokens[b] : bos;
    finished[b] = 0;
    if (n_tokens) n_tokens[b] = L + 1;
  }
}


// ------------------------------------------------------------------------------------------------ beam search
// One CTA per utterance: log-softmax of the K hypothesis rows, candidate score = hypothesis score + log p(token)
// (finished hypotheses: a single candidate, themselves, extended with pad), then the K best of the K x V candidates by
// K rounds of a block-wide arg-max; ties go to the lower flat index k * V + token (the lowest-index rule of argmax,
// model.py:143).  Semantics = oracle/speech_transformer.py: beam_search_kv_cached.
__global__ void __launch_bounds__(256)
beam_select_kernel(const float* __restrict__ logits, int ld, int V, int K, float* score, int32_t* finished,
                   int32_t* parent, int32_t* token, int eos, int pad) {
  extern __shared__ float cand[];                 // [K * V] candidate scores
  __shared__ float s_red[8];
  __shared__ int s_redi[8];
  __shared__ float s_score[16];
  __shared__ int s_fin[16], s_par[16], s_tok[16];
  __shared__ float s_new[16];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < K) {
    s_score[tid] = score[b * K + tid];
    s_fin[tid] = finished[b * K + tid];
  }
  __syncthreads();
  for (int k = warp; k < K; k += 8) {             // one warp per hypothesis row
    const float* row = logits + size_t(b * K + k) * ld;
    float* out = cand + k * V;
    if (s_fin[k]) {
      for (int v = lane; v < V; v += 32) out[v] = (v == pad) ? s_score[k] : -INFINITY;
      continue;
    }
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, row[v]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int v = lane; v < V; v += 32) sum += expf(row[v] - mx);
    const float lse = mx + logf(warp_sum(sum));
    for (int v = lane; v < V; v += 32) out[v] = s_score[k] + (row[v] - lse);   // -inf + x = -inf: dead hypotheses
  }
  __syncthreads();
  const int n = K * V;
  for (int j = 0; j < K; ++j) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < n; i += 256) {
      const float x = cand[i];
      if (x > best || (x == best && i < bi)) {
        best = x;
        bi = i;
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
      s_red[warp] = best;
      s_redi[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_red[w] > best || (s_red[w] == best && s_redi[w] < bi)) {
          best = s_red[w];
          bi = s_redi[w];
        }
      s_par[j] = bi / V;
      s_tok[j] = bi % V;
      s_new[j] = best;
      cand[bi] = __int_as_float(0xffc00000);      // NaN: compares false against everything, never selected again
    }
    __syncthreads();
  }
  if (tid < K) {
    const int par = s_par[tid], tok = s_tok[tid];
    parent[b * K + tid] = par;
    token[b * K + tid] = tok;
    score[b * K + tid] = s_new[tid];
    finished[b * K + tid] = (s_fin[par] || tok == eos) ? 1 : 0;
  }
}

// Survivors inherit their parent's state: self K/V cache rows 0..t of every layer and the token history, gathered into
// the second buffer set (grid: hypothesis row x layer); the new token lands at position t + 1; the device step counter
// (cache append position / attention length of the next step) advances.
__global__ void __launch_bounds__(256)
beam_reorder_kernel(const f16* __restrict__ cache_src, f16* __restrict__ cache_dst, const int32_t* __restrict__ tok_src,
                    int32_t* __restrict__ tok_dst, const int32_t* __restrict__ parent, const int32_t* __restrict__ token,
                    int32_t* step, int R, int K, int L, int D2, int t) {
  const int r = blockIdx.x, l = blockIdx.y;
  const int src_r = (r / K) * K + parent[r];
  const size_t row_elems = size_t(L) * D2;
  const uint4* src = reinterpret_cast<const uint4*>(cache_src + (size_t(l) * R + src_r) * row_elems);
  uint4* dst = reinterpret_cast<uint4*>(cache_dst + (size_t(l) * R + r) * row_elems);
  const int n16 = (t + 1) * D2 / 8;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  if (l == 0) {
    for (int i = threadIdx.x; i <= t; i += blockDim.x) tok_dst[size_t(r) * (L + 1) + i] = tok_src[size_t(src_r) * (L + 1) + i];
    if (threadIdx.x == 0) tok_dst[size_t(r) * (L + 1) + t + 1] = token[r];
    if (r == 0 && threadIdx.x == 0) *step = t + 1;
  }
}

__global__ void beam_init_kernel(int32_t* tok, int ld_tok, float* score, int32_t* finished, int32_t* step, int R, int K,
                                 int bos) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) *step = 0;
  if (r < R) {
    tok[size_t(r) * ld_tok] = bos;
    score[r] = (r % K == 0) ? 0.f : -INFINITY;    // one live hypothesis per utterance at step 0
    finished[r] = 0;
  }
}

struct BeamWs {
  GreedyWs g;
  f16* cache_alt;
  int32_t *tok_a, *tok_b, *parent, *token;
  float* score;
  void carve(Bump& b, const AsrConfig& c, int R, int Tp, int L) {
    g.carve(b, c, R, Tp, L);
    cache_alt = b.take<f16>(size_t(c.decoder_num_layers) * R * ((L + 15) / 16 * 16) * 2 * c.embedding_dim);
    tok_a = b.take<int32_t>(size_t(R) * (L + 1));
    tok_b = b.take<int32_t>(size_t(R) * (L + 1));
    parent = b.take<int32_t>(R);
    token = b.take<int32_t>(R);
    score = b.take<float>(R);
  }
};

}  // namespace

struct AsrHandle {
  AsrConfig cfg;
  AsrWeights w;
  std::vector<AsrEncoderLayerWeights> enc;
  std::vector<AsrDecoderLayerWeights> dec;
  bool loaded = false;
  cudaGraphExec_t graph_exec = nullptr;
  GraphKey graph_key;
  unsigned long long graph_kernels = 0;   // kernels per captured decode step
  cudaStream_t capture_stream = nullptr;  // the caller's stream may be the legacy default stream, which cannot capture
};

namespace {

int check_cfg(const AsrConfig& c) {
  if (c.vocab_size <= 0 || c.input_dim < 7 || c.embedding_dim <= 0 || c.decoder_seq_len <= 0 ||
      c.encoder_seq_len <= 0 || c.encoder_num_layers < 0 || c.decoder_num_layers < 0 || c.num_heads <= 0 ||
      c.ff_dim <= 0)
    return set_error(ASR_E_INVALID, "asr_create: non-positive dimension in config");
  if (c.embedding_dim != 64 * c.num_heads)
    return set_error(ASR_E_UNSUPPORTED, "unsupported config: head_dim = %d/%d, only head_dim 64 is implemented",
                     c.embedding_dim, c.num_heads);
  if (c.embedding_dim % 128 != 0 || c.embedding_dim > 1024)
    return set_error(ASR_E_UNSUPPORTED, "unsupported config: embedding_dim %d must be a multiple of 128, <= 1024",
                     c.embedding_dim);
  if (c.ff_dim % 64 != 0) return set_error(ASR_E_UNSUPPORTED, "unsupported config: ff_dim %d %% 64 != 0", c.ff_dim);
  return 0;
}

// X: the activation operand, [M, K] or, with split, [M, 2K] = [hi | lo]
int gemm(const f16* X, int ldx, const void* W, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s,
         int split = 0) {
  return launch_gemm_tc(X, ldx, static_cast<const f16*>(W), K, M, N, K, ep, s, split);
}

// h_out = X W^T + ep terms, then xn_out = LayerNorm(h_out): ONE launch when the full-row kernel supports the shape
// (N == 256), otherwise GEMM + LayerNorm kernel.  ln_f16: the next GEMM's A operand (hi | lo pair under g_split), ln_f32:
// fp32 LayerNorm output (the encoder's final norm).
int gemm_ln(const f16* X, int ldx, const void* W, int M, int N, int K, const GemmEpilogue& ep, const AsrNormWeights& nw,
            f16* ln_f16, float* ln_f32, cudaStream_t s, int split) {
  LnEpilogue ln;
  ln.gamma = nw.gamma; ln.beta = nw.beta; ln.out_f16 = ln_f16; ln.split = g_split; ln.out_f32 = ln_f32;
  const int rc = launch_gemm_ln(X, ldx, static_cast<const f16*>(W), K, M, N, K, ep, ln, s, split);
  if (rc <= 0) return rc;
  if (int rc2 = launch_gemm_tc(X, ldx, static_cast<const f16*>(W), K, M, N, K, ep, s, split)) return rc2;
  return launch_layernorm(ep.out_f32, nw.gamma, nw.beta, M, N, 1e-5f, ln_f32, ln_f16, s, g_split);
}

// x (f16, normalised) -> self attention block output added to the fp32 residual stream h (in place)
// next_ln: the LayerNorm that follows in the layer (fused into the out projection's epilogue), written to xn_out
int self_attention_block(const AsrConfig& c, const AsrMhaWeights& w, const f16* xn, f16* qkv, f16* att, float* h,
                         int B, int S, int causal, const uint8_t* valid, const int32_t* k_lens, cudaStream_t s,
                         const AsrNormWeights* next_ln = nullptr, f16* xn_out = nullptr) {
  const int D = c.embedding_dim, M = B * S, sp = SP();
  GemmEpilogue e1;
  e1.bias = w.b_qkv; e1.out_f16 = qkv; e1.ld_f16 = 3 * D;
  if (int rc = gemm(xn, sp * D, w.w_qkv, M, 3 * D, D, e1, s, g_split)) return rc;
  AttnParams a;
  a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D;
  a.ldq = a.ldk = a.ldv = 3 * D;
  a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = (long long)S * 3 * D;
  a.out = att; a.ldo = sp * D; a.o_batch_stride = (long long)S * sp * D; a.out_lo_off = g_split ? D : 0;
  a.B = B; a.H = c.num_heads; a.Sq = S; a.Sk = S;
  a.scale = 1.0f / sqrtf((float)D);
  a.causal = causal; a.q_valid = valid; a.k_valid = valid; a.k_lens = k_lens;
  if (int rc = launch_attention_tc(a, s)) return rc;
  GemmEpilogue e2;
  e2.bias = w.b_out; e2.residual = h; e2.ld_res = D; e2.out_f32 = h; e2.ld_f32 = D;
  if (next_ln) return gemm_ln(att, sp * D, w.w_out, M, D, D, e2, *next_ln, xn_out, nullptr, s, g_split);
  return gemm(att, sp * D, w.w_out, M, D, D, e2, s, g_split);
}

// h += FFN(xn) (layers.py:53-58, model.py:24,74).  next_ln (optional): the LayerNorm applied to the block's output, fused
// into the epilogue.  One launch (ffn_fused_kernel: the hidden activation stays on the SM) when D == 256, else two GEMMs.
int ffn_block(const AsrConfig& c, const AsrFfnWeights& w, const f16* xn, f16* ff, float* h, int M, cudaStream_t s,
              const AsrNormWeights* next_ln = nullptr, f16* xn_out = nullptr, float* ln_f32 = nullptr) {
  const int D = c.embedding_dim, FF = c.ff_dim, sp = SP();
  GemmEpilogue e2;
  e2.bias = w.b2; e2.residual = h; e2.ld_res = D; e2.out_f32 = h; e2.ld_f32 = D;
  {
    LnEpilogue ln;
    if (next_ln) {
      ln.gamma = next_ln->gamma; ln.beta = next_ln->beta; ln.out_f16 = xn_out; ln.split = g_split; ln.out_f32 = ln_f32;
    }
    const int rc = launch_ffn_fused(xn, sp * D, static_cast<const f16*>(w.w1), w.b1, static_cast<const f16*>(w.w2), M, D,
                                    FF, e2, ln, s, g_split);
    if (rc <= 0) return rc;
  }
  GemmEpilogue e1;
  e1.bias = w.b1; e1.relu = 1; e1.out_f16 = ff; e1.ld_f16 = sp * FF; e1.f16_lo_off = g_split ? FF : 0;
  if (int rc = gemm(xn, sp * D, w.w1, M, FF, D, e1, s, g_split)) return rc;
  if (next_ln) return gemm_ln(ff, sp * FF, w.w2, M, D, FF, e2, *next_ln, xn_out, ln_f32, s, g_split);
  return gemm(ff, sp * FF, w.w2, M, D, FF, e2, s, g_split);
}

// cross-attention K/V of every decoder layer, once per utterance: ckv[l] = enc * [Wk; Wv]^T + b  (f16 [M, 2D]).
// enc_out fp32 -> enc_sp f16 [M, 2D] = [hi | lo] (always split: the encoder output enters the K/V projections exactly)
int cross_kv(const AsrHandle* h, const float* enc_out, f16* enc_sp, f16* ckv, int M, cudaStream_t s) {
  const int D = h->cfg.embedding_dim;
  if (int rc = launch_f32_to_f16_split(enc_out, enc_sp, size_t(M), D, s)) return rc;
  for (int l = 0; l < h->cfg.decoder_num_layers; ++l) {
    const AsrMhaWeights& w = h->dec[l].cross_attn;
    GemmEpilogue e;
    e.bias = w.b_qkv + D;
    e.out_f16 = ckv + size_t(l) * M * 2 * D;
    e.ld_f16 = 2 * D;
    if (int rc = gemm(enc_sp, 2 * D, static_cast<const f16*>(w.w_qkv) + size_t(D) * D, M, 2 * D, D, e, s, 1)) return rc;
  }
  return 0;
}

int greedy_step(const AsrHandle* h, const GreedyWs& ws, int B, int Tp, int L, int stop_at_eos, int32_t* tokens,
                int32_t* n_tokens, float* step_logits, cudaStream_t s, StepProf* prof = nullptr, bool select = true) {
  const AsrConfig& c = h->cfg;
  const int D = c.embedding_dim, FF = c.ff_dim, H = c.num_heads;
  const float scale = 1.0f / sqrtf((float)D);
  const size_t M = size_t(B) * Tp;
  for (int l = 0; l < c.decoder_num_layers; ++l) {
    const AsrDecoderLayerWeights& w = h->dec[l];
    f16* cache = ws.cache + size_t(l) * B * L * 2 * D;
    const f16* ckv = ws.ckv + size_t(l) * M * 2 * D;
    {  // LN1 -> packed QKV; K/V rows appended to the cache (model.py:67-68, layers.py:16-18)
      DecLinear p;
      p.x = ws.h; p.ldx = D; p.ln_gamma = w.norm1.gamma; p.ln_beta = w.norm1.beta;
      p.w = static_cast<const f16*>(w.self_attn.w_qkv); p.bias = w.self_attn.b_qkv;
      p.B = B; p.N = 3 * D; p.K = D; p.out = ws.qkv; p.ldo = 3 * D;
      p.kv_cache = cache; p.kv_col0 = D; p.kv_rows = L; p.step = ws.step;
      PROF(prof, DC_QKV, launch_dec_linear(p, s));
    }
    {  // causal self attention over the cache rows 0..t
      DecAttn a;
      a.q = ws.qkv; a.ldq = 3 * D; a.k = cache; a.v = cache + D; a.ldkv = 2 * D;
      a.kv_batch_stride = (long long)L * 2 * D; a.n_keys = L; a.step = ws.step;
      a.out = ws.att; a.ldo = D; a.B = B; a.H = H; a.scale = scale;
      PROF(prof, DC_SELF_ATTN, launch_dec_attention(a, s));
    }
    {  // out projection + residual
      DecLinear p;
      p.x = ws.att; p.ldx = D; p.w = static_cast<const f16*>(w.self_attn.w_out); p.bias = w.self_attn.b_out;
      p.B = B; p.N = D; p.K = D; p.out = ws.h; p.ldo = D; p.residual = ws.h; p.ld_res = D;
      PROF(prof, DC_OUT_PROJ, launch_dec_linear(p, s));
    }
    {  // LN2 -> cross-attention query (model.py:70-71)
      DecLinear p;
      p.x = ws.h; p.ldx = D; p.ln_gamma = w.norm2.gamma; p.ln_beta = w.norm2.beta;
      p.w = static_cast<const f16*>(w.cross_attn.w_qkv); p.bias = w.cross_attn.b_qkv;
      p.B = B; p.N = D; p.K = D; p.out = ws.qc; p.ldo = D;
      PROF(prof, DC_CROSS_Q, launch_dec_linear(p, s));
    }
    {  // unmasked cross attention over the precomputed encoder K/V
      DecAttn a;
      a.q = ws.qc; a.ldq = D; a.k = ckv; a.v = ckv + D; a.ldkv = 2 * D;
      a.kv_batch_stride = (long long)Tp * 2 * D; a.n_keys = Tp; a.step = nullptr;
      a.out = ws.att; a.ldo = D; a.B = B; a.H = H; a.scale = scale;
      PROF(prof, DC_CROSS_ATTN, launch_dec_attention(a, s));
    }
    {
      DecLinear p;
      p.x = ws.att; p.ldx = D; p.w = static_cast<const f16*>(w.cross_attn.w_out); p.bias = w.cross_attn.b_out;
      p.B = B; p.N = D; p.K = D; p.out = ws.h; p.ldo = D; p.residual = ws.h; p.ld_res = D;
      PROF(prof, DC_OUT_PROJ, launch_dec_linear(p, s));
    }
    {  // LN3 -> FFN (model.py:73-74)
      DecLinear p;
      p.x = ws.h; p.ldx = D; p.ln_gamma = w.norm3.gamma; p.ln_beta = w.norm3.beta;
      p.w = static_cast<const f16*>(w.ffn.w1); p.bias = w.ffn.b1; p.relu = 1;
      p.B = B; p.N = FF; p.K = D; p.out = ws.ff; p.ldo = FF;
      PROF(prof, DC_FFN1, launch_dec_linear(p, s));
    }
    {
      DecLinear p;
      p.x = ws.ff; p.ldx = FF; p.w = static_cast<const f16*>(w.ffn.w2); p.bias = w.ffn.b2;
      p.B = B; p.N = D; p.K = FF; p.out = ws.h; p.ldo = D; p.residual = ws.h; p.ld_res = D;
      PROF(prof, DC_FFN2, launch_dec_linear(p, s));
    }
  }
  const int vpad = (c.vocab_size + 63) / 64 * 64;
  {  // classifier WITHOUT the final LayerNorm (model.py:142)
    DecLinear p;
    p.x = ws.h; p.ldx = D; p.w = static_cast<const f16*>(h->w.classifier_w);
    p.B = B; p.N = c.vocab_size; p.K = D; p.out = ws.logits; p.ldo = vpad;
    PROF(prof, DC_CLASSIFIER, launch_dec_linear(p, s));
  }
  if (!select) return 0;   // beam search picks the survivors itself (logits are in ws.logits)
  DecSelect sel;
  sel.logits = ws.logits; sel.ld = vpad; sel.V = c.vocab_size; sel.B = B;
  sel.tokens = tokens; sel.ld_tok = L + 1; sel.n_tokens = n_tokens; sel.finished = ws.finished; sel.step = ws.step;
  sel.step_logits = step_logits; sel.L = L; sel.eos = c.eos_token_id; sel.pad = c.pad_token_id;
  sel.stop_at_eos = stop_at_eos;
  PROF(prof, DC_SELECT, launch_dec_select_embed(sel, h->w.embedding, h->w.dec_pe, D, ws.h, s));
  return 0;
}

static bool cluster_available(const AsrHandle* h) {
  const AsrConfig& c = h->cfg;
  ClusterLayout lay;
  return h->w.dec_image && c.decoder_num_layers > 0 &&
         cluster_layout(c.embedding_dim, c.num_heads, c.ff_dim, c.vocab_size, c.decoder_num_layers, &lay) &&
         h->w.dec_image_bytes == lay.total_bytes;
}

static void build_cluster(const AsrHandle* h, const GreedyWs& w, int B, int Tp, int L, int stop_at_eos, int32_t* tokens,
                          int32_t* n_tokens, float* step_logits, ClusterParams& cp) {
  const AsrConfig& c = h->cfg;
  std::memset(&cp, 0, sizeof(cp));
  cp.B = B; cp.D = c.embedding_dim; cp.H = c.num_heads; cp.FF = c.ff_dim; cp.V = c.vocab_size; cp.L = L; cp.Tp = Tp;
  cp.nd = c.decoder_num_layers;
  cp.image = static_cast<const uint8_t*>(h->w.dec_image); cp.image_bytes = h->w.dec_image_bytes;
  cp.emb = h->w.embedding; cp.pe = h->w.dec_pe; cp.h0 = w.h;
  cp.cache = w.cache; cp.ckv = w.ckv;
  cp.tokens = tokens; cp.n_tokens = n_tokens; cp.step_logits = step_logits;
  cp.eos = c.eos_token_id; cp.pad = c.pad_token_id; cp.stop_at_eos = stop_at_eos;
  cp.scale = 1.0f / sqrtf((float)c.embedding_dim);
}

}  // namespace

extern "C" {

const char* asr_last_error(void) { return last_error(); }
int asr_version(void) { return 1; }

int asr_create(const AsrConfig* cfg, AsrHandle** out) {
  if (!cfg || !out) return set_error(ASR_E_INVALID, "asr_create: null argument");
  if (int rc = check_cfg(*cfg)) return rc;
  AsrHandle* h = new (std::nothrow) AsrHandle();
  if (!h) return set_error(ASR_E_INVALID, "asr_create: out of host memory");
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void asr_destroy(AsrHandle* h) {
  if (!h) return;
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  delete h;
}

int asr_load_weights(AsrHandle* h, const AsrWeights* w) {
  if (!h || !w) return set_error(ASR_E_INVALID, "asr_load_weights: null argument");
  if ((h->cfg.encoder_num_layers && !w->enc_layers) || (h->cfg.decoder_num_layers && !w->dec_layers))
    return set_error(ASR_E_INVALID, "asr_load_weights: missing layer tables");
  h->w = *w;
  h->enc.assign(w->enc_layers, w->enc_layers + h->cfg.encoder_num_layers);
  h->dec.assign(w->dec_layers, w->dec_layers + h->cfg.decoder_num_layers);
  h->w.enc_layers = h->enc.data();
  h->w.dec_layers = h->dec.data();
  h->loaded = true;
  if (h->graph_exec) {
    cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
  }
  return 0;
}

int asr_workspace_bytes(const AsrHandle* h, int B, int T, int L, size_t* bytes) {
  if (!h || !bytes || B < 0 || T < 7 || L <= 0) return set_error(ASR_E_INVALID, "asr_workspace_bytes: bad argument");
  const int Tp = conv_len(conv_len(T));
  size_t need = 0;
  { Bump b(nullptr); EncodeWs w; w.carve(b, h->cfg, B, T); need = b.off > need ? b.off : need; }
  { Bump b(nullptr); DecFwdWs w; w.carve(b, h->cfg, B, Tp, L); need = b.off > need ? b.off : need; }
  { Bump b(nullptr); GreedyWs w; w.carve(b, h->cfg, B, Tp, L); need = b.off > need ? b.off : need; }
  *bytes = need + 256;
  return 0;
}

// Transformer.input_layer: the fused kernel when its tile fits shared memory (ASR_B200_CONV=split forces the two-kernel
// path, kept as the fallback for very wide inputs and as the cross-check in the tests)
static int conv_frontend(const float* spec, const float* w1, const float* b1, const f16* w2frag, const float* b2, int B,
                         int F, int T, f16* y1, f16* z, cudaStream_t s) {
  // ASR_B200_CONV: "" (default) tcgen05 front-end where the shape allows, else the mma.sync fused kernel, else two kernels;
  // "fused" / "split" force the legacy paths (cross-checks in tests/test_ops_gpu.py)
  const char* e = std::getenv("ASR_B200_CONV");
  if (!(e && (e[0] == 's' || e[0] == 'f'))) {
    const int rc = launch_conv_tc(spec, w1, b1, w2frag, b2, B, F, T, z, s, g_split);
    if (rc <= 0) return rc;
  }
  if (!(e && e[0] == 's')) {
    const int rc = launch_conv_fused(spec, w1, b1, w2frag, b2, B, F, T, z, s, g_split);
    if (rc <= 0) return rc;
  }
  if (int rc = launch_conv1(spec, w1, b1, B, F, T, y1, s, g_split)) return rc;
  return launch_conv2(y1, w2frag, b2, B, conv_len(F), conv_len(T), z, s, g_split);
}

static int encoder_core(AsrHandle* h, const EncodeWs& w, int B, int T2, const int32_t* enc_lens, float* enc_out,
                        cudaStream_t s) {
  const AsrConfig& c = h->cfg;
  const int F2 = conv_len(conv_len(c.input_dim));
  const int D = c.embedding_dim, M = B * T2, Kin = 64 * F2;
  // Every LayerNorm of the pre-LN stack (model.py:20,23,52) rides in the epilogue of the GEMM that produces its input:
  // _lin_in -> norm1 of layer 0; out projection -> norm2; FFN unsqueeze -> norm1 of the next layer / _norm_out.
  const int nl = c.encoder_num_layers;
  {  // _lin_in + positional encoding (model.py:47)
    GemmEpilogue e;
    e.bias = h->w.lin_in_b; e.rowvec = h->w.enc_pe; e.rowvec_period = T2; e.ld_rowvec = D;
    e.out_f32 = w.h; e.ld_f32 = D;
    const AsrNormWeights& first = nl ? h->enc[0].norm1 : h->w.enc_norm_out;
    if (int rc = gemm_ln(w.z, SP() * Kin, h->w.lin_in_w, M, D, Kin, e, first, nl ? w.xn : nullptr, nl ? nullptr : enc_out, s,
                         g_split))
      return rc;
  }
  for (int l = 0; l < nl; ++l) {
    const AsrEncoderLayerWeights& lw = h->enc[l];
    if (int rc = self_attention_block(c, lw.attn, w.xn, w.qkv, w.att, w.h, B, T2, 0, nullptr, enc_lens, s, &lw.norm2, w.xn))
      return rc;
    const bool last = l + 1 == nl;
    if (int rc = ffn_block(c, lw.ffn, w.xn, w.ff, w.h, M, s, last ? &h->w.enc_norm_out : &h->enc[l + 1].norm1,
                           last ? nullptr : w.xn, last ? enc_out : nullptr))
      return rc;
  }
  return 0;
}

int asr_encode(AsrHandle* h, const float* spectrum, int B, int T, const int32_t* enc_lens, void* ws, size_t ws_bytes,
               float* enc_out, asr_stream_t stream) {
  if (!h || !h->loaded) return set_error(ASR_E_INVALID, "asr_encode: weights not loaded");
  if (B == 0) return 0;
  if (!spectrum || !enc_out || !ws || B < 0) return set_error(ASR_E_INVALID, "asr_encode: null argument");
  const AsrConfig& c = h->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T < 7) return set_error(ASR_E_INVALID, "asr_encode: T=%d too short for two stride-2 convs", T);
  const int F1 = conv_len(c.input_dim), T1 = conv_len(T), T2 = conv_len(T1);
  if (T2 > c.encoder_seq_len)
    return set_error(ASR_E_INVALID, "asr_encode: %d encoder frames exceed encoder_seq_len %d (positional buffer)", T2,
                     c.encoder_seq_len);
  Bump bump(ws);
  EncodeWs w;
  w.carve(bump, c, B, T);
  if (bump.off > ws_bytes) return set_error(ASR_E_WORKSPACE, "asr_encode: workspace %zu < %zu bytes", ws_bytes, bump.off);
  if (int rc = conv_frontend(spectrum, h->w.conv1_w, h->w.conv1_b, static_cast<const f16*>(h->w.conv2_wfrag),
                             h->w.conv2_b, B, c.input_dim, T, w.y1, w.z, s))
    return rc;
  return encoder_core(h, w, B, T2, enc_lens, enc_out, s);
}

int asr_encoder_forward(AsrHandle* h, const void* z_f16, int B, int Tp, const int32_t* enc_lens, void* ws,
                        size_t ws_bytes, float* enc_out, asr_stream_t stream) {
  if (!h || !h->loaded) return set_error(ASR_E_INVALID, "asr_encoder_forward: weights not loaded");
  if (B == 0) return 0;
  if (!z_f16 || !enc_out || !ws || B < 0 || Tp <= 0) return set_error(ASR_E_INVALID, "asr_encoder_forward: bad argument");
  const AsrConfig& c = h->cfg;
  if (Tp > c.encoder_seq_len)
    return set_error(ASR_E_INVALID, "asr_encoder_forward: %d frames exceed encoder_seq_len %d", Tp, c.encoder_seq_len);
  Bump bump(ws);
  EncodeWs w;
  w.carve(bump, c, B, 4 * Tp + 3);   // smallest T with subsampled length Tp; same layout as asr_encode
  if (bump.off > ws_bytes)
    return set_error(ASR_E_WORKSPACE, "asr_encoder_forward: workspace %zu < %zu bytes", ws_bytes, bump.off);
  w.z = const_cast<f16*>(static_cast<const f16*>(z_f16));
  return encoder_core(h, w, B, Tp, enc_lens, enc_out, static_cast<cudaStream_t>(stream));
}

int asr_decoder_forward(AsrHandle* h, const float* enc_out, int B, int Tp, const int32_t* text, const uint8_t* valid,
                        int L, void* ws, size_t ws_bytes, float* logits, asr_stream_t stream) {
  if (!h || !h->loaded) return set_error(ASR_E_INVALID, "asr_decoder_forward: weights not loaded");
  if (B == 0 || L == 0) return 0;
  if (!enc_out || !text || !logits || !ws || B < 0 || Tp <= 0 || L < 0)
    return set_error(ASR_E_INVALID, "asr_decoder_forward: bad argument");
  const AsrConfig& c = h->cfg;
  if (L > c.decoder_seq_len)
    return set_error(ASR_E_INVALID, "asr_decoder_forward: L=%d exceeds decoder_seq_len %d", L, c.decoder_seq_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Bump bump(ws);
  DecFwdWs w;
  w.carve(bump, c, B, Tp, L);
  if (bump.off > ws_bytes)
    return set_error(ASR_E_WORKSPACE, "asr_decoder_forward: workspace %zu < %zu bytes", ws_bytes, bump.off);
  const int D = c.embedding_dim, M = B * Tp, R = B * L;
  if (int rc = cross_kv(h, enc_out, w.enc_f16, w.ckv, M, s)) return rc;
  if (int rc = launch_embed_pe(text, L, h->w.embedding, h->w.dec_pe, B, L, D, c.vocab_size, w.h, s)) return rc;
  // norm1 of layer 0 (or, without layers, the final norm) is the only stand-alone LayerNorm: every other one rides in the
  // epilogue of the GEMM that produces its input (out projections -> norm2 / norm3, FFN -> next norm1 / _norm_layer)
  const int nd = c.decoder_num_layers;
  {
    const AsrNormWeights& first = nd ? h->dec[0].norm1 : h->w.dec_norm;
    if (int rc = launch_layernorm(w.h, first.gamma, first.beta, R, D, 1e-5f, nullptr, w.xn, s, g_split)) return rc;
  }
  for (int l = 0; l < nd; ++l) {
    const AsrDecoderLayerWeights& lw = h->dec[l];
    if (int rc = self_attention_block(c, lw.self_attn, w.xn, w.qkv, w.att, w.h, B, L, 1, valid, nullptr, s, &lw.norm2, w.xn))
      return rc;
    const int sp = SP();
    {  // cross attention: q from the decoder stream, K/V precomputed from the encoder output; never masked
      GemmEpilogue e1;
      e1.bias = lw.cross_attn.b_qkv; e1.out_f16 = w.qc; e1.ld_f16 = D;
      if (int rc = gemm(w.xn, sp * D, lw.cross_attn.w_qkv, R, D, D, e1, s, g_split)) return rc;
      const f16* ckv = w.ckv + size_t(l) * M * 2 * D;
      AttnParams a;
      a.q = w.qc; a.ldq = D; a.q_batch_stride = (long long)L * D;
      a.k = ckv; a.v = ckv + D; a.ldk = a.ldv = 2 * D;
      a.k_batch_stride = a.v_batch_stride = (long long)Tp * 2 * D;
      a.out = w.att; a.ldo = sp * D; a.o_batch_stride = (long long)L * sp * D; a.out_lo_off = g_split ? D : 0;
      a.B = B; a.H = c.num_heads; a.Sq = L; a.Sk = Tp; a.scale = 1.0f / sqrtf((float)D);
      if (int rc = launch_attention_tc(a, s)) return rc;
      GemmEpilogue e2;
      e2.bias = lw.cross_attn.b_out; e2.residual = w.h; e2.ld_res = D; e2.out_f32 = w.h; e2.ld_f32 = D;
      if (int rc = gemm_ln(w.att, sp * D, lw.cross_attn.w_out, R, D, D, e2, lw.norm3, w.xn, nullptr, s, g_split)) return rc;
    }
    const bool last = l + 1 == nd;
    if (int rc = ffn_block(c, lw.ffn, w.xn, w.ff, w.h, R, s, last ? &h->w.dec_norm : &h->dec[l + 1].norm1, w.xn)) return rc;
  }
  GemmEpilogue e;
  e.out_f32 = logits; e.ld_f32 = c.vocab_size; e.n_store = c.vocab_size;
  return gemm(w.xn, SP() * D, h->w.classifier_w, R, c.vocab_size, D, e, s, g_split);
}

// phases: 1 = prepare (cross-attention K/V of every layer, token / state initialisation), 2 = run (the decode loop on
// a prepared workspace), 3 = both
static int decode_greedy_impl(int phases, AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                              const int32_t* first_tokens, const int32_t* enc_lens, void* ws, size_t ws_bytes,
                              int32_t* tokens, int32_t* n_tokens, float* step_logits, asr_stream_t stream) {
  if (!h || !h->loaded) return set_error(ASR_E_INVALID, "asr_decode_greedy: weights not loaded");
  if (B == 0) return 0;
  if (!enc_out || !tokens || !ws || B < 0 || Tp <= 0 || L <= 0)
    return set_error(ASR_E_INVALID, "asr_decode_greedy: bad argument");
  const AsrConfig& c = h->cfg;
  if (L > c.decoder_seq_len)
    return set_error(ASR_E_INVALID, "asr_decode_greedy: L=%d exceeds decoder_seq_len %d", L, c.decoder_seq_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Bump bump(ws);
  GreedyWs w;
  w.carve(bump, c, B, Tp, L);
  if (bump.off > ws_bytes)
    return set_error(ASR_E_WORKSPACE, "asr_decode_greedy: workspace %zu < %zu bytes", ws_bytes, bump.off);
  const int D = c.embedding_dim, M = B * Tp;

  if (phases & 1) {
    if (int rc = cross_kv(h, enc_out, w.enc_f16, w.ckv, M, s)) return rc;
    dec_init_kernel<<<(B + 127) / 128, 128, 0, s>>>(tokens, L + 1, n_tokens, w.finished, w.step, B, L, c.bos_token_id,
                                                 first_tokens);
    ASR_CUDA_OK(cudaGetLastError());
    ASR_LAUNCHED(1);
    if (int rc = launch_dec_embed(tokens, L + 1, w.step, h->w.embedding, h->w.dec_pe, B, D, c.vocab_size, w.h, s))
      return rc;
  }
  if (!(phases & 2)) return 0;

  // Launch mode (ASR_B200_DECODE): "cluster" (default: one launch of dec_cluster_kernel for all L steps), "graph" (one
  // CUDA-graph replay of the per-kernel step per token; the fallback for configurations the cluster kernel is not
  // compiled for) or "eager" (same kernels, launched one by one; bring-up / profiling).
  const char* mode = std::getenv("ASR_B200_DECODE");
  if (mode && mode[0] == 'c' && !cluster_available(h))
    return set_error(ASR_E_UNSUPPORTED, "asr_decode_greedy: cluster decoder requested but unavailable (packed image "
                     "missing or config unsupported: heads must be 2, 4 or 8, ff_dim %% (32 heads) == 0)");
  if ((!mode || !mode[0] || mode[0] == 'c') && cluster_available(h)) {
    ClusterParams cp;
    build_cluster(h, w, B, Tp, L, stop_at_eos, tokens, n_tokens, step_logits, cp);
    cp.enc_lens = enc_lens;
    return launch_dec_cluster(cp, s);
  }
  if (enc_lens)
    return set_error(ASR_E_UNSUPPORTED, "asr_decode_greedy: encoder lengths (cross-attention key padding) need the "
                     "cluster decoder, which is unavailable for this configuration / ASR_B200_DECODE mode");
  if (mode && mode[0] == 'e') {
    for (int t = 0; t < L; ++t)
      if (int rc = greedy_step(h, w, B, Tp, L, stop_at_eos, tokens, n_tokens, step_logits, s)) return rc;
    return 0;
  }
  GraphKey key;
  key.ws = ws; key.enc = enc_out; key.tokens = tokens; key.n_tokens = n_tokens; key.step_logits = step_logits;
  key.first = first_tokens; key.B = B; key.Tp = Tp; key.L = L; key.stop = stop_at_eos;
  int first = 0;
  if (!h->graph_exec || !(h->graph_key == key)) {
    if (h->graph_exec) {
      cudaGraphExecDestroy(h->graph_exec);
      h->graph_exec = nullptr;
    }
    // step 0 runs eagerly (also performs every one-time kernel attribute setup outside of capture)
    if (int rc = greedy_step(h, w, B, Tp, L, stop_at_eos, tokens, n_tokens, step_logits, s)) return rc;
    first = 1;
    if (L > 1) {
      cudaGraph_t graph = nullptr;
      const unsigned long long before = g_kernel_launches;
      if (!h->capture_stream) ASR_CUDA_OK(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
      cudaStream_t cs = h->capture_stream;   // capture records, it does not execute: any capturable stream will do
      ASR_CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
      int rc = greedy_step(h, w, B, Tp, L, stop_at_eos, tokens, n_tokens, step_logits, cs);
      h->graph_kernels = g_kernel_launches - before;
      g_kernel_launches = before;          // captured, not launched
      cudaError_t ce = cudaStreamEndCapture(cs, &graph);
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      ASR_CUDA_OK(ce);
      ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      ASR_CUDA_OK(ce);
      h->graph_key = key;
    }
  }
  for (int t = first; t < L; ++t) ASR_CUDA_OK(cudaGraphLaunch(h->graph_exec, s));
  ASR_LAUNCHED(h->graph_kernels * (unsigned long long)(L - first));
  return 0;
}

int asr_decode_greedy(AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                      const int32_t* first_tokens, const int32_t* enc_lens, void* ws, size_t ws_bytes, int32_t* tokens,
                      int32_t* n_tokens, float* step_logits, asr_stream_t stream) {
  return decode_greedy_impl(3, h, enc_out, B, Tp, L, stop_at_eos, first_tokens, enc_lens, ws, ws_bytes, tokens, n_tokens,
                            step_logits, stream);
}

int asr_decode_prepare(AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                       const int32_t* first_tokens, const int32_t* enc_lens, void* ws, size_t ws_bytes, int32_t* tokens,
                       int32_t* n_tokens, float* step_logits, asr_stream_t stream) {
  return decode_greedy_impl(1, h, enc_out, B, Tp, L, stop_at_eos, first_tokens, enc_lens, ws, ws_bytes, tokens, n_tokens,
                            step_logits, stream);
}

int asr_decode_run(AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                   const int32_t* first_tokens, const int32_t* enc_lens, void* ws, size_t ws_bytes, int32_t* tokens,
                   int32_t* n_tokens, float* step_logits, asr_stream_t stream) {
  return decode_greedy_impl(2, h, enc_out, B, Tp, L, stop_at_eos, first_tokens, enc_lens, ws, ws_bytes, tokens, n_tokens,
                            step_logits, stream);
}

int asr_beam_workspace_bytes(const AsrHandle* h, int B, int beam, int T, int L, size_t* bytes) {
  if (!h || !bytes || B < 0 || beam < 1 || beam > 16 || T < 7 || L <= 0)
    return set_error(ASR_E_INVALID, "asr_beam_workspace_bytes: bad argument (1 <= beam <= 16)");
  Bump b(nullptr);
  BeamWs w;
  w.carve(b, h->cfg, B * beam, conv_len(conv_len(T)), L);
  *bytes = b.off + 256;
  return 0;
}

// Beam search on the KV-cached decode step (the reference lists it as a TODO, README.md:30, so
// the semantics are those of oracle/speech_transformer.py: beam_search_kv_cached).  enc_rep: the encoder output with
// every utterance repeated `beam` times, (B * beam, Tp, D) fp32 - hypotheses are ordinary batch rows of the step
// kernels.  tokens int32 (B, beam, L+1) best first; scores fp32 (B, beam) = sum of token log-probabilities.
int asr_decode_beam(AsrHandle* h, const float* enc_rep, int B, int beam, int Tp, int L, void* ws, size_t ws_bytes,
                    int32_t* tokens, float* scores, asr_stream_t stream) {
  if (!h || !h->loaded) return set_error(ASR_E_INVALID, "asr_decode_beam: weights not loaded");
  if (B == 0) return 0;
  if (!enc_rep || !tokens || !scores || !ws || B < 0 || beam < 1 || beam > 16 || Tp <= 0 || L <= 0)
    return set_error(ASR_E_INVALID, "asr_decode_beam: bad argument (1 <= beam <= 16)");
  const AsrConfig& c = h->cfg;
  if (L > c.decoder_seq_len)
    return set_error(ASR_E_INVALID, "asr_decode_beam: L=%d exceeds decoder_seq_len %d", L, c.decoder_seq_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int R = B * beam, D = c.embedding_dim, V = c.vocab_size;
  Bump bump(ws);
  BeamWs w;
  w.carve(bump, c, R, Tp, L);
  if (bump.off > ws_bytes)
    return set_error(ASR_E_WORKSPACE, "asr_decode_beam: workspace %zu < %zu bytes", ws_bytes, bump.off);
  const size_t smem = size_t(beam) * V * sizeof(float);
  if (smem > 200 * 1024) return set_error(ASR_E_UNSUPPORTED, "asr_decode_beam: beam x vocabulary too large");
  if (int rc = ensure_dyn_smem((const void*)beam_select_kernel, smem)) return rc;
  const int M = R * Tp;
  if (int rc = cross_kv(h, enc_rep, w.g.enc_f16, w.g.ckv, M, s)) return rc;
  int32_t* tok_cur = w.tok_a;
  int32_t* tok_nxt = w.tok_b;
  f16* cache_cur = w.g.cache;
  f16* cache_nxt = w.cache_alt;
  beam_init_kernel<<<(R + 127) / 128, 128, 0, s>>>(tok_cur, L + 1, w.score, w.g.finished, w.g.step, R, beam,
                                                  c.bos_token_id);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  if (int rc = launch_dec_embed(tok_cur, L + 1, w.g.step, h->w.embedding, h->w.dec_pe, R, D, V, w.g.h, s)) return rc;
  const int vpad = (V + 63) / 64 * 64;
  for (int t = 0; t < L; ++t) {
    GreedyWs g = w.g;
    g.cache = cache_cur;
    if (int rc = greedy_step(h, g, R, Tp, L, 0, nullptr, nullptr, nullptr, s, nullptr, /*select=*/false)) return rc;
    beam_select_kernel<<<B, 256, smem, s>>>(g.logits, vpad, V, beam, w.score, w.g.finished, w.parent, w.token,
                                            c.eos_token_id, c.pad_token_id);
    ASR_CUDA_OK(cudaGetLastError());
    beam_reorder_kernel<<<dim3(R, c.decoder_num_layers), 256, 0, s>>>(cache_cur, cache_nxt, tok_cur, tok_nxt, w.parent,
                                                                      w.token, w.g.step, R, beam, L, 2 * D, t);
    ASR_CUDA_OK(cudaGetLastError());
    ASR_LAUNCHED(2);
    std::swap(cache_cur, cache_nxt);
    std::swap(tok_cur, tok_nxt);
    if (t + 1 < L)
      if (int rc = launch_dec_embed(tok_cur, L + 1, w.g.step, h->w.embedding, h->w.dec_pe, R, D, V, w.g.h, s)) return rc;
  }
  ASR_CUDA_OK(cudaMemcpyAsync(tokens, tok_cur, size_t(R) * (L + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  ASR_CUDA_OK(cudaMemcpyAsync(scores, w.score, size_t(R) * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

// Same work as asr_decode_greedy, launched eagerly with a CUDA event pair around every kernel; synchronises the
// stream at the end and returns the summed device time (ms) and launch count per kernel class (DecClass order).
int asr_decode_profile(AsrHandle* h, const float* enc_out, int B, int Tp, int L, void* ws, size_t ws_bytes,
                       int32_t* tokens, float* ms_per_class, int32_t* launches_per_class, long long* phase_cycles,
                       asr_stream_t stream) {
  if (!h || !h->loaded || !enc_out || !tokens || !ws || !ms_per_class || !launches_per_class || B <= 0 || L <= 0)
    return set_error(ASR_E_INVALID, "asr_decode_profile: bad argument");
  const AsrConfig& c = h->cfg;
  if (L > c.decoder_seq_len) return set_error(ASR_E_INVALID, "asr_decode_profile: L exceeds decoder_seq_len");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Bump bump(ws);
  GreedyWs w;
  w.carve(bump, c, B, Tp, L);
  if (bump.off > ws_bytes) return set_error(ASR_E_WORKSPACE, "asr_decode_profile: workspace too small");
  const int D = c.embedding_dim, M = B * Tp;
  if (int rc = cross_kv(h, enc_out, w.enc_f16, w.ckv, M, s)) return rc;
  dec_init_kernel<<<(B + 127) / 128, 128, 0, s>>>(tokens, L + 1, nullptr, w.finished, w.step, B, L, c.bos_token_id,
                                               nullptr);
  ASR_CUDA_OK(cudaGetLastError());
  if (int rc = launch_dec_embed(tokens, L + 1, w.step, h->w.embedding, h->w.dec_pe, B, D, c.vocab_size, w.h, s))
    return rc;
  StepProf prof;
  prof.s = s;
  for (int t = 0; t < L; ++t)
    if (int rc = greedy_step(h, w, B, Tp, L, 0, tokens, nullptr, nullptr, s, &prof)) return rc;
  ASR_CUDA_OK(cudaStreamSynchronize(s));
  for (int i = 0; i <= DC_COUNT; ++i) {
    ms_per_class[i] = 0.f;
    launches_per_class[i] = 0;
  }
  for (size_t i = 0; i < prof.cls.size(); ++i) {
    float ms = 0.f;
    ASR_CUDA_OK(cudaEventElapsedTime(&ms, prof.ev[2 * i], prof.ev[2 * i + 1]));
    ms_per_class[prof.cls[i]] += ms;
    launches_per_class[prof.cls[i]] += 1;
  }
  // slot DC_COUNT: the cluster kernel (one cluster of num_heads CTAs per utterance group, DSMEM all-reduces)
  if (cluster_available(h)) {
    dec_init_kernel<<<(B + 127) / 128, 128, 0, s>>>(tokens, L + 1, nullptr, w.finished, w.step, B, L, c.bos_token_id,
                                                 nullptr);
    ASR_CUDA_OK(cudaGetLastError());
    if (int rc = launch_dec_embed(tokens, L + 1, w.step, h->w.embedding, h->w.dec_pe, B, D, c.vocab_size, w.h, s))
      return rc;
    ClusterParams cp;
    build_cluster(h, w, B, Tp, L, 0, tokens, nullptr, nullptr, cp);
    cp.timing = phase_cycles;
    cudaEvent_t e0, e1;
    ASR_CUDA_OK(cudaEventCreate(&e0));
    ASR_CUDA_OK(cudaEventCreate(&e1));
    ASR_CUDA_OK(cudaEventRecord(e0, s));
    int rc = launch_dec_cluster(cp, s);
    ASR_CUDA_OK(cudaEventRecord(e1, s));
    ASR_CUDA_OK(cudaStreamSynchronize(s));
    if (!rc) {
      ASR_CUDA_OK(cudaEventElapsedTime(&ms_per_class[DC_COUNT], e0, e1));
      launches_per_class[DC_COUNT] = 1;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
  }
  return 0;
}

size_t asr_decoder_image_bytes(const AsrConfig* cfg) {
  ClusterLayout lay;
  if (!cfg || !cluster_layout(cfg->embedding_dim, cfg->num_heads, cfg->ff_dim, cfg->vocab_size, cfg->decoder_num_layers,
                              &lay))
    return 0;
  return lay.total_bytes;
}

unsigned long long asr_launch_count(void) { return g_kernel_launches; }

int asr_split_operands(void) { return g_split; }

// ------------------------------------------------------------------------------------------- operators
int asr_layernorm(const float* x, const float* gamma, const float* beta, int rows, int D, float* y_f32, void* y_f16,
                  asr_stream_t stream) {
  if (rows == 0) return 0;
  if (!x || !gamma || !beta || (!y_f32 && !y_f16) || rows < 0) return set_error(ASR_E_INVALID, "asr_layernorm: bad argument");
  return launch_layernorm(x, gamma, beta, rows, D, 1e-5f, y_f32, static_cast<f16*>(y_f16), static_cast<cudaStream_t>(stream));
}

int asr_f32_to_f16(const float* x, void* y, size_t n, asr_stream_t stream) {
  if (n && (!x || !y)) return set_error(ASR_E_INVALID, "asr_f32_to_f16: null argument");
  return launch_f32_to_f16(x, static_cast<f16*>(y), n, static_cast<cudaStream_t>(stream));
}

int asr_gemm_f16(const void* x, const void* w, const float* bias, const float* residual, const float* pe,
                  int pe_period, int M, int N, int K, int relu, float* y_f32, void* y_f16, int impl,
                  asr_stream_t stream) {
  if (M == 0) return 0;
  if (!x || !w || (!y_f32 && !y_f16) || M < 0 || N <= 0 || K <= 0) return set_error(ASR_E_INVALID, "asr_gemm_f16: bad argument");
  GemmEpilogue e;
  e.bias = bias; e.residual = residual; e.ld_res = N; e.rowvec = pe; e.rowvec_period = pe_period > 0 ? pe_period : 1;
  e.ld_rowvec = N; e.out_f32 = y_f32; e.ld_f32 = N; e.out_f16 = static_cast<f16*>(y_f16); e.ld_f16 = N;
  e.relu = relu; e.n_store = N;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 1) return launch_gemm_naive(static_cast<const f16*>(x), K, static_cast<const f16*>(w), K, M, N, K, e, s);
  return launch_gemm_tc(static_cast<const f16*>(x), K, static_cast<const f16*>(w), K, M, N, K, e, s);
}

int asr_gemm_split(const float* x, const void* w, const float* bias, int M, int N, int K, int relu, float* y_f32,
                   void* y_f16_hilo, void* ws, size_t ws_bytes, asr_stream_t stream) {
  if (M == 0) return 0;
  if (!x || !w || (!y_f32 && !y_f16_hilo) || !ws || M < 0 || N <= 0 || K <= 0 || K % 8 != 0)
    return set_error(ASR_E_INVALID, "asr_gemm_split: bad argument");
  if (ws_bytes < size_t(M) * K * 4) return set_error(ASR_E_WORKSPACE, "asr_gemm_split: workspace < M * K * 4 bytes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  f16* xs = static_cast<f16*>(ws);
  if (int rc = launch_f32_to_f16_split(x, xs, size_t(M), K, s)) return rc;
  GemmEpilogue e;
  e.bias = bias; e.relu = relu; e.out_f32 = y_f32; e.ld_f32 = N; e.n_store = N;
  if (y_f16_hilo) {
    e.out_f16 = static_cast<f16*>(y_f16_hilo); e.ld_f16 = 2 * N; e.f16_lo_off = N;
  }
  return launch_gemm_tc(xs, 2 * K, static_cast<const f16*>(w), K, M, N, K, e, s, 1);
}

int asr_gemm_ln(const float* x, const void* w, const float* bias, const float* residual, const float* pe, int pe_period,
                const float* gamma, const float* beta, int M, int N, int K, float* h_out, float* y_f32, void* y_f16_hilo,
                void* ws, size_t ws_bytes, asr_stream_t stream) {
  if (M == 0) return 0;
  if (!x || !w || !gamma || !beta || (!y_f32 && !y_f16_hilo) || !ws || M < 0 || N <= 0 || K <= 0 || K % 8 != 0)
    return set_error(ASR_E_INVALID, "asr_gemm_ln: bad argument");
  if (ws_bytes < size_t(M) * K * 4) return set_error(ASR_E_WORKSPACE, "asr_gemm_ln: workspace < M * K * 4 bytes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  f16* xs = static_cast<f16*>(ws);
  if (int rc = launch_f32_to_f16_split(x, xs, size_t(M), K, s)) return rc;
  GemmEpilogue e;
  e.bias = bias; e.residual = residual; e.ld_res = N; e.rowvec = pe; e.rowvec_period = pe_period > 0 ? pe_period : 1;
  e.ld_rowvec = N; e.out_f32 = h_out; e.ld_f32 = N;
  LnEpilogue ln;
  ln.gamma = gamma; ln.beta = beta; ln.out_f32 = y_f32; ln.out_f16 = static_cast<f16*>(y_f16_hilo); ln.split = 1;
  const int rc = launch_gemm_ln(xs, 2 * K, static_cast<const f16*>(w), K, M, N, K, e, ln, s, 1);
  if (rc == 1) return set_error(ASR_E_UNSUPPORTED, "asr_gemm_ln: only N == 256 with 16-byte aligned operands is fused");
  return rc;
}

int asr_attention(const void* q, int ldq, long long q_bs, const void* k, int ldk, long long k_bs, const void* v,
                  int ldv, long long v_bs, void* out, int ldo, long long o_bs, int B, int H, int Sq, int Sk,
                  float scale, int causal, const int32_t* k_lens, const uint8_t* q_valid, const uint8_t* k_valid,
                  const uint8_t* dense_mask, int mask_B, int impl, asr_stream_t stream) {
  if (B == 0 || Sq == 0) return 0;
  AttnParams a;
  a.q = static_cast<const f16*>(q); a.ldq = ldq; a.q_batch_stride = q_bs;
  a.k = static_cast<const f16*>(k); a.ldk = ldk; a.k_batch_stride = k_bs;
  a.v = static_cast<const f16*>(v); a.ldv = ldv; a.v_batch_stride = v_bs;
  a.out = static_cast<f16*>(out); a.ldo = ldo; a.o_batch_stride = o_bs;
  a.B = B; a.H = H; a.Sq = Sq; a.Sk = Sk; a.scale = scale; a.causal = causal;
  a.k_lens = k_lens; a.q_valid = q_valid; a.k_valid = k_valid; a.dense_mask = dense_mask; a.mask_B = mask_B;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return impl == 1 ? launch_attention_naive(a, s) : launch_attention_tc(a, s);
}

size_t asr_mha_workspace_bytes(int B, int Sq, int Sk, int D) {
  const size_t rq = size_t(B) * Sq, rk = size_t(B) * Sk;
  return (2 * rq * D + 2 * rk * D + rq * 3 * D + rk * 2 * D + 2 * rq * D) * 2 + 8 * 256;
}

int asr_mha(const float* x, const float* src, const AsrMhaWeights* w, int B, int Sq, int Sk, int D, int H, int causal,
            const uint8_t* q_valid, const uint8_t* k_valid, const uint8_t* dense_mask, int mask_B, void* ws,
            size_t ws_bytes, float* out, asr_stream_t stream) {
  if (B == 0 || Sq == 0) return 0;
  if (!x || !w || !out || !ws || B < 0 || Sq < 0) return set_error(ASR_E_INVALID, "asr_mha: bad argument");
  if (D != 64 * H || D % 128 != 0) return set_error(ASR_E_UNSUPPORTED, "asr_mha: only head_dim 64 and D %% 128 == 0 (D=%d, H=%d)", D, H);
  if (!src) Sk = Sq;
  if (Sk <= 0) return set_error(ASR_E_INVALID, "asr_mha: empty key sequence");
  if (ws_bytes < asr_mha_workspace_bytes(B, Sq, Sk, D)) return set_error(ASR_E_WORKSPACE, "asr_mha: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Rq = B * Sq, Rk = B * Sk, sp = SP();
  Bump b(ws);
  f16* xb = b.take<f16>(size_t(Rq) * D * 2);
  f16* sb = b.take<f16>(size_t(Rk) * D * 2);
  f16* qkv = b.take<f16>(size_t(Rq) * 3 * D);
  f16* kv = b.take<f16>(size_t(Rk) * 2 * D);
  f16* att = b.take<f16>(size_t(Rq) * D * 2);
  // the inputs enter the projections as fp16 hi | lo pairs (or single fp16 with ASR_B200_SPLIT=0), like in the model path
  auto to_operand = [&](const float* src_f32, f16* dst, int rows) {
    return g_split ? launch_f32_to_f16_split(src_f32, dst, size_t(rows), D, s)
                   : launch_f32_to_f16(src_f32, dst, size_t(rows) * D, s);
  };
  if (int rc = to_operand(x, xb, Rq)) return rc;
  AttnParams a;
  if (!src) {
    GemmEpilogue e;
    e.bias = w->b_qkv; e.out_f16 = qkv; e.ld_f16 = 3 * D;
    if (int rc = launch_gemm_tc(xb, sp * D, static_cast<const f16*>(w->w_qkv), D, Rq, 3 * D, D, e, s, g_split)) return rc;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D;
    a.ldq = a.ldk = a.ldv = 3 * D;
    a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = (long long)Sq * 3 * D;
  } else {
    if (int rc = to_operand(src, sb, Rk)) return rc;
    GemmEpilogue e;
    e.bias = w->b_qkv; e.out_f16 = qkv; e.ld_f16 = D;
    if (int rc = launch_gemm_tc(xb, sp * D, static_cast<const f16*>(w->w_qkv), D, Rq, D, D, e, s, g_split)) return rc;
    GemmEpilogue e2;
    e2.bias = w->b_qkv + D; e2.out_f16 = kv; e2.ld_f16 = 2 * D;
    if (int rc = launch_gemm_tc(sb, sp * D, static_cast<const f16*>(w->w_qkv) + size_t(D) * D, D, Rk, 2 * D, D, e2, s,
                                g_split))
      return rc;
    a.q = qkv; a.ldq = D; a.q_batch_stride = (long long)Sq * D;
    a.k = kv; a.v = kv + D; a.ldk = a.ldv = 2 * D;
    a.k_batch_stride = a.v_batch_stride = (long long)Sk * 2 * D;
  }
  a.out = att; a.ldo = sp * D; a.o_batch_stride = (long long)Sq * sp * D; a.out_lo_off = g_split ? D : 0;
  a.B = B; a.H = H; a.Sq = Sq; a.Sk = Sk; a.scale = 1.0f / sqrtf((float)D); a.causal = causal;
  a.q_valid = q_valid; a.k_valid = k_valid; a.dense_mask = dense_mask; a.mask_B = mask_B;
  if (int rc = launch_attention_tc(a, s)) return rc;
  GemmEpilogue e3;
  e3.bias = w->b_out; e3.out_f32 = out; e3.ld_f32 = D;
  return launch_gemm_tc(att, sp * D, static_cast<const f16*>(w->w_out), D, Rq, D, D, e3, s, g_split);
}

size_t asr_ffn_workspace_bytes(int rows, int D, int FF) { return (size_t(rows) * D + size_t(rows) * FF) * 4 + 4 * 256; }

int asr_ffn(const float* x, const AsrFfnWeights* w, int rows, int D, int FF, void* ws, size_t ws_bytes, float* out,
            asr_stream_t stream) {
  if (rows == 0) return 0;
  if (!x || !w || !out || !ws || rows < 0) return set_error(ASR_E_INVALID, "asr_ffn: bad argument");
  if (D % 64 != 0 || FF % 64 != 0) return set_error(ASR_E_UNSUPPORTED, "asr_ffn: D and FF must be multiples of 64");
  if (ws_bytes < asr_ffn_workspace_bytes(rows, D, FF)) return set_error(ASR_E_WORKSPACE, "asr_ffn: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int sp = SP();
  Bump b(ws);
  f16* xb = b.take<f16>(size_t(rows) * D * 2);
  f16* ff = b.take<f16>(size_t(rows) * FF * 2);
  if (int rc = g_split ? launch_f32_to_f16_split(x, xb, size_t(rows), D, s) : launch_f32_to_f16(x, xb, size_t(rows) * D, s))
    return rc;
  GemmEpilogue e2;
  e2.bias = w->b2; e2.out_f32 = out; e2.ld_f32 = D;
  {
    const int rc = launch_ffn_fused(xb, sp * D, static_cast<const f16*>(w->w1), w->b1, static_cast<const f16*>(w->w2), rows, D,
                                    FF, e2, LnEpilogue(), s, g_split);
    if (rc <= 0) return rc;
  }
  GemmEpilogue e1;
  e1.bias = w->b1; e1.relu = 1; e1.out_f16 = ff; e1.ld_f16 = sp * FF; e1.f16_lo_off = g_split ? FF : 0;
  if (int rc = launch_gemm_tc(xb, sp * D, static_cast<const f16*>(w->w1), D, rows, FF, D, e1, s, g_split)) return rc;
  return launch_gemm_tc(ff, sp * FF, static_cast<const f16*>(w->w2), FF, rows, D, FF, e2, s, g_split);
}

size_t asr_conv_workspace_bytes(int B, int F, int T) {
  return size_t(B) * conv_len(T) * conv_len(F) * 64 * 2 * 2 + 512;
}

int asr_conv_frontend(const float* spectrum, const float* conv1_w, const float* conv1_b, const void* conv2_wfrag,
                      const float* conv2_b, int B, int F, int T, void* ws, size_t ws_bytes, void* z_f16,
                      asr_stream_t stream) {
  if (B == 0) return 0;
  if (!spectrum || !conv1_w || !conv1_b || !conv2_wfrag || !conv2_b || !ws || !z_f16 || B < 0)
    return set_error(ASR_E_INVALID, "asr_conv_frontend: bad argument");
  if (conv_len(conv_len(F)) < 1 || conv_len(conv_len(T)) < 1) return set_error(ASR_E_INVALID, "asr_conv_frontend: input too small");
  if (ws_bytes < asr_conv_workspace_bytes(B, F, T)) return set_error(ASR_E_WORKSPACE, "asr_conv_frontend: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Bump b(ws);
  f16* y1 = b.take<f16>(size_t(B) * conv_len(T) * conv_len(F) * 64 * 2);
  return conv_frontend(spectrum, conv1_w, conv1_b, static_cast<const f16*>(conv2_wfrag), conv2_b, B, F, T, y1,
                       static_cast<f16*>(z_f16), s);
}

int asr_spectrogram(const float* audio, int B, int n_samples, int n_fft, int hop, int T, float* spec,
                    asr_stream_t stream) {
  if (B == 0) return 0;
  if (!audio || !spec || B < 0 || n_samples < 0) return set_error(ASR_E_INVALID, "asr_spectrogram: bad argument");
  return launch_spectrogram(audio, B, n_samples, n_fft, hop, T, spec, static_cast<cudaStream_t>(stream));
}

int asr_embed_pe(const int32_t* tokens, const float* emb, const float* pe, int B, int L, int D, int vocab, float* out,
                 asr_stream_t stream) {
  if (B == 0 || L == 0) return 0;
  if (!tokens || !emb || !pe || !out || D % 4 != 0) return set_error(ASR_E_INVALID, "asr_embed_pe: bad argument");
  return launch_embed_pe(tokens, L, emb, pe, B, L, D, vocab, out, static_cast<cudaStream_t>(stream));
}

int asr_dec_linear(const float* x, const float* ln_gamma, const float* ln_beta, const void* w, const float* bias,
                   const float* residual, int B, int N, int K, int relu, float* out, asr_stream_t stream) {
  if (B == 0) return 0;
  if (!x || !w || !out) return set_error(ASR_E_INVALID, "asr_dec_linear: null argument");
  DecLinear p;
  p.x = x; p.ldx = K; p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.w = static_cast<const f16*>(w); p.bias = bias;
  p.B = B; p.N = N; p.K = K; p.relu = relu; p.out = out; p.ldo = N; p.residual = residual; p.ld_res = N;
  return launch_dec_linear(p, static_cast<cudaStream_t>(stream));
}

int asr_dec_attention(const float* q, const void* k, const void* v, int ldkv, long long kv_bs, int n_keys, int B,
                      int H, float scale, float* out, asr_stream_t stream) {
  if (B == 0) return 0;
  if (!q || !k || !v || !out || n_keys <= 0) return set_error(ASR_E_INVALID, "asr_dec_attention: bad argument");
  DecAttn a;
  a.q = q; a.ldq = H * 64; a.k = static_cast<const f16*>(k); a.v = static_cast<const f16*>(v); a.ldkv = ldkv;
  a.kv_batch_stride = kv_bs; a.n_keys = n_keys; a.out = out; a.ldo = H * 64; a.B = B; a.H = H; a.scale = scale;
  return launch_dec_attention(a, static_cast<cudaStream_t>(stream));
}

int asr_umma_probe(const void* a, const void* b, float* d, int N, int b_mn_major, asr_stream_t stream) {
  if (!a || !b || !d) return set_error(ASR_E_INVALID, "asr_umma_probe: null argument");
  return launch_umma_probe(static_cast<const f16*>(a), static_cast<const f16*>(b), d, N, b_mn_major,
                           static_cast<cudaStream_t>(stream));
}

}  // extern "C"
