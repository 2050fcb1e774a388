/* asr_b200.h - C ABI of the B200-native Speech-Transformer forward / greedy-decode hot path.
 *
 * The reference (shockless/asr-transformer) is pure Python: its boundary for this path is the nn.Module surface
 * of modules/Transformer/model.py and layers.py, not an FFI.  This header is the C-ABI directly underneath our
 * drop-in Python modules (asr_transformer_b200/model.py, layers.py); every entry point names the reference call
 * it replaces.  Plain pointers and sizes only: no torch types cross this boundary.
 *
 * Conventions
 *   - every pointer is DEVICE memory owned by the caller unless stated otherwise; the library allocates nothing
 *     on the device: scratch comes from the caller-provided workspace (ws, ws_bytes);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit synchronisation, no host
 *     sync inside asr_decode_greedy (argmax / EOS bookkeeping stay on the device);
 *   - return value 0 = OK, negative = error; asr_last_error() gives the message (thread local);
 *     nothing throws or aborts across the ABI;
 *   - fp32 tensors are row-major contiguous; "f16" buffers are IEEE binary16 (fp16).  Every tensor-core operand of
 *     this path is fp16 with fp32 accumulation: 3 more mantissa bits than bf16 at the same bytes, which is what the
 *     99 %-identical-tokens bar needs; weights that are bf16-representable are exact in fp16
 *     (values below 2^-14 round with an absolute error <= 2^-25); stores saturate to +-65504 instead of overflowing;
 *   - supported shapes: head_dim == 64 (embedding_dim = 64 * num_heads), embedding_dim % 128 == 0,
 *     ff_dim % 64 == 0.  Anything else returns ASR_E_UNSUPPORTED.
 */
#ifndef ASR_B200_H_
#define ASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_OK 0
#define ASR_E_INVALID (-1)
#define ASR_E_UNSUPPORTED (-2)
#define ASR_E_WORKSPACE (-3)
#define ASR_E_CUDA (-100)

typedef struct AsrHandle AsrHandle;
typedef void* asr_stream_t; /* cudaStream_t */

/* ctor arguments of reference Transformer.__init__ (model.py:155-166) + the BOS id of train.py:70 */
typedef struct {
  int32_t vocab_size, input_dim, embedding_dim, decoder_seq_len, encoder_seq_len;
  int32_t encoder_num_layers, decoder_num_layers, num_heads, ff_dim;
  int32_t pad_token_id, eos_token_id, bos_token_id;
} AsrConfig;

/* One multi-head attention block (reference layers.py:31-40), heads packed in index order.
 * w_qkv: f16 [3D, D] = rows [q heads | k heads | v heads]; b_qkv fp32 [3D]; w_out f16 [D, D]; b_out fp32 [D]. */
typedef struct {
  const void* w_qkv;
  const float* b_qkv;
  const void* w_out;
  const float* b_out;
} AsrMhaWeights;

typedef struct { const float* gamma; const float* beta; } AsrNormWeights;       /* nn.LayerNorm, eps 1e-5 */

/* reference layers.py:43-58: w1 = squeeze f16 [FF, D], w2 = unsqueeze f16 [D, FF] */
typedef struct { const void* w1; const float* b1; const void* w2; const float* b2; } AsrFfnWeights;

typedef struct {   /* reference model.py:9-25 (_norm_in is dead and not passed) */
  AsrNormWeights norm1; AsrMhaWeights attn; AsrNormWeights norm2; AsrFfnWeights ffn;
} AsrEncoderLayerWeights;

typedef struct {   /* reference model.py:55-75 */
  AsrNormWeights norm1; AsrMhaWeights self_attn; AsrNormWeights norm2; AsrMhaWeights cross_attn;
  AsrNormWeights norm3; AsrFfnWeights ffn;
} AsrDecoderLayerWeights;

typedef struct {
  /* conv front-end, reference model.py:168-171 */
  const float* conv1_w;      /* fp32 [9 taps (kh*3+kw)][64 out channels] */
  const float* conv1_b;      /* fp32 [64] */
  const void* conv2_wfrag;   /* f16 mma-fragment packing of input_layer.2.weight, 36*8*32*4 f16 (packing.py) */
  const float* conv2_b;      /* fp32 [64] */
  /* encoder, reference model.py:28-52 */
  const void* lin_in_w;      /* f16 [D, 64*F'] with columns permuted from (c*F'+f) to (f*64+c) */
  const float* lin_in_b;     /* fp32 [D] */
  const float* enc_pe;       /* fp32 [encoder_seq_len, D] (buffer encoder._pe.pe) */
  const AsrEncoderLayerWeights* enc_layers;   /* HOST array [encoder_num_layers] */
  AsrNormWeights enc_norm_out;
  /* decoder, reference model.py:78-151 */
  const float* embedding;    /* fp32 [vocab, D] */
  const float* dec_pe;       /* fp32 [decoder_seq_len, D] */
  const AsrDecoderLayerWeights* dec_layers;   /* HOST array [decoder_num_layers] */
  AsrNormWeights dec_norm;   /* decoder._norm_layer (applied by forward, skipped by evaluate: model.py:122 vs 142) */
  const void* classifier_w;  /* f16 [round_up(vocab, 64), D], rows >= vocab zero; no bias (model.py:102) */
  /* optional (NULL disables the cluster decoder): the decoder weights re-packed for the cluster kernel, one contiguous
   * block per head r (= CTA rank in a cluster of num_heads CTAs), consumed front to back by TMA bulk copies.
   * Requires num_heads in {2,4,8}, ff_dim % (32 num_heads) == 0; FFS = ff_dim / num_heads, VS = round_up(ceil(V /
   * num_heads), 16).  Per head r: for each layer
   *   small : fp32 b_qkv rows of head r (q|k|v: 192) | cross b_q rows of head r (64) | b1[r FFS..] (FFS) | b_out (D) |
   *           cross b_out (D) | b2 (D) | norm1 g,b | norm2 g,b | norm3 g,b | norm1 g,b of layer l+1 (zeros after the last
   *           layer) (8 D); zero padded to a multiple of 128 B
   *   Wqkv rows of head r (192 x D) | Wout[:, 64 r..] (D x 64) | cross Wq rows of head r (64 x D) |
   *   cross Wout[:, 64 r..] (D x 64) | W1[r FFS.., :] (FFS x D) | W2[:, r FFS..] (D x FFS)
   * then the classifier rows [r VS, (r+1) VS) (zero padded past vocab).  Every matrix [R x K] is stored f16 in
   * mma.m16n8k16 A-fragment order [K/32 k-blocks][R/16 m-tiles][k-tile s 0..1][g 0..7][tg 0..3][8]: with r = 16 mt + 2 g,
   * c = 32 kb + 8 tg + 4 s the 8 elements are W[r][c..c+1], W[r+1][c..c+1], W[r][c+2..c+3], W[r+1][c+2..c+3].
   * asr_decoder_image_bytes() gives the total size (0 = shape not compiled into the cluster kernel). */
  const void* dec_image;
  size_t dec_image_bytes;
} AsrWeights;

const char* asr_last_error(void);
int asr_version(void);

/* ---- model-level entry points ------------------------------------------------------------------------- */
int asr_create(const AsrConfig* cfg, AsrHandle** out);
void asr_destroy(AsrHandle* h);
/* Copies the pointer table (not the weights); the device buffers must outlive the handle's use. */
int asr_load_weights(AsrHandle* h, const AsrWeights* w);
/* Scratch needed by any of the calls below for batch B, T input frames, decode length L. */
int asr_workspace_bytes(const AsrHandle* h, int B, int T, int L, size_t* bytes);

/* Replaces Transformer.input_layer + Encoder.forward (model.py:203-204 / 41-52).
 * spectrum fp32 (B,1,F,T) -> enc_out fp32 (B,T',D).  enc_lens (nullable, [B], in encoder frames T') enables the
 * key-padding mask the reference root encoder lacks; pass NULL for reference parity. */
int asr_encode(AsrHandle* h, const float* spectrum, int B, int T, const int32_t* enc_lens, void* ws, size_t ws_bytes,
               float* enc_out, asr_stream_t stream);

/* Encoder.forward alone (model.py:41-52) on front-end features already in the kernel layout:
 * z f16 (B,T',S*F'*64), column = f*64 + c [+ F'*64 for the lo parts], S = 1 + asr_split_operands()  (what
 * asr_conv_frontend produces). */
int asr_encoder_forward(AsrHandle* h, const void* z_f16, int B, int Tp, const int32_t* enc_lens, void* ws,
                        size_t ws_bytes, float* enc_out, asr_stream_t stream);

/* Replaces Decoder.forward (model.py:104-123): teacher-forced logits fp32 (B,L,V).
 * text int32 (B,L); valid uint8 (B,L), nonzero = real token (the reference's mask >= 1). */
int asr_decoder_forward(AsrHandle* h, const float* enc_out, int B, int Tp, const int32_t* text, const uint8_t* valid,
                        int L, void* ws, size_t ws_bytes, float* logits, asr_stream_t stream);

/* Beam search on the KV-cached decode step.  The reference has no beam search (README.md:30 lists
 * it as a TODO), so the semantics are defined by oracle/speech_transformer.py: beam_search_kv_cached: candidate score =
 * hypothesis score + log_softmax(logits without the final LayerNorm, model.py:142), no length normalisation, finished
 * hypotheses (EOS emitted) are carried with pad tokens, ties go to the lower (beam, token) index, exactly L steps.
 * enc_rep fp32 (B*beam, Tp, D): the encoder output with every utterance repeated `beam` times; tokens int32
 * (B, beam, L+1) best first; scores fp32 (B, beam).  1 <= beam <= 16; beam == 1 is greedy search that pads after EOS. */
int asr_beam_workspace_bytes(const AsrHandle* h, int B, int beam, int T, int L, size_t* bytes);
int asr_decode_beam(AsrHandle* h, const float* enc_rep, int B, int beam, int Tp, int L, void* ws, size_t ws_bytes,
                    int32_t* tokens, float* scores, asr_stream_t stream);

/* Replaces Decoder.evaluate (model.py:125-151) for the whole batch at once with a device-resident KV cache.
 * tokens int32 (B,L+1) (column 0 = BOS); n_tokens int32 [B] (nullable) = tokens up to and including the first EOS
 * (L+1 if none); step_logits fp32 (B,L,V) nullable: logits (no final LayerNorm, model.py:142) that chose
 * tokens[:,t+1].  stop_at_eos == 0 reproduces the reference (exactly L steps, tokens keep flowing after EOS);
 * stop_at_eos != 0 writes pad_token_id after the first EOS (and a cluster leaves early once all of its utterances
 * have finished).  Launch mode is selected by the environment variable ASR_B200_DECODE: "cluster" (default when
 * AsrWeights.dec_image is set: one thread-block cluster of num_heads CTAs per group of utterances, head-parallel
 * layers, all-reduces over distributed shared memory, everything streamed by TMA), "graph" (CUDA-graph replay of the
 * per-kernel step; the fallback for configurations the cluster kernel is not compiled for) or "eager". */
int asr_decode_greedy(AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                      const int32_t* first_tokens /* [B] nullable: defaults to bos_token_id */,
                      const int32_t* enc_lens /* [B] nullable: valid encoder frames per utterance; cross attention
                                                 masks the rest (the key-padding mask the reference lacks, Q6) */,
                      void* ws, size_t ws_bytes, int32_t* tokens, int32_t* n_tokens, float* step_logits,
                      asr_stream_t stream);

/* asr_decode_greedy in two halves, for pipelined serving: asr_decode_prepare computes the cross-attention K/V of every
 * layer from enc_out and initialises tokens / decoder state in the workspace; asr_decode_run runs the decode loop on a
 * workspace prepared with the same arguments.  The halves may be enqueued on different streams (ordered by an event),
 * so the prepare of batch i+1 can overlap the run of batch i on a second workspace. */
int asr_decode_prepare(AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                       const int32_t* first_tokens, const int32_t* enc_lens, void* ws, size_t ws_bytes, int32_t* tokens,
                       int32_t* n_tokens, float* step_logits, asr_stream_t stream);
int asr_decode_run(AsrHandle* h, const float* enc_out, int B, int Tp, int L, int stop_at_eos,
                   const int32_t* first_tokens, const int32_t* enc_lens, void* ws, size_t ws_bytes, int32_t* tokens,
                   int32_t* n_tokens, float* step_logits, asr_stream_t stream);

/* Profiling aid for bench.py: the same decode launched eagerly with a CUDA-event pair around every kernel.
 * Synchronises `stream` before returning.  ms_per_class / launches_per_class: 10 entries in the order
 * {qkv linear, self attention, out projections, cross q linear, cross attention, ffn1, ffn2, classifier, select}
 * for the per-kernel step, entry 9 = the cluster kernel (all L steps in one launch).  phase_cycles: device
 * [gridDim of the cluster launch (<= 148)][16] int64. */
int asr_decode_profile(AsrHandle* h, const float* enc_out, int B, int Tp, int L, void* ws, size_t ws_bytes,
                       int32_t* tokens, float* ms_per_class, int32_t* launches_per_class,
                       long long* phase_cycles /* nullable: per-CTA clock64 totals of the cluster decoder */,
                       asr_stream_t stream);
/* Size of AsrWeights.dec_image for a configuration (0 = the cluster decoder does not support it). */
size_t asr_decoder_image_bytes(const AsrConfig* cfg);
/* Number of kernels this library has launched in the calling process (graph replays counted per kernel). */
unsigned long long asr_launch_count(void);
/* 1 (default): every activation that feeds a Linear layer or conv2 as a tensor-core operand travels as an fp16
 * hi | lo pair (x = hi + lo to 2^-22) and is multiplied twice, so the linear layers see fp32-accurate inputs (what the
 * 99 %-identical-tokens bar needs); rows of such buffers are [hi (K) | lo (K)].  0 when the process
 * runs with ASR_B200_SPLIT=0 (single fp16 operands: half the MMAs, for A/B measurements).  Affects the layout of the
 * front-end features z below: (B, T', 2 F' 64) = [hi | lo] per row when 1, (B, T', F' 64) when 0. */
int asr_split_operands(void);

/* ---- operator-level entry points (sub-module drop-ins and parity tests) --------------------------------- */
/* nn.LayerNorm call sites (model.py:20,23,52,67,70,73,122). y_f32 / y_f16 nullable. */
int asr_layernorm(const float* x, const float* gamma, const float* beta, int rows, int D, float* y_f32, void* y_f16,
                  asr_stream_t stream);
int asr_f32_to_f16(const float* x, void* y_f16, size_t n, asr_stream_t stream);
/* Y[M,N] = X[M,K] W[N,K]^T (+bias)(ReLU)(+pe[row % period])(+residual); x, w f16; w holds round_up(N,64) rows.
 * impl: 0 = tcgen05 tensor-core kernel (the product path), 1 = CUDA-core cross-check kernel (tests only). */
int asr_gemm_f16(const void* x, const void* w, const float* bias, const float* residual, const float* pe,
                  int pe_period, int M, int N, int K, int relu, float* y_f32, void* y_f16, int impl,
                  asr_stream_t stream);
/* The split-operand GEMM of the model path (asr_split_operands): x fp32 [M,K] is split into fp16 hi | lo halves (ws:
 * M*K*4 bytes) and both are multiplied by w f16 [round_up(N,64), K]: Y = x W^T (+bias)(ReLU) to fp32 accuracy.
 * y_f32 [M,N] and / or y_f16_hilo [M, 2N] = [hi | lo] of Y (the operand layout of a following split GEMM). */
int asr_gemm_split(const float* x, const void* w, const float* bias, int M, int N, int K, int relu, float* y_f32,
                   void* y_f16_hilo, void* ws, size_t ws_bytes, asr_stream_t stream);
/* The full-row GEMM with the following LayerNorm fused into its epilogue (the model path's replacement of every
 * nn.LayerNorm launch but the first; N == 256 only, else ASR_E_UNSUPPORTED): h = x W^T + bias (+pe[row % period])
 * (+residual) -> h_out fp32 [M,N] (nullable; may alias residual); y = LayerNorm(h) gamma + beta -> y_f32 [M,N] and / or
 * y_f16_hilo [M,2N].  x fp32 [M,K] enters as an fp16 hi | lo pair (ws: M*K*4 bytes). */
int asr_gemm_ln(const float* x, const void* w, const float* bias, const float* residual, const float* pe, int pe_period,
                const float* gamma, const float* beta, int M, int N, int K, float* h_out, float* y_f32, void* y_f16_hilo,
                void* ws, size_t ws_bytes, asr_stream_t stream);
/* softmax(mask(q k^T * scale)) v per head (layers.py:20-27); q/k/v/out f16 with row strides ld* (elements) and
 * batch strides; head h lives at column h*64. Masks nullable. impl as above. */
int asr_attention(const void* q, int ldq, long long q_bs, const void* k, int ldk, long long k_bs, const void* v,
                  int ldv, long long v_bs, void* out, int ldo, long long o_bs, int B, int H, int Sq, int Sk,
                  float scale, int causal, const int32_t* k_lens, const uint8_t* q_valid, const uint8_t* k_valid,
                  const uint8_t* dense_mask, int mask_B, int impl, asr_stream_t stream);
/* MHA.forward(x, enc_x, attention_mask) (layers.py:38-40): x fp32 (B,Sq,D), src fp32 (B,Sk,D) or NULL (self). */
size_t asr_mha_workspace_bytes(int B, int Sq, int Sk, int D);
int asr_mha(const float* x, const float* src, const AsrMhaWeights* w, int B, int Sq, int Sk, int D, int H, int causal,
            const uint8_t* q_valid, const uint8_t* k_valid, const uint8_t* dense_mask, int mask_B, void* ws,
            size_t ws_bytes, float* out, asr_stream_t stream);
/* FeedForward.forward (layers.py:53-58): x fp32 (rows,D) -> out fp32 (rows,D). */
size_t asr_ffn_workspace_bytes(int rows, int D, int FF);
int asr_ffn(const float* x, const AsrFfnWeights* w, int rows, int D, int FF, void* ws, size_t ws_bytes, float* out,
            asr_stream_t stream);
/* Transformer.input_layer (model.py:168-171): spectrum fp32 (B,1,F,T) -> z f16 (B,T',S*F'*64), column = f*64 + c
 * [+ F'*64: lo parts], S = 1 + asr_split_operands(). */
size_t asr_conv_workspace_bytes(int B, int F, int T);
int asr_conv_frontend(const float* spectrum, const float* conv1_w, const float* conv1_b, const void* conv2_wfrag,
                      const float* conv2_b, int B, int F, int T, void* ws, size_t ws_bytes, void* z_f16,
                      asr_stream_t stream);
/* Power spectrogram, the step before the hot path: replaces torchaudio.transforms.Spectrogram(n_fft,
 * center=False) of reference modules/dataset.py:34-35,51 (periodic Hann window of n_fft samples, hop = n_fft / 2 in the
 * reference, power 2, one-sided).  audio fp32 (B, n_samples) -> spec fp32 (B, 1, n_fft/2 + 1, T); frames past
 * floor((n_samples - n_fft) / hop) + 1 are zero-filled (the dataset's padding, dataset.py:53-55).  n_fft: a power of
 * two in [64, 2048]. */
int asr_spectrogram(const float* audio, int B, int n_samples, int n_fft, int hop, int T, float* spec,
                    asr_stream_t stream);
/* embedding + positional encoding (model.py:117): out fp32 (B,L,D) */
int asr_embed_pe(const int32_t* tokens, const float* emb, const float* pe, int B, int L, int D, int vocab, float* out,
                 asr_stream_t stream);
/* Decode-step operators (one new token per utterance), exposed for parity tests. */
int asr_dec_linear(const float* x, const float* ln_gamma, const float* ln_beta, const void* w, const float* bias,
                   const float* residual, int B, int N, int K, int relu, float* out, asr_stream_t stream);
int asr_dec_attention(const float* q, const void* k, const void* v, int ldkv, long long kv_bs, int n_keys, int B,
                      int H, float scale, float* out, asr_stream_t stream);
/* tcgen05 descriptor probe: D[128,N] = A[128,64] * B (b_mn_major 0: B is [N,64]; 1: B is [64,64] N-contiguous). */
int asr_umma_probe(const void* a, const void* b, float* d, int N, int b_mn_major, asr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H_ */
