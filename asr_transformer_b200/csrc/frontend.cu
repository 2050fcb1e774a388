// Power spectrogram on the GPU: the step BEFORE the hot path.
//
// Replaces torchaudio.transforms.Spectrogram(n_fft, center=False) of reference modules/dataset.py:34-35,51 (power 2,
// periodic Hann window, win_length = n_fft, hop = n_fft / 2, one-sided) so that raw audio can stay on the device:
// audio fp32 (B, N) -> spectrum fp32 (B, 1, n_fft/2 + 1, T) in exactly the layout conv1 reads, frames past the
// signal zero-filled like the dataset's padding (dataset.py:53-55).
//
// One warp owns one frame: windowed samples are stored bit-reversed into the warp's private shared-memory region,
// log2(n_fft) radix-2 stages run with __syncwarp only (no CTA barrier), twiddles come from a per-CTA table.  16 frames
// per CTA so that, for a fixed frequency bin, the CTA writes 64 contiguous bytes along T.  HBM-bound by design:
// 4 B/sample in (each sample is read by two overlapping frames, the second time from L2), ~2 B/sample out.
#include "kernels.h"
#include "ptx.cuh"

namespace asr {
namespace {

constexpr int SPEC_FRAMES = 16;   // frames (= warps) per CTA (8 when 16 frames of n_fft points do not fit in shared memory)

__global__ void __launch_bounds__(SPEC_FRAMES * 32)
spectrogram_kernel(const float* __restrict__ audio, int n_samples, int log2n, int hop, int n_frames, int T,
                   float* __restrict__ spec) {
  extern __shared__ float2 spec_smem[];
  const int N = 1 << log2n, half = N >> 1, nb = half + 1;
  float2* tw = spec_smem;                        // [N/2] exp(-2 pi i j / N)
  float2* buf = spec_smem + half;                // [frames per CTA][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fpc = blockDim.x >> 5;               // frames per CTA
  const int b = blockIdx.y, f0 = blockIdx.x * fpc;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float s, c;
    sincospif(-2.0f * float(j) / float(N), &s, &c);
    tw[j] = make_float2(c, s);
  }
  const int frame = f0 + warp;
  float2* x = buf + size_t(warp) * N;
  if (frame < n_frames) {
    const float* src = audio + size_t(b) * n_samples + size_t(frame) * hop;
    for (int n = lane; n < N; n += 32) {
      const float w = 0.5f - 0.5f * cospif(2.0f * float(n) / float(N));    // periodic Hann (torch.hann_window default)
      x[__brev(unsigned(n)) >> (32 - log2n)] = make_float2(src[n] * w, 0.f);
    }
  }
  __syncthreads();                               // twiddle table + (per warp) the frame
  if (frame < n_frames) {
    for (int s = 1; s <= log2n; ++s) {
      const int len = 1 << s, hl = len >> 1, tstep = N >> s;
      for (int i = lane; i < half; i += 32) {
        const int j = i & (hl - 1), base = ((i >> (s - 1)) << s) + j;
        const float2 w = tw[j * tstep], a = x[base], c = x[base + hl];
        const float2 t = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
        x[base] = make_float2(a.x + t.x, a.y + t.y);
        x[base + hl] = make_float2(a.x - t.x, a.y - t.y);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // |X[k]|^2 -> spec[b][0][k][f0 + f]; fpc threads write fpc consecutive frames of one bin (64 contiguous bytes)
  const int fl = threadIdx.x & (fpc - 1);
  float* out = spec + size_t(b) * nb * T;
  for (int k = threadIdx.x / fpc; k < nb; k += blockDim.x / fpc) {
    const int f = f0 + fl;
    if (f < T) {
      float v = 0.f;
      if (f < n_frames) {
        const float2 z = buf[size_t(fl) * N + k];
        v = z.x * z.x + z.y * z.y;
      }
      out[size_t(k) * T + f] = v;
    }
  }
}

}  // namespace

int launch_spectrogram(const float* audio, int B, int n_samples, int n_fft, int hop, int T, float* spec,
                       cudaStream_t s) {
  if (B <= 0) return 0;
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  if ((1 << log2n) != n_fft || n_fft < 64 || n_fft > 2048)
    return set_error(-2, "spectrogram: n_fft=%d must be a power of two in [64, 2048]", n_fft);
  if (hop <= 0 || T <= 0) return set_error(-1, "spectrogram: bad hop / frame count");
  const int n_frames = n_samples >= n_fft ? (n_samples - n_fft) / hop + 1 : 0;
  const int fpc = n_fft > 1024 ? SPEC_FRAMES / 2 : SPEC_FRAMES;
  const size_t smem = (size_t(n_fft / 2) + size_t(fpc) * n_fft) * sizeof(float2);
  if (int rc = ensure_dyn_smem((const void*)spectrogram_kernel, smem)) return rc;
  const int frames_covered = T;   // frames >= n_frames are zero-filled by the same kernel
  dim3 grid((frames_covered + fpc - 1) / fpc, B);
  spectrogram_kernel<<<grid, fpc * 32, smem, s>>>(audio, n_samples, log2n, hop, n_frames < T ? n_frames : T, T,
                                                        spec);
  ASR_CUDA_OK(cudaGetLastError());
  ASR_LAUNCHED(1);
  return 0;
}

}  // namespace asr
