// Fused log-mel spectrogram for sm_100a.  Replaces data_utils.py:39-62 (mel_spectrogram:
// reflect-pad -> STFT(1024/256, periodic Hann) -> sqrt(re^2+im^2+1e-9) -> mel basis -> log
// clamp) and data_utils.py:29-30 (dynamic_range_compression_torch).
//
// One CTA (128 threads) owns TWO consecutive frames of one clip: the windowed frames are
// packed as the real and imaginary parts of one 1024-point complex sequence, transformed by a
// radix-4 Stockham autosort FFT in shared memory (5 passes, ping-pong buffers), separated
// with the conjugate-symmetry identity, and reduced by the sparse Slaney filterbank
// (only the [tap_begin, tap_end) bins of each filter are touched: 727 taps instead of 41k).
// Audio is read once from HBM (overlapping frames hit L1/L2), the spectrum never leaves
// shared memory, and only the 80 log-mel values per frame are written.
#include "ssb_common.cuh"
#include <math.h>
#include <mutex>

namespace {

constexpr int NFFT = 1024;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int THREADS = 128;

__device__ float2 g_twiddle[NFFT];  // exp(-2*pi*i*m/1024), filled once per device

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(THREADS)
mel_kernel(const float* __restrict__ y, int64_t y_stride, int S, int hop, int pad, int frames,
           const float* __restrict__ basis, const int* __restrict__ tap_begin,
           const int* __restrict__ tap_end, int num_mels, float clip_val,
           float* __restrict__ out) {
  __shared__ float2 buf0[NFFT];
  __shared__ float2 buf1[NFFT];
  __shared__ float mag[2][NBINS + 3];

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * 2;
  const bool has_b = (f0 + 1) < frames;
  const float* yb = y + (int64_t)b * y_stride;

  // windowed, reflect-padded load (data_utils.py:51,54): frame f covers n = f*hop - pad + k
  for (int k = tid; k < NFFT; k += THREADS) {
    const float w = 0.5f - 0.5f * g_twiddle[k].x;  // periodic Hann
    int n0 = f0 * hop - pad + k;
    int n1 = n0 + hop;
    n0 = n0 < 0 ? -n0 : (n0 >= S ? 2 * (S - 1) - n0 : n0);
    n1 = n1 < 0 ? -n1 : (n1 >= S ? 2 * (S - 1) - n1 : n1);
    const float a = __ldg(yb + n0);
    const float c = has_b ? __ldg(yb + n1) : 0.f;
    buf0[k] = make_float2(w * a, w * c);
  }
  __syncthreads();

  // radix-4 Stockham autosort: Ns = 1, 4, 16, 64, 256
  float2* src = buf0;
  float2* dst = buf1;
#pragma unroll
  for (int pass = 0; pass < 5; ++pass) {
    const int Ns = 1 << (2 * pass);
#pragma unroll
    for (int rep = 0; rep < (NFFT / 4) / THREADS; ++rep) {
      const int j = tid + rep * THREADS;
      const int k = j & (Ns - 1);
      float2 v0 = src[j], v1 = src[j + NFFT / 4], v2 = src[j + NFFT / 2], v3 = src[j + 3 * NFFT / 4];
      if (pass > 0) {
        const int m = k * (NFFT / 4 / Ns);  // angle index of exp(-2*pi*i*k/(4*Ns))
        v1 = cmul(v1, g_twiddle[m]);
        v2 = cmul(v2, g_twiddle[2 * m]);
        v3 = cmul(v3, g_twiddle[3 * m]);
      }
      const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y);
      const float2 d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y);
      const float2 d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
      const int j0 = ((j - k) << 2) + k;
      dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
      dst[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - i*d13
      dst[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
      dst[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i*d13
    }
    __syncthreads();
    float2* t = src;
    src = dst;
    dst = t;
  }
  // 5 passes: result is in `src`

  // separate the two real transforms and take magnitudes (data_utils.py:57)
  for (int q = tid; q < NBINS; q += THREADS) {
    const float2 zp = src[q];
    const float2 zm = src[(NFFT - q) & (NFFT - 1)];
    const float ar = 0.5f * (zp.x + zm.x), ai = 0.5f * (zp.y - zm.y);
    const float br = 0.5f * (zp.y + zm.y), bi = -0.5f * (zp.x - zm.x);
    mag[0][q] = sqrtf(ar * ar + ai * ai + 1e-9f);
    mag[1][q] = sqrtf(br * br + bi * bi + 1e-9f);
  }
  __syncthreads();

  // sparse mel projection + log clamp (data_utils.py:59-60, 29-30)
  for (int idx = tid; idx < 2 * num_mels; idx += THREADS) {
    const int which = idx / num_mels, m = idx - which * num_mels;
    if (which == 1 && !has_b) continue;
    const float* brow = basis + (int64_t)m * NBINS;
    const int lo = tap_begin[m], hi = tap_end[m];
    float acc = 0.f;
    for (int q = lo; q < hi; ++q) acc = fmaf(__ldg(brow + q), mag[which][q], acc);
    out[((int64_t)b * num_mels + m) * frames + f0 + which] = logf(fmaxf(acc, clip_val));
  }
}

std::mutex g_tw_mutex;
bool g_tw_ready[64] = {false};
float2 g_tw_host[NFFT];

int ensure_twiddles(cudaStream_t st) {
  int dev = 0;
  SSB_CUDA(cudaGetDevice(&dev));
  SSB_REQUIRE(dev >= 0 && dev < 64, "mel: unexpected device ordinal %d", dev);
  std::lock_guard<std::mutex> lk(g_tw_mutex);
  if (g_tw_ready[dev]) return SSB_OK;
  for (int m = 0; m < NFFT; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)NFFT;
    g_tw_host[m] = make_float2((float)cos(a), (float)sin(a));
  }
  SSB_CUDA(cudaMemcpyToSymbolAsync(g_twiddle, g_tw_host, sizeof(g_tw_host), 0,
                                   cudaMemcpyHostToDevice, st));
  SSB_CUDA(cudaStreamSynchronize(st));  // one-time per device; host table is static
  g_tw_ready[dev] = true;
  return SSB_OK;
}

}  // namespace

extern "C" {

int64_t ssb_mel_num_frames(int64_t S, int n_fft, int hop) {
  if (S <= 0 || n_fft <= 0 || hop <= 0) return SSB_ERR_ARG;
  const int64_t pad = (n_fft - hop) / 2;
  const int64_t padded = S + 2 * pad;
  if (padded < n_fft) return 0;
  return (padded - n_fft) / hop + 1;
}

int ssb_mel_fwd(const float* y, int64_t B, int64_t S, int64_t y_stride, int n_fft, int hop,
                int win, const float* mel_basis, const int32_t* tap_begin,
                const int32_t* tap_end, int num_mels, float clip_val, float* out,
                void* stream) {
  SSB_REQUIRE(n_fft == NFFT && win == NFFT,
              "mel: only n_fft = win_size = 1024 is built (the reference's only call, "
              "data_utils.py:79); got n_fft=%d win=%d", n_fft, win);
  SSB_REQUIRE(hop > 0 && hop <= n_fft, "mel: bad hop %d", hop);
  SSB_REQUIRE(num_mels >= 1 && num_mels <= 1024, "mel: bad num_mels %d", num_mels);
  SSB_REQUIRE(B >= 0 && S >= 1 && y_stride >= S, "mel: bad shape B=%lld S=%lld stride=%lld",
              (long long)B, (long long)S, (long long)y_stride);
  const int pad = (n_fft - hop) / 2;
  SSB_REQUIRE(S > pad, "mel: reflect padding %d needs more than %d samples (got %lld)", pad, pad,
              (long long)S);
  SSB_REQUIRE(S < (1LL << 30), "mel: clip too long");
  const int64_t frames = ssb_mel_num_frames(S, n_fft, hop);
  if (B == 0 || frames == 0) return SSB_OK;
  SSB_REQUIRE(y && mel_basis && tap_begin && tap_end && out, "mel: null pointer");
  SSB_REQUIRE(B <= 65535, "mel: batch %lld exceeds grid.y", (long long)B);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ensure_twiddles(st)) return rc;
  dim3 grid((unsigned)((frames + 1) / 2), (unsigned)B);
  mel_kernel<<<grid, THREADS, 0, st>>>(y, y_stride, (int)S, hop, pad, (int)frames, mel_basis,
                                       tap_begin, tap_end, num_mels, clip_val, out);
  SSB_LAUNCH_CHECK("mel_kernel");
  return SSB_OK;
}

}  // extern "C"
