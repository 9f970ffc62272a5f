// Fused log-mel spectrogram for sm_100a.  Replaces data_utils.py:39-62 (mel_spectrogram:
// reflect-pad -> STFT(1024/256, periodic Hann) -> sqrt(re^2+im^2+1e-9) -> mel basis -> log
// clamp) and data_utils.py:29-30 (dynamic_range_compression_torch).
//
// One CTA (128 threads) owns FOUR consecutive frames of one clip: each half of the CTA (64
// threads) packs two windowed frames as the real and imaginary parts of one 1024-point complex
// sequence and transforms it as 1024 = 16 x 16 x 4 (Cooley-Tukey, radix-16 butterflies held in
// registers): the samples go from HBM straight into the first butterfly, and the sequence
// crosses shared memory only twice (padded rows, conflict-free) before the spectrum is laid down
// for the conjugate-symmetry split.  (The first version, a 5-pass radix-4 Stockham FFT, was
// bound by shared-memory / L1 traffic: 10 sweeps of the buffer, 8-way conflicts in the first
// passes and 3072 scattered twiddle loads per transform; ncu: L1/TEX 89 %.)  The sparse Slaney
// filterbank then touches only the [tap_begin, tap_end) bins of each filter (727 taps instead
// of 41k).  Audio is read once from HBM (overlapping frames hit L1/L2), the spectrum never
// leaves shared memory, and only the 80 log-mel values per frame are written.
#include "ssb_common.cuh"
#include <math.h>
#include <mutex>

namespace {

constexpr int NFFT = 1024;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int THREADS = 128;
constexpr int FRAMES_PER_CTA = 4;
constexpr int ROW = 68;          // padded row (float2) of the 16 x 64 exchange buffers

__device__ float2 g_twiddle[NFFT];  // exp(-2*pi*i*m/1024), filled once per device

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// in-place 4-point DFT (forward): outputs k = 0..3 land in a, b, c, d
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s02 = make_float2(a.x + c.x, a.y + c.y), d02 = make_float2(a.x - c.x, a.y - c.y);
  const float2 s13 = make_float2(b.x + d.x, b.y + d.y), d13 = make_float2(b.x - d.x, b.y - d.y);
  a = make_float2(s02.x + s13.x, s02.y + s13.y);
  b = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - i*d13
  c = make_float2(s02.x - s13.x, s02.y - s13.y);
  d = make_float2(d02.x - d13.y, d02.y + d13.x);      // d02 + i*d13
}

// in-place 16-point DFT of v[n], n = 4*n1 + m: output X[k] is left in v[4*(k & 3) + (k >> 2)]
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
#pragma unroll
  for (int m = 0; m < 4; ++m) dft4(v[m], v[4 + m], v[8 + m], v[12 + m]);
  // v[4*k1 + m] = A[m][k1]; twiddle by W16^(m*k1) = exp(-2*pi*i*m*k1/16)
  v[5] = cmul(v[5], make_float2(C1, -S1));     // 1
  v[6] = cmul(v[6], make_float2(R2, -R2));     // 2
  v[7] = cmul(v[7], make_float2(S1, -C1));     // 3
  v[9] = cmul(v[9], make_float2(R2, -R2));     // 2
  v[10] = make_float2(v[10].y, -v[10].x);      // 4: -i
  v[11] = cmul(v[11], make_float2(-R2, -R2));  // 6
  v[13] = cmul(v[13], make_float2(S1, -C1));   // 3
  v[14] = cmul(v[14], make_float2(-R2, -R2));  // 6
  v[15] = cmul(v[15], make_float2(-C1, S1));   // 9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

__global__ void __launch_bounds__(THREADS)
mel_kernel(const float* __restrict__ y, int64_t y_stride, int S, int hop, int pad, int frames,
           const float* __restrict__ basis, const int* __restrict__ tap_begin,
           const int* __restrict__ tap_end, int num_mels, float clip_val,
           float* __restrict__ out, unsigned int* __restrict__ range_cell) {
  __shared__ __align__(16) float2 ex1[2][16 * ROW];   // step 1 -> 2a exchange; later the spectrum
  __shared__ __align__(16) float2 ex2[2][16 * ROW];   // step 2a -> 2b exchange
  __shared__ float mag[FRAMES_PER_CTA][NBINS + 3];

  const int tid = threadIdx.x;
  const int g = tid >> 6, t = tid & 63;               // transform (frame pair) of this thread
  const int b = blockIdx.y;
  const int fa = blockIdx.x * FRAMES_PER_CTA + 2 * g;  // frames fa (real part), fa + 1 (imaginary)
  const bool has_a = fa < frames, has_b = fa + 1 < frames;
  const float* yb = y + (int64_t)b * y_stride;
  float2* s1 = ex1[g];
  float2* s2 = ex2[g];
  float2 v[16];
  float smin = 3.4e38f, smax = -3.4e38f;   // range of the samples this thread touches (data_utils.py:40-43)

  // step 1: thread m = t takes samples k = 64*n1 + m (windowed, reflect-padded load:
  // data_utils.py:51,54; frame f covers n = f*hop - pad + k), 16-point DFT over n1,
  // twiddle W1024^(m*k1), row k1 of the exchange buffer
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int k = 64 * n1 + t;
    const float w = 0.5f - 0.5f * g_twiddle[k].x;  // periodic Hann
    int n0 = fa * hop - pad + k;
    int nb = n0 + hop;
    n0 = n0 < 0 ? -n0 : (n0 >= S ? 2 * (S - 1) - n0 : n0);
    nb = nb < 0 ? -nb : (nb >= S ? 2 * (S - 1) - nb : nb);
    const float a = has_a ? __ldg(yb + n0) : 0.f;
    const float c = has_b ? __ldg(yb + nb) : 0.f;
    smin = fminf(smin, fminf(a, c));
    smax = fmaxf(smax, fmaxf(a, c));
    v[n1] = make_float2(w * a, w * c);
  }
  if (range_cell) {
    // order-preserving float -> uint map, so the range check of the reference (min < -1 / max > 1
    // warnings) costs two atomics per warp here instead of a separate reduction and a host read
    smin = ssb::warp_max(-smin);   // = -min
    smax = ssb::warp_max(smax);
    if ((tid & 31) == 0) {
      auto key = [](float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); };
      atomicMax(range_cell, key(smin));       // cell 0: -min
      atomicMax(range_cell + 1, key(smax));   // cell 1: max
    }
  }
  dft16(v);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const float2 x = v[4 * (k1 & 3) + (k1 >> 2)];
    s1[k1 * ROW + t] = k1 == 0 ? x : cmul(x, g_twiddle[(t * k1) & (NFFT - 1)]);
  }
  __syncthreads();

  // step 2a: thread (k1, m2): 16-point DFT over m1 of row k1 at m = 4*m1 + m2, twiddle W64^(m2*q1)
  {
    const int k1 = t >> 2, m2 = t & 3;
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) v[m1] = s1[k1 * ROW + 4 * m1 + m2];
    dft16(v);
#pragma unroll
    for (int q1 = 0; q1 < 16; ++q1) {
      const float2 x = v[4 * (q1 & 3) + (q1 >> 2)];
      s2[k1 * ROW + q1 * 4 + m2] = q1 == 0 ? x : cmul(x, g_twiddle[(16 * m2 * q1) & (NFFT - 1)]);
    }
  }
  __syncthreads();

  // step 2b: 4-point DFT over m2 for (k1, q1); X[k1 + 16*q1 + 256*q2] into the spectrum buffer
  // (aliases the first exchange buffer: every read of it finished before the barrier above)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int u = t + 64 * r, q1 = u >> 4, k1 = u & 15;
    const float4 lo = *reinterpret_cast<const float4*>(s2 + k1 * ROW + q1 * 4);
    const float4 hi = *reinterpret_cast<const float4*>(s2 + k1 * ROW + q1 * 4 + 2);
    float2 a = make_float2(lo.x, lo.y), bq = make_float2(lo.z, lo.w);
    float2 c = make_float2(hi.x, hi.y), d = make_float2(hi.z, hi.w);
    dft4(a, bq, c, d);
    float2* sp = s1 + k1 + 16 * q1;
    sp[0] = a; sp[256] = bq; sp[512] = c; sp[768] = d;
  }
  __syncthreads();

  // separate the two real transforms and take magnitudes (data_utils.py:57)
  for (int q = t; q < NBINS; q += 64) {
    const float2 zp = s1[q];
    const float2 zm = s1[(NFFT - q) & (NFFT - 1)];
    const float ar = 0.5f * (zp.x + zm.x), ai = 0.5f * (zp.y - zm.y);
    const float br = 0.5f * (zp.y + zm.y), bi = -0.5f * (zp.x - zm.x);
    mag[2 * g][q] = sqrtf(ar * ar + ai * ai + 1e-9f);
    mag[2 * g + 1][q] = sqrtf(br * br + bi * bi + 1e-9f);
  }
  __syncthreads();

  // sparse mel projection + log clamp (data_utils.py:59-60, 29-30)
  const int f0 = blockIdx.x * FRAMES_PER_CTA;
  for (int idx = tid; idx < FRAMES_PER_CTA * num_mels; idx += THREADS) {
    const int which = idx / num_mels, m = idx - which * num_mels;
    if (f0 + which >= frames) continue;
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
                void* range_cell, void* stream) {
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
  dim3 grid((unsigned)((frames + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA), (unsigned)B);
  mel_kernel<<<grid, THREADS, 0, st>>>(y, y_stride, (int)S, hop, pad, (int)frames, mel_basis,
                                       tap_begin, tap_end, num_mels, clip_val, out,
                                       (unsigned int*)range_cell);
  SSB_LAUNCH_CHECK("mel_kernel");
  return SSB_OK;
}

}  // extern "C"
