// Element-wise ends of the HiFi-GAN generator (reference: hifi_gan/models.py:75-116, driven by
// vocoder.py:16-36).  The convolutions themselves - conv_pre, the
// stride-phase form of every ConvTranspose1d and the dilated k = 3 / 7 / 11 residual convolutions -
// run on the tcgen05 GEMM engine (gemm_tc.cu, tap step = dilation) whose epilogue already writes
// leaky_relu(result) as the next convolution's operand planes.  What is left are the two places where
// several tensors meet:
//
//   voc_mix_kernel    xs = (x_0 + x_1 + x_2) / num_kernels of the multi-receptive-field fusion
//                     (models.py:101-108), followed by the leaky ReLU in front of the next upsampling
//                     convolution (:99) and written as bf16 split planes: 12 B read + 4 B written
//                     per element, no fp32 round trip.
//   voc_post_kernel   audio[t] = tanh(b + sum_{tap, c} w[tap][c] * leaky_relu(xs[t + tap - pad][c]))
//                     (models.py:109-111: F.leaky_relu default slope 0.01, conv_post with ONE output
//                     channel, tanh) with the same fusion of the three branches in front: a
//                     1-column "GEMM" that would waste a 256-wide tensor-core tile.
//
// Activations are channels-last (rows = samples, C contiguous) like everywhere else in libssb.
#include <cuda_bf16.h>

#include "ssb_common.cuh"

namespace {

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(256)
voc_mix_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
               int64_t n4, float scale, int act, float slope, float* __restrict__ out,
               __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
    if (b) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(b) + i);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (c) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(c) + i);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    if (out) reinterpret_cast<float4*>(out)[i] = v;
    if (planes) {
      if (act) { v.x = lrelu(v.x, slope); v.y = lrelu(v.y, slope); v.z = lrelu(v.z, slope); v.w = lrelu(v.w, slope); }
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
      __nv_bfloat16* dst = planes + 4 * i;
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01),
                                                  *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint2*>(dst + plane_stride) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01),
                                                                 *reinterpret_cast<const uint32_t*>(&l23));
    }
  }
}

constexpr int POST_MAX_W = 16 * 64;   // taps * C floats of the single output channel's filter

// one thread per output sample; the 32 samples of a warp read 32 consecutive rows per tap (each
// 16 B load of a lane hits the line its previous load brought into L1), the filter sits in shared
// memory and is read as a warp-wide broadcast
__global__ void __launch_bounds__(256)
voc_post_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                int64_t rows, int C, int taps, int pad, float scale, float slope,
                const float* __restrict__ w /* [taps][C] */, float bias, float* __restrict__ audio) {
  __shared__ float4 wsh[POST_MAX_W / 4];
  const int nw4 = taps * C / 4;
  for (int i = threadIdx.x; i < nw4; i += blockDim.x) wsh[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows) return;
  const int c4 = C / 4;
  float acc = bias;
  for (int tap = 0; tap < taps; ++tap) {
    const int64_t r = t + tap - pad;
    if (r < 0 || r >= rows) continue;
    const float4* pa = reinterpret_cast<const float4*>(a + r * C);
    const float4* pb = b ? reinterpret_cast<const float4*>(b + r * C) : nullptr;
    const float4* pc = c ? reinterpret_cast<const float4*>(c + r * C) : nullptr;
    for (int j = 0; j < c4; ++j) {
      float4 v = __ldg(pa + j);
      if (pb) { const float4 u = __ldg(pb + j); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
      if (pc) { const float4 u = __ldg(pc + j); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
      const float4 f = wsh[tap * c4 + j];
      acc = fmaf(f.x, lrelu(v.x * scale, slope), acc);
      acc = fmaf(f.y, lrelu(v.y * scale, slope), acc);
      acc = fmaf(f.z, lrelu(v.z * scale, slope), acc);
      acc = fmaf(f.w, lrelu(v.w * scale, slope), acc);
    }
  }
  audio[t] = tanhf(acc);
}

// The shipped generators end on 32 (padded) channels and a 7-tap filter: one warp per run of
// consecutive output samples, lane = channel.  Every row of the three branch tensors is read ONCE,
// as one coalesced 128 B line per tensor; the taps slide through a register window and the channel
// sum is a warp reduction.  (One thread per sample - the generic kernel above - makes every lane walk
// its own rows: 120 us for 153 600 samples, as long as two of the generator's convolutions.)
template <int TAPS, int CPL>
__global__ void __launch_bounds__(256)
voc_post_warp_kernel(const float* __restrict__ a, const float* __restrict__ b,
                     const float* __restrict__ c, int64_t rows, int run, float scale, float slope,
                     const float* __restrict__ w, float bias, float* __restrict__ audio) {
  constexpr int C = 32 * CPL, PAD = TAPS / 2;
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t t0 = wid * run;
  if (t0 >= rows) return;
  const int64_t t1 = min(rows, t0 + (int64_t)run);
  float wr[TAPS][CPL], win[TAPS][CPL];
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      wr[tap][k] = __ldg(w + tap * C + lane + 32 * k);
      win[tap][k] = 0.f;
    }
#pragma unroll 4
  for (int64_t r = t0 - PAD; r < t1 + PAD; ++r) {
#pragma unroll
    for (int tap = 0; tap + 1 < TAPS; ++tap)
#pragma unroll
      for (int k = 0; k < CPL; ++k) win[tap][k] = win[tap + 1][k];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      float v = 0.f;
      if (r >= 0 && r < rows) {
        const int64_t e = r * C + lane + 32 * k;
        v = __ldg(a + e);
        if (b) v += __ldg(b + e);
        if (c) v += __ldg(c + e);
        v = lrelu(v * scale, slope);
      }
      win[TAPS - 1][k] = v;
    }
    const int64_t t = r - PAD;
    if (t >= t0) {
      float acc = 0.f;
#pragma unroll
      for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc = fmaf(wr[tap][k], win[tap][k], acc);
      acc = ssb::warp_sum(acc);
      if (lane == 0) audio[t] = tanhf(acc + bias);
    }
  }
}

}  // namespace

extern "C" {

int ssb_voc_mix(const float* a, const float* b, const float* c, int64_t n, float scale, int lrelu_on,
                float neg_slope, float* out, void* planes, int64_t plane_stride, void* stream) {
  if (n == 0) return SSB_OK;
  SSB_REQUIRE(a && (out || planes) && n > 0 && n % 4 == 0, "voc_mix: n=%lld must be a positive multiple of 4",
              (long long)n);
  SSB_REQUIRE(b || !c, "voc_mix: c without b");
  SSB_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)out) & 15) == 0 &&
                  ((uintptr_t)planes & 7) == 0 && plane_stride % 4 == 0,
              "voc_mix: fp32 pointers must be 16 B aligned, planes 8 B");
  const int64_t n4 = n / 4;
  const int64_t blocks = (n4 + 255) / 256;
  const int cap = ssb::num_sms() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  voc_mix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, c, n4, scale, lrelu_on, neg_slope, out,
                                                         (__nv_bfloat16*)planes, plane_stride);
  SSB_LAUNCH_CHECK("voc_mix_kernel");
  return SSB_OK;
}

int ssb_voc_post(const float* a, const float* b, const float* c, int64_t rows, int64_t C, int64_t taps,
                 float scale, float neg_slope, const float* w, float bias, float* audio, void* stream) {
  if (rows == 0) return SSB_OK;
  SSB_REQUIRE(a && w && audio && rows > 0 && C > 0 && C % 4 == 0 && taps >= 1 && (taps & 1) &&
                  taps * C <= POST_MAX_W,
              "voc_post: C=%lld (multiple of 4), taps=%lld (odd), taps*C <= %d", (long long)C,
              (long long)taps, POST_MAX_W);
  SSB_REQUIRE(b || !c, "voc_post: c without b");
  SSB_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)w) & 15) == 0,
              "voc_post: pointers must be 16 B aligned");
  if (taps == 7 && (C == 32 || C == 64)) {
    const int run = 64;                                  // samples per warp
    const int64_t warps = (rows + run - 1) / run;
    const int64_t wblocks = (warps + 7) / 8;
    SSB_REQUIRE(wblocks < (1LL << 31), "voc_post: too many rows");
    if (C == 32)
      voc_post_warp_kernel<7, 1><<<(unsigned)wblocks, 256, 0, (cudaStream_t)stream>>>(
          a, b, c, rows, run, scale, neg_slope, w, bias, audio);
    else
      voc_post_warp_kernel<7, 2><<<(unsigned)wblocks, 256, 0, (cudaStream_t)stream>>>(
          a, b, c, rows, run, scale, neg_slope, w, bias, audio);
    SSB_LAUNCH_CHECK("voc_post_warp_kernel");
    return SSB_OK;
  }
  const int64_t blocks = (rows + 255) / 256;
  SSB_REQUIRE(blocks < (1LL << 31), "voc_post: too many rows");
  voc_post_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      a, b, c, rows, (int)C, (int)taps, (int)(taps / 2), scale, neg_slope, w, bias, audio);
  SSB_LAUNCH_CHECK("voc_post_kernel");
  return SSB_OK;
}

}  // extern "C"
