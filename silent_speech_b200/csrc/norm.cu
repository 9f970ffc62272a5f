// Normalisation / reduction kernels of the transduction model (sm_100a), channels-last.
//
//   column statistics + BatchNorm1d (training and eval) forward/backward   architecture.py:19-25,32-40
//   residual add + dropout + LayerNorm forward/backward                    transformer.py:55-59
//   column sums (bias gradients)
//
// All tensors are (rows, C) fp32 with C contiguous; per-channel quantities are column
// reductions.  Reductions are two-stage and deterministic: a grid of CTAs writes fp32
// partials per row chunk, a finalize kernel adds them in fp64 in a fixed order.
#include <cuda_bf16.h>

#include "ssb_common.cuh"

namespace {

constexpr int RCH = 256;  // rows per partial chunk

// optional second output of the element-wise kernels: the same 4 values as bf16 split planes
// (x = hi + lo), i.e. already in the operand format of the tcgen05 GEMM that consumes them, so no
// separate fp32 -> planes pass runs.  idx4 = float4 index into the plain tensor.
__device__ __forceinline__ void store_planes4(__nv_bfloat16* planes, int64_t plane_stride,
                                              int64_t idx4, const float4 v) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
  __nv_bfloat16* dst = planes + 4 * idx4;
  *reinterpret_cast<uint2*>(dst) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  *reinterpret_cast<uint2*>(dst + plane_stride) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// ---- column partial sums -------------------------------------------------------------
// MODE 1: sum(x)                                  (bias gradient)
// MODE 2: sum(dz), sum(dz * xhat)   dz = dy * (mask_src > 0 if mask_src)   (BatchNorm backward)
// block = 32 column-lanes (float4 each -> 128 columns) x 8 row-lanes
template <int MODE>
__global__ void __launch_bounds__(256)
col_partials_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                    const float* __restrict__ mask_src, const float* __restrict__ mean,
                    const float* __restrict__ rstd, int64_t rows, int C, int ld,
                    float* __restrict__ partials /* [nchunks][2][C] */) {
  __shared__ float4 red[2][8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cl) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * RCH;
  const int64_t r1 = min(rows, r0 + RCH);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (c < C) {
    float4 mu = s0, rs = s0;
    if (MODE == 2) {
      mu = *reinterpret_cast<const float4*>(mean + c);
      rs = *reinterpret_cast<const float4*>(rstd + c);
    }
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      if (MODE == 1) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
        s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
      } else {
        float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * ld + c));
        if (mask_src) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src + r * ld + c));
          g.x = m.x > 0.f ? g.x : 0.f; g.y = m.y > 0.f ? g.y : 0.f;
          g.z = m.z > 0.f ? g.z : 0.f; g.w = m.w > 0.f ? g.w : 0.f;
        }
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
        s0.x += g.x; s0.y += g.y; s0.z += g.z; s0.w += g.w;
        s1.x = fmaf(g.x, (v.x - mu.x) * rs.x, s1.x); s1.y = fmaf(g.y, (v.y - mu.y) * rs.y, s1.y);
        s1.z = fmaf(g.z, (v.z - mu.z) * rs.z, s1.z); s1.w = fmaf(g.w, (v.w - mu.w) * rs.w, s1.w);
      }
    }
  }
  red[0][rl][cl] = s0;
  red[1][rl][cl] = s1;
  __syncthreads();
  if (rl < 2 && c < C) {
    float4 a = red[rl][0][cl];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 b = red[rl][i][cl];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (rl == 0 || MODE == 2)
      *reinterpret_cast<float4*>(partials + ((int64_t)blockIdx.y * 2 + rl) * C + c) = a;
  }
}

// Single-pass BatchNorm statistics.  Every thread runs Welford's update over its 32 rows of a
// 256-row chunk (mean and centred sum of squares M2 updated per sample: no E[x^2] - mean^2
// cancellation, whatever |mean| / std is), the 8 row-lanes of a column are merged with Chan et
// al.'s pairwise formula in fp64, and the finalize kernel merges the per-chunk (n, mean, M2)
// triples the same way: the tensor is read ONCE.  (Round 1 read it twice - mean, then centred
// squares - after a shifted one-pass sum had lost 1e-2 on near-constant channels.)
// partials: [nchunks][3][C] = chunk mean, chunk M2, (unused).
__global__ void __launch_bounds__(256)
bn_chunk_stats_kernel(const float* __restrict__ x, int64_t rows, int C,
                      float* __restrict__ partials) {
  __shared__ float4 red[2][8][32];
  __shared__ int cnt[8];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cl) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * RCH;
  const int64_t r1 = min(rows, r0 + RCH);
  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), m2 = mu;
  int n = 0;
  if (c < C) {
    // eight independent 16 B loads in flight per thread before the (serial) Welford updates
    // (four left the kernel at 3.8 TB/s: latency-bound)
    constexpr int NL = 8;
    for (int64_t r = r0 + rl; r < r1; r += 8 * NL) {
      float4 v4[NL];
#pragma unroll
      for (int j = 0; j < NL; ++j)
        v4[j] = (r + 8 * j < r1) ? __ldg(reinterpret_cast<const float4*>(x + (r + 8 * j) * C + c))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        if (r + 8 * j >= r1) break;
        const float4 v = v4[j];
        ++n;
        const float inv = __frcp_rn((float)n);
        float d;
        d = v.x - mu.x; mu.x = fmaf(d, inv, mu.x); m2.x = fmaf(d, v.x - mu.x, m2.x);
        d = v.y - mu.y; mu.y = fmaf(d, inv, mu.y); m2.y = fmaf(d, v.y - mu.y, m2.y);
        d = v.z - mu.z; mu.z = fmaf(d, inv, mu.z); m2.z = fmaf(d, v.z - mu.z, m2.z);
        d = v.w - mu.w; mu.w = fmaf(d, inv, mu.w); m2.w = fmaf(d, v.w - mu.w, m2.w);
      }
    }
  } else {
    for (int64_t r = r0 + rl; r < r1; r += 8) ++n;
  }
  red[0][rl][cl] = mu;
  red[1][rl][cl] = m2;
  if (cl == 0) cnt[rl] = n;
  __syncthreads();
  if (rl == 0 && c < C) {
    double na = 0.0, ma[4] = {0, 0, 0, 0}, qa[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double nb = (double)cnt[i];
      if (nb == 0.0) continue;
      const float4 mb4 = red[0][i][cl], qb4 = red[1][i][cl];
      const double mb[4] = {mb4.x, mb4.y, mb4.z, mb4.w}, qb[4] = {qb4.x, qb4.y, qb4.z, qb4.w};
      const double nn = na + nb;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double dlt = mb[j] - ma[j];
        ma[j] += dlt * (nb / nn);
        qa[j] += qb[j] + dlt * dlt * (na * nb / nn);
      }
      na = nn;
    }
    *reinterpret_cast<float4*>(partials + ((int64_t)blockIdx.y * 3) * C + c) =
        make_float4((float)ma[0], (float)ma[1], (float)ma[2], (float)ma[3]);
    *reinterpret_cast<float4*>(partials + ((int64_t)blockIdx.y * 3 + 1) * C + c) =
        make_float4((float)qa[0], (float)qa[1], (float)qa[2], (float)qa[3]);
  }
}

// column sums of a tensor held as bf16 split planes (x = hi + lo): 32 column-lanes x 8 columns
__global__ void __launch_bounds__(256)
col_partials_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                           int64_t rows, int C, int rch, int slots,
                           float* __restrict__ partials /* [nchunks][slots][C] */) {
  __shared__ float red[8][32][9];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cl) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rch;
  const int64_t r1 = min(rows, r0 + rch);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + r * C + c));
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + r * C + c));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[2 * j] += __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
        s[2 * j + 1] += __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(lw[j] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cl][j] = s[j];
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = red[0][cl][j];
#pragma unroll
      for (int i = 1; i < 8; ++i) a += red[i][cl][j];
      partials[((int64_t)blockIdx.y * slots) * C + c + j] = a;
    }
  }
}

// ---- finalize kernels: fp64 accumulation of the chunk partials, fixed order ----------------
// block = (32 channels, 8 chunk-lanes); lane y adds chunks y, y+8, ...; lanes are combined in
// shared memory in a fixed order (deterministic), thread y == 0 finishes the channel.
constexpr int FIN_LANES = 32;   // 8 lanes left 31 dependent fp64 adds per thread on 24 CTAs: 10-20 us
                                // per finalize, 63 of them per step; 32 lanes: 8 adds, then a fixed-order
                                // shared-memory combine
#define FIN_GRID(C) dim3((unsigned)(((C) + 31) / 32)), dim3(32, FIN_LANES)

__device__ __forceinline__ bool chunk_sums(const float* __restrict__ partials, int nchunks, int C,
                                           int stride_slots, int c, double& s0, double& s1) {
  __shared__ double sh[2][FIN_LANES][32];
  double a = 0.0, b = 0.0;
  if (c < C) {
    for (int k = threadIdx.y; k < nchunks; k += FIN_LANES) {
      a += (double)partials[((int64_t)k * stride_slots) * C + c];
      b += (double)partials[((int64_t)k * stride_slots + 1) * C + c];
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = a;
  sh[1][threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return false;
  s0 = 0.0; s1 = 0.0;
#pragma unroll
  for (int y = 0; y < FIN_LANES; ++y) {
    s0 += sh[0][y][threadIdx.x];
    s1 += sh[1][y][threadIdx.x];
  }
  return true;
}

__global__ void __launch_bounds__(32 * FIN_LANES)
colsum_finalize_kernel(const float* __restrict__ partials, int nchunks, int C, int slots,
                                       float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, unused;
  if (!chunk_sums(partials, nchunks, C, slots, c, s, unused)) return;
  out[c] = (accumulate ? out[c] : 0.f) + (float)s;
}

// merge of the per-chunk (n, mean, M2) triples -> mean, rstd, scale, shift + running statistics
__global__ void __launch_bounds__(32 * FIN_LANES)
bn_merge_finalize_kernel(const float* __restrict__ partials, int nchunks, int C,
                                         int64_t rows, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* running_mean,
                                         float* running_var, float momentum, float eps,
                                         float* __restrict__ mean, float* __restrict__ rstd,
                                         float* __restrict__ scale, float* __restrict__ shift) {
  __shared__ double sh[FIN_LANES][32];
  __shared__ double mu_sh[32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const double count = (double)rows;
  // pass 1: global mean = sum_i n_i * mean_i / N
  double a = 0.0;
  if (c < C)
    for (int k = threadIdx.y; k < nchunks; k += FIN_LANES) {
      const double n = (double)min((int64_t)RCH, rows - (int64_t)k * RCH);
      a += n * (double)partials[((int64_t)k * 3) * C + c];
    }
  sh[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < FIN_LANES; ++y) t += sh[y][threadIdx.x];
    mu_sh[threadIdx.x] = t / count;
  }
  __syncthreads();
  const double mu = mu_sh[threadIdx.x];
  // pass 2: M2 = sum_i [ M2_i + n_i * (mean_i - mu)^2 ]
  double b = 0.0;
  if (c < C)
    for (int k = threadIdx.y; k < nchunks; k += FIN_LANES) {
      const double n = (double)min((int64_t)RCH, rows - (int64_t)k * RCH);
      const double d = (double)partials[((int64_t)k * 3) * C + c] - mu;
      b += (double)partials[((int64_t)k * 3 + 1) * C + c] + n * d * d;
    }
  __syncthreads();
  sh[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return;
  double m2 = 0.0;
#pragma unroll
  for (int y = 0; y < FIN_LANES; ++y) m2 += sh[y][threadIdx.x];
  const double var = m2 / count;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mu * sc;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval: running statistics -> (mean, rstd, scale, shift)
__global__ void bn_eval_affine_kernel(int C, const float* __restrict__ gamma,
                                      const float* __restrict__ beta,
                                      const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = 1.f / sqrtf(running_var[c] + eps);
  mean[c] = running_mean[c];
  rstd[c] = r;
  scale[c] = gamma[c] * r;
  shift[c] = beta[c] - running_mean[c] * gamma[c] * r;
}

// y = [relu]( (x-mean)*scale + beta  [+ (x2-mean2)*scale2 + beta2] ),  scale = gamma*rstd.
// Subtracting the mean first (instead of folding it into a shift) avoids cancellation between
// x*scale and mean*scale when |mean| >> std.
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                const float* __restrict__ scale, const float* __restrict__ beta,
                const float* __restrict__ x2, const float* __restrict__ mean2,
                const float* __restrict__ scale2, const float* __restrict__ beta2, int relu,
                int64_t n4, int C4, float* __restrict__ y, __nv_bfloat16* __restrict__ y_planes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    float4 o = make_float4(fmaf(v.x - mu.x, sc.x, be.x), fmaf(v.y - mu.y, sc.y, be.y),
                           fmaf(v.z - mu.z, sc.z, be.z), fmaf(v.w - mu.w, sc.w, be.w));
    if (x2) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(x2) + i);
      const float4 mu2 = *reinterpret_cast<const float4*>(mean2 + c);
      const float4 sc2 = *reinterpret_cast<const float4*>(scale2 + c);
      const float4 be2 = *reinterpret_cast<const float4*>(beta2 + c);
      o.x += fmaf(w.x - mu2.x, sc2.x, be2.x); o.y += fmaf(w.y - mu2.y, sc2.y, be2.y);
      o.z += fmaf(w.z - mu2.z, sc2.z, be2.z); o.w += fmaf(w.w - mu2.w, sc2.w, be2.w);
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    reinterpret_cast<float4*>(y)[i] = o;
    if (y_planes) store_planes4(y_planes, 4 * n4, i, o);
  }
}

// BatchNorm backward, finalize: sums -> dgamma, dbeta and the two per-channel means
__global__ void __launch_bounds__(32 * FIN_LANES)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int nchunks, int C,
                                       double count, int accumulate, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ m_dz,
                                       float* __restrict__ m_dzx) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, sx;
  if (!chunk_sums(partials, nchunks, C, 2, c, s, sx)) return;
  dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s;
  dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)sx;
  m_dz[c] = (float)(s / count);
  m_dzx[c] = (float)(sx / count);
}

// dx = gamma*rstd * (dz - mean(dz) - xhat*mean(dz*xhat));  eval mode: dx = gamma*rstd*dz
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ mask_src,
                    const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ m_dz, const float* __restrict__ m_dzx, int64_t n4,
                    int C4, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_planes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i);
    if (mask_src) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src) + i);
      g.x = m.x > 0.f ? g.x : 0.f; g.y = m.y > 0.f ? g.y : 0.f;
      g.z = m.z > 0.f ? g.z : 0.f; g.w = m.w > 0.f ? g.w : 0.f;
    }
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a;
    if (m_dz) {
      a = *reinterpret_cast<const float4*>(m_dz + c);
      bq = *reinterpret_cast<const float4*>(m_dzx + c);
    }
    float4 o;
    o.x = ga.x * rs.x * (g.x - a.x - (v.x - mu.x) * rs.x * bq.x);
    o.y = ga.y * rs.y * (g.y - a.y - (v.y - mu.y) * rs.y * bq.y);
    o.z = ga.z * rs.z * (g.z - a.z - (v.z - mu.z) * rs.z * bq.z);
    o.w = ga.w * rs.w * (g.w - a.w - (v.w - mu.w) * rs.w * bq.w);
    reinterpret_cast<float4*>(dx)[i] = o;
    if (dx_planes) store_planes4(dx_planes, 4 * n4, i, o);
  }
}

// ---- two-branch BatchNorm backward (ResBlock output: relu(bn2(c2) + res_norm(cr))) -----------
// Both branches receive the SAME dz = dy * (y > 0): one pass reads dy and y once for the four
// column sums, one pass reads them once more and writes both input gradients.
__global__ void __launch_bounds__(256)
bn2_bwd_partials_kernel(const float* __restrict__ dy, const float* __restrict__ mask_src,
                        const float* __restrict__ xa, const float* __restrict__ mean_a,
                        const float* __restrict__ rstd_a, const float* __restrict__ xb,
                        const float* __restrict__ mean_b, const float* __restrict__ rstd_b,
                        int64_t rows, int C, float* __restrict__ partials /* [nchunks][3][C] */) {
  __shared__ float4 red[3][8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cl) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * RCH;
  const int64_t r1 = min(rows, r0 + RCH);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0;
  if (c < C) {
    const float4 ma = *reinterpret_cast<const float4*>(mean_a + c), ra = *reinterpret_cast<const float4*>(rstd_a + c);
    const float4 mb = *reinterpret_cast<const float4*>(mean_b + c), rb = *reinterpret_cast<const float4*>(rstd_b + c);
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * C + c));
      if (mask_src) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src + r * C + c));
        g.x = m.x > 0.f ? g.x : 0.f; g.y = m.y > 0.f ? g.y : 0.f;
        g.z = m.z > 0.f ? g.z : 0.f; g.w = m.w > 0.f ? g.w : 0.f;
      }
      const float4 va = __ldg(reinterpret_cast<const float4*>(xa + r * C + c));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(xb + r * C + c));
      s0.x += g.x; s0.y += g.y; s0.z += g.z; s0.w += g.w;
      s1.x = fmaf(g.x, (va.x - ma.x) * ra.x, s1.x); s1.y = fmaf(g.y, (va.y - ma.y) * ra.y, s1.y);
      s1.z = fmaf(g.z, (va.z - ma.z) * ra.z, s1.z); s1.w = fmaf(g.w, (va.w - ma.w) * ra.w, s1.w);
      s2.x = fmaf(g.x, (vb.x - mb.x) * rb.x, s2.x); s2.y = fmaf(g.y, (vb.y - mb.y) * rb.y, s2.y);
      s2.z = fmaf(g.z, (vb.z - mb.z) * rb.z, s2.z); s2.w = fmaf(g.w, (vb.w - mb.w) * rb.w, s2.w);
    }
  }
  red[0][rl][cl] = s0;
  red[1][rl][cl] = s1;
  red[2][rl][cl] = s2;
  __syncthreads();
  if (rl < 3 && c < C) {
    float4 a = red[rl][0][cl];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 b = red[rl][i][cl];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(partials + ((int64_t)blockIdx.y * 3 + rl) * C + c) = a;
  }
}

__global__ void __launch_bounds__(32 * FIN_LANES)
bn2_bwd_finalize_kernel(const float* __restrict__ partials, int nchunks, int C,
                                        double count, int accumulate, float* __restrict__ dgamma_a,
                                        float* __restrict__ dbeta_a, float* __restrict__ dgamma_b,
                                        float* __restrict__ dbeta_b, float* __restrict__ m_dz,
                                        float* __restrict__ m_dzx_a, float* __restrict__ m_dzx_b) {
  __shared__ double sh[3][FIN_LANES][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0, b = 0.0, d = 0.0;
  if (c < C)
    for (int k = threadIdx.y; k < nchunks; k += FIN_LANES) {
      a += (double)partials[((int64_t)k * 3) * C + c];
      b += (double)partials[((int64_t)k * 3 + 1) * C + c];
      d += (double)partials[((int64_t)k * 3 + 2) * C + c];
    }
  sh[0][threadIdx.y][threadIdx.x] = a;
  sh[1][threadIdx.y][threadIdx.x] = b;
  sh[2][threadIdx.y][threadIdx.x] = d;
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return;
  double s = 0.0, sa = 0.0, sb = 0.0;
#pragma unroll
  for (int y = 0; y < FIN_LANES; ++y) {
    s += sh[0][y][threadIdx.x];
    sa += sh[1][y][threadIdx.x];
    sb += sh[2][y][threadIdx.x];
  }
  dbeta_a[c] = (accumulate ? dbeta_a[c] : 0.f) + (float)s;
  dbeta_b[c] = (accumulate ? dbeta_b[c] : 0.f) + (float)s;
  dgamma_a[c] = (accumulate ? dgamma_a[c] : 0.f) + (float)sa;
  dgamma_b[c] = (accumulate ? dgamma_b[c] : 0.f) + (float)sb;
  m_dz[c] = (float)(s / count);
  m_dzx_a[c] = (float)(sa / count);
  m_dzx_b[c] = (float)(sb / count);
}

__global__ void __launch_bounds__(256)
bn2_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ mask_src,
                     const float* __restrict__ xa, const float* __restrict__ mean_a,
                     const float* __restrict__ rstd_a, const float* __restrict__ gamma_a,
                     const float* __restrict__ xb, const float* __restrict__ mean_b,
                     const float* __restrict__ rstd_b, const float* __restrict__ gamma_b,
                     const float* __restrict__ m_dz, const float* __restrict__ m_dzx_a,
                     const float* __restrict__ m_dzx_b, int training, int64_t n4, int C4,
                     float* __restrict__ dxa, __nv_bfloat16* __restrict__ dxa_planes,
                     float* __restrict__ dxb, __nv_bfloat16* __restrict__ dxb_planes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i);
    if (mask_src) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src) + i);
      g.x = m.x > 0.f ? g.x : 0.f; g.y = m.y > 0.f ? g.y : 0.f;
      g.z = m.z > 0.f ? g.z : 0.f; g.w = m.w > 0.f ? g.w : 0.f;
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), qa = a, qb = a;
    if (training) {
      a = *reinterpret_cast<const float4*>(m_dz + c);
      qa = *reinterpret_cast<const float4*>(m_dzx_a + c);
      qb = *reinterpret_cast<const float4*>(m_dzx_b + c);
    }
    {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xa) + i);
      const float4 mu = *reinterpret_cast<const float4*>(mean_a + c);
      const float4 rs = *reinterpret_cast<const float4*>(rstd_a + c);
      const float4 ga = *reinterpret_cast<const float4*>(gamma_a + c);
      float4 o;
      o.x = ga.x * rs.x * (g.x - a.x - (v.x - mu.x) * rs.x * qa.x);
      o.y = ga.y * rs.y * (g.y - a.y - (v.y - mu.y) * rs.y * qa.y);
      o.z = ga.z * rs.z * (g.z - a.z - (v.z - mu.z) * rs.z * qa.z);
      o.w = ga.w * rs.w * (g.w - a.w - (v.w - mu.w) * rs.w * qa.w);
      reinterpret_cast<float4*>(dxa)[i] = o;
      if (dxa_planes) store_planes4(dxa_planes, 4 * n4, i, o);
    }
    {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xb) + i);
      const float4 mu = *reinterpret_cast<const float4*>(mean_b + c);
      const float4 rs = *reinterpret_cast<const float4*>(rstd_b + c);
      const float4 ga = *reinterpret_cast<const float4*>(gamma_b + c);
      float4 o;
      o.x = ga.x * rs.x * (g.x - a.x - (v.x - mu.x) * rs.x * qb.x);
      o.y = ga.y * rs.y * (g.y - a.y - (v.y - mu.y) * rs.y * qb.y);
      o.z = ga.z * rs.z * (g.z - a.z - (v.z - mu.z) * rs.z * qb.z);
      o.w = ga.w * rs.w * (g.w - a.w - (v.w - mu.w) * rs.w * qb.w);
      reinterpret_cast<float4*>(dxb)[i] = o;
      if (dxb_planes) store_planes4(dxb_planes, 4 * n4, i, o);
    }
  }
}

// ---- residual + dropout + LayerNorm ----------------------------------------------------
// One warp per row; D <= 1024, D % 4 == 0.  z = res + dropout(branch); y = LN(z).
constexpr int LN_MAXV = 8;  // float4 per lane

__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(const float* __restrict__ res, const float* __restrict__ branch,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows,
                  int D, float eps, float drop_p, float drop_scale, uint32_t drop_thresh,
                  uint64_t seed_base, const uint64_t* __restrict__ seed_src, uint32_t site,
                  float* __restrict__ z_out,
                  float* __restrict__ y, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, __nv_bfloat16* __restrict__ y_planes) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint64_t seed = ssb::eff_seed(seed_base, seed_src);
  const int nv = D >> 2;
  float4 z[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      float4 a = __ldg(reinterpret_cast<const float4*>(res + row * D) + v);
      float4 b = __ldg(reinterpret_cast<const float4*>(branch + row * D) + v);
      if (drop_p > 0.f) {
        const uint4 rnd = ssb::dropout_bits4(seed, site, (uint64_t)row * nv + v);
        b.x = rnd.x >= drop_thresh ? b.x * drop_scale : 0.f;
        b.y = rnd.y >= drop_thresh ? b.y * drop_scale : 0.f;
        b.z = rnd.z >= drop_thresh ? b.z * drop_scale : 0.f;
        b.w = rnd.w >= drop_thresh ? b.w * drop_scale : 0.f;
      }
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      z[i] = a;
      s += a.x + a.y + a.z + a.w;
    }
  }
  s = ssb::warp_sum(s);
  const float mu = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      const float dx = z[i].x - mu, dy = z[i].y - mu, dz = z[i].z - mu, dw = z[i].w - mu;
      q += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  q = ssb::warp_sum(q);
  const float rs = rsqrtf(q / (float)D + eps);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + v * 4);
      const float4 b = *reinterpret_cast<const float4*>(beta + v * 4);
      float4 o;
      o.x = (z[i].x - mu) * rs * g.x + b.x;
      o.y = (z[i].y - mu) * rs * g.y + b.y;
      o.z = (z[i].z - mu) * rs * g.z + b.z;
      o.w = (z[i].w - mu) * rs * g.w + b.w;
      reinterpret_cast<float4*>(y + row * D)[v] = o;
      if (y_planes) store_planes4(y_planes, rows * D, row * nv + v, o);
      if (z_out) reinterpret_cast<float4*>(z_out + row * D)[v] = z[i];
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mu;
    if (rstd_out) rstd_out[row] = rs;
  }
}

// backward: d_res = dz, d_branch = dz * dropmask * scale; per-CTA partial dgamma/dbeta
// (Tried and measured slower, r2 sessions 35 / 36, 12 launches per step: 32 rows per CTA - more CTAs
// for the 2-per-SM slots - 756 -> 824 us and a doubled finalize; __launch_bounds__(256, 3) - 80 registers,
// 328 B of spills - 933 us; parameter-gradient partials through shared-memory atomics instead of 48
// accumulator registers, which fits three CTAs per SM without spills - 962 us.  The kernel is
// latency-bound at 2.5 TB/s with 16 warps per SM, and more resident warps do not help it.)
constexpr int LN_ROWS_PER_CTA = 64;

// NV = float4 per lane actually needed (D <= 128 * NV): the register arrays are sized by it, so
// D = 768 runs with 6-wide state instead of the 8-wide maximum (208 -> ~128 registers, two CTAs
// per SM).  gamma is re-read per row (L1-resident) instead of living in registers.
template <int NV>
__global__ void __launch_bounds__(256)
add_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                  const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ gamma, int64_t rows, int D, float drop_p,
                  float drop_scale, uint32_t drop_thresh, uint64_t seed_base,
                  const uint64_t* __restrict__ seed_src, uint32_t site,
                  float* __restrict__ d_res, float* __restrict__ d_branch,
                  __nv_bfloat16* __restrict__ d_branch_planes,
                  float* __restrict__ partials /* [nblk][2][D] : dgamma, dbeta */) {
  __shared__ float sm[2][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed = ssb::eff_seed(seed_base, seed_src);
  const int nv = D >> 2;
  for (int i = threadIdx.x; i < 2 * 1024; i += 256) (&sm[0][0])[i] = 0.f;
  __syncthreads();
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = dg[i];
  }
  const int64_t r0 = (int64_t)blockIdx.x * LN_ROWS_PER_CTA;
  const int64_t r1 = min(rows, r0 + LN_ROWS_PER_CTA);
  for (int64_t row = r0 + warp; row < r1; row += 8) {
    const float mu = mean[row], rs = rstd[row];
    float4 gy[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(dy + row * D) + v);
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z + row * D) + v);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + v);
        xh[i] = make_float4((zz.x - mu) * rs, (zz.y - mu) * rs, (zz.z - mu) * rs,
                            (zz.w - mu) * rs);
        gy[i] = make_float4(d.x * g4.x, d.y * g4.y, d.z * g4.z, d.w * g4.w);
        s1 += gy[i].x + gy[i].y + gy[i].z + gy[i].w;
        s2 += gy[i].x * xh[i].x + gy[i].y * xh[i].y + gy[i].z * xh[i].z + gy[i].w * xh[i].w;
        dg[i].x = fmaf(d.x, xh[i].x, dg[i].x); dg[i].y = fmaf(d.y, xh[i].y, dg[i].y);
        dg[i].z = fmaf(d.z, xh[i].z, dg[i].z); dg[i].w = fmaf(d.w, xh[i].w, dg[i].w);
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      }
    }
    s1 = ssb::warp_sum(s1) / (float)D;
    s2 = ssb::warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        float4 o;
        o.x = rs * (gy[i].x - s1 - xh[i].x * s2);
        o.y = rs * (gy[i].y - s1 - xh[i].y * s2);
        o.z = rs * (gy[i].z - s1 - xh[i].z * s2);
        o.w = rs * (gy[i].w - s1 - xh[i].w * s2);
        reinterpret_cast<float4*>(d_res + row * D)[v] = o;
        if (drop_p > 0.f) {
          const uint4 rnd = ssb::dropout_bits4(seed, site, (uint64_t)row * nv + v);
          o.x = rnd.x >= drop_thresh ? o.x * drop_scale : 0.f;
          o.y = rnd.y >= drop_thresh ? o.y * drop_scale : 0.f;
          o.z = rnd.z >= drop_thresh ? o.z * drop_scale : 0.f;
          o.w = rnd.w >= drop_thresh ? o.w * drop_scale : 0.f;
        }
        if (d_branch) reinterpret_cast<float4*>(d_branch + row * D)[v] = o;
        if (d_branch_planes) store_planes4(d_branch_planes, rows * D, row * nv + v, o);
      }
    }
  }
  // cross-warp reduction of the per-lane column partials
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      atomicAdd(&sm[0][v * 4 + 0], dg[i].x); atomicAdd(&sm[0][v * 4 + 1], dg[i].y);
      atomicAdd(&sm[0][v * 4 + 2], dg[i].z); atomicAdd(&sm[0][v * 4 + 3], dg[i].w);
      atomicAdd(&sm[1][v * 4 + 0], db[i].x); atomicAdd(&sm[1][v * 4 + 1], db[i].y);
      atomicAdd(&sm[1][v * 4 + 2], db[i].z); atomicAdd(&sm[1][v * 4 + 3], db[i].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    partials[((int64_t)blockIdx.x * 2 + 0) * D + c] = sm[0][c];
    partials[((int64_t)blockIdx.x * 2 + 1) * D + c] = sm[1][c];
  }
}

__global__ void __launch_bounds__(32 * FIN_LANES)
ln_param_grad_finalize_kernel(const float* __restrict__ partials, int nblk, int D,
                                              int accumulate, float* __restrict__ dgamma,
                                              float* __restrict__ dbeta) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double a, b;
  if (!chunk_sums(partials, nblk, D, 2, c, a, b)) return;
  dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)a;
  dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)b;
}

int check_rows_c(const void* p, int64_t rows, int64_t C, const char* what) {
  SSB_REQUIRE(p != nullptr, "%s: null pointer", what);
  SSB_REQUIRE(rows >= 1 && C >= 4 && C % 4 == 0 && C < (1 << 24), "%s: bad shape rows=%lld C=%lld",
              what, (long long)rows, (long long)C);
  SSB_REQUIRE(((uintptr_t)p & 15) == 0, "%s: pointer must be 16 B aligned", what);
  return SSB_OK;
}

inline int nchunks_for(int64_t rows) { return (int)((rows + RCH - 1) / RCH); }

inline uint32_t thresh_of(float p) {
  const double th = (double)p * 4294967296.0;
  return th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
}

}  // namespace

extern "C" {

int64_t ssb_col_partials_bytes(int64_t rows, int64_t C) {
  if (rows < 0 || C < 0) return SSB_ERR_ARG;
  // up to 3 slots per 256-row chunk (BatchNorm), or one slot per 64-row chunk plus the row the
  // two-value finalize reads past the last one (narrow column sums of split planes)
  return (int64_t)nchunks_for(rows > 0 ? rows : 1) * 6 * C * 4;
}

int ssb_colsum(const float* x, int64_t rows, int64_t C, float* out, int accumulate,
               void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_rows_c(x, rows, C, "colsum")) return rc;
  SSB_REQUIRE(out && workspace && workspace_bytes >= ssb_col_partials_bytes(rows, C),
              "colsum: bad out/workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = nchunks_for(rows);
  dim3 grid((unsigned)((C + 127) / 128), (unsigned)nch);
  col_partials_kernel<1><<<grid, 256, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, rows, (int)C,
                                               (int)C, (float*)workspace);
  SSB_LAUNCH_CHECK("col_partials<1>");
  colsum_finalize_kernel<<<FIN_GRID(C), 0, st>>>((const float*)workspace, nch, (int)C, 2, out,
                                                 accumulate);
  SSB_LAUNCH_CHECK("colsum_finalize");
  return SSB_OK;
}

int ssb_colsum_planes(const void* planes, int64_t plane_stride, int64_t rows, int64_t C,
                      float* out, int accumulate, void* workspace, int64_t workspace_bytes,
                      void* stream) {
  SSB_REQUIRE(planes && rows > 0 && C > 0 && C % 8 == 0 && ((uintptr_t)planes & 15) == 0 &&
                  plane_stride % 8 == 0,
              "colsum_planes: rows=%lld C=%lld (C %% 8 == 0, 16 B aligned planes)", (long long)rows,
              (long long)C);
  SSB_REQUIRE(out && workspace && workspace_bytes >= ssb_col_partials_bytes(rows, C),
              "colsum_planes: bad out/workspace");
  cudaStream_t st = (cudaStream_t)stream;
  // A 768-column tensor has only 3 column blocks: with 256-row chunks 16000 rows give 189 CTAs and the
  // loads in flight cover a third of the HBM latency-bandwidth product (21 us for 49 MB).  Narrow
  // tensors use 64-row chunks, one partial slot each (4 x the CTAs).
  const bool narrow = C <= 1024 && rows >= 2048;
  const int rch = narrow ? 64 : RCH, slots = narrow ? 1 : 2;
  const int nch = (int)((rows + rch - 1) / rch);
  dim3 grid((unsigned)((C + 255) / 256), (unsigned)nch);
  const __nv_bfloat16* hi = (const __nv_bfloat16*)planes;
  col_partials_planes_kernel<<<grid, 256, 0, st>>>(hi, hi + plane_stride, rows, (int)C, rch, slots,
                                                   (float*)workspace);
  SSB_LAUNCH_CHECK("col_partials_planes");
  colsum_finalize_kernel<<<FIN_GRID(C), 0, st>>>((const float*)workspace, nch, (int)C, slots, out,
                                                 accumulate);
  SSB_LAUNCH_CHECK("colsum_finalize");
  return SSB_OK;
}

int ssb_bn_stats(const float* x, int64_t rows, int64_t C, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps,
                 int training, float* mean, float* rstd, float* scale, float* shift,
                 void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_rows_c(x, rows, C, "bn_stats")) return rc;
  SSB_REQUIRE(gamma && beta && mean && rstd && scale && shift, "bn_stats: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned cb = (unsigned)((C + 127) / 128);
  if (!training) {
    SSB_REQUIRE(running_mean && running_var, "bn_stats: eval mode needs running statistics");
    bn_eval_affine_kernel<<<cb, 128, 0, st>>>((int)C, gamma, beta, running_mean, running_var, eps,
                                              mean, rstd, scale, shift);
    SSB_LAUNCH_CHECK("bn_eval_affine");
    return SSB_OK;
  }
  SSB_REQUIRE(workspace && workspace_bytes >= ssb_col_partials_bytes(rows, C),
              "bn_stats: workspace too small");
  const int nch = nchunks_for(rows);
  dim3 grid(cb, (unsigned)nch);
  bn_chunk_stats_kernel<<<grid, 256, 0, st>>>(x, rows, (int)C, (float*)workspace);
  SSB_LAUNCH_CHECK("bn_chunk_stats");
  bn_merge_finalize_kernel<<<FIN_GRID(C), 0, st>>>((const float*)workspace, nch, (int)C, rows, gamma,
                                                   beta, running_mean, running_var, momentum, eps,
                                                   mean, rstd, scale, shift);
  SSB_LAUNCH_CHECK("bn_finalize");
  return SSB_OK;
}

int ssb_bn_apply(const float* x, const float* mean, const float* scale, const float* beta,
                 const float* x2, const float* mean2, const float* scale2, const float* beta2,
                 int relu, int64_t rows, int64_t C, float* y, void* y_planes, void* stream) {
  if (int rc = check_rows_c(x, rows, C, "bn_apply")) return rc;
  SSB_REQUIRE(mean && scale && beta && y && (!x2 || (mean2 && scale2 && beta2)),
              "bn_apply: null pointer");
  const int64_t n4 = rows * C / 4;
  const int64_t blocks = (n4 + 255) / 256;
  const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  bn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, mean, scale, beta, x2, mean2, scale2,
                                                          beta2, relu, n4, (int)(C / 4), y,
                                                          (__nv_bfloat16*)y_planes);
  SSB_LAUNCH_CHECK("bn_apply");
  return SSB_OK;
}

int ssb_bn_bwd(const float* dy, const float* mask_src, const float* x, const float* mean,
               const float* rstd, const float* gamma, int training, int64_t rows, int64_t C,
               float* dx, void* dx_planes, float* dgamma, float* dbeta, int accumulate,
               void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_rows_c(x, rows, C, "bn_bwd")) return rc;
  SSB_REQUIRE(dy && mean && rstd && gamma && dx && dgamma && dbeta, "bn_bwd: null pointer");
  const int64_t need = ssb_col_partials_bytes(rows, C) + 3 * C * 4;
  SSB_REQUIRE(workspace && workspace_bytes >= need, "bn_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = nchunks_for(rows);
  const unsigned cb = (unsigned)((C + 127) / 128);
  float* partials = (float*)workspace;
  float* m_dz = partials + (int64_t)nch * 3 * C;
  float* m_dzx = m_dz + C;
  dim3 grid(cb, (unsigned)nch);
  col_partials_kernel<2><<<grid, 256, 0, st>>>(x, dy, mask_src, mean, rstd, rows, (int)C, (int)C,
                                               partials);
  SSB_LAUNCH_CHECK("col_partials<2>");
  bn_bwd_finalize_kernel<<<FIN_GRID(C), 0, st>>>(partials, nch, (int)C, (double)rows, accumulate,
                                             dgamma, dbeta, m_dz, m_dzx);
  SSB_LAUNCH_CHECK("bn_bwd_finalize");
  const int64_t n4 = rows * C / 4;
  const int64_t blocks = (n4 + 255) / 256;
  const int g2 = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  bn_bwd_apply_kernel<<<g2, 256, 0, st>>>(dy, mask_src, x, mean, rstd, gamma,
                                          training ? m_dz : nullptr, training ? m_dzx : nullptr,
                                          n4, (int)(C / 4), dx, (__nv_bfloat16*)dx_planes);
  SSB_LAUNCH_CHECK("bn_bwd_apply");
  return SSB_OK;
}

int ssb_bn_bwd2(const float* dy, const float* mask_src, const float* xa, const float* mean_a,
                const float* rstd_a, const float* gamma_a, const float* xb, const float* mean_b,
                const float* rstd_b, const float* gamma_b, int training, int64_t rows, int64_t C,
                float* dxa, void* dxa_planes, float* dxb, void* dxb_planes, float* dgamma_a,
                float* dbeta_a, float* dgamma_b, float* dbeta_b, int accumulate, void* workspace,
                int64_t workspace_bytes, void* stream) {
  if (int rc = check_rows_c(xa, rows, C, "bn_bwd2")) return rc;
  if (int rc = check_rows_c(xb, rows, C, "bn_bwd2")) return rc;
  SSB_REQUIRE(dy && mean_a && rstd_a && gamma_a && mean_b && rstd_b && gamma_b && dxa && dxb &&
                  dgamma_a && dbeta_a && dgamma_b && dbeta_b,
              "bn_bwd2: null pointer");
  const int64_t need = ssb_col_partials_bytes(rows, C) + 3 * C * 4;
  SSB_REQUIRE(workspace && workspace_bytes >= need, "bn_bwd2: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = nchunks_for(rows);
  const unsigned cb = (unsigned)((C + 127) / 128);
  float* partials = (float*)workspace;
  float* m_dz = partials + (int64_t)nch * 3 * C;
  float* m_a = m_dz + C;
  float* m_b = m_a + C;
  bn2_bwd_partials_kernel<<<dim3(cb, (unsigned)nch), 256, 0, st>>>(
      dy, mask_src, xa, mean_a, rstd_a, xb, mean_b, rstd_b, rows, (int)C, partials);
  SSB_LAUNCH_CHECK("bn2_bwd_partials");
  bn2_bwd_finalize_kernel<<<FIN_GRID(C), 0, st>>>(partials, nch, (int)C, (double)rows, accumulate,
                                                  dgamma_a, dbeta_a, dgamma_b, dbeta_b, m_dz, m_a, m_b);
  SSB_LAUNCH_CHECK("bn2_bwd_finalize");
  const int64_t n4 = rows * C / 4;
  const int64_t blocks = (n4 + 255) / 256;
  const int g2 = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  bn2_bwd_apply_kernel<<<g2, 256, 0, st>>>(dy, mask_src, xa, mean_a, rstd_a, gamma_a, xb, mean_b,
                                           rstd_b, gamma_b, m_dz, m_a, m_b, training, n4,
                                           (int)(C / 4), dxa, (__nv_bfloat16*)dxa_planes, dxb,
                                           (__nv_bfloat16*)dxb_planes);
  SSB_LAUNCH_CHECK("bn2_bwd_apply");
  return SSB_OK;
}

int ssb_add_dropout_ln_fwd(const float* res, const float* branch, const float* gamma,
                           const float* beta, int64_t rows, int64_t D, float eps, float drop_p,
                           uint64_t seed, uint32_t site, float* z_out, float* y, float* mean,
                           float* rstd, void* y_planes, void* stream) {
  if (int rc = check_rows_c(res, rows, D, "add_dropout_ln_fwd")) return rc;
  SSB_REQUIRE(D <= 1024, "add_dropout_ln_fwd: D=%lld > 1024 not built", (long long)D);
  SSB_REQUIRE(branch && gamma && beta && y, "add_dropout_ln_fwd: null pointer");
  SSB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "add_dropout_ln_fwd: bad p");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  add_ln_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      res, branch, gamma, beta, rows, (int)D, eps, drop_p, drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f,
      thresh_of(drop_p), seed, ssb::seed_source(), site, z_out, y, mean, rstd,
      (__nv_bfloat16*)y_planes);
  SSB_LAUNCH_CHECK("add_ln_fwd");
  return SSB_OK;
}

int64_t ssb_add_dropout_ln_bwd_workspace_bytes(int64_t rows, int64_t D) {
  if (rows < 0 || D < 0) return SSB_ERR_ARG;
  const int64_t nblk = (rows + LN_ROWS_PER_CTA - 1) / LN_ROWS_PER_CTA;
  return (nblk > 0 ? nblk : 1) * 2 * D * 4;
}

int ssb_add_dropout_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                           const float* gamma, int64_t rows, int64_t D, float drop_p,
                           uint64_t seed, uint32_t site, float* d_res, float* d_branch,
                           void* d_branch_planes, float* dgamma, float* dbeta, int accumulate,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_rows_c(dy, rows, D, "add_dropout_ln_bwd")) return rc;
  SSB_REQUIRE(D <= 1024, "add_dropout_ln_bwd: D=%lld > 1024 not built", (long long)D);
  SSB_REQUIRE(z && mean && rstd && gamma && d_res && (d_branch || d_branch_planes) && dgamma && dbeta,
              "add_dropout_ln_bwd: null pointer");
  SSB_REQUIRE(workspace && workspace_bytes >= ssb_add_dropout_ln_bwd_workspace_bytes(rows, D),
              "add_dropout_ln_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (int)((rows + LN_ROWS_PER_CTA - 1) / LN_ROWS_PER_CTA);
#define SSB_LN_BWD(NV)                                                                        \
  add_ln_bwd_kernel<NV><<<nblk, 256, 0, st>>>(dy, z, mean, rstd, gamma, rows, (int)D, drop_p,       \
                                              drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f,           \
                                              thresh_of(drop_p), seed, ssb::seed_source(), site,   \
                                              d_res, d_branch, (__nv_bfloat16*)d_branch_planes,   \
                                              (float*)workspace)
  const int nvl = (int)((D / 4 + 31) / 32);
  if (nvl <= 2) SSB_LN_BWD(2);
  else if (nvl <= 4) SSB_LN_BWD(4);
  else if (nvl <= 6) SSB_LN_BWD(6);
  else SSB_LN_BWD(8);
#undef SSB_LN_BWD
  SSB_LAUNCH_CHECK("add_ln_bwd");
  ln_param_grad_finalize_kernel<<<FIN_GRID(D), 0, st>>>((const float*)workspace, nblk, (int)D,
                                                        accumulate, dgamma, dbeta);
  SSB_LAUNCH_CHECK("ln_param_grad_finalize");
  return SSB_OK;
}

}  // extern "C"
