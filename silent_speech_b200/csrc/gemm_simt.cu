// fp32 CUDA-core GEMM family with gathered operands (sm_100a).
//
// One tile engine (128x128x16, 256 threads, 8x8 register micro-tiles, double-buffered shared
// memory) serves every dense contraction of the model that is not (yet) on the tcgen05 path:
//   NN  C[m,n]  = epi( sum_k A(m,k) * W[k,n] )          forward of Linear / Conv1d
//   NT  C[m,j]  = epi( sum_k A(m,k) * W[row(j,k), col(k)] )   data gradient (W read transposed)
//   TN  C[k,n] += sum_m A(m,k) * G[m,n]                  weight gradient (split over m)
// A is never materialised: A(m,k) is gathered on the fly from a channels-last activation
// tensor (B, L, C) with  m -> (b, t),  k -> (tap, c),  source row  t*s_t + tap*s_tap + off,
// rows outside [0, L) read as zero.  That one rule expresses nn.Linear (taps = 1), the k3/p1
// convolutions with stride 1 or 2, the k1/s2 residual convolution and all of their data
// gradients (architecture.py:18-24) without im2col buffers or padded copies.
// Output rows go through the same (b, t) -> b*batch_stride + (t*d_t + d_off)*ld map, which
// lets the stride-2 data gradients write the even / odd input rows directly.
#include "ssb_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;

struct Gather {          // A(m, k)
  const float* base;
  int64_t batch_stride;  // elements between batch items of the source
  int rows_per_batch;    // m -> b = m / rows_per_batch, t = m % rows_per_batch
  int C;                 // channels per tap (k -> tap = k / C, c = k % C)
  int L_src;             // valid source rows per batch item
  int ld;                // elements between consecutive source rows
  int s_t, s_tap, off;   // source row = t*s_t + tap*s_tap + off
};

struct Scatter {         // C(m, n) destination
  float* base;
  int64_t batch_stride;
  int rows_per_batch;
  int ld;
  int d_t, d_off;        // dest row = t*d_t + d_off
};

struct WeightNT {        // B(n', k') = W[(tapmap[k'/Cb]*Nn + n') * ld + k' % Cb]
  const float* base;
  int ld, Cb, Nn;
  int tapmap[3];
};

struct Epilogue {
  Scatter out;
  const float* bias;       // [N] or null
  const float* mask_src;   // plain (M, N) row-major, ld = N: keep where mask_src > 0
  float mask_scale;
  int relu;
  int accumulate;          // C += result
  float drop_p;            // dropout applied after relu (0 = off)
  float drop_scale;
  uint32_t drop_thresh;
  uint64_t seed;
  const uint64_t* seed_src;
  uint32_t site;
};

__device__ __forceinline__ const float* gather_ptr(const Gather& g, int m, int k, bool& valid) {
  const int b = m / g.rows_per_batch;
  const int t = m - b * g.rows_per_batch;
  const int tap = k / g.C;
  const int c = k - tap * g.C;
  const int ts = t * g.s_t + tap * g.s_tap + g.off;
  valid = (unsigned)ts < (unsigned)g.L_src;
  return g.base + (int64_t)b * g.batch_stride + (int64_t)ts * g.ld + c;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// ---------------------------------------------------------------------------------------
// MODE 0: NN   (B = W[k, n], n contiguous)
// MODE 1: NT   (B = WeightNT, k contiguous)
// MODE 2: TN   (reduction over m; A gathered, B = G[m, n]); output plain (K, N)
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(THREADS)
gemm_kernel(const Gather ga, const float* __restrict__ Bplain, int ldb, const WeightNT wnt,
            const Epilogue ep, int M, int N, int K, int red_per_split) {
  // "o" = output index of the operand (m for A, n for B; for TN: k for A, n for B)
  // "r" = reduction index (k for NN/NT; m for TN)
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int o0a = blockIdx.y * BM;   // A-side output origin (m, or k for TN)
  const int o0b = blockIdx.x * BN;   // B-side output origin (n)
  const int OA = (MODE == 2) ? K : M;   // extent of A-side output index
  const int RED = (MODE == 2) ? M : K;  // reduction extent
  const int red_begin = blockIdx.z * red_per_split;
  const int red_end = min(RED, red_begin + red_per_split);

  // --- per-thread load coordinates -------------------------------------------------
  // RED-contiguous operand tiles: o = tid % 128, r4 = (tid / 128) * 4 (+8 second half)
  // OUT-contiguous operand tiles: o4 = (tid % 32) * 4, r = tid / 32 (+8 second half)
  const int rc_o = tid & 127, rc_r4 = (tid >> 7) << 2;
  const int oc_o4 = (tid & 31) << 2, oc_r = tid >> 5;

  float4 ra[2], rb[2];

  auto load_tiles = [&](int r0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // ---- A ----
      if (MODE != 2) {  // A(m, k): k contiguous
        const int m = o0a + rc_o, k = r0 + rc_r4 + 8 * h;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && k < red_end) {
          bool ok;
          const float* p = gather_ptr(ga, m, k, ok);
          if (ok) v = ldg4(p);
        }
        ra[h] = v;
      } else {          // TN: tile rows are m (reduction), columns are k (contiguous)
        const int m = r0 + oc_r + 8 * h, k = o0a + oc_o4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < red_end && k < K) {
          bool ok;
          const float* p = gather_ptr(ga, m, k, ok);
          if (ok) v = ldg4(p);
        }
        ra[h] = v;
      }
      // ---- B ----
      if (MODE == 0) {  // W[k, n], n contiguous
        const int k = r0 + oc_r + 8 * h, n = o0b + oc_o4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < red_end && n < N) v = ldg4(Bplain + (int64_t)k * ldb + n);
        rb[h] = v;
      } else if (MODE == 1) {  // B(n', k'), k' contiguous inside a tap block
        const int n = o0b + rc_o, k = r0 + rc_r4 + 8 * h;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N && k < red_end) {
          const int tap = k / wnt.Cb, c = k - tap * wnt.Cb;
          v = ldg4(wnt.base + ((int64_t)wnt.tapmap[tap] * wnt.Nn + n) * wnt.ld + c);
        }
        rb[h] = v;
      } else {  // TN: G[m, n]
        const int m = r0 + oc_r + 8 * h, n = o0b + oc_o4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < red_end && n < N) v = ldg4(Bplain + (int64_t)m * ldb + n);
        rb[h] = v;
      }
    }
  };

  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (MODE != 2) {
        const int r = rc_r4 + 8 * h;
        As[buf][r + 0][rc_o] = ra[h].x;
        As[buf][r + 1][rc_o] = ra[h].y;
        As[buf][r + 2][rc_o] = ra[h].z;
        As[buf][r + 3][rc_o] = ra[h].w;
      } else {
        *reinterpret_cast<float4*>(&As[buf][oc_r + 8 * h][oc_o4]) = ra[h];
      }
      if (MODE == 1) {
        const int r = rc_r4 + 8 * h;
        Bs[buf][r + 0][rc_o] = rb[h].x;
        Bs[buf][r + 1][rc_o] = rb[h].y;
        Bs[buf][r + 2][rc_o] = rb[h].z;
        Bs[buf][r + 3][rc_o] = rb[h].w;
      } else {
        *reinterpret_cast<float4*>(&Bs[buf][oc_r + 8 * h][oc_o4]) = rb[h];
      }
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int ntiles = (red_end > red_begin) ? (red_end - red_begin + BK - 1) / BK : 0;
  if (ntiles > 0) {
    load_tiles(red_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int tIdx = 0; tIdx < ntiles; ++tIdx) {
    const int buf = tIdx & 1;
    if (tIdx + 1 < ntiles) load_tiles(red_begin + (tIdx + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (tIdx + 1 < ntiles) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // --- epilogue -----------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int oa = o0a + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (oa >= OA) continue;
    float* row;
    if (MODE == 2) {
      row = ep.out.base + (int64_t)oa * ep.out.ld;
    } else {
      const int b = oa / ep.out.rows_per_batch, t = oa - b * ep.out.rows_per_batch;
      row = ep.out.base + (int64_t)b * ep.out.batch_stride +
            (int64_t)(t * ep.out.d_t + ep.out.d_off) * ep.out.ld;
    }
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = o0b + jh * 64 + tx * 4;
      if (n >= N) continue;
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2],
                             acc[i][jh * 4 + 3]);
      if (MODE == 2) {
        if (gridDim.z > 1) {
          atomicAdd(row + n + 0, v.x);
          atomicAdd(row + n + 1, v.y);
          atomicAdd(row + n + 2, v.z);
          atomicAdd(row + n + 3, v.w);
        } else {
          if (ep.accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(row + n);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *reinterpret_cast<float4*>(row + n) = v;
        }
        continue;
      }
      if (ep.bias) {
        const float4 bv = ldg4(ep.bias + n);
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      }
      if (ep.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      if (ep.drop_p > 0.f) {
        const uint64_t e = (uint64_t)oa * (uint64_t)N + (uint64_t)n;  // n % 4 == 0
        const uint4 rnd = ssb::dropout_bits4(ssb::eff_seed(ep.seed, ep.seed_src), ep.site, e >> 2);
        v.x = rnd.x >= ep.drop_thresh ? v.x * ep.drop_scale : 0.f;
        v.y = rnd.y >= ep.drop_thresh ? v.y * ep.drop_scale : 0.f;
        v.z = rnd.z >= ep.drop_thresh ? v.z * ep.drop_scale : 0.f;
        v.w = rnd.w >= ep.drop_thresh ? v.w * ep.drop_scale : 0.f;
      }
      if (ep.mask_src) {
        const float4 mk = ldg4(ep.mask_src + (int64_t)oa * N + n);
        v.x = mk.x > 0.f ? v.x * ep.mask_scale : 0.f;
        v.y = mk.y > 0.f ? v.y * ep.mask_scale : 0.f;
        v.z = mk.z > 0.f ? v.z * ep.mask_scale : 0.f;
        v.w = mk.w > 0.f ? v.w * ep.mask_scale : 0.f;
      }
      if (ep.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(row + n);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *reinterpret_cast<float4*>(row + n) = v;
    }
  }
}


// ---------------------------------------------------------------------------------------
// Thin reductions (K <= 32): the first ResBlock reads the 8-channel EMG (architecture.py:18,24:
// K = 3*8 and K = 8).  The 128x128 tile engine above wastes 80-95 % of its A tile there and the
// weight gradient (a 24 x 768 result reduced over 64 000 rows) ran at 5 TF/s.  Here a thread owns
// 4 output columns x ALL K: the K x 4 weight block (NN) or accumulator block (TN) lives in
// registers, A rows are warp-broadcast loads, and the wide operand streams through once.
// block = 64 column lanes (float4 -> 256 columns) x 4 row lanes.
// ---------------------------------------------------------------------------------------
constexpr int THIN_ROWS = 64;   // A rows staged per shared-memory chunk

// cooperative gather of rows [mc, mc + THIN_ROWS) of A into shared memory (zero beyond m1 / K):
// the index arithmetic of the gather and the global-load latency are paid once per chunk by the
// whole CTA instead of per row by every warp (148-register kernels run 8 warps per SM).
template <int K4>   // K padded to 4*K4
__device__ __forceinline__ void thin_stage_a(const Gather& ga, int mc, int m1, int K,
                                             float4 (*As)[K4]) {
  for (int i = threadIdx.x; i < THIN_ROWS * K4; i += blockDim.x) {
    const int r = i / K4, q = i - r * K4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mc + r < m1 && 4 * q < K) {
      bool valid;
      const float* src = gather_ptr(ga, mc + r, 4 * q, valid);
      if (valid) v = ldg4(src);
    }
    As[r][q] = v;
  }
}

template <int K4>
__global__ void __launch_bounds__(256)
thin_nn_kernel(const Gather ga, const float* __restrict__ W, int ldw, const Epilogue ep, int M,
               int N, int K, int rows_per_cta) {
  __shared__ float4 As[THIN_ROWS][K4];
  const int nl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int n = (blockIdx.x * 64 + nl) * 4;
  const bool n_ok = n < N;
  float4 w[4 * K4];
#pragma unroll
  for (int k = 0; k < 4 * K4; ++k)
    w[k] = (n_ok && k < K) ? ldg4(W + (int64_t)k * ldw + n) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ep.bias && n_ok) bias = ldg4(ep.bias + n);
  const int m0 = blockIdx.y * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  for (int mc = m0; mc < m1; mc += THIN_ROWS) {
    __syncthreads();
    thin_stage_a<K4>(ga, mc, m1, K, As);
    __syncthreads();
    if (!n_ok) continue;
    const int rend = min(THIN_ROWS, m1 - mc);
    for (int r = rl; r < rend; r += 4) {
      // packed fp32 FMAs (FFMA2, sm_100): two columns per instruction, same rounding as two FFMAs.
      // The kernel is issue-bound (ncu, r2 session 30: 60 M warp instructions, 62 % of them FFMA,
      // IPC 2.0 at 8 warps per SM), so halving the FMA count is the lever: 125.7 -> 113.9 us at cfg-1.
      float2 v01 = make_float2(bias.x, bias.y), v23 = make_float2(bias.z, bias.w);
#pragma unroll
      for (int q = 0; q < K4; ++q) {
        const float4 a = As[r][q];
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ww = w[4 * q + j];
          const float2 aa = make_float2(av[j], av[j]);
          v01 = __ffma2_rn(aa, make_float2(ww.x, ww.y), v01);
          v23 = __ffma2_rn(aa, make_float2(ww.z, ww.w), v23);
        }
      }
      float4 v = make_float4(v01.x, v01.y, v23.x, v23.y);
      if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      const int m = mc + r;
      const int b = m / ep.out.rows_per_batch, t = m - b * ep.out.rows_per_batch;
      float* row = ep.out.base + (int64_t)b * ep.out.batch_stride +
                   (int64_t)(t * ep.out.d_t + ep.out.d_off) * ep.out.ld;
      *reinterpret_cast<float4*>(row + n) = v;
    }
  }
}

template <int K4>
__global__ void __launch_bounds__(256)
thin_tn_kernel(const Gather ga, const float* __restrict__ G, int ldg, float* __restrict__ dW,
               int lddw, int M, int N, int K, int rows_per_cta) {
  __shared__ float4 As[THIN_ROWS][K4];
  __shared__ float4 red[3][64];
  const int nl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int n = (blockIdx.x * 64 + nl) * 4;
  const bool n_ok = n < N;
  float4 acc[4 * K4];
#pragma unroll
  for (int k = 0; k < 4 * K4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int m0 = blockIdx.y * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  for (int mc = m0; mc < m1; mc += THIN_ROWS) {
    __syncthreads();
    thin_stage_a<K4>(ga, mc, m1, K, As);
    __syncthreads();
    if (!n_ok) continue;
    const int rend = min(THIN_ROWS, m1 - mc);
    // the wide operand: up to 16 independent 16 B loads per thread in flight
    float4 g[THIN_ROWS / 4];
#pragma unroll
    for (int i = 0; i < THIN_ROWS / 4; ++i) {
      const int r = rl + 4 * i;
      g[i] = r < rend ? ldg4(G + (int64_t)(mc + r) * ldg + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < THIN_ROWS / 4; ++i) {
      const int r = rl + 4 * i;      // rows >= rend were staged as zeros
#pragma unroll
      for (int q = 0; q < K4; ++q) {
        const float4 a = As[r][q];
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // (scalar FFMAs on purpose: the packed form that helps thin_nn_kernel made this kernel,
          // which waits on its global loads with 96 live accumulators, 50 % slower: 190 -> 285 us)
          float4& c = acc[4 * q + j];
          c.x = fmaf(av[j], g[i].x, c.x); c.y = fmaf(av[j], g[i].y, c.y);
          c.z = fmaf(av[j], g[i].z, c.z); c.w = fmaf(av[j], g[i].w, c.w);
        }
      }
    }
  }
  // combine the 4 row lanes (fixed order), then one vector atomic per (k, 4 columns) and CTA
#pragma unroll
  for (int k = 0; k < 4 * K4; ++k) {
    if (k < K) {   // uniform
      if (rl > 0) red[rl - 1][nl] = acc[k];
      __syncthreads();
      if (rl == 0 && n_ok) {
        float4 c = acc[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 o = red[i][nl];
          c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
        }
        atomicAdd(reinterpret_cast<float4*>(dW + (int64_t)k * lddw + n), c);
      }
      __syncthreads();
    }
  }
}

// Weight gradient of a thin reduction, K <= 32 (the first convolution: K = 3 taps x 8
// channels, dW (24 x 768) reduced over 64 000 rows).  thin_tn_kernel keeps all K x 4 accumulators
// AND a 16-row register prefetch of G per thread: 217 registers, 8 warps per SM, 190 us at cfg-1 with
// the warps waiting on their global loads (ncu, r2 session 30: IPC 1.3, long-scoreboard stalls).
// Here the K range is split over two thread halves (2 x K4 accumulator float4 per thread) and G is
// staged through shared memory by cp.async, double-buffered, so the loads of chunk c+1 fly under the
// FMAs of chunk c and ~24 warps per SM are resident.  block = 64 column lanes x 2 K halves x 2 row
// lanes; a chunk is 16 rows x 256 columns of G (16 KB) + 16 rows of A.
constexpr int T2_ROWS = 16;

template <int K4>   // K padded to 4*K4, K4 even; each half owns K4/2 float4 of every A row
__global__ void __launch_bounds__(256)
thin_tn2_kernel(const Gather ga, const float* __restrict__ G, int ldg, float* __restrict__ dW,
                int lddw, int M, int N, int K, int rows_per_cta) {
  constexpr int KH4 = K4 / 2;
  __shared__ float4 Gs[2][T2_ROWS][64];
  __shared__ float4 As[2][T2_ROWS][K4];
  __shared__ float4 red[64];
  const int nl = threadIdx.x & 63, kh = (threadIdx.x >> 6) & 1, rl = threadIdx.x >> 7;
  const int n = (blockIdx.x * 64 + nl) * 4;
  const bool n_ok = n < N;
  float4 acc[4 * KH4];
#pragma unroll
  for (int k = 0; k < 4 * KH4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int m0 = blockIdx.y * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  const int nchunk = (m1 - m0 + T2_ROWS - 1) / T2_ROWS;
  auto stage = [&](int c) {
    const int mc = m0 + c * T2_ROWS, buf = c & 1;
    // G: 16 rows x 64 float4, 4 per thread, zero-filled outside [m1) x [N)
    for (int i = threadIdx.x; i < T2_ROWS * 64; i += 256) {
      const int r = i >> 6, q = i & 63;
      const int col = (blockIdx.x * 64 + q) * 4;
      const bool ok = mc + r < m1 && col < N;
      ssb::cp_async_16(ssb::smem_u32(&Gs[buf][r][q]), ok ? G + (int64_t)(mc + r) * ldg + col : G, ok ? 16 : 0);
    }
    ssb::cp_async_commit();
    for (int i = threadIdx.x; i < T2_ROWS * K4; i += 256) {
      const int r = i / K4, q = i - r * K4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mc + r < m1 && 4 * q < K) {
        bool valid;
        const float* src = gather_ptr(ga, mc + r, 4 * q, valid);
        if (valid) v = ldg4(src);
      }
      As[buf][r][q] = v;
    }
  };
  if (nchunk > 0) stage(0);
  for (int c = 0; c < nchunk; ++c) {
    ssb::cp_async_wait<0>();
    __syncthreads();                 // chunk c landed; every thread is done with chunk c-1's buffer
    if (c + 1 < nchunk) stage(c + 1);
    const int buf = c & 1;
#pragma unroll 2
    for (int r = rl; r < T2_ROWS; r += 2) {      // rows past m1 were staged as zeros
      const float4 g = Gs[buf][r][nl];
#pragma unroll
      for (int q = 0; q < KH4; ++q) {
        const float4 a = As[buf][r][kh * KH4 + q];
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4& cc = acc[4 * q + j];
          cc.x = fmaf(av[j], g.x, cc.x); cc.y = fmaf(av[j], g.y, cc.y);
          cc.z = fmaf(av[j], g.z, cc.z); cc.w = fmaf(av[j], g.w, cc.w);
        }
      }
    }
  }
  // combine the 2 row lanes (fixed order), then one vector atomic per (k, 4 columns) and CTA
#pragma unroll
  for (int k = 0; k < 4 * KH4; ++k) {
    const int kk = kh * 4 * KH4 + k;
#pragma unroll
    for (int h = 0; h < 2; ++h) {      // the two K halves take turns on the 64-entry exchange buffer
      __syncthreads();
      if (kh == h && rl == 1) red[nl] = acc[k];
      __syncthreads();
      if (kh == h && rl == 0 && n_ok && kk < K) {
        const float4 o = red[nl];
        float4 cv = acc[k];
        cv.x += o.x; cv.y += o.y; cv.z += o.z; cv.w += o.w;
        atomicAdd(reinterpret_cast<float4*>(dW + (int64_t)kk * lddw + n), cv);
      }
    }
  }
}

bool thin_tn2_enabled() {   // SSB_THIN_TN2=0: the register-prefetch kernel (A/B)
  static const bool on = [] {
    const char* e = getenv("SSB_THIN_TN2");
    return !(e && e[0] == '0');
  }();
  return on;
}

int thin_rows_per_cta(int64_t M, int64_t N) {
  const int col_blocks = (int)((N + 255) / 256);
  int row_blocks = (4 * ssb::num_sms() + col_blocks - 1) / col_blocks;
  int64_t per = (M + row_blocks - 1) / row_blocks;
  per = (per + 3) / 4 * 4;
  return (int)(per < 4 ? 4 : per);
}

int check_gather(const ssb_gather_t* g, int64_t M, int64_t K) {
  SSB_REQUIRE(g && g->base, "gemm: null A");
  SSB_REQUIRE(g->rows_per_batch > 0 && g->C > 0 && g->L_src > 0 && g->ld >= g->C,
              "gemm: bad gather geometry");
  SSB_REQUIRE(g->C % 4 == 0 && g->ld % 4 == 0 && g->batch_stride % 4 == 0 &&
                  ((uintptr_t)g->base & 15) == 0,
              "gemm: A must be 16 B aligned with channel counts divisible by 4");
  SSB_REQUIRE(K % 4 == 0 && K <= (int64_t)3 * g->C, "gemm: K=%lld incompatible with C=%d",
              (long long)K, g->C);
  SSB_REQUIRE(M < (1LL << 31) && K < (1LL << 31), "gemm: dimension too large");
  return SSB_OK;
}

Gather to_gather(const ssb_gather_t* g) {
  Gather r;
  r.base = g->base; r.batch_stride = g->batch_stride; r.rows_per_batch = g->rows_per_batch;
  r.C = g->C; r.L_src = g->L_src; r.ld = g->ld; r.s_t = g->s_t; r.s_tap = g->s_tap; r.off = g->off;
  return r;
}

int fill_epilogue(const ssb_epilogue_t* e, int64_t M, int64_t N, Epilogue* out) {
  SSB_REQUIRE(e && e->out.base, "gemm: null output");
  SSB_REQUIRE(!e->planes_out && !e->mask_planes && !e->mask_bits && !e->mask_bits_out,
              "gemm: split-plane epilogue operands exist on the tcgen05 engine only");
  SSB_REQUIRE(e->out.rows_per_batch > 0 && e->out.ld >= N && e->out.ld % 4 == 0 &&
                  e->out.batch_stride % 4 == 0 && ((uintptr_t)e->out.base & 15) == 0,
              "gemm: bad output geometry / alignment");
  SSB_REQUIRE(N % 4 == 0, "gemm: N=%lld must be a multiple of 4", (long long)N);
  SSB_REQUIRE(e->drop_p >= 0.f && e->drop_p < 1.f, "gemm: bad dropout p");
  out->out.base = e->out.base; out->out.batch_stride = e->out.batch_stride;
  out->out.rows_per_batch = e->out.rows_per_batch; out->out.ld = e->out.ld;
  out->out.d_t = e->out.d_t; out->out.d_off = e->out.d_off;
  out->bias = e->bias; out->mask_src = e->mask_src; out->mask_scale = e->mask_scale;
  out->relu = e->relu; out->accumulate = e->accumulate;
  out->drop_p = e->drop_p;
  out->drop_scale = e->drop_p > 0.f ? 1.f / (1.f - e->drop_p) : 1.f;
  const double th = (double)e->drop_p * 4294967296.0;
  out->drop_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
  out->seed = e->seed; out->seed_src = ssb::seed_source(); out->site = e->site;
  return SSB_OK;
}

}  // namespace

extern "C" {

int ssb_gemm_nn(const ssb_gather_t* A, const float* W, int64_t ldw, const ssb_epilogue_t* epi,
                int64_t M, int64_t N, int64_t K, void* stream) {
  if (M == 0 || N == 0) return SSB_OK;
  if (int rc = check_gather(A, M, K)) return rc;
  SSB_REQUIRE(W && ldw >= N && ldw % 4 == 0 && ((uintptr_t)W & 15) == 0, "gemm_nn: bad W");
  Epilogue ep;
  if (int rc = fill_epilogue(epi, M, N, &ep)) return rc;
  if (K <= 32 && M >= 4096 && ep.drop_p == 0.f && !ep.mask_src && !ep.accumulate) {
    const int per = thin_rows_per_cta(M, N);
    dim3 tg((unsigned)((N + 255) / 256), (unsigned)((M + per - 1) / per));
    cudaStream_t st = (cudaStream_t)stream;
    const Gather ga = to_gather(A);
    switch ((K + 3) / 4 > 4 ? ((K + 7) / 8) * 2 : (int)((K + 3) / 4)) {
      case 1: case 2: thin_nn_kernel<2><<<tg, 256, 0, st>>>(ga, W, (int)ldw, ep, (int)M, (int)N, (int)K, per); break;
      case 3: case 4: thin_nn_kernel<4><<<tg, 256, 0, st>>>(ga, W, (int)ldw, ep, (int)M, (int)N, (int)K, per); break;
      case 6: thin_nn_kernel<6><<<tg, 256, 0, st>>>(ga, W, (int)ldw, ep, (int)M, (int)N, (int)K, per); break;
      default: thin_nn_kernel<8><<<tg, 256, 0, st>>>(ga, W, (int)ldw, ep, (int)M, (int)N, (int)K, per); break;
    }
    SSB_LAUNCH_CHECK("thin_nn");
    return SSB_OK;
  }
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), 1);
  SSB_REQUIRE(grid.y <= 65535, "gemm_nn: M too large for grid.y");
  WeightNT dummy = {};
  gemm_kernel<0><<<grid, THREADS, 0, (cudaStream_t)stream>>>(to_gather(A), W, (int)ldw, dummy, ep,
                                                             (int)M, (int)N, (int)K, (int)K);
  SSB_LAUNCH_CHECK("gemm_nn");
  return SSB_OK;
}

int ssb_gemm_nt(const ssb_gather_t* A, const float* W, int64_t ldw, int64_t Cb, int tap0, int tap1,
                int tap2, const ssb_epilogue_t* epi, int64_t M, int64_t N, int64_t K,
                void* stream) {
  if (M == 0 || N == 0) return SSB_OK;
  if (int rc = check_gather(A, M, K)) return rc;
  SSB_REQUIRE(W && ldw >= Cb && ldw % 4 == 0 && Cb % 4 == 0 && ((uintptr_t)W & 15) == 0 &&
                  K % Cb == 0 && K / Cb <= 3,
              "gemm_nt: bad W geometry");
  Epilogue ep;
  if (int rc = fill_epilogue(epi, M, N, &ep)) return rc;
  WeightNT w;
  w.base = W; w.ld = (int)ldw; w.Cb = (int)Cb; w.Nn = (int)N;
  w.tapmap[0] = tap0; w.tapmap[1] = tap1; w.tapmap[2] = tap2;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), 1);
  SSB_REQUIRE(grid.y <= 65535, "gemm_nt: M too large for grid.y");
  gemm_kernel<1><<<grid, THREADS, 0, (cudaStream_t)stream>>>(to_gather(A), nullptr, 0, w, ep,
                                                             (int)M, (int)N, (int)K, (int)K);
  SSB_LAUNCH_CHECK("gemm_nt");
  return SSB_OK;
}

int ssb_gemm_tn(const ssb_gather_t* A, const float* G, int64_t ldg, float* dW, int64_t lddw,
                int accumulate, int64_t M, int64_t N, int64_t K, void* stream) {
  if (K == 0 || N == 0) return SSB_OK;
  if (int rc = check_gather(A, M, K)) return rc;
  SSB_REQUIRE(G && dW && ldg >= N && lddw >= N && ldg % 4 == 0 && lddw % 4 == 0 && N % 4 == 0 &&
                  ((uintptr_t)G & 15) == 0 && ((uintptr_t)dW & 15) == 0,
              "gemm_tn: bad G / dW geometry");
  cudaStream_t st = (cudaStream_t)stream;
  if (K <= 32 && M >= 4096) {
    if (!accumulate)
      SSB_CUDA(cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)N * 4, (size_t)K, st));
    const int per = thin_rows_per_cta(M, N);
    dim3 tg((unsigned)((N + 255) / 256), (unsigned)((M + per - 1) / per));
    const Gather ga = to_gather(A);
    if (thin_tn2_enabled()) {
      switch ((int)((K + 7) / 8) * 2) {   // K4 even
        case 2: thin_tn2_kernel<2><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
        case 4: thin_tn2_kernel<4><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
        case 6: thin_tn2_kernel<6><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
        default: thin_tn2_kernel<8><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
      }
      SSB_LAUNCH_CHECK("thin_tn2");
      return SSB_OK;
    }
    switch ((K + 3) / 4 > 4 ? ((K + 7) / 8) * 2 : (int)((K + 3) / 4)) {
      case 1: case 2: thin_tn_kernel<2><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
      case 3: case 4: thin_tn_kernel<4><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
      case 6: thin_tn_kernel<6><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
      default: thin_tn_kernel<8><<<tg, 256, 0, st>>>(ga, G, (int)ldg, dW, (int)lddw, (int)M, (int)N, (int)K, per); break;
    }
    SSB_LAUNCH_CHECK("thin_tn");
    return SSB_OK;
  }
  const int tiles = (int)(((N + BN - 1) / BN) * ((K + BM - 1) / BM));
  int splits = 1;
  const int sms = ssb::num_sms();
  if (M > 4 * BK) {
    splits = (2 * sms + tiles - 1) / tiles;
    const int64_t max_splits = (M + 8 * BK - 1) / (8 * BK);
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    if (splits > 1024) splits = 1024;
  }
  int64_t per = (M + splits - 1) / splits;
  per = ((per + BK - 1) / BK) * BK;
  splits = (int)((M + per - 1) / per);
  if (splits < 1) splits = 1;
  if (splits > 1 && !accumulate)
    SSB_CUDA(cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)N * 4, (size_t)K, st));
  Epilogue ep = {};
  ep.out.base = dW; ep.out.ld = (int)lddw; ep.out.rows_per_batch = 1; ep.accumulate = accumulate;
  WeightNT dummy = {};
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((K + BM - 1) / BM), (unsigned)splits);
  gemm_kernel<2><<<grid, THREADS, 0, st>>>(to_gather(A), G, (int)ldg, dummy, ep, (int)M, (int)N,
                                           (int)K, (int)per);
  SSB_LAUNCH_CHECK("gemm_tn");
  return SSB_OK;
}

}  // extern "C"
