// Element-wise stages of the tensor-core attention path (sm_100a).
//
// The contractions of attention (Q K^T, Q E^T, P V and the four gradient products) run as
// batched tcgen05 GEMMs (gemm_tc.cu, one batch item per (b, h)); this file holds what sits
// between them.  Same mathematics as attn.cu (transformer.py:99-110 with the relative-position
// logits of :162-297 in closed form), different schedule: logits are formed per (b, h) as a
// dense (T x T) tile on the tensor cores and masked to the exact band |k - q| <= W here.
//
//   pad_split_heads        fp32 (rows, G*dh)          -> bf16 planes (2, rows, G, 128), zero padded
//   transpose_split_heads  fp32 k / v (B*T, ...)      -> bf16 planes (2, B*H, 128, Tp)  [d][key]
//   attn_softmax_fwd       S (in place -> P), R band  -> P fp32, dropout(P) planes
//   attn_ds_bwd            P, dP                      -> scaled dS planes, dS band planes
#include "ssb_common.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>

namespace {

constexpr int HP = 128;  // padded head dimension (one 128-column group = two 64-wide k-blocks)

__device__ __forceinline__ void split2(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// x: (rows, ld_in) fp32, head g occupies columns [col_off + g*dh, +dh).  out: (2, rows, G, HP).
__global__ void __launch_bounds__(256)
pad_split_heads_kernel(const float* __restrict__ x, int64_t rows, int ld_in, int col_off, int G,
                       int dh, __nv_bfloat16* __restrict__ out) {
  const int64_t n8 = rows * G * (HP / 8);
  const int64_t plane = rows * G * HP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int a8 = (int)(i % (HP / 8)) * 8;
    const int64_t rg = i / (HP / 8);
    const int g = (int)(rg % G);
    const int64_t r = rg / G;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int a = a8 + j;
      const float v = a < dh ? __ldg(x + r * ld_in + col_off + g * dh + a) : 0.f;
      split2(v, hi[j], lo[j]);
    }
    *reinterpret_cast<uint4*>(out + rg * HP + a8) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(out + plane + rg * HP + a8) = *reinterpret_cast<const uint4*>(lo);
  }
}

// out[pl][(b*H + h)][a][t] = split(x[(b*T + t), col_off + h*dh + a]), zero for a >= dh or t >= T.
// grid: (Tp/32, HP/32, B*H), block (32, 8): 32x32 tile transpose through shared memory.
__global__ void __launch_bounds__(256)
transpose_split_heads_kernel(const float* __restrict__ x, int ld_in, int col_off, int T, int H,
                             int dh, int Tp, __nv_bfloat16* __restrict__ out, int64_t plane) {
  __shared__ float tile[32][33];
  const int bh = blockIdx.z, b = bh / H, h = bh - b * H;
  const int t0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i, a = a0 + threadIdx.x;
    tile[i][threadIdx.x] =
        (t < T && a < dh) ? __ldg(x + ((int64_t)b * T + t) * ld_in + col_off + h * dh + a) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int a = a0 + i, t = t0 + threadIdx.x;
    __nv_bfloat16 hi, lo;
    split2(tile[threadIdx.x][i], hi, lo);
    const int64_t o = ((int64_t)bh * HP + a) * Tp + t;
    out[o] = hi;
    out[plane + o] = lo;
  }
}

struct SmParams {
  float* S;             // (BH, T, Tp) in: raw q.k ; out: P (softmax), columns >= T zero
  const float* R;       // (BH, T, RW) positional logits
  __nv_bfloat16* Pd;    // (2, BH, T, Tp) dropout(P) planes
  const float* dP;      // bwd: (BH, T, Tp)
  __nv_bfloat16* dSp;   // bwd: (2, BH, T, Tp) planes of scale * dS
  __nv_bfloat16* dSb;   // bwd: (2, B*T, H, RWp) planes of dS in band layout
  int BH, H, T, Tp, W, RW, RWp;
  float scale, drop_p, drop_scale;
  uint32_t drop_thresh;
  uint64_t seed;
  const uint64_t* seed_src;
  uint32_t site;
};

constexpr int SM_MAXV = 8;  // float4 per lane: Tp <= 1024

// one warp per (bh, q) row
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(const SmParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (int64_t)p.BH * p.T) return;
  const uint64_t seed = ssb::eff_seed(p.seed, p.seed_src);
  const int q = (int)(row % p.T);
  float* Srow = p.S + row * p.Tp;
  const float* Rrow = p.R + row * p.RW;
  const int nv = p.Tp >> 2;
  float4 v[SM_MAXV];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int c4 = lane + i * 32;
    if (c4 < nv) {
      float4 s = *reinterpret_cast<const float4*>(Srow + 4 * c4);
      float e[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * c4 + j, rel = k - q + p.W;
        e[j] = (k < p.T && rel >= 0 && rel <= 2 * p.W) ? fmaf(e[j], p.scale, __ldg(Rrow + rel))
                                                       : -CUDART_INF_F;
        m = fmaxf(m, e[j]);
      }
      v[i] = make_float4(e[0], e[1], e[2], e[3]);
    }
  }
  m = ssb::warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int c4 = lane + i * 32;
    if (c4 < nv) {
      v[i].x = expf(v[i].x - m); v[i].y = expf(v[i].y - m);
      v[i].z = expf(v[i].z - m); v[i].w = expf(v[i].w - m);
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  sum = ssb::warp_sum(sum);
  const float inv = 1.f / sum;
  const int64_t plane = (int64_t)p.BH * p.T * p.Tp;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int c4 = lane + i * 32;
    if (c4 < nv) {
      float4 pr = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
      *reinterpret_cast<float4*>(Srow + 4 * c4) = pr;
      if (p.drop_p > 0.f) {
        const uint4 rnd = ssb::dropout_bits4(seed, p.site, (uint64_t)row * nv + c4);
        pr.x = rnd.x >= p.drop_thresh ? pr.x * p.drop_scale : 0.f;
        pr.y = rnd.y >= p.drop_thresh ? pr.y * p.drop_scale : 0.f;
        pr.z = rnd.z >= p.drop_thresh ? pr.z * p.drop_scale : 0.f;
        pr.w = rnd.w >= p.drop_thresh ? pr.w * p.drop_scale : 0.f;
      }
      __align__(8) __nv_bfloat16 hi[4], lo[4];
      split2(pr.x, hi[0], lo[0]); split2(pr.y, hi[1], lo[1]);
      split2(pr.z, hi[2], lo[2]); split2(pr.w, hi[3], lo[3]);
      *reinterpret_cast<uint2*>(p.Pd + row * p.Tp + 4 * c4) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(p.Pd + plane + row * p.Tp + 4 * c4) =
          *reinterpret_cast<const uint2*>(lo);
    }
  }
}

// dPm = dP * dropmask / (1-p);  dS = P * (dPm - sum_k P*dPm)
// outputs: planes of scale*dS (dense, for dQ and dK) and planes of dS in band layout (for the
// positional part of dQ, where logits are not scaled).
__global__ void __launch_bounds__(256) attn_ds_bwd_kernel(const SmParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (int64_t)p.BH * p.T) return;
  const uint64_t seed = ssb::eff_seed(p.seed, p.seed_src);
  const int q = (int)(row % p.T);
  const int64_t bh = row / p.T;
  const int b = (int)(bh / p.H), h = (int)(bh % p.H);
  const float* Prow = p.S + row * p.Tp;
  const float* dProw = p.dP + row * p.Tp;
  const int nv = p.Tp >> 2;
  float4 pv[SM_MAXV], dv[SM_MAXV];
  float delta = 0.f;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int c4 = lane + i * 32;
    if (c4 < nv) {
      pv[i] = *reinterpret_cast<const float4*>(Prow + 4 * c4);
      float4 d = *reinterpret_cast<const float4*>(dProw + 4 * c4);
      if (p.drop_p > 0.f) {
        const uint4 rnd = ssb::dropout_bits4(seed, p.site, (uint64_t)row * nv + c4);
        d.x = rnd.x >= p.drop_thresh ? d.x * p.drop_scale : 0.f;
        d.y = rnd.y >= p.drop_thresh ? d.y * p.drop_scale : 0.f;
        d.z = rnd.z >= p.drop_thresh ? d.z * p.drop_scale : 0.f;
        d.w = rnd.w >= p.drop_thresh ? d.w * p.drop_scale : 0.f;
      }
      // columns >= T of dP were never written by the GEMM: keep garbage (NaN) out
      if (4 * c4 + 0 >= p.T) d.x = 0.f;
      if (4 * c4 + 1 >= p.T) d.y = 0.f;
      if (4 * c4 + 2 >= p.T) d.z = 0.f;
      if (4 * c4 + 3 >= p.T) d.w = 0.f;
      dv[i] = d;
      delta += pv[i].x * d.x + pv[i].y * d.y + pv[i].z * d.z + pv[i].w * d.w;
    }
  }
  delta = ssb::warp_sum(delta);
  const int64_t plane = (int64_t)p.BH * p.T * p.Tp;
  const int64_t band_row = (((int64_t)b * p.T + q) * p.H + h) * p.RWp;
  const int64_t band_plane = (int64_t)(p.BH / p.H) * p.T * p.H * p.RWp;
  // zero the band row first (entries outside [0, 2W] and beyond the sequence stay zero)
  for (int r = lane; r < p.RWp; r += 32) {
    p.dSb[band_row + r] = __float2bfloat16_rn(0.f);
    p.dSb[band_plane + band_row + r] = __float2bfloat16_rn(0.f);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int c4 = lane + i * 32;
    if (c4 < nv) {
      const float ds[4] = {pv[i].x * (dv[i].x - delta), pv[i].y * (dv[i].y - delta),
                           pv[i].z * (dv[i].z - delta), pv[i].w * (dv[i].w - delta)};
      __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split2(ds[j] * p.scale, hi[j], lo[j]);
        const int k = 4 * c4 + j, rel = k - q + p.W;
        if (k < p.T && rel >= 0 && rel <= 2 * p.W) {   // P is exactly 0 outside the band
          __nv_bfloat16 bhv, blv;
          split2(ds[j], bhv, blv);
          p.dSb[band_row + rel] = bhv;
          p.dSb[band_plane + band_row + rel] = blv;
        }
      }
      *reinterpret_cast<uint2*>(p.dSp + row * p.Tp + 4 * c4) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(p.dSp + plane + row * p.Tp + 4 * c4) =
          *reinterpret_cast<const uint2*>(lo);
    }
  }
}

int fill(SmParams* p, int64_t B, int64_t H, int64_t T, int64_t Tp, int64_t W, int64_t RW,
         int64_t RWp, int64_t dh, float drop_p, uint64_t seed, uint32_t site) {
  SSB_REQUIRE(B >= 1 && H >= 1 && T >= 1 && Tp >= T && Tp % 64 == 0 && Tp <= 1024,
              "attn_tc: T=%lld Tp=%lld (need Tp %% 64 == 0, Tp <= 1024)", (long long)T,
              (long long)Tp);
  SSB_REQUIRE(W >= 0 && RW >= 2 * W + 1 && RWp >= RW && RWp % 64 == 0, "attn_tc: bad band sizes");
  SSB_REQUIRE(dh >= 1 && dh <= HP, "attn_tc: head dim %lld > 128", (long long)dh);
  SSB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attn_tc: bad dropout p");
  p->BH = (int)(B * H); p->H = (int)H; p->T = (int)T; p->Tp = (int)Tp; p->W = (int)W;
  p->RW = (int)RW; p->RWp = (int)RWp;
  p->scale = 1.0f / sqrtf((float)dh);
  p->drop_p = drop_p; p->drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const double th = (double)drop_p * 4294967296.0;
  p->drop_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
  p->seed = seed; p->seed_src = ssb::seed_source(); p->site = site;
  return SSB_OK;
}

}  // namespace

extern "C" {

int ssb_pad_split_heads(const float* x, int64_t rows, int64_t ld_in, int64_t col_off, int64_t G,
                        int64_t dh, void* planes, void* stream) {
  SSB_REQUIRE(x && planes && rows >= 1 && G >= 1 && dh >= 1 && dh <= HP && ld_in >= col_off + G * dh,
              "pad_split_heads: bad arguments");
  SSB_REQUIRE(((uintptr_t)planes & 15) == 0, "pad_split_heads: planes must be 16 B aligned");
  const int64_t n8 = rows * G * (HP / 8);
  const int64_t blocks = (n8 + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  pad_split_heads_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, (int)ld_in, (int)col_off,
                                                                 (int)G, (int)dh,
                                                                 (__nv_bfloat16*)planes);
  SSB_LAUNCH_CHECK("pad_split_heads");
  return SSB_OK;
}

int ssb_transpose_split_heads(const float* x, int64_t ld_in, int64_t col_off, int64_t B, int64_t T,
                              int64_t H, int64_t dh, int64_t Tp, void* planes, void* stream) {
  SSB_REQUIRE(x && planes && B >= 1 && T >= 1 && H >= 1 && dh >= 1 && dh <= HP && Tp >= T &&
                  Tp % 32 == 0 && B * H <= 65535,
              "transpose_split_heads: bad arguments");
  dim3 grid((unsigned)(Tp / 32), HP / 32, (unsigned)(B * H));
  transpose_split_heads_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
      x, (int)ld_in, (int)col_off, (int)T, (int)H, (int)dh, (int)Tp, (__nv_bfloat16*)planes,
      B * H * HP * Tp);
  SSB_LAUNCH_CHECK("transpose_split_heads");
  return SSB_OK;
}

int ssb_attn_softmax_fwd(float* S, const float* R, int64_t B, int64_t H, int64_t T, int64_t Tp,
                         int64_t W, int64_t RW, int64_t dh, float drop_p, uint64_t seed,
                         uint32_t site, void* Pd_planes, void* stream) {
  SmParams p = {};
  if (int rc = fill(&p, B, H, T, Tp, W, RW, 64 * ((RW + 63) / 64), dh, drop_p, seed, site)) return rc;
  SSB_REQUIRE(S && R && Pd_planes, "attn_softmax_fwd: null pointer");
  p.S = S; p.R = R; p.Pd = (__nv_bfloat16*)Pd_planes;
  const int64_t rows = B * H * T;
  attn_softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  SSB_LAUNCH_CHECK("attn_softmax_fwd");
  return SSB_OK;
}

int ssb_attn_ds_bwd(const float* P, const float* dP, int64_t B, int64_t H, int64_t T, int64_t Tp,
                    int64_t W, int64_t RWp, int64_t dh, float drop_p, uint64_t seed,
                    uint32_t site, void* dS_planes, void* dSband_planes, void* stream) {
  SmParams p = {};
  if (int rc = fill(&p, B, H, T, Tp, W, 2 * W + 1, RWp, dh, drop_p, seed, site)) return rc;
  SSB_REQUIRE(P && dP && dS_planes && dSband_planes, "attn_ds_bwd: null pointer");
  p.S = const_cast<float*>(P); p.dP = dP; p.dSp = (__nv_bfloat16*)dS_planes;
  p.dSb = (__nv_bfloat16*)dSband_planes;
  const int64_t rows = B * H * T;
  attn_ds_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  SSB_LAUNCH_CHECK("attn_ds_bwd");
  return SSB_OK;
}

}  // extern "C"
