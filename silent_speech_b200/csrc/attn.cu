// Banded relative-position self-attention, forward and backward (sm_100a, fp32 CUDA cores).
//
// Replaces the attention core of transformer.py:99-110 together with the relative-position
// machinery of transformer.py:162-297.  Closed form of the reference:
//     logits[q,k] = (q.k)/sqrt(dh) + q.E[h, k-q+W]      if |k-q| <= W      (W = 99)
//                 = -1e8  (=> softmax probability exactly 0 in fp32)        otherwise
// so attention is EXACTLY banded: only the 2W+1 keys around each query contribute.
//
// Data layout.  qkv: (B*T, 3*D) rows = tokens (b-major), columns [q | k | v], head h at
// h*dh.  The positional logits  R[b,t,h,r] = q[b,t,h,:].E[h,r,:]  (r = k-q+W, row stride
// H*RW, RW = 2W+2 = 200) are produced by the GEMM engine (one NT call per head) and consumed
// here; the backward pass hands dS back in the same band layout so that the positional part
// of dQ is again one GEMM per head.  Nothing of size T x T is ever materialised: the
// reference's 256 MB logits + 511 MB position tensors become 2 x 100 MB band tensors.
//
// Kernels (one CTA = one (b, h, tile of TQ=32 queries or keys), 256 threads):
//   band_attn_fwd:  S = scale*Q K^T + R -> softmax -> P (saved, band layout) -> dropout -> O = P V
//   band_attn_bwd_q: dP = dO V^T -> dS = P*(dP - sum(P*dP)) (band layout, saved) -> dQ = scale*dS K
//   band_attn_bwd_kv: dK = scale * dS^T Q,  dV = Pdrop^T dO   (gathers the band transposed)
#include "ssb_common.cuh"
#include <math_constants.h>

namespace {

constexpr int TQ = 32;         // queries (or keys) per CTA
constexpr int MAXW = 99;       // band half-width supported by the shared-memory plan
constexpr int NKW = TQ + 2 * MAXW;  // 230 window rows
constexpr int NKW_PAD = 232;
constexpr int THREADS = 256;

struct AttnParams {
  const float* qkv;   // (B*T, 3D)
  const float* R;     // (B*T, H, RW) positional logits (fwd) ; unused in bwd_kv
  float* P;           // (B*T, H, RW) softmax probabilities, band layout (fwd out / bwd in)
  float* O;           // (B*T, D)    fwd out
  const float* dO;    // (B*T, D)    bwd in
  float* dS;          // (B*T, H, RW) bwd: written by bwd_q, read by bwd_kv
  float* dqkv;        // (B*T, 3D)   bwd out
  int B, T, H, dh, D, W, RW;
  float scale;
  float drop_p, drop_scale;
  uint32_t drop_thresh;
  uint64_t seed;
  const uint64_t* seed_src;
  uint32_t site;
};

__device__ __forceinline__ int smem_ld(int dh) { return dh + 4; }  // keeps LDS.128 conflict-free

// dropout keep bits for the 4 band elements rel = 4*g .. 4*g+3 of (token row, head):
// one Philox call per group (RW % 4 == 0 keeps groups aligned with the element index).
__device__ __forceinline__ uint32_t keep_bits4(const AttnParams& p, int64_t token, int h, int g) {
  const uint64_t e4 = (((uint64_t)token * p.H + h) * p.RW >> 2) + g;
  const uint4 r = ssb::dropout_bits4(ssb::eff_seed(p.seed, p.seed_src), p.site, e4);
  return (r.x >= p.drop_thresh ? 1u : 0u) | (r.y >= p.drop_thresh ? 2u : 0u) |
         (r.z >= p.drop_thresh ? 4u : 0u) | (r.w >= p.drop_thresh ? 8u : 0u);
}

// Load `nrows` rows of one head (dh floats each) of `src` (row stride ld) starting at token
// row `t0` of batch item b into smem[r][0..dh); rows outside [0, T) are zero-filled.
__device__ __forceinline__ void load_rows(float* sm, int ldS, const float* src, int64_t ld,
                                          int64_t tok_base, int t0, int nrows, int T, int dh,
                                          float mul) {
  const int nd4 = dh >> 2;
  for (int i = threadIdx.x; i < nrows * nd4; i += THREADS) {
    const int r = i / nd4, c = (i - r * nd4) << 2;
    const int t = t0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) {
      v = __ldg(reinterpret_cast<const float4*>(src + (tok_base + t) * ld + c));
      v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
    }
    *reinterpret_cast<float4*>(sm + r * ldS + c) = v;
  }
}

// S[qi][kw] (32 x NKW) = sum_d Qs[qi][d] * Ks[kw][d]; thread -> 4 window rows x 8 queries.
// Window rows kw = kk + 64*j (j < 4), queries qg*8 .. qg*8+7.
__device__ __forceinline__ void qk_rect(const float* Qs, const float* Ks, int ldS, int dh,
                                        float (&acc)[4][8]) {
  const int kk = threadIdx.x & 63, qg = threadIdx.x >> 6;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  for (int d = 0; d < dh; d += 4) {
    float4 kv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kw = kk + 64 * j;
      kv[j] = kw < NKW ? *reinterpret_cast<const float4*>(Ks + kw * ldS + d)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 qv = *reinterpret_cast<const float4*>(Qs + (qg * 8 + i) * ldS + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][i] = fmaf(qv.x, kv[j].x, acc[j][i]);
        acc[j][i] = fmaf(qv.y, kv[j].y, acc[j][i]);
        acc[j][i] = fmaf(qv.z, kv[j].z, acc[j][i]);
        acc[j][i] = fmaf(qv.w, kv[j].w, acc[j][i]);
      }
    }
  }
}

// out[qi][d] (32 x dh) = sum_kw Sc[qi][kw] * Vs[kw][d]; thread -> 4 queries x 4 d.
__device__ __forceinline__ void pv_rect(const float* Sc, const float* Vs, int ldS, int dh,
                                        int kw_lo, int kw_hi, float4 (&acc)[4], int& d4,
                                        int& qg, bool& active) {
  const int nd4 = dh >> 2;
  d4 = (threadIdx.x % nd4) << 2;
  qg = threadIdx.x / nd4;  // 0..7 when active
  active = qg < 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!active) return;
  for (int kw = kw_lo; kw < kw_hi; ++kw) {
    const float4 v = *reinterpret_cast<const float4*>(Vs + kw * ldS + d4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float pw = Sc[(qg * 4 + i) * NKW_PAD + kw];
      acc[i].x = fmaf(pw, v.x, acc[i].x);
      acc[i].y = fmaf(pw, v.y, acc[i].y);
      acc[i].z = fmaf(pw, v.z, acc[i].z);
      acc[i].w = fmaf(pw, v.w, acc[i].w);
    }
  }
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) band_attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ldS = smem_ld(p.dh);
  float* Qs = smem;                     // [TQ][ldS]
  float* KVs = Qs + TQ * ldS;           // [NKW][ldS]
  float* Sc = KVs + NKW * ldS;          // [TQ][NKW_PAD]

  const int q0 = blockIdx.x * TQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int64_t tok_base = (int64_t)b * p.T;
  const int64_t ld = 3 * (int64_t)p.D;
  const int kbase = q0 - p.W;  // window row kw <-> key kbase + kw;  rel = kw - qi + (MAXW - W)... (W == MAXW layout)

  load_rows(Qs, ldS, p.qkv + h * p.dh, ld, tok_base, q0, TQ, p.T, p.dh, p.scale);
  load_rows(KVs, ldS, p.qkv + p.D + h * p.dh, ld, tok_base, kbase, TQ + 2 * p.W, p.T, p.dh, 1.f);
  __syncthreads();

  {
    float acc[4][8];
    qk_rect(Qs, KVs, ldS, p.dh, acc);
    const int kk = threadIdx.x & 63, qg = threadIdx.x >> 6;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kw = kk + 64 * j;
      if (kw >= NKW_PAD) continue;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int qi = qg * 8 + i;
        const int rel = kw - qi;           // = key - query + W
        const int key = kbase + kw, qt = q0 + qi;
        float s = -CUDART_INF_F;
        if (rel >= 0 && rel <= 2 * p.W && key >= 0 && key < p.T && qt < p.T)
          s = acc[j][i] + __ldg(p.R + ((tok_base + qt) * p.H + h) * p.RW + rel);
        Sc[qi * NKW_PAD + kw] = s;
      }
    }
  }
  __syncthreads();

  // softmax per query row (one warp per row, 4 rows per warp); P saved pre-dropout
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int qi = warp; qi < TQ; qi += 8) {
      const int qt = q0 + qi;
      float* row = Sc + qi * NKW_PAD;
      if (qt >= p.T) {
        for (int kw = lane; kw < NKW_PAD; kw += 32) row[kw] = 0.f;
        continue;
      }
      float m = -CUDART_INF_F;
      for (int kw = lane; kw < NKW_PAD; kw += 32) m = fmaxf(m, row[kw]);
      m = ssb::warp_max(m);
      float sum = 0.f;
      for (int kw = lane; kw < NKW_PAD; kw += 32) {
        const float e = expf(row[kw] - m);  // masked entries: exp(-inf) = 0
        row[kw] = e;
        sum += e;
      }
      sum = ssb::warp_sum(sum);
      const float inv = 1.f / sum;
      __syncwarp();
      float* Prow = p.P ? p.P + ((tok_base + qt) * p.H + h) * p.RW : nullptr;
      // normalise, save P (pre-dropout, band layout), apply dropout; 4 rels per lane-step
      for (int g = lane; g < (p.RW >> 2); g += 32) {
        const uint32_t keep = p.drop_p > 0.f ? keep_bits4(p, tok_base + qt, h, g) : 0xfu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rel = 4 * g + j;
          if (rel > 2 * p.W) {
            if (Prow) Prow[rel] = 0.f;  // padding columns of the band tensor
            continue;
          }
          const int kw = rel + qi;
          float pr = row[kw] * inv;
          if (Prow) Prow[rel] = pr;
          row[kw] = ((keep >> j) & 1u) ? pr * p.drop_scale : 0.f;
        }
      }
      // entries left of the band (kw < qi) and right of it are not read by the PV product
      for (int kw = lane; kw < NKW_PAD; kw += 32) {
        const int rel = kw - qi;
        if (rel < 0 || rel > 2 * p.W) row[kw] = 0.f;
      }
    }
  }
  __syncthreads();  // everyone is done with K in KVs and P is complete

  load_rows(KVs, ldS, p.qkv + 2 * p.D + h * p.dh, ld, tok_base, kbase, TQ + 2 * p.W, p.T, p.dh, 1.f);
  __syncthreads();
  {
    float4 acc[4];
    int d4, qg;
    bool active;
    pv_rect(Sc, KVs, ldS, p.dh, 0, TQ + 2 * p.W, acc, d4, qg, active);
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qt = q0 + qg * 4 + i;
        if (qt < p.T)
          *reinterpret_cast<float4*>(p.O + (tok_base + qt) * p.D + h * p.dh + d4) = acc[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward, query side: dS (band) and dQ
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) band_attn_bwd_q_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ldS = smem_ld(p.dh);
  float* Qs = smem;                     // dO tile, later unused
  float* KVs = Qs + TQ * ldS;           // V window, then K window
  float* Sc = KVs + NKW * ldS;          // dP -> dS

  const int q0 = blockIdx.x * TQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int64_t tok_base = (int64_t)b * p.T;
  const int64_t ld = 3 * (int64_t)p.D;
  const int kbase = q0 - p.W;

  load_rows(Qs, ldS, p.dO + h * p.dh, p.D, tok_base, q0, TQ, p.T, p.dh, 1.f);
  load_rows(KVs, ldS, p.qkv + 2 * p.D + h * p.dh, ld, tok_base, kbase, TQ + 2 * p.W, p.T, p.dh, 1.f);
  __syncthreads();
  {
    float acc[4][8];
    qk_rect(Qs, KVs, ldS, p.dh, acc);  // dPdrop[qi][kw] = dO[qi] . V[kw]
    const int kk = threadIdx.x & 63, qg = threadIdx.x >> 6;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kw = kk + 64 * j;
      if (kw >= NKW_PAD) continue;
#pragma unroll
      for (int i = 0; i < 8; ++i) Sc[(qg * 8 + i) * NKW_PAD + kw] = acc[j][i];
    }
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int qi = warp; qi < TQ; qi += 8) {
      const int qt = q0 + qi;
      float* row = Sc + qi * NKW_PAD;
      if (qt >= p.T) {
        for (int kw = lane; kw < NKW_PAD; kw += 32) row[kw] = 0.f;
        continue;
      }
      const float* Prow = p.P + ((tok_base + qt) * p.H + h) * p.RW;
      float* dSrow = p.dS + ((tok_base + qt) * p.H + h) * p.RW;
      constexpr int NG = 2;  // rel groups per lane: RW/4 <= 64
      float pr[NG][4], dp[NG][4];
      float delta = 0.f;
#pragma unroll
      for (int it = 0; it < NG; ++it) {
        const int g = lane + it * 32;
        uint32_t keep = 0xfu;
        if (g < (p.RW >> 2) && p.drop_p > 0.f) keep = keep_bits4(p, tok_base + qt, h, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rel = 4 * g + j, kw = rel + qi, key = kbase + kw;
          float pv = 0.f, dv = 0.f;
          if (g < (p.RW >> 2) && rel <= 2 * p.W && key >= 0 && key < p.T) {
            pv = __ldg(Prow + rel);
            dv = ((keep >> j) & 1u) ? row[kw] * p.drop_scale : 0.f;
          }
          pr[it][j] = pv;
          dp[it][j] = dv;
          delta = fmaf(pv, dv, delta);
        }
      }
      delta = ssb::warp_sum(delta);
      __syncwarp();
      for (int kw = lane; kw < NKW_PAD; kw += 32) {
        const int rel = kw - qi;
        if (rel < 0 || rel > 2 * p.W) row[kw] = 0.f;
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < NG; ++it) {
        const int g = lane + it * 32;
        if (g >= (p.RW >> 2)) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rel = 4 * g + j;
          const float ds = pr[it][j] * (dp[it][j] - delta);
          dSrow[rel] = rel <= 2 * p.W ? ds : 0.f;
          if (rel <= 2 * p.W) row[rel + qi] = ds;
        }
      }
    }
  }
  __syncthreads();
  load_rows(KVs, ldS, p.qkv + p.D + h * p.dh, ld, tok_base, kbase, TQ + 2 * p.W, p.T, p.dh, 1.f);
  __syncthreads();
  {
    float4 acc[4];
    int d4, qg;
    bool active;
    pv_rect(Sc, KVs, ldS, p.dh, 0, TQ + 2 * p.W, acc, d4, qg, active);  // dQ = scale * dS K
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qt = q0 + qg * 4 + i;
        if (qt < p.T) {
          float4 o = acc[i];
          o.x *= p.scale; o.y *= p.scale; o.z *= p.scale; o.w *= p.scale;
          *reinterpret_cast<float4*>(p.dqkv + (tok_base + qt) * ld + h * p.dh + d4) = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward, key side: dK = scale * dS^T Q,  dV = Pdrop^T dO.  CTA = 32 keys; window = queries.
// Window row qw <-> query qbase + qw (qbase = k0 - W); rel = key - query + W = ki - qw + 2W.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) band_attn_bwd_kv_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ldS = smem_ld(p.dh);
  float* Ws = smem;                     // [NKW][ldS]  Q window, then dO window
  float* Sc = Ws + NKW * ldS;           // [TQ][NKW_PAD] dS^T, then Pdrop^T

  const int k0 = blockIdx.x * TQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int64_t tok_base = (int64_t)b * p.T;
  const int64_t ld = 3 * (int64_t)p.D;
  const int qbase = k0 - p.W;
  const int nwin = TQ + 2 * p.W;

  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: Sc = dS^T, window = Q (scaled)  -> dK ;  pass 1: Sc = Pdrop^T, window = dO -> dV
    if (pass == 0)
      load_rows(Ws, ldS, p.qkv + h * p.dh, ld, tok_base, qbase, nwin, p.T, p.dh, p.scale);
    else
      load_rows(Ws, ldS, p.dO + h * p.dh, p.D, tok_base, qbase, nwin, p.T, p.dh, 1.f);
    const float* src = pass == 0 ? p.dS : p.P;
    for (int i = threadIdx.x; i < TQ * NKW_PAD; i += THREADS) Sc[i] = 0.f;
    __syncthreads();
    // window row qw needs rels [2W - qw, 2W - qw + TQ) clipped to [0, 2W]: <= 9 aligned groups
    for (int i = threadIdx.x; i < nwin * 9; i += THREADS) {
      const int qw = i / 9, gi = i - qw * 9;
      const int qt = qbase + qw;
      if (qt < 0 || qt >= p.T) continue;
      const int rel_lo = max(2 * p.W - qw, 0);
      const int g = (rel_lo >> 2) + gi;
      if (4 * g > min(2 * p.W - qw + TQ - 1, 2 * p.W)) continue;
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(
          src + ((tok_base + qt) * p.H + h) * p.RW + 4 * g));
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      uint32_t keep = 0xfu;
      if (pass == 1 && p.drop_p > 0.f) keep = keep_bits4(p, tok_base + qt, h, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rel = 4 * g + j;
        const int ki = rel + qw - 2 * p.W;
        if (rel > 2 * p.W || ki < 0 || ki >= TQ || k0 + ki >= p.T) continue;
        float v = vv[j];
        if (pass == 1) v = ((keep >> j) & 1u) ? v * p.drop_scale : 0.f;
        Sc[ki * NKW_PAD + qw] = v;
      }
    }
    __syncthreads();
    float4 acc[4];
    int d4, kg;
    bool active;
    pv_rect(Sc, Ws, ldS, p.dh, 0, nwin, acc, d4, kg, active);
    if (active) {
      const int64_t col = (pass == 0 ? p.D : 2 * (int64_t)p.D) + h * p.dh + d4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kt = k0 + kg * 4 + i;
        if (kt < p.T) *reinterpret_cast<float4*>(p.dqkv + (tok_base + kt) * ld + col) = acc[i];
      }
    }
    __syncthreads();
  }
}

int make_params(const float* qkv, const float* R, float* P, float* O, const float* dO, float* dS,
                float* dqkv, int64_t B, int64_t T, int64_t H, int64_t dh, int64_t W, int64_t RW,
                float drop_p, uint64_t seed, uint32_t site, AttnParams* p) {
  SSB_REQUIRE(B >= 1 && T >= 1 && H >= 1 && B <= 65535 && H <= 65535, "band_attn: bad B/T/H");
  SSB_REQUIRE(dh >= 4 && dh <= 128 && dh % 4 == 0, "band_attn: head dim %lld not in [4,128] step 4",
              (long long)dh);
  SSB_REQUIRE(W >= 0 && W <= MAXW, "band_attn: band half-width %lld > %d not built", (long long)W,
              MAXW);
  SSB_REQUIRE(RW >= 2 * W + 1 && RW % 4 == 0, "band_attn: RW=%lld must be >= 2W+1 and %% 4 == 0",
              (long long)RW);
  SSB_REQUIRE(B * T < (1LL << 31), "band_attn: too many tokens");
  SSB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "band_attn: bad dropout p");
  SSB_REQUIRE(((uintptr_t)qkv & 15) == 0, "band_attn: qkv must be 16 B aligned");
  p->qkv = qkv; p->R = R; p->P = P; p->O = O; p->dO = dO; p->dS = dS; p->dqkv = dqkv;
  p->B = (int)B; p->T = (int)T; p->H = (int)H; p->dh = (int)dh; p->D = (int)(H * dh);
  p->W = (int)W; p->RW = (int)RW;
  p->scale = 1.0f / sqrtf((float)dh);
  p->drop_p = drop_p;
  p->drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const double th = (double)drop_p * 4294967296.0;
  p->drop_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
  p->seed = seed; p->seed_src = ssb::seed_source(); p->site = site;
  return SSB_OK;
}

}  // namespace

extern "C" {

int ssb_band_attn_fwd(const float* qkv, const float* R, int64_t B, int64_t T, int64_t H,
                      int64_t dh, int64_t W, int64_t RW, float drop_p, uint64_t seed,
                      uint32_t site, float* P, float* O, void* stream) {
  AttnParams p;
  if (int rc = make_params(qkv, R, P, O, nullptr, nullptr, nullptr, B, T, H, dh, W, RW, drop_p,
                           seed, site, &p))
    return rc;
  SSB_REQUIRE(qkv && R && O, "band_attn_fwd: null pointer");
  const int ldS = (int)dh + 4;
  const size_t smem = (size_t)(TQ * ldS + NKW * ldS + TQ * NKW_PAD) * sizeof(float);
  SSB_CUDA(cudaFuncSetAttribute(band_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid((unsigned)((T + TQ - 1) / TQ), (unsigned)H, (unsigned)B);
  band_attn_fwd_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(p);
  SSB_LAUNCH_CHECK("band_attn_fwd");
  return SSB_OK;
}

int ssb_band_attn_bwd(const float* qkv, const float* P, const float* dO, int64_t B, int64_t T,
                      int64_t H, int64_t dh, int64_t W, int64_t RW, float drop_p, uint64_t seed,
                      uint32_t site, float* dS, float* dqkv, void* stream) {
  AttnParams p;
  if (int rc = make_params(qkv, nullptr, const_cast<float*>(P), nullptr, dO, dS, dqkv, B, T, H, dh,
                           W, RW, drop_p, seed, site, &p))
    return rc;
  SSB_REQUIRE(qkv && P && dO && dS && dqkv, "band_attn_bwd: null pointer");
  const int ldS = (int)dh + 4;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((T + TQ - 1) / TQ), (unsigned)H, (unsigned)B);
  const size_t smem_q = (size_t)(TQ * ldS + NKW * ldS + TQ * NKW_PAD) * sizeof(float);
  SSB_CUDA(cudaFuncSetAttribute(band_attn_bwd_q_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
  band_attn_bwd_q_kernel<<<grid, THREADS, smem_q, st>>>(p);
  SSB_LAUNCH_CHECK("band_attn_bwd_q");
  const size_t smem_kv = (size_t)(NKW * ldS + TQ * NKW_PAD) * sizeof(float);
  SSB_CUDA(cudaFuncSetAttribute(band_attn_bwd_kv_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
  band_attn_bwd_kv_kernel<<<grid, THREADS, smem_kv, st>>>(p);
  SSB_LAUNCH_CHECK("band_attn_bwd_kv");
  return SSB_OK;
}

}  // extern "C"
