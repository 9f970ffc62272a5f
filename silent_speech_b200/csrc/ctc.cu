// Fused log_softmax + CTC loss, forward and gradient w.r.t. the LOGITS (sm_100a).   SURVEY.md §8 f2
//
// Replaces recognition_model.py:96-101:  F.log_softmax(pred, 2) -> pad_sequence ->
// F.ctc_loss(pred, y, lengths, text_int_lengths, blank=n_chars)  and its autograd backward.
// Three launches (S = 2L + 1 <= 1024 states of the blank-extended target l'):
//   ctc_lp_kernel     parallel over frames: lse[t] = logsumexp_c logits[t, c] and the table
//                     lp[t][s] = logits[t, l'_s] - lse[t]
//   ctc_recur_kernel  the only sequential part (T = 750 frames at cfg-5): one 4-warp CTA per
//                     (utterance, direction) - alpha forwards and beta backwards run concurrently
//                     on different SMs - with the states in registers, neighbours by shuffle (one
//                     tiny mailbox + ONE barrier per frame across warps), the lp rows prefetched
//                     8 frames ahead (no dependent global load on the chain) and MUFU exp / log.  (Round 1 walked alpha then beta in
//                     one CTA with two block barriers and an L2 round trip per frame: 1.84 ms at
//                     cfg-5 against 0.64 ms for ATen's log_softmax + ctc_loss forward.)
//                     nll = -logsumexp(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
//   ctc_grad_kernel   parallel over frames: occ_t(c) = sum_{s: l'_s = c} exp(alpha_t(s) +
//                     beta_t(s) - lp_t(s) + nll) (<= 1) through shared-memory atomics and
//                     d nll / d logits[t, c] = softmax_t(c) - occ_t(c)   (0 for t >= input length)
// fp32 throughout (torch's ctc_loss also runs its log-space recursion in the input dtype).
#include "ssb_common.cuh"
#include <math_constants.h>

namespace {

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m));
}
struct CtcParams {
  const float* logits;            // (N, T, C)
  const int64_t* targets;         // (N, Lmax)
  const int64_t* in_len;          // (N)
  const int64_t* tgt_len;         // (N)
  int T, C, Lmax, blank, Sp;      // Sp = 128 * KS: row pitch of the per-frame state tables
  int mean_reduction;             // gradient scaled by 1 / (N * max(L, 1))
  int N;
  // workspace tables
  float* lp;                      // (N, T, Sp)  log-softmax of frame t at the label of state s
  float* alpha;                   // (N, T, Sp)  alpha_t(s) relative to zoff[t]
  float* beta;                    // (N, T, Sp)  beta_t(s) relative to yoff[t]
  float* lse;                     // (N, T)
  double* zoff;                   // (N, T)
  double* yoff;                   // (N, T)
  double* nlld;                   // (N)
  float* nll;                     // (N)
  float* grad;                    // (N, T, C) or null
};

constexpr int RENORM = 8;   // steps between renormalisations of the log-space recursions

__device__ __forceinline__ int ext_label(const int64_t* tg, int s, int blank) {
  return (s & 1) ? (int)__ldg(tg + (s >> 1)) : blank;
}

// ---- phase 1 (parallel over frames): lse[t] and lp[t][s] -------------------------------------
__global__ void __launch_bounds__(256) ctc_lp_kernel(const CtcParams p) {
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int L = (int)min((int64_t)p.Lmax, max((int64_t)0, p.tgt_len[n]));
  const int S = 2 * L + 1;
  const int Tn = (int)min((int64_t)p.T, max((int64_t)0, p.in_len[n]));
  if (t >= Tn) return;
  const float* x = p.logits + ((int64_t)n * p.T + t) * p.C;
  float m = -CUDART_INF_F;
  for (int c = lane; c < p.C; c += 32) m = fmaxf(m, __ldg(x + c));
  m = ssb::warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < p.C; c += 32) sum += expf(__ldg(x + c) - m);
  sum = ssb::warp_sum(sum);
  const float lse = m + logf(sum);
  if (lane == 0) p.lse[(int64_t)n * p.T + t] = lse;
  const int64_t* tg = p.targets + (int64_t)n * p.Lmax;
  float* row = p.lp + ((int64_t)n * p.T + t) * p.Sp;
  for (int s = lane; s < p.Sp; s += 32)
    row[s] = s < S ? __ldg(x + ext_label(tg, s, p.blank)) - lse : 0.f;
}

// ---- phase 2 (sequential over frames): one 4-warp CTA per (utterance, direction) ---------------
// Thread j keeps states j*KS .. j*KS + KS-1 of the blank-extended target in registers (KS = 1, 2,
// 4 or 8 covers S <= 1024 with 128 threads).  The two neighbours a step needs from the adjacent
// thread come by shuffle inside a warp and through a double-buffered 4-entry shared-memory
// mailbox across warps: ONE block barrier per frame.  The lp rows are prefetched PFD frames ahead
// into registers, so no global load sits on the dependence chain, and exp / log use the MUFU
// intrinsics (arguments are differences to the running maximum: |error| ~ 1e-7 per step on values
// kept O(10)).  Log-space values drift by ~log(1/C) per frame (|alpha| ~ 2500 after 750 frames,
// one fp32 ulp = 2.4e-4 there), so both recursions are kept RELATIVE to a running offset,
// accumulated in double and bumped to the current maximum every RENORM frames.
constexpr int RW = 4;            // warps per recursion CTA

__device__ __forceinline__ float lse3_fast(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

template <int KS>
__global__ void __launch_bounds__(RW * 32) ctc_recur_kernel(const CtcParams p) {
  constexpr int PFD = 8;
  __shared__ float edge[2][RW][2];    // [parity][warp]: the two states a neighbouring warp needs
  __shared__ float wmax[2][RW];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_beta = blockIdx.y != 0;
  const int L = (int)min((int64_t)p.Lmax, max((int64_t)0, p.tgt_len[n]));
  const int S = 2 * L + 1;
  const int Tn = (int)min((int64_t)p.T, max((int64_t)0, p.in_len[n]));
  const float NEG = -CUDART_INF_F;
  const int64_t* tg = p.targets + (int64_t)n * p.Lmax;
  const float* LP = p.lp + (int64_t)n * p.T * p.Sp + tid * KS;
  float* OUT = (is_beta ? p.beta : p.alpha) + (int64_t)n * p.T * p.Sp + tid * KS;
  double* OFF = (is_beta ? p.yoff : p.zoff) + (int64_t)n * p.T;
  if (Tn == 0) {
    if (!is_beta && tid == 0) { p.nll[n] = CUDART_INF_F; p.nlld[n] = (double)CUDART_INF_F; }
    return;
  }
  // which states may take the skip transition (from s-2 for alpha, to s+2 for beta)
  uint32_t skip = 0, valid = 0;
#pragma unroll
  for (int i = 0; i < KS; ++i) {
    const int s = tid * KS + i;
    if (s < S) {
      valid |= 1u << i;
      if (s & 1) {
        const int lab = (int)__ldg(tg + (s >> 1));
        if (!is_beta && s >= 3 && lab != (int)__ldg(tg + (s >> 1) - 1)) skip |= 1u << i;
        if (is_beta && s + 2 < S && lab != (int)__ldg(tg + (s >> 1) + 1)) skip |= 1u << i;
      }
    }
  }
  const int t0 = is_beta ? Tn - 1 : 0, dt = is_beta ? -1 : 1;
  float pf[PFD][KS];
#pragma unroll
  for (int d = 0; d < PFD; ++d) {
    const int t = t0 + dt * d;
    const bool ok = t >= 0 && t < Tn;
#pragma unroll
    for (int i = 0; i < KS; ++i) pf[d][i] = ok ? LP[(int64_t)t * p.Sp + i] : 0.f;
  }
  if (tid < 2 * RW * 2) (&edge[0][0][0])[tid] = NEG;
  float a[KS];
#pragma unroll
  for (int i = 0; i < KS; ++i) a[i] = NEG;
  double Z = 0.0;
  __syncthreads();
  // The frame loop is unrolled by the prefetch depth so that ring slot d is a fixed set of
  // registers: a slot is consumed, then immediately refilled with the row PFD frames ahead.
  for (int k0 = 0; k0 < Tn; k0 += PFD) {
#pragma unroll
    for (int d = 0; d < PFD; ++d) {
      const int k = k0 + d;
      if (k >= Tn) break;
      const int t = t0 + dt * k;
      float lp[KS];
#pragma unroll
      for (int i = 0; i < KS; ++i) lp[i] = pf[d][i];
      {
        const int tn = t + dt * PFD;
        if (tn >= 0 && tn < Tn) {
#pragma unroll
          for (int i = 0; i < KS; ++i) pf[d][i] = LP[(int64_t)tn * p.Sp + i];
        }
      }
      float nv[KS];
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < KS; ++i) {
          const int s = tid * KS + i;
          const bool on = is_beta ? (s < S && s >= S - 2) : (s < 2 && s < S);
          nv[i] = on ? lp[i] : NEG;
        }
      } else {
        // neighbours held by the adjacent thread: n1 = one state away, n2 = two states away.
        // Mailbox of the previous frame: edge[(k-1)&1][w] = {nearest, second nearest} state of
        // warp w as seen from the warp that follows it in the direction of the dependence.
        const float* mb = &edge[(k - 1) & 1][0][0];
        float n1, n2;
        if (!is_beta) {      // states tid*KS - 1 and tid*KS - 2
          n1 = __shfl_up_sync(0xffffffffu, a[KS - 1], 1);
          n2 = KS >= 2 ? __shfl_up_sync(0xffffffffu, a[KS >= 2 ? KS - 2 : 0], 1)
                       : __shfl_up_sync(0xffffffffu, a[0], 2);
          if (lane == 0) {
            n1 = warp > 0 ? mb[(warp - 1) * 2] : NEG;
            n2 = warp > 0 ? mb[(warp - 1) * 2 + 1] : NEG;
          }
          if (KS == 1 && lane == 1) n2 = warp > 0 ? mb[(warp - 1) * 2] : NEG;
        } else {             // states (tid+1)*KS and (tid+1)*KS + 1
          n1 = __shfl_down_sync(0xffffffffu, a[0], 1);
          n2 = KS >= 2 ? __shfl_down_sync(0xffffffffu, a[KS >= 2 ? 1 : 0], 1)
                       : __shfl_down_sync(0xffffffffu, a[0], 2);
          if (lane == 31) {
            n1 = warp + 1 < RW ? mb[(warp + 1) * 2] : NEG;
            n2 = warp + 1 < RW ? mb[(warp + 1) * 2 + 1] : NEG;
          }
          if (KS == 1 && lane == 30) n2 = warp + 1 < RW ? mb[(warp + 1) * 2] : NEG;
        }
#pragma unroll
        for (int i = 0; i < KS; ++i) {
          float x1, x2;
          if (!is_beta) {
            x1 = i >= 1 ? a[i >= 1 ? i - 1 : 0] : n1;
            x2 = i >= 2 ? a[i >= 2 ? i - 2 : 0] : (i == 1 ? n1 : n2);
          } else {
            x1 = i + 1 < KS ? a[i + 1 < KS ? i + 1 : 0] : n1;
            x2 = i + 2 < KS ? a[i + 2 < KS ? i + 2 : 0] : (i + 1 < KS ? n1 : n2);
          }
          const float v = lse3_fast(a[i], x1, ((skip >> i) & 1u) ? x2 : NEG) + lp[i];
          nv[i] = ((valid >> i) & 1u) ? v : NEG;
        }
      }
      if (t % RENORM == 0 && (is_beta || k > 0)) {   // uniform: block-wide maximum
        float m = NEG;
#pragma unroll
        for (int i = 0; i < KS; ++i) m = fmaxf(m, nv[i]);
        m = ssb::warp_max(m);
        if (lane == 0) wmax[k & 1][warp] = m;
        __syncthreads();
        m = fmaxf(fmaxf(wmax[k & 1][0], wmax[k & 1][1]), fmaxf(wmax[k & 1][2], wmax[k & 1][3]));
        if (m > NEG) {
#pragma unroll
          for (int i = 0; i < KS; ++i) nv[i] -= m;
          Z += (double)m;
        }
      }
#pragma unroll
      for (int i = 0; i < KS; ++i) {
        a[i] = nv[i];
        OUT[(int64_t)t * p.Sp + i] = nv[i];
      }
      // publish what the neighbouring warp needs next frame
      if (!is_beta && lane == 31) {          // the warp's LAST two states
        edge[k & 1][warp][0] = a[KS - 1];
        if (KS >= 2) edge[k & 1][warp][1] = a[KS >= 2 ? KS - 2 : 0];
      }
      if (!is_beta && KS == 1 && lane == 30) edge[k & 1][warp][1] = a[0];
      if (is_beta && lane == 0) {            // the warp's FIRST two states
        edge[k & 1][warp][0] = a[0];
        if (KS >= 2) edge[k & 1][warp][1] = a[KS >= 2 ? 1 : 0];
      }
      if (is_beta && KS == 1 && lane == 1) edge[k & 1][warp][1] = a[0];
      if (tid == 0) OFF[t] = Z;
      __syncthreads();
    }
  }
  if (!is_beta) {   // nll = -logsumexp(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
    __shared__ float tail2[2];
    if (tid < 2) tail2[tid] = NEG;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < KS; ++i) {
      const int s = tid * KS + i;
      if (s == S - 1) tail2[0] = a[i];
      if (s == S - 2) tail2[1] = a[i];
    }
    __syncthreads();
    if (tid == 0) {
      const float tail = lse2(tail2[0], tail2[1]);
      const double nll = tail > NEG ? -(Z + (double)tail) : (double)CUDART_INF_F;
      p.nll[n] = (float)nll;
      p.nlld[n] = nll;
    }
  }
}

// ---- phase 3 (parallel over frames): class occupancies and the gradient -----------------------
//   occ_t(c) = sum_{s: l'_s = c} exp(alpha_t(s) + beta_t(s) - lp_t(s) + nll)   (<= 1)
//   d nll / d logits[t, c] = softmax_t(c) - occ_t(c)      (0 for t >= input length)
__global__ void __launch_bounds__(128) ctc_grad_kernel(const CtcParams p) {
  extern __shared__ float occ[];
  const int n = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
  const int L = (int)min((int64_t)p.Lmax, max((int64_t)0, p.tgt_len[n]));
  const int S = 2 * L + 1;
  const int Tn = (int)min((int64_t)p.T, max((int64_t)0, p.in_len[n]));
  float* G = p.grad + ((int64_t)n * p.T + t) * p.C;
  const double nll = p.nlld[n];
  const bool feasible = nll < (double)CUDART_INF_F;
  if (t >= Tn || !feasible) {
    for (int c = tid; c < p.C; c += blockDim.x) G[c] = 0.f;
    return;
  }
  for (int c = tid; c < p.C; c += blockDim.x) occ[c] = 0.f;
  __syncthreads();
  const int64_t row = (int64_t)n * p.T + t;
  const float off = (float)(p.zoff[row] + p.yoff[row] + nll);
  const int64_t* tg = p.targets + (int64_t)n * p.Lmax;
  for (int s = tid; s < S; s += blockDim.x) {
    const float v = p.alpha[row * p.Sp + s] + p.beta[row * p.Sp + s] - p.lp[row * p.Sp + s] + off;
    if (v > -80.f) atomicAdd(occ + ext_label(tg, s, p.blank), expf(fminf(v, 0.f)));
  }
  __syncthreads();
  const float scale = p.mean_reduction ? 1.f / ((float)p.N * (float)max(L, 1)) : 1.f;
  const float lse = p.lse[row];
  const float* x = p.logits + row * p.C;
  for (int c = tid; c < p.C; c += blockDim.x) G[c] = (expf(__ldg(x + c) - lse) - occ[c]) * scale;
}

int ks_for(int64_t Lmax) {   // states per thread of the 128-thread recursion CTA
  const int64_t need = (2 * Lmax + 1 + 127) / 128;
  int ks = 1;
  while (ks < need) ks <<= 1;
  return ks;   // 1, 2, 4, 8
}

int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace

extern "C" {

int64_t ssb_ctc_workspace_bytes(int64_t N, int64_t T, int64_t Lmax) {
  if (N < 0 || T < 0 || Lmax < 0 || 2 * Lmax + 1 > 1024) return SSB_ERR_ARG;
  const int64_t Sp = 128 * ks_for(Lmax);
  return 3 * align256(N * T * Sp * 4) + align256(N * T * 4) + 2 * align256(N * T * 8) +
         align256(N * 8) + 256;
}

int ssb_ctc_loss_fused(const float* logits, int64_t N, int64_t T, int64_t C, const int64_t* targets,
                       int64_t Lmax, const int64_t* input_lengths, const int64_t* target_lengths,
                       int64_t blank, int mean_reduction, float* nll, float* grad_logits,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  if (N == 0) return SSB_OK;
  SSB_REQUIRE(logits && input_lengths && target_lengths && nll && workspace,
              "ctc_loss_fused: null argument");
  SSB_REQUIRE(targets || Lmax == 0, "ctc_loss_fused: null targets");
  SSB_REQUIRE(N > 0 && N <= 65535 && T > 0 && C > 0 && C <= 1024 && blank >= 0 && blank < C,
              "ctc_loss_fused: N=%lld T=%lld C=%lld blank=%lld", (long long)N, (long long)T,
              (long long)C, (long long)blank);
  SSB_REQUIRE(2 * Lmax + 1 <= 1024, "ctc_loss_fused: targets longer than 511 symbols (Lmax=%lld)",
              (long long)Lmax);
  SSB_REQUIRE(workspace_bytes >= ssb_ctc_workspace_bytes(N, T, Lmax),
              "ctc_loss_fused: workspace too small");
  SSB_REQUIRE(((uintptr_t)workspace & 255) == 0, "ctc_loss_fused: workspace must be 256 B aligned");
  static const int64_t dummy_target = 0;   // Lmax == 0: never dereferenced (no odd state exists)
  (void)dummy_target;
  const int ks = ks_for(Lmax);
  CtcParams p;
  p.logits = logits; p.targets = targets; p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.T = (int)T; p.C = (int)C; p.Lmax = (int)Lmax; p.blank = (int)blank;
  p.Sp = 128 * ks;
  p.mean_reduction = mean_reduction; p.N = (int)N;
  char* w = (char*)workspace;
  const int64_t tab = align256(N * T * p.Sp * 4);
  p.lp = (float*)w; w += tab;
  p.alpha = (float*)w; w += tab;
  p.beta = (float*)w; w += tab;
  p.lse = (float*)w; w += align256(N * T * 4);
  p.zoff = (double*)w; w += align256(N * T * 8);
  p.yoff = (double*)w; w += align256(N * T * 8);
  p.nlld = (double*)w;
  p.nll = nll; p.grad = grad_logits;
  cudaStream_t st = (cudaStream_t)stream;
  ctc_lp_kernel<<<dim3((unsigned)((T + 7) / 8), (unsigned)N), 256, 0, st>>>(p);
  SSB_LAUNCH_CHECK("ctc_lp_kernel");
  const dim3 rgrid((unsigned)N, grad_logits ? 2u : 1u);   // alpha only when no gradient is wanted
  switch (ks) {
    case 1: ctc_recur_kernel<1><<<rgrid, RW * 32, 0, st>>>(p); break;
    case 2: ctc_recur_kernel<2><<<rgrid, RW * 32, 0, st>>>(p); break;
    case 4: ctc_recur_kernel<4><<<rgrid, RW * 32, 0, st>>>(p); break;
    default: ctc_recur_kernel<8><<<rgrid, RW * 32, 0, st>>>(p); break;
  }
  SSB_LAUNCH_CHECK("ctc_recur_kernel");
  if (grad_logits) {
    ctc_grad_kernel<<<dim3((unsigned)T, (unsigned)N), 128, (size_t)C * 4, st>>>(p);
    SSB_LAUNCH_CHECK("ctc_grad_kernel");
  }
  return SSB_OK;
}

}  // extern "C"
