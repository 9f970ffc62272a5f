// Fused log_softmax + CTC loss, forward and gradient w.r.t. the LOGITS (sm_100a).   SURVEY.md §8 f2
//
// Replaces recognition_model.py:96-101:  F.log_softmax(pred, 2) -> pad_sequence ->
// F.ctc_loss(pred, y, lengths, text_int_lengths, blank=n_chars)  and its autograd backward.
// One CTA per utterance, one thread per position s of the blank-extended target l'
// (S = 2L + 1 <= 1024).  The recursion over time is sequential (T = 750 at cfg-5) and latency
// bound; everything that is parallel (row log-sum-exp, the S states, the C classes) is spread
// over the CTA:
//   phase 1  lse[t] = logsumexp_c logits[t, c]                      (one warp per row)
//   phase 2  alpha_t(s) in log space, stored to the workspace        (1 barrier per step)
//   phase 3  nll = -logsumexp(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
//   phase 4  beta_t(s) backwards in time; per step the class sums
//              occ_t(c) = sum_{s: l'_s = c} exp(alpha_t(s) + beta_t(s) - lp_t(c) + nll)   (<= 1)
//            go through shared-memory atomics and
//              d nll / d logits[t, c] = softmax_t(c) - occ_t(c)      (0 for t >= input length)
// fp32 throughout (torch's ctc_loss also runs its log-space recursion in the input dtype).
#include "ssb_common.cuh"
#include <math_constants.h>

namespace {

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

struct CtcParams {
  const float* logits;            // (N, T, C)
  const int64_t* targets;         // (N, Lmax)
  const int64_t* in_len;          // (N)
  const int64_t* tgt_len;         // (N)
  int T, C, Lmax, blank, Sp;      // Sp = padded row length of the alpha workspace
  int mean_reduction;             // gradient scaled by 1 / (N * max(L, 1))
  int N;
  float* alpha;                   // workspace (N, T, Sp)
  float* nll;                     // (N)
  float* grad;                    // (N, T, C) or null
};

// block-wide max of v over the threads with `active`; every thread gets the result.
// wm: [32] floats of shared memory.  Two barriers.
__device__ __forceinline__ float block_max(float v, float* wm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = ssb::warp_max(v);
  if (lane == 0) wm[warp] = v;
  __syncthreads();
  float m = lane < nw ? wm[lane] : -CUDART_INF_F;
  m = ssb::warp_max(m);
  __syncthreads();
  return m;
}

constexpr int RENORM = 8;   // steps between renormalisations of the log-space recursions

__global__ void ctc_fused_kernel(const CtcParams p) {
  extern __shared__ double smem_d[];
  const int n = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int L = (int)min((int64_t)p.Lmax, max((int64_t)0, p.tgt_len[n]));
  const int S = 2 * L + 1;
  const int Tn = (int)min((int64_t)p.T, max((int64_t)0, p.in_len[n]));
  const int Cp = (p.C + 31) / 32 * 32;
  // Log-space values drift by ~log(1/C) per frame: after 750 frames |alpha| ~ 2500, where one
  // fp32 ulp is 2.4e-4 and the occupancies exp(alpha + beta + nll) lose 3 digits.  Both
  // recursions are therefore kept RELATIVE to a running offset (zoff[t], accumulated in double
  // and bumped to the current maximum every RENORM frames), so the stored values stay O(10).
  double* zoff = smem_d;                          // [T] offset of the stored alpha_t
  float* lse = reinterpret_cast<float*>(zoff + p.T);   // [T]
  float* buf = lse + p.T;                         // [2][Sp + 4]
  float* occ = buf + 2 * (p.Sp + 4);              // [2][Cp]
  float* wm = occ + 2 * Cp;                       // [32]
  const int bstride = p.Sp + 4;
  const float* X = p.logits + (int64_t)n * p.T * p.C;
  float* A = p.alpha + (int64_t)n * p.T * p.Sp;
  const float NEG = -CUDART_INF_F;

  // labels of the extended sequence
  int label = p.blank;
  bool skip_in = false, skip_out = false;         // may come from s-2 / may go to s+2
  if (tid < S && (tid & 1)) {
    const int64_t* tg = p.targets + (int64_t)n * p.Lmax;
    label = (int)tg[tid >> 1];
    skip_in = tid >= 3 && label != (int)tg[(tid >> 1) - 1];
    skip_out = tid + 2 < S && label != (int)tg[(tid >> 1) + 1];
  }
  // phase 1: row log-sum-exp
  for (int t = warp; t < Tn; t += nwarps) {
    float m = NEG;
    for (int c = lane; c < p.C; c += 32) m = fmaxf(m, X[(int64_t)t * p.C + c]);
    m = ssb::warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < p.C; c += 32) sum += expf(X[(int64_t)t * p.C + c] - m);
    sum = ssb::warp_sum(sum);
    if (lane == 0) lse[t] = m + logf(sum);
  }
  for (int i = tid; i < 2 * bstride; i += nthr) buf[i] = NEG;
  for (int i = tid; i < 2 * Cp; i += nthr) occ[i] = 0.f;
  __syncthreads();

  double nll = (double)CUDART_INF_F;
  if (Tn > 0) {
    // phase 2: alpha (states s at index s + 2: two -inf pads in front)
    double Z = 0.0;
    if (tid < S) {
      const float a0 = tid < 2 ? X[label] - lse[0] : NEG;
      buf[tid + 2] = a0;
      A[tid] = a0;
    }
    if (tid == 0) zoff[0] = 0.0;
    __syncthreads();
    for (int t = 1; t < Tn; ++t) {
      const float* prev = buf + ((t - 1) & 1) * bstride;
      float* cur = buf + (t & 1) * bstride;
      float a = NEG;
      if (tid < S)
        a = lse3(prev[tid + 2], prev[tid + 1], skip_in ? prev[tid] : NEG) +
            (X[(int64_t)t * p.C + label] - lse[t]);
      if (t % RENORM == 0) {
        const float m = block_max(a, wm);
        if (m > NEG) { a -= m; Z += (double)m; }
      }
      if (tid < S) {
        cur[tid + 2] = a;
        A[(int64_t)t * p.Sp + tid] = a;
      }
      if (tid == 0) zoff[t] = Z;
      __syncthreads();
    }
    // phase 3
    const float* last = buf + ((Tn - 1) & 1) * bstride;
    const float tail = lse2(last[S - 1 + 2], S >= 2 ? last[S - 2 + 2] : NEG);
    nll = tail > NEG ? -(Z + (double)tail) : (double)CUDART_INF_F;
  }
  __syncthreads();
  if (tid == 0) p.nll[n] = (float)nll;
  if (!p.grad) return;
  float* G = p.grad + (int64_t)n * p.T * p.C;
  for (int i = Tn * p.C + tid; i < p.T * p.C; i += nthr) G[i] = 0.f;   // padding frames
  if (Tn == 0) return;
  const bool feasible = nll < (double)CUDART_INF_F;
  const float scale = p.mean_reduction ? 1.f / ((float)p.N * (float)max(L, 1)) : 1.f;
  // phase 4: beta (states at index s: two -inf pads BEHIND), class occupancies, gradient
  for (int i = tid; i < 2 * bstride; i += nthr) buf[i] = NEG;
  __syncthreads();
  double Y = 0.0;
  for (int t = Tn - 1; t >= 0; --t) {
    const float* nxt = buf + ((t + 1) & 1) * bstride;
    float* cur = buf + (t & 1) * bstride;
    float* oc = occ + (t & 1) * Cp;
    float b = NEG, lp = 0.f;
    if (tid < S) {
      lp = X[(int64_t)t * p.C + label] - lse[t];
      if (t == Tn - 1) b = tid >= S - 2 ? lp : NEG;
      else b = lse3(nxt[tid], nxt[tid + 1], skip_out ? nxt[tid + 2] : NEG) + lp;
    }
    if (t % RENORM == 0) {
      const float m = block_max(b, wm);
      if (m > NEG) { b -= m; Y += (double)m; }
    }
    if (tid < S) {
      cur[tid] = b;
      if (feasible) {
        // alpha_t(s) beta_t(s) / (y_t(l'_s) p(l|x)) with the offsets put back in double
        const float off = (float)(zoff[t] + Y + nll);
        const float v = A[(int64_t)t * p.Sp + tid] + b - lp + off;
        if (v > -80.f) atomicAdd(oc + label, expf(fminf(v, 0.f)));
      }
    }
    __syncthreads();
    if (tid < p.C) {
      const float sm = expf(X[(int64_t)t * p.C + tid] - lse[t]);
      G[(int64_t)t * p.C + tid] = feasible ? (sm - oc[tid]) * scale : 0.f;
      oc[tid] = 0.f;   // reused at step t - 2, after the barrier of step t - 1
    }
  }
}

}  // namespace

extern "C" {

int64_t ssb_ctc_workspace_bytes(int64_t N, int64_t T, int64_t Lmax) {
  if (N < 0 || T < 0 || Lmax < 0) return SSB_ERR_ARG;
  const int64_t Sp = (2 * Lmax + 1 + 31) / 32 * 32;
  return N * T * Sp * 4;
}

int ssb_ctc_loss_fused(const float* logits, int64_t N, int64_t T, int64_t C, const int64_t* targets,
                       int64_t Lmax, const int64_t* input_lengths, const int64_t* target_lengths,
                       int64_t blank, int mean_reduction, float* nll, float* grad_logits,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  if (N == 0) return SSB_OK;
  SSB_REQUIRE(logits && input_lengths && target_lengths && nll && workspace,
              "ctc_loss_fused: null argument");
  SSB_REQUIRE(targets || Lmax == 0, "ctc_loss_fused: null targets");
  SSB_REQUIRE(N > 0 && T > 0 && C > 0 && C <= 1024 && blank >= 0 && blank < C,
              "ctc_loss_fused: N=%lld T=%lld C=%lld blank=%lld", (long long)N, (long long)T,
              (long long)C, (long long)blank);
  SSB_REQUIRE(2 * Lmax + 1 <= 1024, "ctc_loss_fused: targets longer than 511 symbols (Lmax=%lld)",
              (long long)Lmax);
  SSB_REQUIRE(workspace_bytes >= ssb_ctc_workspace_bytes(N, T, Lmax),
              "ctc_loss_fused: workspace too small");
  CtcParams p;
  p.logits = logits; p.targets = targets; p.in_len = input_lengths; p.tgt_len = target_lengths;
  p.T = (int)T; p.C = (int)C; p.Lmax = (int)Lmax; p.blank = (int)blank;
  p.Sp = (int)((2 * Lmax + 1 + 31) / 32 * 32);
  p.mean_reduction = mean_reduction; p.N = (int)N;
  p.alpha = (float*)workspace; p.nll = nll; p.grad = grad_logits;
  int threads = p.Sp;
  const int cpad = (int)((C + 31) / 32 * 32);
  if (threads < cpad) threads = cpad;
  if (threads < 128) threads = 128;
  const size_t smem = (size_t)T * 8 + ((size_t)T + 2 * (p.Sp + 4) + 2 * cpad + 32) * 4;
  SSB_REQUIRE(smem <= 200 * 1024, "ctc_loss_fused: T=%lld too long for shared memory", (long long)T);
  if (smem > 48 * 1024)
    SSB_CUDA(cudaFuncSetAttribute(ctc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  ctc_fused_kernel<<<(unsigned)N, threads, smem, (cudaStream_t)stream>>>(p);
  SSB_LAUNCH_CHECK("ctc_fused");
  return SSB_OK;
}

}  // extern "C"
