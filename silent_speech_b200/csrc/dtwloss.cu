// On-device transduction loss around the DTW kernel: replaces the
// per-utterance Python loop of transduction_model.py:98-157 —
//   silent:  costs = cdist(pred, y) + w * (-log_softmax(pred_phone)[:, y_phone])      (:116-124)
//            alignment = align_from_distances(costs.T)                                 (:126)
//            loss = costs[alignment, range(Tg)].sum()                                  (:128)
//   voiced:  loss = pairwise_distance(y, pred).sum() + w * cross_entropy(sum)          (:141-145)
// — for ALL utterances of a batch, ragged shapes included, with three launches:
//   dtw_cost_kernel       cost matrices of every silent utterance (one tile grid)
//   (ssb_dtw_align_ragged, dtw.cu: fill + backtrace of all of them)
//   dtw_loss_rows_kernel  per predicted frame: its loss terms AND its gradient rows.  The loss is
//                         a gather along a monotone path, so its gradient w.r.t. a predicted frame
//                         only involves the contiguous run of target frames aligned to it: the
//                         backward is sparse by construction (no dense T_pred x T_tgt gradient).
// Distances are evaluated directly, sqrt(sum (a-b)^2), not through the |a|^2+|b|^2-2ab expansion
// ATen's cdist uses (closer to the exact value; alignments can differ from the reference's only
// where two paths tie to within fp32 rounding).
#include "ssb_common.cuh"
#include <math_constants.h>

namespace {

constexpr int TILE = 64;          // predicted frames x target frames per CTA
constexpr int MAXF = 128;         // feature width limit (80 in the reference)
constexpr int MAXP = 64;          // phoneme classes limit (48 in the reference)

// ---- cost matrices ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dtw_cost_kernel(const float* __restrict__ pred, const float* __restrict__ phon,
                const float* __restrict__ tgt, const int64_t* __restrict__ tgt_phone,
                const ssb_utt_t* __restrict__ table, int F, int NP, float w,
                float* __restrict__ cost_base) {
  extern __shared__ float sm[];
  const ssb_utt_t u = table[blockIdx.y];
  if (!u.silent) return;
  const int tiles_t = (u.Tg + TILE - 1) / TILE, tiles_p = (u.Tp + TILE - 1) / TILE;
  if ((int)blockIdx.x >= tiles_t * tiles_p) return;
  const int p0 = ((int)blockIdx.x / tiles_t) * TILE, t0 = ((int)blockIdx.x % tiles_t) * TILE;
  const int FS = F + 1, PS = NP + 1;                   // odd-ish strides: conflict-free columns
  float* A = sm;                                       // [TILE][FS]   predicted frames
  float* Bt = A + TILE * FS;                           // [TILE][FS]   target frames
  float* PH = Bt + TILE * FS;                          // [TILE][PS]   phoneme logits
  float* lse = PH + TILE * PS;                         // [TILE]
  int* yph = reinterpret_cast<int*>(lse + TILE);       // [TILE]
  const int tid = threadIdx.x;
  for (int i = tid; i < TILE * F; i += 256) {
    const int r = i / F, c = i - r * F;
    A[r * FS + c] = (p0 + r < u.Tp) ? __ldg(pred + (u.pred_row + p0 + r) * F + c) : 0.f;
    Bt[r * FS + c] = (t0 + r < u.Tg) ? __ldg(tgt + (u.tgt_row + t0 + r) * F + c) : 0.f;
  }
  for (int i = tid; i < TILE * NP; i += 256) {
    const int r = i / NP, c = i - r * NP;
    PH[r * PS + c] = (p0 + r < u.Tp) ? __ldg(phon + (u.pred_row + p0 + r) * NP + c) : 0.f;
  }
  if (tid < TILE) {
    int y = (t0 + tid < u.Tg) ? (int)__ldg(tgt_phone + u.tgt_row + t0 + tid) : 0;
    yph[tid] = min(max(y, 0), NP - 1);
  }
  __syncthreads();
  if (tid < TILE) {                                    // log-sum-exp of each predicted frame
    float m = -CUDART_INF_F;
    for (int c = 0; c < NP; ++c) m = fmaxf(m, PH[tid * PS + c]);
    float s = 0.f;
    for (int c = 0; c < NP; ++c) s += expf(PH[tid * PS + c] - m);
    lse[tid] = m + logf(s);
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;              // rows 4*ty .. +3, cols tx + 16*c
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int k = 0; k < F; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = A[(4 * ty + r) * FS + k];
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = Bt[(tx + 16 * c) * FS + k];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float d = a[r] - b[c];
        acc[r][c] = fmaf(d, d, acc[r][c]);
      }
  }
  float* cost = cost_base + u.cost_off;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pl = 4 * ty + r, p = p0 + pl;
    if (p >= u.Tp) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int tl = tx + 16 * c, t = t0 + tl;
      if (t < u.Tg)
        cost[(int64_t)p * u.pitch + t] = sqrtf(acc[r][c]) + w * (lse[pl] - PH[pl * PS + yph[tl]]);
    }
  }
}

// ---- per-row loss and gradient ---------------------------------------------------------------
// One warp per row of the flattened (rows_total, F) prediction tensor.
__global__ void __launch_bounds__(256)
dtw_loss_rows_kernel(const float* __restrict__ pred, const float* __restrict__ phon,
                     const float* __restrict__ tgt, const int64_t* __restrict__ tgt_phone,
                     const ssb_utt_t* __restrict__ table, int n_utt,
                     const int32_t* __restrict__ path, int path_pitch, int64_t rows_total, int F,
                     int NP, float w, float pd_eps, float* __restrict__ row_loss,
                     float* __restrict__ grad_pred, float* __restrict__ grad_phon) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  // utterance owning this row: last u with pred_row <= row (pred_row ascending)
  int lo = 0, hi = n_utt;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (table[mid].pred_row <= row) lo = mid; else hi = mid;
  }
  constexpr int FV = MAXF / 32, PV = MAXP / 32;
  float gp[FV], gq[PV];
#pragma unroll
  for (int i = 0; i < FV; ++i) gp[i] = 0.f;
#pragma unroll
  for (int i = 0; i < PV; ++i) gq[i] = 0.f;
  float loss = 0.f;
  ssb_utt_t u;
  u.pred_row = 0; u.tgt_row = 0; u.cost_off = 0; u.Tp = 0; u.Tg = 0; u.pitch = 0; u.silent = 0;
  u.pair = -1; u.reserved_ = 0;
  if (n_utt > 0) u = table[lo];
  const int pl = (int)(row - u.pred_row);
  if (n_utt > 0 && row >= u.pred_row && pl < u.Tp) {
    float a[FV], q[PV];
#pragma unroll
    for (int i = 0; i < FV; ++i) a[i] = (lane + 32 * i < F) ? __ldg(pred + row * F + lane + 32 * i) : 0.f;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      q[i] = (lane + 32 * i < NP) ? __ldg(phon + row * NP + lane + 32 * i) : -CUDART_INF_F;
      m = fmaxf(m, q[i]);
    }
    m = ssb::warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PV; ++i) s += (lane + 32 * i < NP) ? expf(q[i] - m) : 0.f;
    s = ssb::warp_sum(s);
    const float lse = m + logf(s);
    int t_lo, t_hi;
    if (u.silent) {   // target frames aligned to this predicted frame: a run of the monotone path
      const int32_t* pa = path + (int64_t)u.pair * path_pitch;
      int l = 0, h = u.Tg;
      while (l < h) { const int mid = (l + h) >> 1; if (__ldg(pa + mid) < pl) l = mid + 1; else h = mid; }
      t_lo = l;
      h = u.Tg;
      while (l < h) { const int mid = (l + h) >> 1; if (__ldg(pa + mid) <= pl) l = mid + 1; else h = mid; }
      t_hi = l;
    } else {
      t_lo = pl;
      t_hi = pl + 1;
    }
    float cnt[PV];
#pragma unroll
    for (int i = 0; i < PV; ++i) cnt[i] = 0.f;
    for (int t = t_lo; t < t_hi; ++t) {
      const float* y = tgt + (u.tgt_row + t) * F;
      float d[FV], ss = 0.f;
#pragma unroll
      for (int i = 0; i < FV; ++i) {
        // silent: cdist(pred, y);  voiced: pairwise_distance(y, pred) = ||y - pred + eps||
        const float yv = (lane + 32 * i < F) ? __ldg(y + lane + 32 * i) : 0.f;
        d[i] = (lane + 32 * i < F) ? (u.silent ? a[i] - yv : (yv - a[i]) + pd_eps) : 0.f;
        ss = fmaf(d[i], d[i], ss);
      }
      ss = ssb::warp_sum(ss);
      const float dist = sqrtf(ss);
      const float inv = dist > 0.f ? 1.f / dist : 0.f;
      const float sgn = u.silent ? inv : -inv;          // d(dist)/d(pred)
#pragma unroll
      for (int i = 0; i < FV; ++i) gp[i] = fmaf(d[i], sgn, gp[i]);
      const int yp = min(max((int)__ldg(tgt_phone + u.tgt_row + t), 0), NP - 1);
      float qy = 0.f;
#pragma unroll
      for (int i = 0; i < PV; ++i) {
        const bool hit = (lane + 32 * i == yp);
        cnt[i] += hit ? 1.f : 0.f;
        qy += hit ? q[i] : 0.f;
      }
      qy = ssb::warp_sum(qy);
      loss += dist + w * (lse - qy);
    }
    const float n = (float)(t_hi - t_lo);
#pragma unroll
    for (int i = 0; i < PV; ++i)
      gq[i] = (lane + 32 * i < NP) ? w * (n * expf(q[i] - lse) - cnt[i]) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < FV; ++i)
    if (lane + 32 * i < F) grad_pred[row * F + lane + 32 * i] = gp[i];
#pragma unroll
  for (int i = 0; i < PV; ++i)
    if (lane + 32 * i < NP) grad_phon[row * NP + lane + 32 * i] = gq[i];
  if (lane == 0) row_loss[row] = loss;
}

}  // namespace

extern "C" {

int ssb_dtw_cost_batch(const float* pred, const float* phon, const float* tgt,
                       const int64_t* tgt_phone, const ssb_utt_t* table_dev, int64_t n_utt,
                       int64_t max_Tp, int64_t max_Tg, int64_t F, int64_t NP, float w,
                       float* cost_base, void* stream) {
  SSB_REQUIRE(n_utt >= 0 && n_utt <= 65535, "dtw cost: bad utterance count %lld", (long long)n_utt);
  if (n_utt == 0 || max_Tp <= 0 || max_Tg <= 0) return SSB_OK;
  SSB_REQUIRE(pred && phon && tgt && tgt_phone && table_dev && cost_base, "dtw cost: null pointer");
  SSB_REQUIRE(F >= 1 && F <= MAXF && NP >= 1 && NP <= MAXP, "dtw cost: F=%lld (<= %d), NP=%lld (<= %d)",
              (long long)F, MAXF, (long long)NP, MAXP);
  const int64_t tiles = ((max_Tp + TILE - 1) / TILE) * ((max_Tg + TILE - 1) / TILE);
  SSB_REQUIRE(tiles < (1LL << 31), "dtw cost: too many tiles");
  const int smem = (int)((2 * TILE * (F + 1) + TILE * (NP + 1) + 2 * TILE) * sizeof(float));
  SSB_CUDA(cudaFuncSetAttribute(dtw_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dtw_cost_kernel<<<dim3((unsigned)tiles, (unsigned)n_utt), 256, smem, (cudaStream_t)stream>>>(
      pred, phon, tgt, tgt_phone, table_dev, (int)F, (int)NP, w, cost_base);
  SSB_LAUNCH_CHECK("dtw_cost_kernel");
  return SSB_OK;
}

int ssb_dtw_loss_rows(const float* pred, const float* phon, const float* tgt,
                      const int64_t* tgt_phone, const ssb_utt_t* table_dev, int64_t n_utt,
                      const int32_t* path, int64_t path_pitch, int64_t rows_total, int64_t F,
                      int64_t NP, float w, float pairwise_eps, float* row_loss, float* grad_pred,
                      float* grad_phon, void* stream) {
  if (rows_total <= 0) return SSB_OK;
  SSB_REQUIRE(n_utt >= 0 && n_utt < (1 << 30), "dtw loss: bad utterance count");
  SSB_REQUIRE(pred && phon && row_loss && grad_pred && grad_phon, "dtw loss: null pointer");
  SSB_REQUIRE(n_utt == 0 || (tgt && tgt_phone && table_dev), "dtw loss: null pointer");
  SSB_REQUIRE(F >= 1 && F <= MAXF && NP >= 1 && NP <= MAXP, "dtw loss: F=%lld (<= %d), NP=%lld (<= %d)",
              (long long)F, MAXF, (long long)NP, MAXP);
  SSB_REQUIRE(rows_total < (1LL << 31) * 8, "dtw loss: too many rows");
  const int wpb = 8;
  const int64_t grid = (rows_total + wpb - 1) / wpb;
  dtw_loss_rows_kernel<<<(unsigned)grid, wpb * 32, 0, (cudaStream_t)stream>>>(
      pred, phon, tgt, tgt_phone, table_dev, (int)n_utt, path, (int)path_pitch, rows_total, (int)F,
      (int)NP, w, pairwise_eps, row_loss, grad_pred, grad_phon);
  SSB_LAUNCH_CHECK("dtw_loss_rows_kernel");
  return SSB_OK;
}

}  // extern "C"
