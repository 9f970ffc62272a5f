/*
 * ORACLE — test infrastructure, not product code.
 *
 * CPU restatement (plain C) of the reference's DTW alignment.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * it; the product path (silent_speech_b200/) never does.
 *
 * Pinned against the executed reference: tests/golden/dtw_*.npz were produced by running
 * /root/reference/align.py (numba) itself — see tests/golden/make_golden.py — and
 * tests/test_oracle_dtw.py checks this file against them bit for bit.
 *
 * Follows, line by line:
 *   time_warp             align.py:5-14   (dtw[0,0]=0, first row/col +inf, one add per cell)
 *   align_from_distances  align.py:16-34  (backtrace from (N-1,M-1); Python min() over
 *                                          [(i-1,j),(i,j-1),(i-1,j-1)] keeps the FIRST minimum)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AT(p, i, j, si, sj) ((p)[(int64_t)(i) * (si) + (int64_t)(j) * (sj)])

#define DEFINE_DTW(T, SUF)                                                                      \
  /* align.py:5-14.  dtw is written C-contiguous (N, M). */                                     \
  void ssb_oracle_time_warp_##SUF(const T* costs, int64_t N, int64_t M, int64_t si, int64_t sj, \
                                  T* dtw) {                                                     \
    for (int64_t k = 0; k < N * M; ++k) dtw[k] = (T)0;             /* np.zeros_like  :6 */      \
    for (int64_t j = 1; j < M; ++j) dtw[j] = (T)INFINITY;          /* dtw[0,1:]=inf  :7 */      \
    for (int64_t i = 1; i < N; ++i) dtw[i * M] = (T)INFINITY;      /* dtw[1:,0]=inf  :8 */      \
    for (int64_t i = 1; i < N; ++i) {                              /* :11 */                    \
      for (int64_t j = 1; j < M; ++j) {                            /* :12 */                    \
        T best = dtw[(i - 1) * M + j];                             /* min(up, left, diag) :13 */\
        const T left = dtw[i * M + j - 1];                                                      \
        const T diag = dtw[(i - 1) * M + j - 1];                                                \
        if (left < best) best = left;                                                           \
        if (diag < best) best = diag;                                                           \
        dtw[i * M + j] = AT(costs, i, j, si, sj) + best;                                        \
      }                                                                                         \
    }                                                                                           \
  }                                                                                             \
  /* align.py:16-34.  results has N entries; dtw is (N, M) scratch. */                          \
  void ssb_oracle_align_##SUF(const T* costs, int64_t N, int64_t M, int64_t si, int64_t sj,     \
                              int32_t* results, T* dtw) {                                       \
    ssb_oracle_time_warp_##SUF(costs, N, M, si, sj, dtw);                                       \
    int64_t i = N - 1, j = M - 1;                                  /* :19-20 */                 \
    for (int64_t k = 0; k < N; ++k) results[k] = 0;                /* :21 */                    \
    while (i > 0 && j > 0) {                                       /* :22 */                    \
      results[i] = (int32_t)j;                                     /* :23 */                    \
      int64_t bi = i - 1, bj = j;                                  /* candidates in list order */\
      T best = dtw[(i - 1) * M + j];                                                            \
      const T left = dtw[i * M + j - 1];                                                        \
      const T diag = dtw[(i - 1) * M + j - 1];                                                  \
      if (left < best) { best = left; bi = i; bj = j - 1; }                                     \
      if (diag < best) { best = diag; bi = i - 1; bj = j - 1; }                                 \
      i = bi; j = bj;                                              /* :24 */                    \
    }                                                                                           \
  }

DEFINE_DTW(float, f32)
DEFINE_DTW(double, f64)

/* Batched fp32 driver used by the CPU baseline: pairs are independent, so all host cores
 * can be used (OpenMP when compiled with -fopenmp).  Layout as ssb_dtw_align_batch. */
int ssb_oracle_align_batch_f32(const float* cost, int64_t npairs, int64_t pair_stride, int64_t N,
                               int64_t M, int64_t si, int64_t sj, int32_t* path, int threads) {
  int rc = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
#endif
  {
    float* dtw = (float*)malloc((size_t)N * M * sizeof(float));
    if (!dtw) {
      rc = 1;
    } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
      for (int64_t p = 0; p < npairs; ++p)
        ssb_oracle_align_f32(cost + p * pair_stride, N, M, si, sj, path + p * N, dtw);
      free(dtw);
    }
  }
  return rc;
}
