/*
 * libssb — C ABI of the B200-native silent_speech transduction hot path.
 *
 * Every entry point is `extern "C" int ssb_<op>(..., void* stream)`:
 *   - pointers are raw DEVICE pointers unless the parameter name ends in `_host`;
 *   - sizes / strides are in ELEMENTS (int64_t) unless the name ends in `_bytes`;
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 *   - the library never allocates or frees device memory and never synchronises the
 *     host: callers own outputs and workspaces (query with the *_workspace_bytes twins);
 *   - return 0 on success, <0 for argument / shape / alignment errors (SSB_ERR_*),
 *     >0 for a cudaError_t; `ssb_last_error()` returns a thread-local message.
 *
 * The reference (dgaddy/silent_speech, pure Python) has no FFI of its own; the
 * interface each entry point replaces is the Python symbol cited beside it
 * (paths relative to the reference checkout). INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 */
#ifndef SSB_H_
#define SSB_H_

#include <stdint.h>

#if defined(__GNUC__)
#define SSB_API __attribute__((visibility("default")))
#else
#define SSB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_OK 0
#define SSB_ERR_ARG (-1)       /* bad argument / shape */
#define SSB_ERR_ALIGN (-2)     /* pointer or stride alignment */
#define SSB_ERR_UNSUPPORTED (-3)
#define SSB_ERR_WORKSPACE (-4) /* workspace too small */

/* ---- library ------------------------------------------------------------ */
/* Bumped whenever a struct layout or a signature in this header changes; the Python binding
 * (silent_speech_b200/_lib.py ABI_VERSION) refuses to load a library reporting another value. */
#define SSB_ABI_VERSION 212
SSB_API int ssb_version(void);               /* == SSB_ABI_VERSION of the header it was built from */
/* sizeof() of the descriptor structs below as compiled into the library (0: ssb_gather_t,
 * 1: ssb_scatter_t, 2: ssb_epilogue_t, 3: ssb_tc_operand_t, 4: ssb_dtw_pair_t, 5: ssb_utt_t,
 * 6: ssb_prep_entry_t, 7: ssb_emg_rec_t;
 * -1 otherwise), so a foreign-language
 * binding can verify its mirror of the layouts at load time. */
SSB_API int64_t ssb_sizeof(int which);
SSB_API const char* ssb_last_error(void);    /* thread-local, never NULL */
SSB_API int ssb_device_sm_count(void);       /* SM count of the current device (148 on B200), <0 on error */
/* Dropout keys for CUDA-graph replay.  Every dropout site keys Philox4x32-10 with the by-value
 * `seed` argument of its entry point (one draw per forward pass: the reference's nn.Dropout,
 * transformer.py:28-41, draws fresh masks per call).  A captured graph would freeze that value,
 * so a process may register ONE device-resident 64-bit offset here: kernels launched afterwards
 * use seed + *dev_seed_offset, read at execution time.  The caller owns the cell, bumps it
 * between replays (a node of the same graph) and passes NULL to detach.  Process-global. */
SSB_API int ssb_set_seed_source(const uint64_t* dev_seed_offset);

/* ---- DTW alignment -------------------------------------------------------
 * Replaces align.py:5-14 (`time_warp`) and align.py:16-34 (`align_from_distances`),
 * called from transduction_model.py:88,126,131.
 *
 * A batch of `npairs` cost matrices of DTW shape (N rows, M columns), fp32.
 * Element (i, j) of pair p lives at  cost[p*pair_stride + i*stride_i + j*stride_j];
 * exactly one of stride_i / stride_j must be 1 (the reference passes `costs.T`,
 * an F-ordered view: stride_i == 1, stride_j == N_alloc).
 *
 * ssb_dtw_align_batch writes, for every pair, `path[p*N + i]` = the smallest column
 * visited in row i by the reference's backtrace (align.py:21-26; tie order
 * up, left, diagonal; path[0] == 0), bit-exact with the reference for finite and
 * +inf costs.
 */
SSB_API int64_t ssb_dtw_workspace_bytes(int64_t npairs, int64_t N, int64_t M, int64_t stride_i,
                                int64_t stride_j);
SSB_API int ssb_dtw_align_batch(const float* cost, int64_t npairs, int64_t pair_stride, int64_t N,
                        int64_t M, int64_t stride_i, int64_t stride_j, int32_t* path,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* Same fill, additionally materialising the accumulated-cost matrix exactly as
 * align.py:5-14 returns it (dtw[0,0]=0, first row/col +inf), into `dtw` with the
 * same strides as `cost`. Debug / API-compat path; slower than ssb_dtw_align_batch. */
SSB_API int ssb_dtw_time_warp_batch(const float* cost, int64_t npairs, int64_t pair_stride, int64_t N,
                            int64_t M, int64_t stride_i, int64_t stride_j, float* dtw,
                            int32_t* path, void* workspace, int64_t workspace_bytes,
                            void* stream);
/* Ragged batches: ONE launch aligns pairs of different shapes (every silent utterance of a real
 * SizeAwareSampler batch has its own (T_target, T_pred): transduction_model.py:111-128 loops over
 * them with one host round trip each).  Matrix p is stored like the reference's `costs` tensor,
 * (M_p = T_pred rows) x (N_p = T_target columns) row-major with row pitch pitch_p >= N_p, starting
 * at cost_base + cost_off_p, and is aligned as its `.T` view exactly as transduction_model.py:126
 * passes it: DTW rows i = target frames.  ssb_dtw_ragged_plan is a pure HOST function: it fills
 * the host table (which the caller uploads; device copy = table_dev), returns the workspace size
 * in bytes (negative on error) and the batch maxima {max N, max M}.  path is (npairs, max_N) int32,
 * rows >= N_p are set to 0.  vectorized != 0 promises that cost_base is 16 B aligned and every
 * cost_off_p and pitch_p is a multiple of 4 floats (16 B loads). */
typedef struct ssb_dtw_pair {
  int64_t cost_off;   /* float offset of the matrix from cost_base */
  int64_t dirs_off;   /* filled by the plan: word offset into the workspace */
  int64_t pitch;      /* floats between consecutive T_pred rows */
  int32_t N, M;       /* DTW rows (target frames), DTW columns (predicted frames) */
  int32_t nbands, nch;/* filled by the plan */
} ssb_dtw_pair_t;
SSB_API int64_t ssb_dtw_ragged_plan(int64_t npairs, const int64_t* N, const int64_t* M,
                                    const int64_t* cost_off, const int64_t* pitch,
                                    ssb_dtw_pair_t* table_out, int64_t* max_dims_out);
SSB_API int ssb_dtw_align_ragged(const float* cost_base, int64_t npairs,
                                 const ssb_dtw_pair_t* table_dev, int64_t max_N, int64_t max_M,
                                 int vectorized, int32_t* path, void* workspace,
                                 int64_t workspace_bytes, void* stream);
/* ---- on-device transduction loss around the DTW kernel (csrc/dtwloss.cu) --------
 * Replaces the per-utterance loop of transduction_model.py:98-157 for a whole batch, ragged shapes
 * included.  Utterance u owns rows [pred_row, pred_row + Tp) of the flattened (rows_total, F)
 * predictions / (rows_total, NP) phoneme logits (decollate_tensor order, data_utils.py:169-178) and
 * rows [tgt_row, tgt_row + Tg) of the concatenated targets (sum Tg, F) / target phonemes (int64).
 * The table is built by the caller (ascending pred_row) and lives in device memory. */
typedef struct ssb_utt {
  int64_t pred_row;   /* first row of the utterance in the flattened prediction tensors */
  int64_t tgt_row;    /* first row in the concatenated target tensors */
  int64_t cost_off;   /* silent: float offset of its (Tp x pitch) cost matrix from cost_base */
  int32_t Tp, Tg;     /* predicted frames, target frames (voiced: Tp == Tg) */
  int32_t pitch;      /* silent: floats between consecutive cost rows (>= Tg) */
  int32_t silent;     /* 1: DTW-aligned (transduction_model.py:115-128), 0: frame-synchronous (:138-145) */
  int32_t pair;       /* silent: row of `path` / index in the ragged DTW batch; voiced: -1 */
  int32_t reserved_;
} ssb_utt_t;
/* cost[p, t] = ||pred[p] - tgt[t]||_2 + w * (logsumexp(phon[p]) - phon[p, tgt_phone[t]]) for every
 * silent utterance (transduction_model.py:116-124), written at cost_base + cost_off. */
SSB_API int ssb_dtw_cost_batch(const float* pred, const float* phon, const float* tgt,
                               const int64_t* tgt_phone, const ssb_utt_t* table_dev, int64_t n_utt,
                               int64_t max_Tp, int64_t max_Tg, int64_t F, int64_t NP, float w,
                               float* cost_base, void* stream);
/* Per row of the flattened predictions: row_loss = its share of the batch loss (silent: the sum
 * over the target frames the path aligned to it of cost[p, t], :128; voiced: ||y - pred + eps|| +
 * w * CE, :141-145; rows outside every utterance: 0) and the gradient of the SUMMED loss with
 * respect to that row of pred / phon.  path: (n_silent, path_pitch) int32 from
 * ssb_dtw_align_ragged (may be NULL when no utterance is silent). */
SSB_API int ssb_dtw_loss_rows(const float* pred, const float* phon, const float* tgt,
                              const int64_t* tgt_phone, const ssb_utt_t* table_dev, int64_t n_utt,
                              const int32_t* path, int64_t path_pitch, int64_t rows_total, int64_t F,
                              int64_t NP, float w, float pairwise_eps, float* row_loss,
                              float* grad_pred, float* grad_phon, void* stream);
/* float64 variant (align.py:6 follows the caller's dtype: a float64 matrix is accumulated and
 * compared in float64).  `dtw` (same strides as `cost`) is required: it is output and workspace.
 * Any positive strides.  Interface parity only - the training path passes fp32. */
SSB_API int ssb_dtw_time_warp_batch_f64(const double* cost, int64_t npairs, int64_t pair_stride,
                                int64_t N, int64_t M, int64_t stride_i, int64_t stride_j,
                                double* dtw, int32_t* path, void* stream);

/* ---- log-mel spectrogram ---------------------------------------------------
 * Replaces data_utils.py:39-62 (`mel_spectrogram`, center=False) together with
 * data_utils.py:29-34 (`dynamic_range_compression_torch` / `spectral_normalize_torch`).
 *
 * y: (B, S) fp32 clips, rows `y_stride` apart.  The (n_fft-hop)/2 reflect padding of
 * data_utils.py:51 is applied inside the kernel.  mel_basis: dense (num_mels, n_fft/2+1)
 * fp32 filterbank exactly as the reference holds it (data_utils.py:47-48); tap_begin/tap_end:
 * int32[num_mels], the half-open range of non-zero bins of each filter (host-computed).
 * out: (B, num_mels, frames) fp32, frames = ssb_mel_num_frames(S, n_fft, hop);
 * out = log(max(mel_basis @ sqrt(re^2+im^2+1e-9), clip_val)).
 * Built for n_fft == win == 1024 (the reference's only configuration, data_utils.py:79).
 */
SSB_API int64_t ssb_mel_num_frames(int64_t S, int n_fft, int hop);
/* range_cell (nullable): two device uint32 words, initialised to 0 by the caller, that receive
 * order-preserving keys of -min(y) and max(y) over every sample the kernel reads (the reference
 * prints a warning when y leaves [-1, 1], data_utils.py:40-43): key(f) = f >= 0 ? bits | 2^31 :
 * ~bits.  Lets the caller test the range without a separate reduction pass. */
SSB_API int ssb_mel_fwd(const float* y, int64_t B, int64_t S, int64_t y_stride, int n_fft, int hop,
                        int win, const float* mel_basis, const int32_t* tap_begin,
                        const int32_t* tap_end, int num_mels, float clip_val, float* out,
                        void* range_cell, void* stream);

/* ---- dense contractions with gathered operands ------------------------------
 * Replaces nn.Linear (architecture.py:51,55,59; transformer.py:32,34), the three
 * nn.Conv1d shapes of ResBlock (architecture.py:18,20,24: k3/p1 stride 1|2, k1 stride 2),
 * the four einsum projections of MultiHeadAttention (transformer.py:96-98,111), the
 * relative-position einsum (transformer.py:249-253) and all of their gradients.
 *
 * A(m,k) is gathered from a channels-last tensor:  m -> (b = m / rows_per_batch,
 * t = m % rows_per_batch),  k -> (tap = k / C, c = k % C),  source row
 * ts = t*s_t + tap*s_tap + off  (zero outside [0, L_src)),  element
 * base[b*batch_stride + ts*ld + c].   A plain row-major matrix is rows_per_batch = M,
 * C = K, L_src = M, s_t = 1, s_tap = 0, off = 0.
 * Output row m goes to  out.base[b*batch_stride + (t*d_t + d_off)*ld + n].
 * Weights are always in "GEMM layout" W[k, n] (n contiguous).
 */
typedef struct {
  const float* base;
  int64_t batch_stride;
  int32_t rows_per_batch, C, L_src, ld, s_t, s_tap, off;
} ssb_gather_t;

typedef struct {
  float* base;
  int64_t batch_stride;
  int32_t rows_per_batch, ld, d_t, d_off;
  int64_t batch_stride_hi;  /* second batch level (tcgen05 batched GEMMs): batch = hi*div + lo */
} ssb_scatter_t;

typedef struct {
  ssb_scatter_t out;
  const float* bias;      /* [N] added before relu, or NULL */
  const float* mask_src;  /* plain (M, N), ld == N: result *= (mask_src > 0) * mask_scale, or NULL */
  float mask_scale;
  int32_t relu;           /* max(x, 0) */
  int32_t accumulate;     /* out += result */
  float drop_p;           /* inverted dropout after relu; 0 disables */
  uint64_t seed;          /* Philox key */
  uint32_t site;          /* Philox stream id of this dropout site */
  /* tcgen05 engine only (ssb_gemm_tc_kmajor; the CUDA-core engine rejects them): */
  void* planes_out;       /* also write the result as bf16 split planes, plain (M, N) per plane;
                             with out.base == NULL the fp32 result is not written at all */
  int64_t planes_stride;  /* elements between the hi and lo output planes */
  const void* mask_planes; /* instead of mask_src: bf16 hi plane (M, N) of the mask source
                              (hi = bf16(x) keeps the sign and zero-ness of x) */
  const void* mask_bits;   /* instead of mask_src: (M, N / 8) bytes, bit n % 8 of byte (m, n / 8) set
                              where the mask source is > 0 (N % 8 == 0) */
  void* mask_bits_out;     /* with planes_out: also write that bit mask of the RESULT (the forward
                              FFN GEMM hands it to the data gradient: 1/16 of the hi plane's bytes) */
  int32_t planes_lrelu;    /* with planes_out: the PLANE copy is leaky_relu(result, planes_neg_slope)
                              while the fp32 output (if any) keeps the result itself: the operand of the
                              next convolution of a HiFi-GAN residual block (hifi_gan/models.py:40-44:
                              x = x + c2(lrelu(c1(lrelu(x)))) needs x in fp32 and lrelu(x) as operand) */
  float planes_neg_slope;
} ssb_epilogue_t;

/* C[m,n] = epi( sum_k A(m,k) * W[k*ldw + n] ) */
SSB_API int ssb_gemm_nn(const ssb_gather_t* A, const float* W, int64_t ldw,
                        const ssb_epilogue_t* epi, int64_t M, int64_t N, int64_t K, void* stream);
/* C[m,j] = epi( sum_k A(m,k) * W[(tapK*N + j)*ldw + k % Cb] ),  tapK = tap{0,1,2}[k / Cb]
 * (data gradients: W is read transposed, tap blocks optionally permuted) */
SSB_API int ssb_gemm_nt(const ssb_gather_t* A, const float* W, int64_t ldw, int64_t Cb, int tap0,
                        int tap1, int tap2, const ssb_epilogue_t* epi, int64_t M, int64_t N,
                        int64_t K, void* stream);
/* dW[k*lddw + n] (+)= sum_m A(m,k) * G[m*ldg + n]   (weight gradients; split over m) */
SSB_API int ssb_gemm_tn(const ssb_gather_t* A, const float* G, int64_t ldg, float* dW,
                        int64_t lddw, int accumulate, int64_t M, int64_t N, int64_t K,
                        void* stream);

/* ---- per-channel reductions, BatchNorm1d, residual+dropout+LayerNorm -----------
 * All tensors (rows, C) fp32, C contiguous and a multiple of 4.
 */
SSB_API int64_t ssb_col_partials_bytes(int64_t rows, int64_t C);
/* out[c] (+)= sum_r x[r,c]   (bias gradients) */
SSB_API int ssb_colsum(const float* x, int64_t rows, int64_t C, float* out, int accumulate,
                       void* workspace, int64_t workspace_bytes, void* stream);
/* same, x given as bf16 split planes [2][rows][C] (x = hi + lo); C a multiple of 8 */
SSB_API int ssb_colsum_planes(const void* planes, int64_t plane_stride, int64_t rows, int64_t C,
                              float* out, int accumulate, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* nn.BatchNorm1d statistics (architecture.py:19,21,25).  training != 0: batch mean / biased var
 * over rows, running stats updated in place (momentum, unbiased var); else running stats.
 * Produces mean, rstd and the fused affine  scale = gamma*rstd, shift = beta - mean*scale. */
SSB_API int ssb_bn_stats(const float* x, int64_t rows, int64_t C, const float* gamma,
                         const float* beta, float* running_mean, float* running_var,
                         float momentum, float eps, int training, float* mean, float* rstd,
                         float* scale, float* shift, void* workspace, int64_t workspace_bytes,
                         void* stream);
/* y = [relu]( (x-mean)*scale + beta [+ (x2-mean2)*scale2 + beta2] )   (architecture.py:32-40)
 * The *_planes arguments of this family are optional (NULL) second outputs: the same result as
 * bf16 split planes [2][rows][C] (ssb_split_bf16 format) for the tcgen05 GEMM that consumes it. */
SSB_API int ssb_bn_apply(const float* x, const float* mean, const float* scale, const float* beta,
                         const float* x2, const float* mean2, const float* scale2,
                         const float* beta2, int relu, int64_t rows, int64_t C, float* y,
                         void* y_planes, void* stream);
/* BatchNorm backward through an optional ReLU mask (dz = dy * (mask_src > 0)).
 * accumulate != 0: dgamma / dbeta are ADDED to (they point into a gradient bucket).
 * workspace >= ssb_col_partials_bytes(rows, C) + 12*C bytes. */
SSB_API int ssb_bn_bwd(const float* dy, const float* mask_src, const float* x, const float* mean,
                       const float* rstd, const float* gamma, int training, int64_t rows,
                       int64_t C, float* dx, void* dx_planes, float* dgamma, float* dbeta,
                       int accumulate, void* workspace, int64_t workspace_bytes, void* stream);
/* The two normalised branches of a ResBlock output, relu(bn2(c2) + res_norm(cr))
 * (architecture.py:37-40), share dz: one call reads dy and the mask once per pass and produces
 * both input gradients and both pairs of parameter gradients.  Same workspace rule. */
SSB_API int ssb_bn_bwd2(const float* dy, const float* mask_src, const float* xa, const float* mean_a,
                        const float* rstd_a, const float* gamma_a, const float* xb,
                        const float* mean_b, const float* rstd_b, const float* gamma_b,
                        int training, int64_t rows, int64_t C, float* dxa, void* dxa_planes,
                        float* dxb, void* dxb_planes, float* dgamma_a, float* dbeta_a,
                        float* dgamma_b, float* dbeta_b, int accumulate, void* workspace,
                        int64_t workspace_bytes, void* stream);
/* z = res + dropout(branch); y = LayerNorm(z)*gamma + beta   (transformer.py:55-56,58-59) */
SSB_API int ssb_add_dropout_ln_fwd(const float* res, const float* branch, const float* gamma,
                                   const float* beta, int64_t rows, int64_t D, float eps,
                                   float drop_p, uint64_t seed, uint32_t site, float* z_out,
                                   float* y, float* mean, float* rstd, void* y_planes,
                                   void* stream);
/* backward: d_branch (fp32) may be NULL when d_branch_planes is given (the consumer is a tensor-core
 * GEMM and nothing reads the fp32 copy). */
SSB_API int64_t ssb_add_dropout_ln_bwd_workspace_bytes(int64_t rows, int64_t D);
SSB_API int ssb_add_dropout_ln_bwd(const float* dy, const float* z, const float* mean,
                                   const float* rstd, const float* gamma, int64_t rows, int64_t D,
                                   float drop_p, uint64_t seed, uint32_t site, float* d_res,
                                   float* d_branch, void* d_branch_planes, float* dgamma,
                                   float* dbeta, int accumulate, void* workspace,
                                   int64_t workspace_bytes, void* stream);

/* ---- banded relative-position attention ------------------------------------------
 * Replaces transformer.py:99-110 (logits, softmax, dropout, PV) with the relative-position
 * logits of transformer.py:162-297 folded in (closed form: band |k-q| <= W exact, W <= 99).
 * qkv: (B*T, 3*H*dh) [q|k|v]; R, P, dS: (B*T, H, RW) band tensors indexed by r = k-q+W
 * (RW % 4 == 0, RW >= 2W+1; padding columns are written as 0); O, dO: (B*T, H*dh).
 * P receives the pre-dropout probabilities (may be NULL for inference).
 * Backward writes dS (for the positional dQ GEMM) and all of dqkv (content part of dq).
 */
SSB_API int ssb_band_attn_fwd(const float* qkv, const float* R, int64_t B, int64_t T, int64_t H,
                              int64_t dh, int64_t W, int64_t RW, float drop_p, uint64_t seed,
                              uint32_t site, float* P, float* O, void* stream);
SSB_API int ssb_band_attn_bwd(const float* qkv, const float* P, const float* dO, int64_t B,
                              int64_t T, int64_t H, int64_t dh, int64_t W, int64_t RW,
                              float drop_p, uint64_t seed, uint32_t site, float* dS, float* dqkv,
                              void* stream);

/* ---- tcgen05 tensor-core GEMM with fp32-class accuracy (bf16 hi/lo split, 3 MMAs) -----
 * Same contractions as ssb_gemm_{nn,nt,tn} for the large shapes, on 5th-gen tensor cores.
 * Operands are "split planes": [2][...] bf16 with x = hi + lo, produced by ssb_split_bf16
 * (n % 8 == 0).  An operand is a channels-last tensor (batches, L_src, C) per plane:
 *   element (b, row, c) of plane p at  planes[p*plane_stride + b*batch_stride + row*ld + c]
 * and is consumed through TMA as the im2col matrix  A(m=(b,t), k=(tap,c)) =
 * x[b, t*s_t + tap*s_tap + off, c]  (rows outside [0, L_src) read as zero), t < rows_out,
 * K = taps*C.  A plain (M, K) matrix is batches = 1, rows_out = L_src = M, C = ld = K.
 */
typedef struct {
  const void* planes;       /* bf16 [2][...] */
  int64_t plane_stride;     /* elements between the hi and lo planes */
  int64_t batch_stride;     /* elements between batch items */
  int32_t batches, rows_out, L_src, C, ld, s_t, s_tap, off;
  int32_t batch_div;        /* > 0: batch b = hi*batch_div + lo with strides (batch_stride, batch_stride_hi) */
  int32_t reserved_;
  int64_t batch_stride_hi;
} ssb_tc_operand_t;

SSB_API int ssb_split_bf16(const float* x, int64_t n, void* planes /* bf16 [2][n] */, void* stream);
/* transposed: planes[p][c][r] = split(x[r][c]) of a row-major (rows, cols) fp32 matrix (both
 * multiples of 8): the W^T operand of an nn.Linear data gradient without an fp32 transpose. */
/* strided: planes[p][r*ld_out + c] = split(x[r*ld_in + c]) for c < cols (cols % 8 == 0); the lo plane
 * starts plane_stride elements after the hi plane. */
SSB_API int ssb_split_bf16_2d(const float* x, int64_t rows, int64_t cols, int64_t ld_in, void* planes,
                              int64_t ld_out, int64_t plane_stride, void* stream);
SSB_API int ssb_split_bf16_t(const float* x, int64_t rows, int64_t cols,
                             void* planes /* bf16 [2][cols][rows] */, void* stream);
/* C[(b,t), n] = epi( sum_k A((b,t), k) * B[n, k] );  B planes: [2][N][K] bf16 (K contiguous).
 * K % 64 == 0, C % 64 == 0.  epi->out.rows_per_batch must equal A->rows_out. */
SSB_API int ssb_gemm_tc_kmajor(const ssb_tc_operand_t* A, const void* Bplanes, int64_t N, int64_t K,
                               const ssb_epilogue_t* epi, void* stream);
/* Stream-K remainder for ssb_gemm_tc_kmajor.  With T output tiles on W persistent workers (CTA pairs)
 * the last wave holds only T mod W tiles; when a workspace is registered, those tiles are cut along K
 * into equal pieces over the idle workers, the piece holding a tile's last k-block adds the others'
 * fp32 partial accumulators (through the workspace) and runs the epilogue.  Deterministic (fixed
 * schedule and summation order per shape); results differ from the classic schedule by fp32
 * re-association only.  The workspace is caller-owned device memory of
 * ssb_gemm_tc_streamk_workspace_bytes() bytes, 256 B aligned, ZERO-FILLED once by the caller and left
 * zero in its flag area by every launch; one per device (the current device at the call), shared by
 * all ssb_gemm_tc_kmajor launches on that device, which therefore must be stream-ordered with respect
 * to each other.  NULL detaches (classic schedule).  Opt-in: ignored unless SSB_STREAMK=1 is set in the
 * environment (on B200 the partial exchange costs more than the idle tail it removes for every shape of
 * the transduction step; csrc/gemm_tc.cu has the numbers). */
SSB_API int64_t ssb_gemm_tc_streamk_workspace_bytes(void);
SSB_API int ssb_gemm_tc_set_streamk_workspace(void* workspace, int64_t workspace_bytes);
/* dW[k, n] (+)= sum_(b,t) X((b,t), k) * G[(b,t), n];  G planes: [2][batches*rows_out][N] bf16.
 * K % 128 == 0, X->C % 128 == 0, N % 8 == 0.
 * group_w > 0 (needs accumulate): element (k, n) goes to dW[(n / group_w) * group_stride + k * lddw +
 * n % group_w] - the fused QKV weight gradient written straight into the (H, D, dh) parameters of
 * transformer.py:71-75 (group_w = dh, lddw = dh, group_stride = D * dh). */
SSB_API int ssb_gemm_tc_wgrad(const ssb_tc_operand_t* X, const void* Gplanes, int64_t g_plane_stride,
                              int64_t N, int64_t K, float* dW, int64_t lddw, int64_t group_w,
                              int64_t group_stride, int accumulate, void* stream);
/* Batched forms (attention): one GEMM per batch item b, all in one launch.
 *   C_b[t, n] = epi( sum_k A_b[t, k] * B_b'[n, k] ),   b' = b (b_mode 1), lo(b) (b_mode 2), 0 (b_mode 0)
 *   B: rows = N (L_src), inner = K (C).  Output row (b, t) at
 *   out.base + lo*batch_stride + hi*batch_stride_hi + (t*d_t + d_off)*ld,  (lo, hi) = (b % div, b / div),
 *   div = A->batch_div. */
SSB_API int ssb_gemm_tc_batched(const ssb_tc_operand_t* A, const ssb_tc_operand_t* B, int b_mode,
                                int64_t N, int64_t K, const ssb_epilogue_t* epi, void* stream);
/*   C_b[f, n] = sum_r X_b[r, f] * G_b[r, n]   (r < X->rows_out; f < K = X->C; n < N = G->C);
 *   both operands are read transposed in place (UMMA MN-major).  Output row (b, f) as above. */
SSB_API int ssb_gemm_tc_batched_tn(const ssb_tc_operand_t* X, const ssb_tc_operand_t* G, int64_t N,
                                   int64_t K, const ssb_epilogue_t* epi, void* stream);

/* ---- element-wise stages of the tensor-core attention path (csrc/attn_tc.cu) ------------
 * Together with ssb_gemm_tc_batched{,_tn} they replace transformer.py:99-110 (+ :162-297)
 * on the tensor cores: per (b, h) dense (T x T) logits, masked to the exact band here.
 * Head dim is zero-padded to 128 in every bf16 plane tensor; Tp = T rounded up to 64. */
SSB_API int ssb_pad_split_heads(const float* x, int64_t rows, int64_t ld_in, int64_t col_off,
                                int64_t G, int64_t dh, void* planes /* (2, rows, G, 128) */,
                                void* stream);
SSB_API int ssb_transpose_split_heads(const float* x, int64_t ld_in, int64_t col_off, int64_t B,
                                      int64_t T, int64_t H, int64_t dh, int64_t Tp,
                                      void* planes /* (2, B*H, 128, Tp) */, void* stream);
/* S (B*H, T, Tp): in raw q.k, out P = softmax(scale*S + R[k-q+W] inside the band); R (B*H, T, RW);
 * Pd_planes (2, B*H, T, Tp) = split(dropout(P)). */
SSB_API int ssb_attn_softmax_fwd(float* S, const float* R, int64_t B, int64_t H, int64_t T,
                                 int64_t Tp, int64_t W, int64_t RW, int64_t dh, float drop_p,
                                 uint64_t seed, uint32_t site, void* Pd_planes, void* stream);
/* dS = P*(dPm - sum(P*dPm)), dPm = dropout-masked dP.  dS_planes (2, B*H, T, Tp) = split(scale*dS);
 * dSband_planes (2, B*T, H, RWp) = split(dS) re-indexed by k-q+W (zero elsewhere). */
SSB_API int ssb_attn_ds_bwd(const float* P, const float* dP, int64_t B, int64_t H, int64_t T,
                            int64_t Tp, int64_t W, int64_t RWp, int64_t dh, float drop_p,
                            uint64_t seed, uint32_t site, void* dS_planes, void* dSband_planes,
                            void* stream);

/* ---- fused banded relative-position attention (csrc/attn_fused.cu) ----------------------------
 * Replaces the body of MultiHeadAttention.forward (transformer.py:99-110) and its autograd
 * backward between the QKV projection and the output projection, with the relative-position
 * logits of LearnedRelativePositionalEmbedding (transformer.py:162-297) in closed form.
 *   qkv_planes : bf16 (2, B*T, 3H, 128) split planes of [q heads | k heads | v heads], zero
 *                padded to 128 per head (ssb_pad_split_heads)
 *   R          : fp32 (B*H, T, RW) positional logits q . E[rel], rel = k - q + W
 *   O          : fp32 (B*T, H*dh);  stat_m / stat_linv : fp32 (B*H, T) row max and 1 / sum exp
 * dh in {32, 64, 96}, W <= 99, RW <= 200 (RW % 4 == 0).  Dropout keys as everywhere else
 * (seed, site, element), identical masks in forward and backward. */
/* head_stride / do_head_stride: elements between consecutive heads in the plane tensors: 128 (or
 * 0) = heads zero-padded to 128 columns (ssb_pad_split_heads); dh = PACKED, i.e. the plain split
 * planes of the (B*T, 3*H*dh) projection / (B*T, H*dh) output gradient exactly as a GEMM epilogue
 * writes them (`planes_out`), so no re-layout pass runs between the GEMM and the attention kernel. */
SSB_API int ssb_attn_fused_fwd(const void* qkv_planes, const float* R, int64_t B, int64_t T, int64_t H,
                       int64_t dh, int64_t W, int64_t RW, float drop_p, uint64_t seed, uint32_t site,
                       float* O, void* O_planes, float* stat_m, float* stat_linv, int64_t head_stride,
                       void* stream);
/* delta[b*H+h, q] = sum_d dO * O over the head's dh columns */
/* zero_out (nullable): additionally clears H*dh floats of every row of a (B*T, zero_ld) matrix -
 * the content-dQ accumulator of ssb_attn_fused_bwd - in the same pass. */
SSB_API int ssb_attn_delta(const float* O, const float* dO, int64_t B, int64_t T, int64_t H, int64_t dh,
                   float* delta, float* zero_out, int64_t zero_ld, void* stream);
/* dqkv: rows dq_ld floats apart; the dQ columns [0, H*dh) must be zero on entry (the content part is
 * accumulated with red.global.add).  dkv_planes == NULL: dq_ld == 3*H*dh and dK / dV overwrite the
 * other two thirds of the (B*T, 3*H*dh) buffer.  dkv_planes != NULL: a bf16 (2, B*T, 3*H*dh) planes
 * buffer whose dK / dV thirds are written instead (the operand format of the QKV data-gradient
 * GEMM; its dQ third is the caller's to fill, ssb_split_bf16_2d), and dqkv may be (B*T, H*dh).
 * dSband_planes: bf16 (2, B*T, H, RWp), zero on entry; receives dS in band layout for the positional
 * part of dQ (a tensor-core GEMM with E).
 * ssb_attn_fused_fwd: O_planes (nullable) additionally receives O as bf16 (2, B*T, H*dh) planes. */
SSB_API int ssb_attn_fused_bwd(const void* qkv_planes, const void* dO_planes, const float* R,
                       const float* stat_m, const float* stat_linv, const float* delta, int64_t B,
                       int64_t T, int64_t H, int64_t dh, int64_t W, int64_t RW, float drop_p,
                       uint64_t seed, uint32_t site, float* dqkv, int64_t dq_ld, void* dkv_planes,
                       void* dSband_planes, int64_t RWp, int64_t head_stride, int64_t do_head_stride,
                       void* stream);

/* ---- weight operand preparation (csrc/prep.cu) -----------------------------------------------
 * One launch writes the bf16 hi/lo split planes of every parameter in every layout the tensor-core
 * GEMMs of a step consume (replaces the per-use `.t().contiguous()` / cat / split passes over
 * nn.Linear, nn.Conv1d and the attention projection weights: architecture.py:18-24,51-59,
 * transformer.py:32-34,71-78).  Entry = a strided 2-D view of one fp32 parameter,
 *   view(r, c) = src[(r / RL) * s_rhi + (r % RL) * s_rlo + (c / CL) * s_chi + (c % CL) * s_clo],
 * written as planes to dst_n[r * ld_n + c] (hi; lo at + plane_n elements) and / or transposed to
 * dst_t[c * ld_t + r] (lo at + plane_t).  ssb_prep_plan (HOST) fills tile0 / tiles_c and returns
 * the total tile count (negative on error); the caller uploads the table. */
typedef struct ssb_prep_entry {
  const float* src;
  void* dst_n;            /* bf16, nullable */
  void* dst_t;            /* bf16, nullable */
  int64_t plane_n, plane_t;
  int64_t s_rhi, s_rlo, s_chi, s_clo;
  int32_t rows, cols, RL, CL;
  int32_t ld_n, ld_t;
  int32_t tile0, tiles_c; /* filled by ssb_prep_plan */
} ssb_prep_entry_t;
SSB_API int64_t ssb_prep_plan(ssb_prep_entry_t* table_host, int64_t n_entries);
SSB_API int ssb_prep_planes(const ssb_prep_entry_t* table_dev, int64_t n_entries,
                            int64_t total_tiles, void* stream);

/* ---- fused log_softmax + CTC loss (csrc/ctc.cu) ------------------------------------------------
 * Replaces recognition_model.py:96-101: F.log_softmax(pred, 2) -> pad_sequence ->
 * F.ctc_loss(pred, y, lengths, text_int_lengths, blank) and its backward.
 *   logits (N, T, C) fp32, batch-first, frames >= input_lengths[n] are padding;
 *   targets (N, Lmax) int64 (device), entries >= target_lengths[n] ignored; lengths int64 (device);
 *   nll (N): negative log-likelihood per utterance (+inf when no alignment exists);
 *   grad_logits (N, T, C) or NULL: d nll[n] / d logits (softmax - occupancy; 0 on padding frames
 *   and for infeasible utterances), scaled by 1 / (N * max(L_n, 1)) when mean_reduction != 0
 *   (= the gradient of torch's default reduction='mean').
 * 2*Lmax + 1 <= 1024.  workspace: ssb_ctc_workspace_bytes (the alpha table). */
SSB_API int64_t ssb_ctc_workspace_bytes(int64_t N, int64_t T, int64_t Lmax);
SSB_API int ssb_ctc_loss_fused(const float* logits, int64_t N, int64_t T, int64_t C,
                               const int64_t* targets, int64_t Lmax, const int64_t* input_lengths,
                               const int64_t* target_lengths, int64_t blank, int mean_reduction,
                               float* nll, float* grad_logits, void* workspace,
                               int64_t workspace_bytes, void* stream);

/* ---- fused AdamW on flat buffers (csrc/optim.cu) ------------------------------------------------
 * Replaces optim.step() of torch.optim.AdamW(model.parameters(), weight_decay=l2) with the
 * per-iteration learning rate of transduction_model.py:178-189,210.
 * p, g, m, v: flat fp32 buffers of n elements (parameters, gradients = the data-parallel bucket,
 * exp_avg, exp_avg_sq), 16 B aligned.  torch.optim.AdamW arithmetic (decoupled weight decay,
 * bias-corrected moments); g is multiplied by grad_scale first (1 / world_size of the mean).
 * lr_cell (float) and step_cell (int64) are DEVICE cells owned by the caller: the call first
 * increments *step_cell, then updates with *lr_cell, both read at execution time, so a captured
 * CUDA graph replays a valid optimiser step. */
SSB_API int ssb_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n,
                           const float* lr_cell, int64_t* step_cell, float beta1, float beta2,
                           float eps, float weight_decay, float grad_scale, void* stream);

/* ---- EMG signal conditioning (csrc/emg.cu) ------------------------------
 * Replaces, for a whole batch of recordings, the per-channel scipy / numpy chain of
 * read_emg.py:27-45 as load_utterance applies it (read_emg.py:62-67):
 *   notch_harmonics(x, 60, 1000) -> remove_drift(x, 1000)      = a cascade of scipy.signal.filtfilt
 *   subsample(x, new_freq, old_freq)                           = np.interp onto a uniform grid
 * in float64 with the reference's operation order: results are bit-identical to scipy / numpy.
 * Signals are (rows, C) float64 row-major with the recordings stacked along rows; recording r owns
 * rows [off, off + n) of the input / filtered arrays and rows [out_off, out_off + n_out) of the
 * resampled array (out_off dense and ascending).  The table lives in device memory.
 *
 * ssb_emg_filtfilt_chain: y = filtfilt_S(... filtfilt_1(x)) per recording and channel (defaults of
 * scipy.signal.filtfilt: odd extension by 3 * ntaps samples, lfilter_zi initial state).  Stage s is
 * described by 11 doubles at coef_host + 11 s = { b[0..3], a[0..3] (a[0] == 1), zi[0..2] } and
 * ntaps_host[s] in {3, 4}; unused entries are ignored.  Coefficients are design-time data computed
 * by the caller (scipy.signal.iirnotch / butter / lfilter_zi, exactly as the reference does).
 * min_n = the smallest n of the table (scipy requires n > padlen).  x and y may not alias.
 * ssb_emg_subsample: out[out_off + i] = np.interp(i * step, arange(n) / old_freq, y) for i < n_out;
 * step = 1 / new_freq as the host computed it; out is float64, or float32 when out_f32 != 0. */
typedef struct ssb_emg_rec {
  int64_t off;        /* first row of the recording in the (rows, C) signal arrays */
  int64_t out_off;    /* first row in the resampled output */
  int32_t n;          /* samples */
  int32_t n_out;      /* resampled samples */
} ssb_emg_rec_t;
SSB_API int64_t ssb_emg_filtfilt_workspace_bytes(int64_t rows_total, int64_t n_rec, int C);
SSB_API int ssb_emg_filtfilt_chain(const double* x, double* y, const ssb_emg_rec_t* table_dev,
                                   int64_t n_rec, int64_t rows_total, int min_n, int C,
                                   const double* coef_host, const int32_t* ntaps_host, int n_stages,
                                   void* workspace, int64_t workspace_bytes, void* stream);
SSB_API int ssb_emg_subsample(const double* x, const ssb_emg_rec_t* table_dev, int64_t n_rec,
                              int64_t out_rows_total, int C, double old_freq, double step,
                              void* out, int out_f32, void* stream);

/* ---- HiFi-GAN generator inference (csrc/vocoder.cu) ---------------------
 * Replaces the element-wise ends of hifi_gan/models.py:96-112 (Generator.forward, called by
 * vocoder.py:28-36); every convolution of the generator runs through ssb_gemm_tc_kmajor with
 * s_tap = dilation (host side: silent_speech_b200/vocoder.py).  Tensors are channels-last
 * (rows = samples, C contiguous).
 *
 * ssb_voc_mix: v = scale * (a + b + c) over n elements (b, c nullable) - the multi-receptive-field
 * fusion xs / num_kernels of models.py:101-108.  Written as fp32 to `out` (nullable) and / or as
 * bf16 split planes (hi at planes, lo plane_stride elements later; nullable) of
 * leaky_relu(v, neg_slope) when lrelu_on != 0 (models.py:99), of v otherwise.  n % 4 == 0.
 *
 * ssb_voc_post: audio[t] = tanh(bias + sum_{tap, ch} w[tap * C + ch] *
 *     leaky_relu(scale * (a + b + c)[t + tap - taps / 2, ch], neg_slope)), rows outside [0, rows)
 * reading as zero padding - models.py:109-111 (F.leaky_relu default slope, conv_post with its
 * single output channel, tanh) fused with the last stage's fusion.  taps odd, taps * C <= 1024. */
SSB_API int ssb_voc_mix(const float* a, const float* b, const float* c, int64_t n, float scale,
                        int lrelu_on, float neg_slope, float* out, void* planes,
                        int64_t plane_stride, void* stream);
SSB_API int ssb_voc_post(const float* a, const float* b, const float* c, int64_t rows, int64_t C,
                         int64_t taps, float scale, float neg_slope, const float* w, float bias,
                         float* audio, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H_ */
