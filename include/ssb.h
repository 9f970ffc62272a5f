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
SSB_API int ssb_version(void);               /* 10000*major + 100*minor + patch */
SSB_API const char* ssb_last_error(void);    /* thread-local, never NULL */
SSB_API int ssb_device_sm_count(void);       /* SM count of the current device (148 on B200), <0 on error */

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
SSB_API int ssb_mel_fwd(const float* y, int64_t B, int64_t S, int64_t y_stride, int n_fft, int hop,
                        int win, const float* mel_basis, const int32_t* tap_begin,
                        const int32_t* tap_end, int num_mels, float clip_val, float* out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H_ */
