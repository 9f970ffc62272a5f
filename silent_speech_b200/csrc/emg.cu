// EMG signal conditioning for a batch of recordings: the scipy / numpy
// chain of read_emg.py:27-45,62-67
//
//   x = apply_to_all(notch_harmonics, x, 60, 1000)     seven filtfilt passes of iirnotch sections
//   x = apply_to_all(remove_drift, x, 1000)            filtfilt of a 3rd-order Butterworth high-pass
//   emg_orig = apply_to_all(subsample, x, 689.06, 1000)   np.interp onto the model's sample grid
//
// in float64 with the reference's operation order, so the result is BIT-IDENTICAL to scipy /
// numpy (every product and sum is an explicit round-to-nearest intrinsic: no FMA contraction).
//
// scipy.signal.filtfilt(b, a, x) with its defaults is, per channel:
//   ext = odd extension of x by padlen = 3 * max(len(a), len(b)) samples at both ends
//   zi  = lfilter_zi(b, a)                              (host: coefficients are design-time data)
//   y   = lfilter(b, a, ext, zi = zi * ext[0])          direct form II transposed, forward
//   y   = lfilter(b, a, y[::-1], zi = zi * y[-1])[::-1] the same filter, backward
//   return y[padlen:-padlen]
// and lfilter's recurrence (scipy/signal/_lfilter.c.in) is, with a[0] == 1,
//   y[n] = z[0] + b[0] x[n];  z[k] = z[k+1] + x[n] b[k+1] - y[n] a[k+1];  z[last] = x[n] b[last] - y[n] a[last].
//
// Parallelism.  The recurrence is a chain of dependent float64 operations along time; being
// bit-exact rules out re-associating it (a blocked scan over time changes the rounding).  What is
// independent is the (recording, channel) pair: one thread per pair walks the whole cascade, the
// 8 channels of a recording sit in one 64 B row so a warp's loads coalesce per recording, and a
// batch of recordings (a whole session directory) fills the machine.  The kernel is latency-bound
// by construction: ~3 dependent DP operations per sample and pass, 16 passes per recording.
#include "ssb_common.cuh"

namespace {

constexpr int MAX_STAGES = 16;
constexpr int MAX_PAD = 12;       // 3 * 4 taps (order 3)

struct Stage {
  double b[4], a[4], zi[3];
  int ntaps;                      // 3 (biquad) or 4 (order 3); padlen = 3 * ntaps
};

struct ChainParams {
  Stage st[MAX_STAGES];
  int n_stages;
  const double* x;
  double* y;
  double* scratch;                // forward-pass output, (n + 2 * MAX_PAD) rows per recording
  const ssb_emg_rec_t* table;
  int64_t n_rec;
  int C;
};

// one lfilter step; NT = taps
template <int NT>
__device__ __forceinline__ double df2t(const Stage& s, double (&z)[3], double xn) {
  const double yn = __dadd_rn(z[0], __dmul_rn(s.b[0], xn));
#pragma unroll
  for (int k = 0; k < NT - 2; ++k)
    z[k] = __dsub_rn(__dadd_rn(z[k + 1], __dmul_rn(xn, s.b[k + 1])), __dmul_rn(yn, s.a[k + 1]));
  z[NT - 2] = __dsub_rn(__dmul_rn(xn, s.b[NT - 1]), __dmul_rn(yn, s.a[NT - 1]));
  return yn;
}

// Runs the recurrence over indices i0, i0 + dir, ..., (count of them), reading x through `load(i)`
// and handing y to `store(i, y)`.  The loads do not depend on the recurrence: a block of BLK
// samples is fetched while the previous block is being filtered (registers, double-buffered), so
// the chain of dependent DP operations never waits for global memory.
constexpr int BLK = 8;
template <int NT, class Load, class Store>
__device__ __forceinline__ void run_pass(const Stage& s, double (&z)[3], int i0, int count, int dir,
                                         Load load, Store store) {
  double cur[BLK], nxt[BLK];
#pragma unroll
  for (int k = 0; k < BLK; ++k) cur[k] = k < count ? load(i0 + dir * k) : 0.0;
  for (int done = 0; done < count; done += BLK) {
    const int base = i0 + dir * done, left = count - done - BLK;
#pragma unroll
    for (int k = 0; k < BLK; ++k) nxt[k] = k < left ? load(base + dir * (BLK + k)) : 0.0;
#pragma unroll
    for (int k = 0; k < BLK; ++k)
      if (done + k < count) store(base + dir * k, df2t<NT>(s, z, cur[k]));
#pragma unroll
    for (int k = 0; k < BLK; ++k) cur[k] = nxt[k];
  }
}

template <int NT>
__device__ void filtfilt_stage(const Stage& s, const double* src, double* dst, double* fwd, int n,
                               int C) {
  constexpr int PAD = 3 * NT;
  const int m = n + 2 * PAD;
  // forward over the odd extension: 2 x[0] - x[PAD..1], x, 2 x[n-1] - x[n-2..n-1-PAD]
  const double x0 = src[0], xl = src[(int64_t)(n - 1) * C];
  const double two_x0 = __dmul_rn(2.0, x0), two_xl = __dmul_rn(2.0, xl);
  const double ext0 = __dsub_rn(two_x0, src[(int64_t)PAD * C]);
  double z[3];
#pragma unroll
  for (int k = 0; k < NT - 1; ++k) z[k] = __dmul_rn(s.zi[k], ext0);
  auto put_fwd = [&](int i, double y) { fwd[(int64_t)i * C] = y; };
  run_pass<NT>(s, z, 0, PAD, 1,
               [&](int i) { return __dsub_rn(two_x0, src[(int64_t)(PAD - i) * C]); }, put_fwd);
  run_pass<NT>(s, z, PAD, n, 1, [&](int i) { return src[(int64_t)(i - PAD) * C]; }, put_fwd);
  run_pass<NT>(s, z, PAD + n, PAD, 1,
               [&](int i) { return __dsub_rn(two_xl, src[(int64_t)(n - 2 - (i - PAD - n)) * C]); },
               put_fwd);
  // backward over the forward result; only the n central samples are kept.  (src may be dst: the
  // forward pass above has consumed it completely.)
  const double y0 = fwd[(int64_t)(m - 1) * C];
#pragma unroll
  for (int k = 0; k < NT - 1; ++k) z[k] = __dmul_rn(s.zi[k], y0);
  auto get_fwd = [&](int i) { return fwd[(int64_t)i * C]; };
  run_pass<NT>(s, z, m - 1, PAD, -1, get_fwd, [](int, double) {});
  run_pass<NT>(s, z, PAD + n - 1, n, -1, get_fwd,
               [&](int i, double y) { dst[(int64_t)(i - PAD) * C] = y; });
}

__global__ void __launch_bounds__(64) emg_filtfilt_kernel(const __grid_constant__ ChainParams p) {
  const int64_t chain = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= p.n_rec * p.C) return;
  const int64_t rec = chain / p.C;
  const int c = (int)(chain - rec * p.C);
  const ssb_emg_rec_t t = p.table[rec];
  const double* x = p.x + t.off * p.C + c;
  double* y = p.y + t.off * p.C + c;
  double* fwd = p.scratch + (t.off + rec * 2 * MAX_PAD) * p.C + c;
  for (int s = 0; s < p.n_stages; ++s) {
    const double* src = s == 0 ? x : y;
    if (p.st[s].ntaps == 3) filtfilt_stage<3>(p.st[s], src, y, fwd, t.n, p.C);
    else filtfilt_stage<4>(p.st[s], src, y, fwd, t.n, p.C);
  }
}

// np.interp(sample_times, times, signal) of read_emg.py:40-45 with times[j] = j / old_freq and
// sample_times[i] = i * step (step = 1 / new_freq as the host computed it): one thread per output
// sample and channel.  numpy (compiled_base.c): slope = (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]);
// out = slope * (x - xp[j]) + fp[j], with out = fp[j] when x == xp[j].
template <typename OutT>
__global__ void __launch_bounds__(256) emg_subsample_kernel(const double* __restrict__ x,
                                                           OutT* __restrict__ out,
                                                           const ssb_emg_rec_t* __restrict__ table,
                                                           int64_t n_rec, int C, double old_freq,
                                                           double step, int64_t work_total) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < work_total;
       w += (int64_t)gridDim.x * blockDim.x) {
    // w indexes (output row over all recordings, channel); find the recording by bisection
    const int64_t row = w / C;
    const int c = (int)(w - row * C);
    int64_t lo = 0, hi = n_rec - 1;
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (table[mid].out_off <= row) lo = mid; else hi = mid - 1;
    }
    const ssb_emg_rec_t t = table[lo];
    const int64_t i = row - t.out_off;
    if (i >= t.n_out) continue;           // (cannot happen: out_off is dense)
    const double xv = __dmul_rn((double)i, step);
    int64_t j = (int64_t)floor(xv * old_freq);
    if (j > t.n - 2) j = t.n - 2;
    if (j < 0) j = 0;
    // times[j] <= xv < times[j+1] with the reference's own grid values
    while (j > 0 && __ddiv_rn((double)j, old_freq) > xv) --j;
    while (j < t.n - 2 && __ddiv_rn((double)(j + 1), old_freq) <= xv) ++j;
    const double xj = __ddiv_rn((double)j, old_freq), xj1 = __ddiv_rn((double)(j + 1), old_freq);
    const double fj = x[(t.off + j) * C + c], fj1 = x[(t.off + j + 1) * C + c];
    double r;
    if (xj == xv) {
      r = fj;
    } else {
      const double slope = __ddiv_rn(__dsub_rn(fj1, fj), __dsub_rn(xj1, xj));
      r = __dadd_rn(__dmul_rn(slope, __dsub_rn(xv, xj)), fj);
    }
    out[row * C + c] = (OutT)r;
  }
}

}  // namespace

extern "C" {

int64_t ssb_emg_filtfilt_workspace_bytes(int64_t rows_total, int64_t n_rec, int C) {
  if (rows_total < 0 || n_rec < 0 || C < 1) return SSB_ERR_ARG;
  const int64_t b = (rows_total + n_rec * 2 * MAX_PAD) * C * (int64_t)sizeof(double);
  return b > 16 ? b : 16;
}

int ssb_emg_filtfilt_chain(const double* x, double* y, const ssb_emg_rec_t* table_dev, int64_t n_rec,
                           int64_t rows_total, int min_n, int C, const double* coef_host,
                           const int32_t* ntaps_host, int n_stages, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  SSB_REQUIRE(n_rec >= 0 && C >= 1 && rows_total >= 0, "emg: bad sizes");
  if (n_rec == 0) return SSB_OK;
  SSB_REQUIRE(x && y && table_dev && coef_host && ntaps_host && workspace, "emg: null pointer");
  SSB_REQUIRE(n_stages >= 1 && n_stages <= MAX_STAGES, "emg: n_stages %d outside [1, %d]", n_stages,
              MAX_STAGES);
  const int64_t need = ssb_emg_filtfilt_workspace_bytes(rows_total, n_rec, C);
  if (workspace_bytes < need) {
    ssb::set_error("emg: workspace %lld B < required %lld B", (long long)workspace_bytes,
                   (long long)need);
    return SSB_ERR_WORKSPACE;
  }
  ChainParams p;
  int max_pad = 0;
  for (int s = 0; s < n_stages; ++s) {
    const int nt = ntaps_host[s];
    SSB_REQUIRE(nt == 3 || nt == 4, "emg: stage %d has %d taps (3 or 4 supported)", s, nt);
    const double* c = coef_host + (int64_t)s * 11;
    SSB_REQUIRE(c[4] == 1.0, "emg: stage %d is not normalised (a[0] = %g)", s, c[4]);
    for (int k = 0; k < 4; ++k) { p.st[s].b[k] = c[k]; p.st[s].a[k] = c[4 + k]; }
    for (int k = 0; k < 3; ++k) p.st[s].zi[k] = c[8 + k];
    p.st[s].ntaps = nt;
    max_pad = 3 * nt > max_pad ? 3 * nt : max_pad;
  }
  // scipy raises "The length of the input vector x must be greater than padlen"
  SSB_REQUIRE(min_n > max_pad, "emg: a recording has %d samples, filtfilt needs more than padlen = %d",
              min_n, max_pad);
  p.n_stages = n_stages;
  p.x = x; p.y = y; p.scratch = (double*)workspace;
  p.table = table_dev; p.n_rec = n_rec; p.C = C;
  const int64_t chains = n_rec * C;
  const int threads = 64;
  emg_filtfilt_kernel<<<(unsigned)((chains + threads - 1) / threads), threads, 0,
                        (cudaStream_t)stream>>>(p);
  SSB_LAUNCH_CHECK("emg_filtfilt_kernel");
  return SSB_OK;
}

int ssb_emg_subsample(const double* x, const ssb_emg_rec_t* table_dev, int64_t n_rec,
                      int64_t out_rows_total, int C, double old_freq, double step, void* out,
                      int out_f32, void* stream) {
  SSB_REQUIRE(n_rec >= 0 && C >= 1 && out_rows_total >= 0, "emg: bad sizes");
  if (n_rec == 0 || out_rows_total == 0) return SSB_OK;
  SSB_REQUIRE(x && table_dev && out, "emg: null pointer");
  SSB_REQUIRE(old_freq > 0.0 && step > 0.0, "emg: bad rates");
  const int64_t work = out_rows_total * C;
  const int threads = 256;
  int64_t grid = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)ssb::num_sms() * 16;
  if (grid > cap) grid = cap;
  if (out_f32)
    emg_subsample_kernel<float><<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(
        x, (float*)out, table_dev, n_rec, C, old_freq, step, work);
  else
    emg_subsample_kernel<double><<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(
        x, (double*)out, table_dev, n_rec, C, old_freq, step, work);
  SSB_LAUNCH_CHECK("emg_subsample_kernel");
  return SSB_OK;
}

}  // extern "C"
