/* ORACLE - test infrastructure, not product code.
 *
 * The recurrence inside scipy.signal.lfilter (third-party: scipy, unpinned in the reference's
 * environment.yml; algorithm: scipy/signal/_lfilter.c.in, DOUBLE_filt - direct form II transposed)
 * as read_emg.py:27-33 reaches it through scipy.signal.filtfilt, and numpy's np.interp
 * (numpy/core/src/multiarray/compiled_base.c, arr_interp) as read_emg.py:40-45 calls it.
 * Compiled with -ffp-contract=off: every product and sum rounds separately, as in the reference's
 * own binaries. */
#include <math.h>
#include <stdint.h>

/* y[n] = z[0] + b[0] x[n]; z[k] = z[k+1] + x[n] b[k+1] - y[n] a[k+1]; z[last] = x[n] b[last] - y[n] a[last]
 * b, a: ntaps coefficients with a[0] == 1; z: ntaps-1 delays, updated in place; stride in elements. */
void ssb_oracle_lfilter(const double* b, const double* a, int ntaps, const double* x, int64_t n,
                        int64_t stride, double* z, double* y) {
  for (int64_t i = 0; i < n; ++i) {
    const double xn = x[i * stride];
    const double yn = z[0] + b[0] * xn;
    for (int k = 0; k < ntaps - 2; ++k) z[k] = z[k + 1] + xn * b[k + 1] - yn * a[k + 1];
    z[ntaps - 2] = xn * b[ntaps - 1] - yn * a[ntaps - 1];
    y[i] = yn;
  }
}

/* out[i] = np.interp(i * step, arange(n) / old_freq, f), i < n_out  (x strictly inside the grid) */
void ssb_oracle_interp_uniform(const double* f, int64_t n, double old_freq, double step,
                               int64_t n_out, double* out) {
  for (int64_t i = 0; i < n_out; ++i) {
    const double xv = (double)i * step;
    int64_t j = (int64_t)floor(xv * old_freq);
    if (j > n - 2) j = n - 2;
    if (j < 0) j = 0;
    while (j > 0 && (double)j / old_freq > xv) --j;
    while (j < n - 2 && (double)(j + 1) / old_freq <= xv) ++j;
    const double xj = (double)j / old_freq, xj1 = (double)(j + 1) / old_freq;
    if (xj == xv) {
      out[i] = f[j];
    } else {
      const double slope = (f[j + 1] - f[j]) / (xj1 - xj);
      out[i] = slope * (xv - xj) + f[j];
    }
  }
}
