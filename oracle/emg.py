"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference's EMG signal conditioning (read_emg.py:27-51) in float64:
`remove_drift`, `notch`, `notch_harmonics`, `subsample`, `apply_to_all`.  The reference delegates
the arithmetic to third-party code that is not under /root/reference — scipy.signal.filtfilt
(scipy, unpinned: environment.yml) and np.interp (numpy) — so their published algorithms are
restated here (filtfilt's odd extension / lfilter_zi start / forward-backward structure in numpy,
the lfilter and interp inner loops in oracle/emg_oracle.c) and pinned against outputs of the
reference's own functions: tests/golden/emg_golden.npz, made by tests/golden/make_golden_emg.py.
Filter DESIGN (iirnotch, butter, lfilter_zi) is taken from scipy, as the reference takes it.
"""
import ctypes

import numpy as np

from . import lib as _lib

_i64 = ctypes.c_int64
_dbl = ctypes.c_double


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def lfilter(b, a, x, zi):
    """scipy.signal.lfilter(b, a, x, zi=zi)[0] for a 1-D float64 signal, a[0] == 1."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    nt = max(len(a), len(b))
    bb, aa = np.zeros(nt), np.zeros(nt)
    bb[:len(b)], aa[:len(a)] = b, a
    assert aa[0] == 1.0
    x = np.ascontiguousarray(x, dtype=np.float64)
    z = np.array(zi, dtype=np.float64).copy()
    y = np.empty_like(x)
    _lib().ssb_oracle_lfilter(_p(bb), _p(aa), ctypes.c_int(nt), _p(x), _i64(len(x)), _i64(1), _p(z), _p(y))
    return y


def filtfilt(b, a, x):
    """scipy.signal.filtfilt(b, a, x) with its defaults (padtype='odd', padlen=3*ntaps, method='pad')."""
    import scipy.signal
    b = np.asarray(b, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    if a[0] != 1.0:
        b, a = b / a[0], a / a[0]
    x = np.asarray(x, dtype=np.float64)
    edge = 3 * max(len(a), len(b))
    assert x.shape[0] > edge, "The length of the input vector x must be greater than padlen"
    ext = np.concatenate([2 * x[0] - x[edge:0:-1], x, 2 * x[-1] - x[-2:-edge - 2:-1]])
    zi = scipy.signal.lfilter_zi(b, a)
    y = lfilter(b, a, ext, zi * ext[0])
    y = lfilter(b, a, y[::-1], zi * y[-1])[::-1]
    return y[edge:-edge].copy()


def remove_drift(signal, fs):
    """read_emg.py:27-29"""
    import scipy.signal
    b, a = scipy.signal.butter(3, 2, 'highpass', fs=fs)
    return filtfilt(b, a, signal)


def notch(signal, freq, sample_frequency):
    """read_emg.py:31-33"""
    import scipy.signal
    b, a = scipy.signal.iirnotch(freq, 30, sample_frequency)
    return filtfilt(b, a, signal)


def notch_harmonics(signal, freq, sample_frequency):
    """read_emg.py:35-38"""
    for harmonic in range(1, 8):
        signal = notch(signal, freq * harmonic, sample_frequency)
    return signal


def subsample(signal, new_freq, old_freq):
    """read_emg.py:40-45"""
    signal = np.ascontiguousarray(signal, dtype=np.float64)
    n = len(signal)
    step = 1 / new_freq
    n_out = max(int(np.ceil(((n - 1) / old_freq) / step)), 0)
    out = np.empty(n_out)
    _lib().ssb_oracle_interp_uniform(_p(signal), _i64(n), _dbl(old_freq), _dbl(step), _i64(n_out), _p(out))
    return out


def apply_to_all(function, signal_array, *args, **kwargs):
    """read_emg.py:47-51"""
    return np.stack([function(signal_array[:, i], *args, **kwargs)
                     for i in range(signal_array.shape[1])], 1)


def condition(before, current, after, rates=(689.06, 516.79)):
    """read_emg.py:62-67: filter the concatenation, cut the current recording out, resample."""
    x = np.concatenate([before, current, after], 0).astype(np.float64)
    x = apply_to_all(notch_harmonics, x, 60, 1000)
    x = apply_to_all(remove_drift, x, 1000)
    x = x[before.shape[0]:x.shape[0] - after.shape[0], :]
    return [apply_to_all(subsample, x, r, 1000) for r in rates]
