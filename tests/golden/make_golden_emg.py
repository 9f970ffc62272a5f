"""Generate tests/golden/emg_golden.npz by running the reference's OWN read_emg.py functions
(remove_drift, notch_harmonics, subsample, apply_to_all: read_emg.py:27-51, i.e. scipy.signal.filtfilt
and np.interp as the reference calls them) in this container, on synthetic EMG-like recordings.
Run: python tests/golden/make_golden_emg.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _reference_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emg_golden.npz")


def make_recording(seed, n, C=8):
    """Raw-EMG-like float64: broadband noise + 60 Hz mains with harmonics + slow drift, O(100) scale."""
    rs = np.random.RandomState(seed)
    t = np.arange(n) / 1000.0
    x = 40.0 * rs.randn(n, C)
    for h in (1, 2, 3, 5):
        x += (30.0 / h) * np.sin(2 * np.pi * 60.0 * h * t[:, None] + rs.rand(1, C) * 6.28)
    x += 200.0 * np.sin(2 * np.pi * 0.2 * t[:, None] + rs.rand(1, C)) + 500.0 * rs.rand(1, C)
    return x


# (seed, n_before, n, n_after): short / long, with and without neighbours (read_emg.py:56-61)
CASES = [(1, 0, 700, 0), (2, 300, 1200, 250), (3, 0, 50, 40), (4, 120, 2001, 0)]


def main():
    (re_,) = import_reference("read_emg")
    out = {"meta": np.array([",".join(map(str, c)) for c in CASES])}
    for i, (seed, nb, n, na) in enumerate(CASES):
        full = make_recording(seed, nb + n + na)
        # read_emg.py:62-67 verbatim
        x = re_.apply_to_all(re_.notch_harmonics, full, 60, 1000)
        x = re_.apply_to_all(re_.remove_drift, x, 1000)
        filtered = x.copy()
        x = x[nb:x.shape[0] - na, :]
        emg_orig = re_.apply_to_all(re_.subsample, x, 689.06, 1000)
        emg = re_.apply_to_all(re_.subsample, x, 516.79, 1000)
        # the raw input is regenerated from the seed by the tests (make_recording above)
        if i in (0, 2):
            out[f"filtered_{i}"] = filtered
        out[f"orig_{i}"] = emg_orig
        out[f"emg_{i}"] = emg
    # single functions on one channel
    sig = make_recording(9, 900, 1)[:, 0]
    out["sig_notch"] = re_.notch(sig, 180, 1000)
    out["sig_drift"] = re_.remove_drift(sig, 1000)
    out["sig_sub"] = re_.subsample(sig, 689.06, 1000)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k != "meta"})


if __name__ == "__main__":
    main()
