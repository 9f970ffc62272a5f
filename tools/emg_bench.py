"""Timing of the EMG conditioning chain (read_emg.py:62-67) on the GPU vs the reference formulation
on the host cores.  Usage: python tools/emg_bench.py [n_recordings] [samples]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import emg_signal as es  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12000      # before + current + after at 1 kHz
    rs = np.random.RandomState(0)
    recs = [(np.zeros((0, 8)), 50.0 * rs.randn(n, 8), np.zeros((0, 8))) for _ in range(R)]
    es.condition_utterances(recs[:2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = es.condition_utterances(recs)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    # kernels only (inputs resident)
    cat = [np.concatenate(r, 0) for r in recs]
    x, offs, C = es._stack(cat, torch.device("cuda"))
    st = es.notch_stages(60, 1000) + es.drift_stages(1000)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    y, _ = es.filtfilt_cascade([x[o:o + n] for o in offs], st)
    ev[1].record()
    es.subsample_rows(y, [(o, n) for o in offs], 689.06, 1000)
    ev[2].record()
    torch.cuda.synchronize()
    # one recording alone (the per-item DataLoader call of the reference)
    t0 = time.perf_counter()
    es.condition_utterances(recs[:1])
    torch.cuda.synchronize()
    t_one = time.perf_counter() - t0
    # the reference formulation on the host (scipy), 4 recordings
    import scipy.signal

    def ref_chain(full):
        x = full
        for c in range(1):
            pass
        cols = []
        for i in range(full.shape[1]):
            s = full[:, i]
            for h in range(1, 8):
                b, a = scipy.signal.iirnotch(60 * h, 30, 1000)
                s = scipy.signal.filtfilt(b, a, s)
            b, a = scipy.signal.butter(3, 2, 'highpass', fs=1000)
            s = scipy.signal.filtfilt(b, a, s)
            times = np.arange(len(s)) / 1000
            cols.append(np.interp(np.arange(0, times[-1], 1 / 689.06), times, s))
        return np.stack(cols, 1)
    t0 = time.perf_counter()
    for r in recs[:4]:
        ref_chain(r[1])
    t_ref = (time.perf_counter() - t0) / 4
    print(json.dumps({"recordings": R, "samples": n, "channels": 8,
                      "gpu_batch_s": t_all, "gpu_per_recording_ms": 1e3 * t_all / R,
                      "gpu_filtfilt_kernel_ms": ev[0].elapsed_time(ev[1]),
                      "gpu_subsample_kernel_ms": ev[1].elapsed_time(ev[2]),
                      "gpu_single_recording_ms": 1e3 * t_one,
                      "cpu_scipy_per_recording_ms": 1e3 * t_ref,
                      "checksum": float(out[0][0].sum().item())}))


if __name__ == "__main__":
    main()
