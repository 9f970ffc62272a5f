"""EMG signal conditioning on the GPU.

Mirrors the signal functions of the reference's read_emg.py with the same names, arguments and
return conventions, so `load_utterance` (read_emg.py:53-100) can run on them unchanged:

    remove_drift(signal, fs)                       read_emg.py:27-29   filtfilt(butter(3, 2, 'highpass'))
    notch(signal, freq, sample_frequency)          read_emg.py:31-33   filtfilt(iirnotch(freq, 30))
    notch_harmonics(signal, freq, sample_frequency) read_emg.py:35-38  harmonics 1..7
    subsample(signal, new_freq, old_freq)          read_emg.py:40-45   np.interp on a uniform grid
    apply_to_all(function, signal_array, *args)    read_emg.py:47-51   all channels at once here

Each accepts what the reference accepts (a 1-D float signal) and ALSO a 2-D (samples, channels)
array, which `apply_to_all` uses to run all channels in one launch.  `condition_utterances` is
the batched form of read_emg.py:62-67 for many recordings (a session directory) per launch.

Arithmetic is float64 in scipy's / numpy's operation order: outputs are bit-identical to the
reference's (tests/test_emg_gpu.py).  Filter COEFFICIENTS are design-time data and come from the
same scipy.signal calls the reference makes (iirnotch, butter, lfilter_zi); the filtering itself
runs in csrc/emg.cu.  There is no CPU fallback: without a CUDA device these functions raise.
"""
import ctypes
import math
from functools import lru_cache

import numpy as np
import torch

from . import _lib

_F64 = torch.float64


@lru_cache(maxsize=None)
def _stage(kind, freq, fs):
    """(b[4], a[4], zi[3], ntaps) of one filtfilt stage, zero-padded."""
    import scipy.signal
    if kind == "notch":
        b, a = scipy.signal.iirnotch(freq, 30, fs)              # read_emg.py:32
    elif kind == "drift":
        b, a = scipy.signal.butter(3, 2, 'highpass', fs=fs)     # read_emg.py:28
    else:
        raise ValueError(kind)
    b = np.asarray(b, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    if a[0] != 1.0:                                             # lfilter normalises by a[0] first
        b, a = b / a[0], a / a[0]
    zi = scipy.signal.lfilter_zi(b, a)
    nt = max(len(a), len(b))
    out = np.zeros(11, dtype=np.float64)
    out[:len(b)] = b
    out[4:4 + len(a)] = a
    out[8:8 + len(zi)] = zi
    return out, nt


def notch_stages(freq, sample_frequency):
    return [_stage("notch", float(freq * h), float(sample_frequency)) for h in range(1, 8)]


def drift_stages(fs):
    return [_stage("drift", 0.0, float(fs))]


def _device():
    if not torch.cuda.is_available():
        raise _lib.SSBError(-3, "EMG conditioning runs in csrc/emg.cu and needs a CUDA device; "
                                "libssb has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _table(entries, dev):
    arr = (_lib.EmgRec * len(entries))(*[_lib.EmgRec(*e) for e in entries])
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(dev), arr


def _stack(signals, dev):
    """list of (n_i, C) arrays / tensors -> one (rows, C) float64 device tensor, offsets, C."""
    C = signals[0].shape[1]
    parts, offs, off = [], [], 0
    for s in signals:
        t = torch.as_tensor(np.ascontiguousarray(s) if isinstance(s, np.ndarray) else s)
        if t.dim() != 2 or t.shape[1] != C:
            raise ValueError("every recording must be (samples, channels) with the same channel count")
        parts.append(t.to(device=dev, dtype=_F64, non_blocking=True))
        offs.append(off)
        off += t.shape[0]
    x = parts[0].contiguous() if len(parts) == 1 else torch.cat(parts, 0)
    return x, offs, C


def filtfilt_cascade(signals, stages):
    """Apply scipy.signal.filtfilt stage after stage to every channel of every recording.
    signals: list of (n_i, C) float arrays (numpy or torch).  -> ((rows, C) float64 device tensor,
    row offsets).  One kernel launch for the whole batch."""
    lib = _lib.load()
    dev = _device()
    x, offs, C = _stack(signals, dev)
    ns = [int(s.shape[0]) for s in signals]
    rows = int(x.shape[0])
    table, _keep = _table([(o, 0, n, 0) for o, n in zip(offs, ns)], dev)
    coef = np.ascontiguousarray(np.stack([c for c, _ in stages]))
    ntaps = np.asarray([nt for _, nt in stages], dtype=np.int32)
    y = torch.empty_like(x)
    ws_bytes = lib.ssb_emg_filtfilt_workspace_bytes(rows, len(ns), C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.ssb_emg_filtfilt_chain(
        x.data_ptr(), y.data_ptr(), table.data_ptr(), len(ns), rows, min(ns), C,
        coef.ctypes.data_as(ctypes.c_void_p), ntaps.ctypes.data_as(ctypes.c_void_p), len(stages),
        ws.data_ptr(), ws_bytes, _lib.current_stream()))
    return y, offs


def resampled_length(n, new_freq, old_freq):
    """len(np.arange(0, (n - 1) / old_freq, 1 / new_freq))   (read_emg.py:41-42)"""
    stop = (n - 1) / old_freq
    return max(int(math.ceil(stop / (1 / new_freq))), 0)


def subsample_rows(x, segments, new_freq, old_freq, out_dtype=torch.float64):
    """np.interp resampling of row segments [(first_row, n)] of a (rows, C) float64 device tensor.
    -> (out (sum n_out, C), [(out_off, n_out)])."""
    lib = _lib.load()
    dev = x.device
    C = x.shape[1]
    entries, places, out_off = [], [], 0
    for off, n in segments:
        n_out = resampled_length(n, new_freq, old_freq)
        entries.append((off, out_off, n, n_out))
        places.append((out_off, n_out))
        out_off += n_out
    out = torch.empty((out_off, C), dtype=out_dtype, device=dev)
    if out_off:
        table, _keep = _table(entries, dev)
        _lib.check(lib.ssb_emg_subsample(x.data_ptr(), table.data_ptr(), len(entries), out_off, C,
                                         float(old_freq), 1 / new_freq, out.data_ptr(),
                                         int(out_dtype == torch.float32), _lib.current_stream()))
    return out, places


# ---- the reference's function names --------------------------------------------------------------
def _as2d(signal):
    a = np.asarray(signal)
    return (a[:, None], True) if a.ndim == 1 else (a, False)


def _run_filter(signal, stages):
    a, squeeze = _as2d(signal)
    y, _ = filtfilt_cascade([a], stages)
    out = y.cpu().numpy()
    return out[:, 0] if squeeze else out


def remove_drift(signal, fs):
    """read_emg.py:27-29"""
    return _run_filter(signal, drift_stages(fs))


def notch(signal, freq, sample_frequency):
    """read_emg.py:31-33"""
    return _run_filter(signal, [_stage("notch", float(freq), float(sample_frequency))])


def notch_harmonics(signal, freq, sample_frequency):
    """read_emg.py:35-38"""
    return _run_filter(signal, notch_stages(freq, sample_frequency))


def subsample(signal, new_freq, old_freq):
    """read_emg.py:40-45"""
    a, squeeze = _as2d(signal)
    dev = _device()
    x = torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=_F64)
    out, _ = subsample_rows(x, [(0, a.shape[0])], new_freq, old_freq)
    out = out.cpu().numpy()
    return out[:, 0] if squeeze else out


_BATCHED = {"remove_drift": remove_drift, "notch": notch, "notch_harmonics": notch_harmonics,
            "subsample": subsample}


def apply_to_all(function, signal_array, *args, **kwargs):
    """read_emg.py:47-51.  The reference loops over channels; the functions above take the whole
    (samples, channels) array in one launch.  Any other callable is applied per channel."""
    batched = _BATCHED.get(getattr(function, "__name__", None))
    if batched is not None and not kwargs:
        return batched(signal_array, *args)
    results = []
    for i in range(signal_array.shape[1]):
        results.append(function(signal_array[:, i], *args, **kwargs))
    return np.stack(results, 1)


def condition_utterances(recordings, rates=(689.06, 516.79), line_freq=60, fs=1000,
                         out_dtype=torch.float64):
    """read_emg.py:62-67 for a batch.  recordings: list of (before, current, after) raw EMG arrays
    (samples, channels); before / after may have 0 rows (read_emg.py:56-61 concatenates the
    neighbouring recordings so the filters settle).  Returns one list per rate in `rates` with the
    resampled current part of every recording, as device tensors (n_out, C): rates[0] is the
    reference's `emg_orig` (before its final float32 cast), rates[1] its `emg`."""
    cat = [np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, np.asarray(c).shape[1]),
                           np.asarray(c, dtype=np.float64),
                           np.asarray(a, dtype=np.float64).reshape(-1, np.asarray(c).shape[1])], 0)
           for b, c, a in recordings]
    y, offs = filtfilt_cascade(cat, notch_stages(line_freq, fs) + drift_stages(fs))
    segs = [(o + np.asarray(b).shape[0], np.asarray(c).shape[0]) for o, (b, c, a) in zip(offs, recordings)]
    outs = []
    for r in rates:
        flat, places = subsample_rows(y, segs, r, fs, out_dtype)
        outs.append([flat[o:o + n] for o, n in places])
    return outs


def patch_reference_module(mod):
    """Point a loaded reference `read_emg` module at the functions above (dropin/read_emg.py does
    this when a CUDA device is present): `load_utterance` looks the names up in its module
    globals at call time, so the reference file itself stays untouched."""
    for name in ("remove_drift", "notch", "notch_harmonics", "subsample", "apply_to_all"):
        if hasattr(mod, name):
            setattr(mod, "_cpu_" + name, getattr(mod, name))
            setattr(mod, name, globals()[name])
    return mod
