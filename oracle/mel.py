"""ORACLE — test infrastructure, not product code.

numpy restatement of the reference's log-mel extraction, data_utils.py:39-62 (+ :29-34),
and of librosa.filters.mel (third-party, absent from the reference checkout, unpinned in
environment.yml:17; call site data_utils.py:47) from its published Slaney definition.
Pinned by tests/golden/mel_golden.npz, produced by the executed reference
(tests/golden/make_golden_mel.py).
"""
import numpy as np


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    out = f / (200.0 / 3.0)
    big = f >= 1000.0
    out = np.where(big, 15.0 + np.log(np.where(big, f, 1000.0) / 1000.0) / (np.log(6.4) / 27.0), out)
    return out


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    out = m * (200.0 / 3.0)
    big = m >= 15.0
    return np.where(big, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), out)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=) defaults: Slaney scale + norm."""
    freqs = np.arange(n_fft // 2 + 1, dtype=np.float64) * (sr / float(n_fft))
    edges = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fb = np.zeros((n_mels, freqs.size))
    for m in range(n_mels):
        lo, ce, hi = edges[m], edges[m + 1], edges[m + 2]
        up = (freqs - lo) / (ce - lo)
        down = (hi - freqs) / (hi - ce)
        fb[m] = np.clip(np.minimum(up, down), 0.0, None) * (2.0 / (hi - lo))
    return fb.astype(np.float32)


def mel_spectrogram(y, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
                    win_size=1024, fmin=0, fmax=8000):
    """data_utils.py:39-62 with center=False.  y: (B, S) float32 -> (B, num_mels, frames)."""
    y = np.asarray(y, dtype=np.float32)
    assert y.ndim == 2 and win_size == n_fft
    pad = int((n_fft - hop_size) / 2)                                  # :51
    yp = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")
    frames = 1 + (yp.shape[1] - n_fft) // hop_size                     # torch.stft, center=False
    n = np.arange(win_size, dtype=np.float64)
    window = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_size)).astype(np.float32)  # periodic Hann
    idx = np.arange(frames)[:, None] * hop_size + np.arange(n_fft)[None, :]
    seg = yp[:, idx] * window[None, None, :]                           # (B, frames, n_fft) fp32
    spec = np.fft.rfft(seg.astype(np.float32), axis=-1)                # (B, frames, bins)
    mag = np.sqrt(spec.real.astype(np.float32) ** 2 + spec.imag.astype(np.float32) ** 2
                  + np.float32(1e-9))                                  # :57
    fb = mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)    # :47
    mel = np.einsum("mk,bfk->bmf", fb, mag.astype(np.float32)).astype(np.float32)  # :59
    return np.log(np.maximum(mel, np.float32(1e-5))).astype(np.float32)  # :60, :29-30
