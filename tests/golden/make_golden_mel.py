"""Generate tests/golden/mel_golden.npz by running the reference's data_utils.mel_spectrogram
(torch.stft path, data_utils.py:39-62) in this container.  librosa.filters.mel is supplied by
torchaudio's Slaney filterbank (see _reference_import.py).  Run: python tests/golden/make_golden_mel.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _reference_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mel_golden.npz")


def make_clip(seed, B, S, kind="uniform"):
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":                       # SURVEY.md §8d synthetic mel input
        return ((torch.rand(B, S, generator=g) * 2 - 1) * 0.5).numpy()
    t = torch.arange(S, dtype=torch.float64) / 22050.0   # tones + noise: non-flat spectrum
    y = 0.4 * torch.sin(2 * np.pi * 220.0 * t) + 0.2 * torch.sin(2 * np.pi * 3000.0 * t)
    y = y[None].repeat(B, 1) + 0.01 * torch.randn(B, S, generator=g, dtype=torch.float64)
    return y.to(torch.float32).clamp(-1, 1).numpy()


CASES = [(1234, 2, 4000, "uniform"), (7, 1, 22050, "tones"), (8, 3, 1000, "uniform"),
         (9, 1, 256 * 5 + 17, "tones"), (10, 2, 385, "uniform")]


def main():
    (du,) = import_reference("data_utils")
    out = {"meta": np.array([f"{s},{B},{S},{k}" for s, B, S, k in CASES])}
    for i, (seed, B, S, kind) in enumerate(CASES):
        y = make_clip(seed, B, S, kind)
        ref = du.mel_spectrogram(torch.from_numpy(y), 1024, 80, 22050, 256, 1024, 0, 8000,
                                 center=False).numpy()
        out[f"mel_{i}"] = ref.astype(np.float32)
        out[f"y_{i}"] = y.astype(np.float32)
    # the filterbank the reference used (torchaudio == librosa to 6.8e-8)
    import librosa
    out["basis"] = librosa.filters.mel(sr=22050, n_fft=1024, n_mels=80, fmin=0, fmax=8000)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))

    # full-size check of the oracle restatement against the reference (not stored)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import mel as omel
    y = make_clip(1234, 4, 220500, "uniform")
    ref = du.mel_spectrogram(torch.from_numpy(y), 1024, 80, 22050, 256, 1024, 0, 8000).numpy()
    got = omel.mel_spectrogram(y)
    print("full-size oracle vs reference: shape", got.shape, "max abs", np.abs(got - ref).max(),
          "rel-L2", np.linalg.norm(got - ref) / np.linalg.norm(ref))


if __name__ == "__main__":
    main()
