"""CPU: the numpy mel oracle and the package's host-side Slaney filterbank against golden
vectors produced by the executed reference (tests/golden/make_golden_mel.py)."""
import os

import numpy as np
import pytest

from make_golden_mel import CASES, make_clip
from oracle import mel as omel
from silent_speech_b200 import data_utils as du


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "mel_golden.npz"))


def rel_err(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b), np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_mel_matches_reference(golden, i):
    seed, B, S, kind = CASES[i]
    y = make_clip(seed, B, S, kind)
    np.testing.assert_array_equal(y.astype(np.float32), golden[f"y_{i}"])
    got = omel.mel_spectrogram(y)
    want = golden[f"mel_{i}"]
    assert got.shape == want.shape == (B, 80, S // 256)
    l2, mx = rel_err(got, want)
    assert l2 < 2e-5 and mx < 2e-5, (l2, mx)    # fp32 FFT rounding only (tolerance bar: 1e-3)


def test_filterbanks_match_reference_basis(golden):
    basis = golden["basis"]
    assert basis.shape == (80, 513)
    for fb in (omel.mel_filterbank(22050, 1024, 80, 0, 8000),
               du.slaney_mel_filterbank(22050, 1024, 80, 0, 8000)):
        assert fb.dtype == np.float32
        assert np.abs(fb - basis).max() < 2e-7
        # sparsity facts from SURVEY.md §8 a7: 727 non-zeros, highest bin 371
        assert (fb != 0).sum() == 727 and np.nonzero(fb.any(0))[0].max() == 371


def test_batching_glue_roundtrip():
    import torch
    xs = [torch.arange(n * 3, dtype=torch.float32).view(n, 3) for n in (5, 9, 2)]
    c = du.combine_fixed_length(xs, 4)
    assert c.shape == (4, 4, 3) and c.view(-1, 3)[16:].abs().sum() == 0
    back = du.decollate_tensor(c, [5, 9, 2])
    for a, b in zip(xs, back):
        assert torch.equal(a, b)
    assert len(du.phoneme_inventory) == 48 and du.phoneme_inventory[-1] == 'sil'
