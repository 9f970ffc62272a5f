"""GPU parity: csrc/mel.cu vs the numpy oracle and the reference-generated golden vectors.
Tolerance (north star): 1e-3 relative fp32; we assert 1e-4 rel-L2 and 5e-4 max-abs/max-ref."""
import os

import numpy as np
import pytest
import torch

from make_golden_mel import CASES, make_clip
from oracle import mel as omel
from silent_speech_b200 import data_utils as du

pytestmark = pytest.mark.gpu
REL_L2, MAX_REL = 1e-4, 5e-4


def errs(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b), np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "mel_golden.npz"))


@pytest.mark.parametrize("i", range(len(CASES)))
def test_golden(golden, i):
    seed, B, S, kind = CASES[i]
    y = torch.from_numpy(make_clip(seed, B, S, kind)).cuda()
    got = du.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000, center=False)
    assert got.is_cuda and got.shape == (B, 80, S // 256) and got.dtype == torch.float32
    l2, mx = errs(got.cpu().numpy(), golden[f"mel_{i}"])
    assert l2 < REL_L2 and mx < MAX_REL, (l2, mx)


def test_full_size_vs_oracle():
    y = make_clip(1234, 32, 220500, "uniform")          # SURVEY.md §8d mel workload
    got = du.mel_spectrogram(torch.from_numpy(y).cuda(), 1024, 80, 22050, 256, 1024, 0, 8000)
    assert got.shape == (32, 80, 861)
    l2, mx = errs(got[:4].cpu().numpy(), omel.mel_spectrogram(y[:4]))
    assert l2 < REL_L2 and mx < MAX_REL, (l2, mx)
    # batch independence: clip 5 alone == clip 5 in the batch (bitwise)
    alone = du.mel_spectrogram(torch.from_numpy(y[5:6]).cuda(), 1024, 80, 22050, 256, 1024, 0, 8000)
    assert torch.equal(alone[0], got[5])


def test_edge_cases():
    # odd frame count, strided rows, silence (hits the 1e-5 clamp -> log(1e-5))
    y = torch.zeros(2, 256 * 3, device="cuda")
    out = du.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
    assert out.shape == (2, 80, 3)
    # sqrt(1e-9) * sum(filter) is far below the clamp for most filters
    assert torch.allclose(out.min(), torch.log(torch.tensor(1e-5)))
    big = torch.rand(3, 5000, device="cuda") - 0.5
    sub = big[:, :4100]                                   # non-contiguous rows
    a = du.mel_spectrogram(sub, 1024, 80, 22050, 256, 1024, 0, 8000)
    b = du.mel_spectrogram(sub.contiguous(), 1024, 80, 22050, 256, 1024, 0, 8000)
    assert torch.equal(a, b)
    with pytest.raises(Exception):
        du.mel_spectrogram(torch.zeros(1, 4000), 1024, 80, 22050, 256, 1024, 0, 8000)  # CPU tensor
    with pytest.raises(Exception):
        du.mel_spectrogram(y, 512, 80, 22050, 128, 512, 0, 8000)                        # not built


def test_out_of_range_warning_is_reported_without_a_blocking_check(capsys):
    """data_utils.py:40-43 prints 'min value is' / 'max value is' for samples outside [-1, 1]; here
    the kernel tracks the range and the report comes from flush_range_warnings()."""
    y = torch.zeros(2, 4096, device="cuda")
    y[0, 100] = -1.75
    y[1, 3000] = 2.5
    du.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
    du.flush_range_warnings(block=True)
    out = capsys.readouterr().out
    assert "min value is  -1.75" in out and "max value is  2.5" in out
    du.mel_spectrogram(y.clamp(-1, 1), 1024, 80, 22050, 256, 1024, 0, 8000)
    du.flush_range_warnings(block=True)
    assert capsys.readouterr().out == ""
