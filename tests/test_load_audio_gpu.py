"""GPU: `load_audio` (data_utils.py:64-83), the caller of the mel kernel at dataset-load time
(read_emg.py:80), both as the package exports it and as the REFERENCE's own load_audio running
through dropin/data_utils.py (its call to mel_spectrogram lands on csrc/mel.cu).  soundfile is
absent from this image, so `sf.read` is a fake returning seeded audio; the arithmetic after it
is what is compared: channel select, slicing, clipping, log-mel, transpose, max_frames."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from oracle import mel as omel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_audio(seed=0, n=30000, stereo=True, amp=0.7):
    rs = np.random.RandomState(seed)
    a = (rs.rand(n, 2 if stereo else 1) * 2 - 1) * amp
    a[100:110] = 1.5          # out-of-range samples: load_audio clips to [-1, 1] (:77)
    return a if stereo else a[:, 0]


def errs(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b), np.abs(a - b).max() / np.abs(b).max()


def test_load_audio_matches_oracle(monkeypatch):
    from silent_speech_b200 import data_utils as du
    audio = fake_audio()
    sf = types.ModuleType("soundfile")
    sf.read = lambda fn: (audio.copy(), 22050)
    monkeypatch.setitem(sys.modules, "soundfile", sf)
    want = omel.mel_spectrogram(np.clip(audio[:, 0], -1, 1)[None].astype(np.float32))[0].T
    got = du.load_audio("x.flac")
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == want.shape
    l2, mx = errs(got, want)
    assert l2 < 1e-4 and mx < 5e-4, (l2, mx)
    # start / end slicing and max_frames truncation (:70-71, :81-82)
    got2 = du.load_audio("x.flac", start=1000, end=21000, max_frames=50)
    want2 = omel.mel_spectrogram(np.clip(audio[1000:21000, 0], -1, 1)[None].astype(np.float32))[0].T[:50]
    assert got2.shape == (50, 80) and errs(got2, want2)[0] < 1e-4
    # volume renormalisation (:19-27, :73-74): rms target 0.2 over centred 2048/512 frames
    got3 = du.load_audio("x.flac", renormalize_volume=True)
    a = audio[:, 0]
    p = np.pad(a, 1024)
    rms = max(np.sqrt(np.mean(p[i * 512:i * 512 + 2048] ** 2)) for i in range(1 + (len(p) - 2048) // 512))
    a3 = a * (0.2 / (rms + 0.01))
    if np.abs(a3).max() > 1.0:
        a3 = a3 / np.abs(a3).max()
    want3 = omel.mel_spectrogram(np.clip(a3, -1, 1)[None].astype(np.float32))[0].T
    assert errs(got3, want3)[0] < 1e-4
    sf.read = lambda fn: (audio.copy(), 44100)
    with pytest.raises(AssertionError):                       # :78 `assert r == 22050`
        du.load_audio("x.flac")


def test_reference_load_audio_through_the_dropin_uses_the_gpu_mel():
    sys.path.insert(0, ROOT)
    from baseline import refenv
    ref = refenv.reference_dir()
    if ref is None:
        pytest.skip("no reference copy")
    code = f"""
import sys, types
import numpy as np
sys.path.insert(0, {ROOT!r})
from baseline import refenv
refenv.install_stubs()
rs = np.random.RandomState(3)
audio = (rs.rand(25000) * 2 - 1) * 0.6
sys.modules["soundfile"].read = lambda fn: (audio.copy(), 22050)
sys.path[:0] = [{os.path.join(ROOT, 'dropin')!r}, {ref!r}]
import data_utils
assert data_utils.load_audio.__module__ == "_shadowed_data_utils"     # the reference's function
from silent_speech_b200 import _lib
n0 = _lib.launch_count
got = data_utils.load_audio("clip.flac")
assert _lib.launch_count == n0 + 1, "mel kernel did not run"
from oracle import mel as omel
want = omel.mel_spectrogram(audio[None].astype(np.float32))[0].T
assert got.shape == want.shape
print("ERR", np.linalg.norm(got - want) / np.linalg.norm(want))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    err = float([l for l in r.stdout.splitlines() if l.startswith("ERR")][-1].split()[1])
    assert err < 1e-4, err
