"""Import the UNMODIFIED reference (dgaddy/silent_speech at /root/reference) in this
container so that golden vectors can be generated from the real code.

Only used by tests/golden/make_golden*.py (run by hand here; /root/reference does not exist
on the GPU box).  Missing third-party modules that the hot path never touches are stubbed
(SURVEY.md §8c); `librosa.filters.mel` is provided by torchaudio's Slaney filterbank, which
SURVEY.md §8c verified equal to librosa's to 6.8e-8.
"""
import os
import sys
import types
from unittest import mock

REF = os.environ.get("SSB_REFERENCE", "/root/reference")


def _stub(name):
    m = mock.MagicMock(name=name)
    m.__name__ = name
    m.__path__ = []
    sys.modules[name] = m
    return m


def install_stubs():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "soundfile", "textgrids",
                 "jiwer", "unidecode", "deepspeech", "librosa.util"]:
        if name not in sys.modules:
            _stub(name)
    if "librosa" not in sys.modules or isinstance(sys.modules["librosa"], mock.MagicMock):
        import torchaudio

        librosa = types.ModuleType("librosa")
        filters = types.ModuleType("librosa.filters")

        def mel(sr, n_fft, n_mels, fmin, fmax):
            fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax),
                                                       n_mels, sr, norm="slaney",
                                                       mel_scale="slaney")
            return fb.T.contiguous().numpy()

        filters.mel = mel
        librosa.filters = filters
        librosa.util = sys.modules.get("librosa.util") or _stub("librosa.util")
        sys.modules["librosa"] = librosa
        sys.modules["librosa.filters"] = filters


def import_reference(*names):
    """import_reference('align', 'architecture') -> list of reference modules."""
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference checkout not found at {REF}")
    install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    mods = []
    for n in names:
        if n in sys.modules and not getattr(sys.modules[n], "__file__", "").startswith(REF):
            del sys.modules[n]
        mods.append(__import__(n))
    return mods


def fix_transformer_shim(model):
    """torch >= 2.1 reads layers[0].self_attn.batch_first."""
    for layer in model.transformer.layers:
        layer.self_attn.batch_first = False
    return model
