"""CPU: the drop-in boundary.  The dropin/ shims expose exactly the names the reference's
scripts import (transduction_model.py:13-18, recognition_model.py:14-16), and — when the
reference checkout is present (this container only) — the UNMODIFIED transduction_model.py
resolves its hot-path imports to our modules."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_shims_export_the_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for mod, names in {
            "architecture": ["Model", "ResBlock", "FLAGS"],
            "transformer": ["TransformerEncoderLayer", "MultiHeadAttention",
                            "LearnedRelativePositionalEmbedding"],
            "align": ["align_from_distances", "time_warp"],
            "data_utils": ["phoneme_inventory", "decollate_tensor", "combine_fixed_length",
                           "mel_spectrogram", "FeatureNormalizer", "load_audio", "TextTransform"],
            "read_emg": ["EMGDataset", "SizeAwareSampler"],
        }.items():
            sys.modules.pop(mod, None)
            m = importlib.import_module(mod)
            assert os.path.dirname(m.__file__).endswith("dropin"), m.__file__
            for n in names:
                assert hasattr(m, n), f"{mod}.{n}"
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for mod in ("architecture", "transformer", "align", "data_utils", "read_emg"):
            sys.modules.pop(mod, None)


def test_model_interface_and_flags():
    from absl import flags
    from silent_speech_b200 import architecture as A
    for f in ("model_size", "num_layers", "dropout"):
        assert f in flags.FLAGS
    F = flags.FLAGS
    if not F.is_parsed():
        F(["t"])
    F.model_size, F.num_layers = 32, 1
    m = A.Model(112, 80)                      # num_aux_outs optional, as in the reference
    assert not m.has_aux_out and not hasattr(m, "w_aux")
    m2 = A.Model(112, 80, 48)
    sd = m2.state_dict()
    assert sd["transformer.layers.0.self_attn.w_q"].shape == (8, 32, 4)
    assert sd["transformer.layers.0.self_attn.relative_positional.embeddings"].shape == (8, 199, 4, 1)
    assert sd["conv_blocks.0.conv1.weight"].shape == (32, 8, 3)
    import torch
    with pytest.raises(Exception):            # no CPU path: fails loudly, never falls back
        m2(None, torch.zeros(1, 64, 8), None)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_unmodified_reference_script_imports_our_modules():
    code = f"""
import sys, types
from unittest import mock
for n in ["matplotlib","matplotlib.pyplot","matplotlib.pylab","soundfile","textgrids","jiwer","unidecode","deepspeech","librosa","librosa.util","librosa.filters"]:
    if n not in sys.modules:
        try: __import__(n)
        except Exception:
            m = mock.MagicMock(); m.__path__ = []; sys.modules[n] = m
import os
os.chdir({REF!r})
sys.path[:0] = [{os.path.join(ROOT, 'dropin')!r}, {REF!r}, {os.path.join(REF, 'hifi_gan')!r}]
import transduction_model as tm
import architecture, align, data_utils, read_emg
for m in (architecture, align, data_utils, read_emg):
    assert "dropin" in m.__file__, m.__file__
assert tm.Model is architecture.Model and tm.align_from_distances is align.align_from_distances
assert tm.__file__.startswith({REF!r})
print("OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_opt_in_pieces_have_no_cpu_path():
    """§8 f2 / f3 entry points follow the same rule as the rest of the package: CPU tensors raise,
    nothing falls back to torch."""
    import pytest
    import torch
    from silent_speech_b200 import _lib
    from silent_speech_b200.losses import ctc_loss
    from silent_speech_b200.optim import FlatAdamW
    from silent_speech_b200.training import GradientBucket
    with pytest.raises(_lib.SSBError):
        ctc_loss(torch.randn(2, 10, 5, requires_grad=True), torch.zeros(2, 3, dtype=torch.int64),
                 [10, 10], [3, 2], blank=4)
    m = torch.nn.Linear(4, 4)
    with pytest.raises(TypeError):
        FlatAdamW(GradientBucket(m))
