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
sys.path.insert(0, ROOT)
from baseline import refenv  # noqa: E402

REF = refenv.reference_dir()      # /root/reference here, baseline/_ref on the GPU box, or None


def test_shims_export_the_reference_names(monkeypatch):
    monkeypatch.setenv("SSB_SYNTHETIC_CORPUS", "1")     # no corpus / reference on this path
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


_need_ref = pytest.mark.skipif(REF is None, reason="no reference checkout / baseline/_ref copy")


def _run(code, **env):
    e = {k: v for k, v in os.environ.items() if k != "SSB_SYNTHETIC_CORPUS"}
    e.update(env)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                          timeout=300, env=e)


@_need_ref
def test_unmodified_reference_script_imports_our_modules():
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
from baseline import refenv
tm = refenv.import_transduction_model_with_dropin(synthetic_corpus=True)
import architecture, align, data_utils, read_emg
for m in (architecture, align, data_utils, read_emg):
    assert "dropin" in m.__file__, m.__file__
assert tm.Model is architecture.Model and tm.align_from_distances is align.align_from_distances
assert "silent_speech_b200" in tm.Model.__module__
assert tm.__file__.startswith({REF!r})
print("OK")
"""
    r = _run(code)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@_need_ref
def test_dataset_side_is_the_references_own_unless_synthetic_is_requested():
    """following INTEGRATION.md must never silently train on random tensors.  Without
    SSB_SYNTHETIC_CORPUS=1 the read_emg shim hands back the reference's own module (real corpus
    loader), and the data_utils shim keeps every non-hot-path name the reference's scripts import
    (read_emg.py:18, evaluate.py:10-15) while the hot-path names are ours."""
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
from baseline import refenv
refenv.install_stubs()
sys.path[:0] = [{os.path.join(ROOT, 'dropin')!r}, {REF!r}]
import read_emg, data_utils
assert "dropin" in read_emg.__file__ and "dropin" in data_utils.__file__
assert read_emg.EMGDataset.__module__ == "_shadowed_read_emg", read_emg.EMGDataset.__module__
assert read_emg.load_utterance.__module__ == "_shadowed_read_emg"
for n in ("get_emg_features", "read_phonemes", "print_confusion", "FeatureNormalizer",
          "TextTransform", "load_audio", "splice_audio", "phoneme_inventory"):
    assert hasattr(data_utils, n), n
from absl import flags
assert "normalizers_file" in flags.FLAGS and "testset_file" in flags.FLAGS
assert "silent_speech_b200" in data_utils.combine_fixed_length.__module__
assert "silent_speech_b200" in data_utils.decollate_tensor.__module__
assert data_utils.mel_spectrogram.__module__ == "data_utils"            # the GPU wrapper
import _shadowed_data_utils as ref_du
assert ref_du.mel_spectrogram is data_utils.mel_spectrogram             # load_audio -> GPU mel
import pickle                                                           # normalizers.pkl contract
with open({os.path.join(REF, 'normalizers.pkl')!r}, "rb") as f:
    mfcc_norm, emg_norm = pickle.load(f)
assert mfcc_norm.feature_means.shape == (1, 80)
print("OK")
"""
    r = _run(code)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_read_emg_shim_refuses_to_guess(tmp_path):
    """No reference read_emg.py on the path and no explicit opt-in: ImportError, not noise."""
    code = f"""
import sys, os
os.chdir({str(tmp_path)!r})
sys.path[:] = [p for p in sys.path if p and os.path.isdir(p)
               and not os.path.exists(os.path.join(p, "read_emg.py"))]
sys.path.insert(0, {os.path.join(ROOT, 'dropin')!r})
try:
    import read_emg
except ImportError as e:
    assert "SSB_SYNTHETIC_CORPUS" in str(e)
    print("OK")
"""
    r = _run(code)
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
