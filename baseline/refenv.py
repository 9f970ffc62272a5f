"""Import the UNMODIFIED reference (dgaddy/silent_speech) for the checks and the reference arm.

Test / benchmark infrastructure, not product code: only tests/, tests/golden/make_golden*.py
and `bench.py --impl reference` import this module; silent_speech_b200/ never does.

The reference is found at /root/reference (the build container) or at baseline/_ref (the copy
`baseline/install_ref.py` makes, which travels to the GPU box).  Nothing here edits it; what
this module provides is the ENVIRONMENT the reference needs under torch 2.11 without its
optional third-party packages:
  * stub modules for matplotlib, soundfile, textgrids, jiwer, unidecode, deepspeech — imported
    at module top by the reference but never touched by the hot path;
  * `librosa.filters.mel` supplied by the Slaney filterbank (torchaudio's, equal to librosa's
    to 6.8e-8) because librosa is absent;
  * the one-attribute shim `self_attn.batch_first = False` that torch >= 2.1's
    nn.TransformerEncoder reads from the reference's custom attention module.
"""
import contextlib
import importlib
import importlib.util
import os
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_HOT = ("architecture", "transformer", "align", "data_utils", "read_emg")


def reference_dir():
    """Directory holding the reference's *.py, or None."""
    for d in (os.environ.get("SSB_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if d and os.path.isfile(os.path.join(d, "transduction_model.py")):
            return d
    return None


def _stub(name):
    m = mock.MagicMock(name=name)
    m.__name__ = name
    m.__path__ = []
    m.__spec__ = None
    sys.modules[name] = m
    return m


def slaney_mel(sr, n_fft, n_mels, fmin, fmax):
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr,
                                               norm="slaney", mel_scale="slaney")
    return fb.T.contiguous().numpy()


def install_stubs():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "soundfile", "textgrids",
                 "jiwer", "unidecode", "deepspeech", "librosa.util"]:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "librosa" not in sys.modules or isinstance(sys.modules["librosa"], mock.MagicMock):
        try:
            importlib.import_module("librosa")
            return
        except Exception:
            pass
        librosa = types.ModuleType("librosa")
        filters = types.ModuleType("librosa.filters")
        filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: slaney_mel(sr, n_fft, n_mels, fmin, fmax)
        librosa.filters = filters
        librosa.util = sys.modules.get("librosa.util") or _stub("librosa.util")
        sys.modules["librosa"] = librosa
        sys.modules["librosa.filters"] = filters


@contextlib.contextmanager
def tolerant_flags():
    """absl raises DuplicateFlagError when the reference's module and its drop-in twin define the
    same flag in one process; inside this context a second definition keeps the first."""
    from absl import flags
    saved = {}
    for n in ("DEFINE_integer", "DEFINE_float", "DEFINE_string", "DEFINE_list", "DEFINE_boolean",
              "DEFINE_bool"):
        fn = getattr(flags, n)
        saved[n] = fn

        def wrap(*a, _fn=fn, **k):
            try:
                return _fn(*a, **k)
            except flags.DuplicateFlagError:
                return None
        setattr(flags, n, wrap)
    try:
        yield
    finally:
        for n, fn in saved.items():
            setattr(flags, n, fn)


def _load_as(ref, name, alias, inject=None):
    """Execute <ref>/<name>.py as module `alias`; `inject` temporarily maps import names to
    modules while it executes (e.g. 'transformer' -> the reference's own transformer)."""
    spec = importlib.util.spec_from_file_location(alias, os.path.join(ref, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    saved = {}
    for k, v in (inject or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    sys.modules[alias] = mod
    try:
        with tolerant_flags():
            spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_reference_model_modules():
    """The reference's own transformer.py / architecture.py under the aliases `ref_transformer`,
    `ref_architecture`, importable next to the drop-in modules of the same base names."""
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference not found (neither /root/reference nor baseline/_ref)")
    install_stubs()
    if "ref_architecture" in sys.modules:
        return sys.modules["ref_architecture"], sys.modules["ref_transformer"]
    rt = _load_as(ref, "transformer", "ref_transformer")
    ra = _load_as(ref, "architecture", "ref_architecture", inject={"transformer": rt})
    return ra, rt


def fix_transformer_shim(model):
    """torch >= 2.1 reads layers[0].self_attn.batch_first."""
    for layer in model.transformer.layers:
        layer.self_attn.batch_first = False
    return model


def import_reference(*names):
    """Plain import of reference modules BY THEIR OWN NAMES from the reference directory
    (for a process that does not load the drop-in).  Returns the list of modules."""
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference not found (neither /root/reference nor baseline/_ref)")
    install_stubs()
    for p in (os.path.join(ref, "hifi_gan"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    mods = []
    for n in names:
        cur = sys.modules.get(n)
        if cur is not None and not (getattr(cur, "__file__", "") or "").startswith(ref):
            del sys.modules[n]
        with tolerant_flags():
            mods.append(importlib.import_module(n))
    return mods


def import_transduction_model_with_dropin(synthetic_corpus=True):
    """`import transduction_model` — the reference's file, untouched — with dropin/ ahead of the
    reference on sys.path, so its `from architecture import Model`, `from align import ...`,
    `from data_utils import ...` (and, for tests without a corpus, `from read_emg import ...`)
    resolve to the B200 hot path.  Returns the module."""
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference not found (neither /root/reference nor baseline/_ref)")
    install_stubs()
    if synthetic_corpus:
        os.environ["SSB_SYNTHETIC_CORPUS"] = "1"
    dropin = os.path.join(ROOT, "dropin")
    for n in _HOT + ("transduction_model",):
        sys.modules.pop(n, None)
    for p in (os.path.join(ref, "hifi_gan"), ref, dropin):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    with tolerant_flags():
        tm = importlib.import_module("transduction_model")
    assert (tm.__file__ or "").startswith(ref), tm.__file__
    return tm
