"""Drop-in module `read_emg`.

The dataset (EMGDataset, SizeAwareSampler, file IO, feature extraction) is NOT part of the
accelerated hot path: with dropin/ ahead of the reference checkout on sys.path this shim hands the
reference's OWN read_emg.py back, so `python transduction_model.py` trains on the real corpus (its
`from data_utils import load_audio, ...` then picks up the GPU mel-spectrogram through
dropin/data_utils.py).  The one thing re-pointed inside it is the EMG signal conditioning of
read_emg.py:27-51, see below.

Only when SSB_SYNTHETIC_CORPUS=1 is set — tests and benchmarks on a box without the Zenodo
corpus — does it export the synthetic-utterance mirror `silent_speech_b200.read_emg`, which
honours the same item / collate_raw / sampler contract with seeded random tensors and ignores
the data-directory flags.  It says so on stderr.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(1, _root)

if _os.environ.get("SSB_SYNTHETIC_CORPUS") == "1":
    print("dropin/read_emg: SSB_SYNTHETIC_CORPUS=1 -> SYNTHETIC utterances (random tensors), the "
          "data-directory flags are ignored", file=_sys.stderr)
    from silent_speech_b200 import read_emg as _impl  # noqa: E402
else:
    from _defer import load_shadowed as _load_shadowed  # noqa: E402
    _impl = _load_shadowed("read_emg")
    if _impl is None:
        raise ImportError(
            "dropin/read_emg: the dataset loader is not part of the B200 hot path and no reference "
            "read_emg.py was found further down sys.path.  Run from the reference checkout (or put "
            "it on PYTHONPATH after dropin/), or set SSB_SYNTHETIC_CORPUS=1 to use the synthetic "
            "corpus mirror for tests and benchmarks.")

    # the signal conditioning inside load_utterance (read_emg.py:62-67:
    # notch_harmonics / remove_drift / subsample through apply_to_all) runs on the GPU when one is
    # present -- bit-identical to the scipy / numpy chain; the reference file itself is untouched
    # (its module globals are re-pointed).  SSB_EMG_GPU=0 keeps the reference's CPU functions.
    if _os.environ.get("SSB_EMG_GPU", "1") != "0":
        import torch as _torch
        if _torch.cuda.is_available():
            from silent_speech_b200.emg_signal import patch_reference_module as _patch
            _patch(_impl)

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
