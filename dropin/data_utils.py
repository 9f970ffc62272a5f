"""Drop-in module `data_utils`.

Hot-path symbols come from `silent_speech_b200.data_utils` (GPU log-mel csrc/mel.cu,
combine_fixed_length / decollate_tensor); everything else the reference's scripts import from
data_utils (get_emg_features, read_phonemes, print_confusion, splice_audio, the
`normalizers_file` flag, ...) is handed back from the reference's OWN data_utils.py, found
further down sys.path, so read_emg.py / evaluate.py keep importing what they expect.  The
reference's `load_audio` (data_utils.py:64-83) is kept as is — file IO, resampling and volume
normalisation are dataset preparation — but its call to `mel_spectrogram` lands on the GPU
kernel: host tensors are copied to the device, computed there and copied back to the caller's
device (there is no CPU arithmetic path).  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(1, _root)

from silent_speech_b200 import data_utils as _impl  # noqa: E402
from _defer import load_shadowed as _load_shadowed  # noqa: E402

try:
    _ref = _load_shadowed("data_utils")
except ImportError as _e:     # the reference's optional third-party imports (librosa, ...) missing
    print(f"dropin/data_utils: reference data_utils.py not importable ({_e}); exporting the "
          "hot-path symbols only", file=_sys.stderr)
    _ref = None

if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                    center=False):
    """data_utils.py:39-62 on csrc/mel.cu; result on y.device like the reference's.  A host tensor
    (what load_audio passes, data_utils.py:79) makes one H2D / D2H round trip."""
    if y.is_cuda:
        return _impl.mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin,
                                     fmax, center)
    return _impl.mel_spectrogram(y.cuda(), n_fft, num_mels, sampling_rate, hop_size, win_size,
                                 fmin, fmax, center).cpu()


_HOT = {"mel_spectrogram": mel_spectrogram,
        "dynamic_range_compression_torch": _impl.dynamic_range_compression_torch,
        "spectral_normalize_torch": _impl.spectral_normalize_torch,
        "combine_fixed_length": _impl.combine_fixed_length,
        "decollate_tensor": _impl.decollate_tensor,
        "phoneme_inventory": _impl.phoneme_inventory}
if _ref is None:
    globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
globals().update(_HOT)
if _ref is not None:
    # the reference's own functions (load_audio) call these through their module globals
    for _k in ("mel_spectrogram", "combine_fixed_length", "decollate_tensor"):
        setattr(_ref, _k, _HOT[_k])
