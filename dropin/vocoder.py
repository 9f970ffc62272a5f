"""Drop-in module `vocoder`: with this directory ahead of the reference checkout on sys.path the
reference's unmodified transduction_model.py / evaluate.py (`from vocoder import Vocoder`,
transduction_model.py:18,176; evaluate.py:15,59) synthesise audio with the HiFi-GAN generator running
on libssb (silent_speech_b200/vocoder.py) instead of hifi_gan/models.py on stock PyTorch.  Same flag
(--hifigan_checkpoint), same checkpoint and config.json, same call: Vocoder()(mel (seq_len, 80)) ->
1-D audio tensor.  SSB_VOCODER=ref hands the reference's own vocoder.py back.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(1, _root)

if _os.environ.get("SSB_VOCODER", "b200") == "ref":
    from _defer import load_shadowed as _load_shadowed  # noqa: E402
    _impl = _load_shadowed("vocoder")
    if _impl is None:
        raise ImportError("dropin/vocoder: SSB_VOCODER=ref but no reference vocoder.py further down sys.path")
    globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
else:
    from absl import flags as _flags  # noqa: E402

    FLAGS = _flags.FLAGS
    if "hifigan_checkpoint" not in FLAGS:
        # vocoder.py:12 of the reference
        _flags.DEFINE_string('hifigan_checkpoint', None, 'filename of hifi-gan generator checkpoint')

    from silent_speech_b200.vocoder import AttrDict, Generator, Vocoder  # noqa: F401,E402
