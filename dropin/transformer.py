"""Drop-in module `transformer`: put this directory ahead of the reference checkout on sys.path and the
reference's unmodified transduction_model.py / recognition_model.py import the B200 hot path
(`from transformer import ...`) instead of their own transformer.py.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(1, _root)

from silent_speech_b200.transformer import *  # noqa: F401,F403,E402
from silent_speech_b200 import transformer as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
