"""Import the UNMODIFIED reference (dgaddy/silent_speech at /root/reference) in this
container so that golden vectors can be generated from the real code.

Only used by tests/golden/make_golden*.py (run by hand here).  The environment handling
(stubs for absent third-party modules, the Slaney `librosa.filters.mel`, the torch >= 2.1
`batch_first` shim) lives in baseline/refenv.py, shared with the GPU
integration test and `bench.py --impl reference`.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from baseline.refenv import (fix_transformer_shim, import_reference, install_stubs,  # noqa: E402,F401
                             reference_dir)

REF = reference_dir()
