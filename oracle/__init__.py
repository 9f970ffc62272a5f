"""ORACLE — test infrastructure, not product code.

CPU restatements of the reference's hot path (dgaddy/silent_speech), each citing the
reference file:line it follows:
  oracle/dtw_oracle.c   align.py:5-34            (plain C, via ctypes: oracle.dtw)
  oracle/mel.py         data_utils.py:29-62      (numpy)
  oracle/model.py       architecture.py, transformer.py  (torch fp32, CPU)
  oracle/step.py        transduction_model.py:98-157,196-212 (loss + train step, CPU)
  oracle/ctc.py         recognition_model.py:96-101      (numpy fp64 log-softmax + CTC)
  oracle/emg.py + emg_oracle.c   read_emg.py:27-51        (filtfilt cascade + np.interp, float64)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package.  silent_speech_b200/ never does.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "dtw_oracle.c")
_SRCS = [_SRC, os.path.join(_HERE, "emg_oracle.c")]
_lib = None


def build(force=False):
    """gcc -O2 (no fast-math) -> oracle/liboracle.so"""
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in _SRCS)):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB_PATH,
           *_SRCS, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib
