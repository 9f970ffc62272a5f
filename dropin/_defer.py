"""Helper of the drop-in shims: load "the module this shim shadows" — the next file of the same
name further down sys.path (the reference checkout the user runs from) — under an alias, so a
shim can re-export the reference's non-hot-path symbols untouched and override only the hot path."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def find_shadowed(name):
    """Path of the first <name>.py on sys.path (or the CWD) outside dropin/, else None."""
    seen = set()
    for p in list(sys.path) + [os.getcwd()]:
        d = os.path.abspath(p or os.getcwd())
        if d in seen or d == _HERE:
            continue
        seen.add(d)
        f = os.path.join(d, name + ".py")
        if os.path.isfile(f) and "silent_speech_b200" not in d:
            return f
    return None


def load_shadowed(name):
    """Execute the shadowed module as `_shadowed_<name>` and return it (None if there is none)."""
    alias = "_shadowed_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    f = find_shadowed(name)
    if f is None:
        return None
    spec = importlib.util.spec_from_file_location(alias, f)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(alias, None)
        raise
    return mod
