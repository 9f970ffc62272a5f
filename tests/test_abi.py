"""CPU: libssb.so builds, loads, and exports every symbol include/ssb.h declares; argument
validation paths that do not touch the GPU behave."""
import os
import re

from silent_speech_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ssb.h")).read()
    return sorted(set(re.findall(r"SSB_API[^;(]*?\b(ssb_\w+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 6
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in ssb.h but not exported by libssb.so"
    # the ctypes table mirrors the header one to one
    assert sorted(_lib.signatures()) == syms


def test_version_and_error_string():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "ssb.h")).read()
    abi = int(re.search(r"#define\s+SSB_ABI_VERSION\s+(\d+)", hdr).group(1))
    assert lib.ssb_version() == abi == _lib.ABI_VERSION       # header, library, binding agree
    assert isinstance(_lib.last_error(), str)


def test_struct_mirrors_match_the_compiled_layouts():
    """a .so built from an older ssb.h must not load silently; the ctypes mirrors of
    the descriptor structs are checked against sizeof() as compiled (also at every load())."""
    import ctypes
    lib = _lib.load()
    for which, cls in enumerate((_lib.Gather, _lib.Scatter, _lib.Epilogue, _lib.TcOperand,
                                 _lib.DtwPair, _lib.Utt, _lib.PrepEntry, _lib.EmgRec)):
        assert lib.ssb_sizeof(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.ssb_sizeof(99) == -1


def test_staleness_is_decided_by_source_content():
    from silent_speech_b200 import build
    assert not build.is_stale()
    h = build.source_hash()
    assert open(build.HASH_PATH).read().strip() == h and len(h) == 64


def test_dtw_workspace_query_and_validation():
    lib = _lib.load()
    # cfg-2 geometry: 600x500 F-ordered
    b = lib.ssb_dtw_workspace_bytes(10000, 600, 500, 1, 600)
    assert b == 10000 * 5 * 34 * 128 * 4
    assert lib.ssb_dtw_workspace_bytes(1, 600, 500, 3, 7) < 0       # no unit stride
    assert "stride" in _lib.last_error()
    assert lib.ssb_dtw_workspace_bytes(1, 0, 5, 1, 5) < 0
