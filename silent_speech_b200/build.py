"""In-tree build of libssb.so (the C-ABI CUDA library) with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ behind `include/ssb.h`.
`python -m silent_speech_b200.build` rebuilds what is stale; `build(force=True)` rebuilds all.
The .so stays in the package directory (git-ignored, but it travels to the GPU box).
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_PATH = os.path.join(PKG_DIR, "libssb.so")
HASH_PATH = LIB_PATH + ".srchash"      # sha256 of the sources the in-tree .so was built from

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-warn-spills",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libssb.so cannot be built")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "ssb.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile_one(src, obj, verbose):
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout.strip() or r.stderr.strip()):
        print(r.stdout, r.stderr, file=sys.stderr)


def source_hash():
    """sha256 over csrc/* and include/ssb.h (content, not mtimes: the snapshot that ships the tree
    to the GPU box does not preserve a usable mtime order)."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                   if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(ROOT, "include", "ssb.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale():
    """True when libssb.so is missing or was built from different sources than the tree holds."""
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(HASH_PATH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libssb.so. Returns the library path.  Serialised across processes by
    a file lock (torchrun ranks may all find the library stale at once)."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB_PATH          # another process built it while we waited
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    hdr_t = _deps_mtime()
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src)
                or os.path.getmtime(obj) < hdr_t):
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile_one(j[0], j[1], verbose), jobs))
    if jobs or force or not os.path.exists(LIB_PATH):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode",
               "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-cudart", "static"]
        tmp = LIB_PATH + ".tmp"
        cmd[cmd.index(LIB_PATH)] = tmp
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB_PATH)        # atomic: a concurrent loader never maps a half-written .so
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
