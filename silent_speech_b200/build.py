"""In-tree build of libssb.so (the C-ABI CUDA library) with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ behind `include/ssb.h`.
`python -m silent_speech_b200.build` rebuilds what is stale; `build(force=True)` rebuilds all.
The .so stays in the package directory (git-ignored, but it travels to the GPU box).
"""
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


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources()) or _deps_mtime() > t


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libssb.so. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
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
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
