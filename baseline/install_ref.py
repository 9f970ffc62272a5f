"""Recipe that makes the UNMODIFIED reference (dgaddy/silent_speech) travel to the GPU box.

The reference is ~1.6 kLoC of plain Python with no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build ("neither setup.py nor
pyproject.toml found").  Its "installation" is a checkout on sys.path, which is what this
recipe produces: the first-party *.py files (plus the vendored hifi_gan/*.py that vocoder.py
imports, normalizers.pkl and the test-set lists) are copied verbatim into the git-ignored
`baseline/_ref/`.  `.gpurunignore` does not list it, so it ships with the snapshot exactly like
the built libssb.so.  Nothing under baseline/_ref is ever committed or edited.

Run by `__graft_entry__.build()` whenever /root/reference is present (the build container);
the GPU box only uses the copied files.  Consumers: tests/test_reference_integration_gpu.py
(the reference's transduction_model.py running against the drop-in) and
`bench.py --impl reference` (the reference's own CPU step).
"""
import filecmp
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SSB_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def install(src=SRC, dst=DST, verbose=False):
    """Copy the reference's Python into baseline/_ref. Returns dst, or None if src is absent."""
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else None
    files = sorted(glob.glob(os.path.join(src, "*.py")))
    files += [os.path.join(src, f) for f in ("normalizers.pkl", "testset_largedev.json",
                                             "testset_origdev.json", "environment.yml", "LICENSE")
              if os.path.exists(os.path.join(src, f))]
    files += sorted(glob.glob(os.path.join(src, "hifi_gan", "*.py")))
    files += sorted(glob.glob(os.path.join(src, "hifi_gan", "config_v*.json")))
    n = 0
    for f in files:
        rel = os.path.relpath(f, src)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if not os.path.exists(out) or not filecmp.cmp(f, out, shallow=False):
            shutil.copyfile(f, out)
            n += 1
    if verbose:
        print(f"baseline/_ref: {len(files)} files ({n} copied) from {src}", file=sys.stderr)
    return dst


if __name__ == "__main__":
    print(install(verbose=True))
