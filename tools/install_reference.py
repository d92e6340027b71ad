"""Put the UNMODIFIED reference where the GPU box can import it: /root/reference/src -> baseline/_ref/src.

``pip install --target baseline/_ref /root/reference`` fails ("neither setup.py nor pyproject.toml": the reference is
a script tree, not a package), so the install is a verbatim copy of its ``src/`` python files.  ``baseline/_ref`` is
git-ignored (never part of this repository's history) but NOT gpurun-ignored, so it travels to the GPU box with the
snapshot; /root/reference itself does not exist there.  Run by ``__graft_entry__.build()`` whenever /root/reference
is present.

``ref_env.import_reference()`` (tools/ref_env.py) makes the copy importable as ``src`` with the three stubs the survey
lists (torch._six, h5py, collections.Mapping) -- nothing inside the reference is patched.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref", "src")


def install(verbose: bool = False) -> bool:
    """Returns True when baseline/_ref/src is present and (if /root/reference exists) byte-identical to it."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    n = 0
    for dirpath, dirnames, filenames in os.walk(SRC):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, SRC)
        out = os.path.normpath(os.path.join(DST, rel))
        os.makedirs(out, exist_ok=True)
        for f in filenames:
            if not f.endswith(".py"):
                continue
            s, d = os.path.join(dirpath, f), os.path.join(out, f)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print(f"baseline/_ref/src: {n} file(s) refreshed from {SRC}")
    return True


if __name__ == "__main__":
    ok = install(verbose=True)
    sys.exit(0 if ok else 1)
