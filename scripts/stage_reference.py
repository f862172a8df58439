"""Stage the UNMODIFIED reference package for the binding tests: /root/reference/{torch_points3d,conf} -> baseline/_ref/.

`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference` is what the
task contract names; it fails in this image ("No module named 'poetry'": the pyproject's build backend is absent and
there is no network).  The package is pure Python, so a --target install would place exactly these directories;
this script does that copy.  baseline/_ref is git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.
Run by __graft_entry__.build() whenever /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PGS_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, "torch_points3d")):
        return None
    for sub in ("torch_points3d", "conf"):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            if not force:
                continue
            shutil.rmtree(d)
        shutil.copytree(os.path.join(SRC, sub), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
