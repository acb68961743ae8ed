"""Stage the UNMODIFIED reference under oracle/_ref/ so that it travels to the GPU box (dev container only).

    python -m oracle.build_ref

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's hot path is pure Python: ``torch/classes.py`` and the
modules it imports by bare name (``quaternion.py``, ``helpers.py``, and through ``helpers.py`` ``models.py``).  The files
are copied byte for byte from ``/root/reference/torch`` -- nothing is edited -- into ``oracle/_ref/torch/``; three
import-only stand-ins (``h5py``, ``torchsummary``, ``matplotlib``: packages the reference imports at module level,
never touches on the loss path, and that are not installed in this image) are written next to them in
``oracle/_ref/shims/``.  ``oracle/_ref/`` is git-ignored (reference sources never enter this repo's history) but not
gpurun-ignored, so the box sees it.  ``oracle/ref_import.load()`` then imports the reference classes from here:
``bench.py --impl reference`` and ``cpu_baseline`` time them on the box's host cores (``kind: "reference"``), and a
``-m gpu`` test compares the CUDA path with them directly.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SQ_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("classes.py", "quaternion.py", "helpers.py", "models.py")

SHIMS = {
    "h5py.py": '"""import-only stand-in (the reference imports h5py at torch/classes.py:3; the losses never use it)"""\n',
    "torchsummary.py": '"""import-only stand-in (torch/classes.py:12)"""\n\n\ndef summary(*args, **kwargs):\n    return None\n',
    "matplotlib/__init__.py": '"""import-only stand-in (torch/classes.py:15, torch/helpers.py:5-7, torch/quaternion.py:4)"""\n',
    "matplotlib/pyplot.py": "",
    "matplotlib/pylab.py": "",
    "matplotlib/lines.py": "class Line2D:\n    def __init__(self, *args, **kwargs):\n        pass\n",
}


def build(verbose: bool = True) -> bool:
    """Copy the files; returns False (and does nothing) when the reference tree is not mounted."""
    src_dir = os.path.join(SRC, "torch")
    if not all(os.path.isfile(os.path.join(src_dir, f)) for f in FILES):
        return False
    os.makedirs(os.path.join(DST, "torch"), exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(DST, "torch", f))
        manifest[f] = hashlib.sha256(open(os.path.join(src_dir, f), "rb").read()).hexdigest()
    for rel, text in SHIMS.items():
        path = os.path.join(DST, "shims", rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_dir, "sha256": manifest, "note": "verbatim copies; do not edit"}, fh, indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
