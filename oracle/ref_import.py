"""Import the UNMODIFIED reference loss classes from /root/reference (dev container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` does not exist on the GPU
box; there the verbatim copies staged by ``oracle/build_ref.py`` under ``oracle/_ref/`` are imported
instead (``bench.py --impl reference``, ``cpu_baseline`` and one ``-m gpu`` test).  ``torch/classes.py`` imports three packages that are not installed here
(``h5py`` :3, ``torchsummary`` :12, ``matplotlib`` :15 and ``helpers.py:5-7``); none is touched
by the loss classes, so empty stand-in modules satisfy the imports.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("SQ_REFERENCE_ROOT", "/root/reference")
STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/build_ref.py: travels to the GPU box


def available() -> bool:
    """The full reference tree (dev container): classes AND data/example_imgs."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "torch", "classes.py"))


def staged() -> bool:
    """The verbatim copies under oracle/_ref/ (GPU box)."""
    return os.path.isfile(os.path.join(STAGED, "torch", "classes.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        __import__(name)
        return sys.modules[name]
    except Exception:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m


def load():
    """Return the reference ``classes`` and ``quaternion`` modules (imported by bare name, as the reference does): from
    /root/reference when the tree is mounted, else from the staged copies."""
    if available():
        tdir = os.path.join(REFERENCE_ROOT, "torch")
    elif staged():
        tdir = os.path.join(STAGED, "torch")
    else:
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT} nor staged under {STAGED}")
    _stub("h5py")
    _stub("torchsummary", summary=lambda *a, **k: None)
    mpl = _stub("matplotlib")
    for sub in ("pyplot", "pylab"):
        m = _stub(f"matplotlib.{sub}")
        setattr(mpl, sub, m)
    lines = _stub("matplotlib.lines", Line2D=object)
    mpl.lines = lines
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import classes as ref_classes          # noqa: E402  (reference module, bare name)
        import quaternion as ref_quaternion    # noqa: E402
    return ref_classes, ref_quaternion


def example_fixtures():
    """(imgs[10,1,256,256] float32 in [0,1], labels[10,12] float32) from data/example_imgs.

    Labels follow torch/helpers.py:188-218 (a and t divided by 255, last four columns = quaternion),
    images follow torch/test.py:29-30 (8-bit BMP / 255).
    """
    import cv2
    import numpy as np
    d = os.path.join(REFERENCE_ROOT, "data", "example_imgs")
    rows, imgs = [], []
    with open(os.path.join(d, "labels.txt")) as f:
        lines = [l for l in f.read().split("\n") if l][1:]      # skip the header line
    for line in lines:
        s = line.split(",")
        v = [float(s[i]) / 255.0 if i in (1, 2, 3, 6, 7, 8) else float(s[i]) for i in range(1, 9)]
        v += [float(s[i]) for i in range(-4, 0)]
        rows.append(np.array(v, dtype=np.float32))
        imgs.append(cv2.imread(os.path.join(d, s[0]), 0).astype(np.float32) / 255.0)
    return np.stack(imgs)[:, None], np.stack(rows)
