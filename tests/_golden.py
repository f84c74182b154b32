"""Access to the committed origin-generated fixtures (tests/golden/*.npz, see make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix):
    out = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "_*.npz")))
    assert out, "no golden fixtures for %s" % prefix
    return out


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def inp(self, key, dtype=None):
        k = "in_" + key
        if k not in self.z.files:
            return None
        a = self.z[k]
        if dtype is not None and a.dtype.kind == "f":
            a = a.astype(dtype)
        return a

    def attr(self, key):
        if "attrnone_" + key in self.z.files:
            return None
        return float(self.z["attr_" + key])

    def out(self, key, prec):
        return self.z["out%d_%s" % (prec, key)]

    def grad(self, key, prec):
        return self.z["grad%d_%s" % (prec, key)]


def rel_err(a, b):
    """max|a-b| / max(1, max|b|)  -- the norm-relative metric of SURVEY.md 8c."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def grad_err(a, b):
    """max|a-b| / max|b|: gradients of mean-type losses carry 1/(T*B) (or 1/B), so the floor of 1 in ``rel_err`` would
    make a 1e-5 bar vacuous for them -- they are compared relative to their own largest entry instead."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = float(np.max(np.abs(b)))
    if scale == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / scale)
