"""Loader for tests/golden/mask/*.npz (made with SciPy by tests/golden/mask/make_golden_mask.py)."""
import glob
import os

import numpy as np

MASK_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mask")


def _names(prefix_rule):
    out = []
    for p in sorted(glob.glob(os.path.join(MASK_DIR, "*.npz"))):
        n = os.path.splitext(os.path.basename(p))[0]
        if n.startswith("rule_") == prefix_rule:
            out.append(n)
    return out


def mask_names():
    return _names(False)


def rule_names():
    return _names(True)


def load_mask_case(name):
    f = np.load(os.path.join(MASK_DIR, name + ".npz"))
    shape = tuple(int(s) for s in f["shape"])
    n = int(np.prod(shape))
    g = {"shape": shape}
    if "mask" in f.files:
        g["mask"] = np.unpackbits(f["mask"])[:n].reshape(shape).astype(bool)
        g["edt"] = f["edt"]
        g["labels"] = f["labels"]
        g["n_components"] = int(f["n_components"])
    else:
        g["vesselness"] = f["k"].astype(np.float64) / int(f["quantum"])
        g["brain"] = np.unpackbits(f["brain"])[:n].reshape(shape).astype(np.uint8)
        g["vessel_mask"] = np.unpackbits(f["vessel_mask"])[:n].reshape(shape).astype(np.uint8)
        g["min_size"] = int(f["min_size"])
        g["brain_edt"] = f["brain_edt"]
    return g
