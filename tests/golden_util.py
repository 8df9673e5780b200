"""Loader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    f = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: f[k] for k in f.files}
    q = int(g["quantum"])
    if q == 0:  # continuous intensities are stored as they are
        k = g["data_f64"]
        g["data"] = k.astype(np.float64)
    else:
        k = g["k"].astype(np.int64)
        g["data"] = k if bool(g["data_is_int"]) else k.astype(np.float64) / q
    g["value_map_in"] = g["value_map_in"].astype(np.int64)
    ms = int(g["max_segment_size"])
    g["max_segment_size"] = (k.size + 1) if ms < 0 else ms
    g["H"] = float(g["H"])
    g["iterations"] = int(g["iterations"])
    g["seg_bool"] = np.unpackbits(g["seg"])[: k.size].reshape(k.shape).astype(bool)
    return g
