"""Loader for the strict-mode fixtures (tests/golden/strict/*.npz, made by make_golden_strict.py from the unmodified reference)."""
import glob
import os

import numpy as np

STRICT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "strict")


def strict_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(STRICT_DIR, "*.npz")))


def load_strict(name):
    f = np.load(os.path.join(STRICT_DIR, name + ".npz"))
    g = {k: f[k] for k in f.files}
    k = g["k"].astype(np.int64)
    g["data"] = k.astype(np.float64) / int(g["quantum"])
    g["value_map_in"] = g["value_map_in"].astype(np.int64)
    ms = int(g["max_segment_size"])
    g["max_segment_size"] = (k.size + 1) if ms < 0 else ms
    g["H"] = float(g["H"])
    g["iterations"] = int(g["iterations"])
    g["segmented"] = g["segmented"].astype(np.int64)
    ends = np.cumsum(g["band_n"])
    g["bands"] = [(g["band_idx"][e - n:e].astype(np.int64), g["band_pin"][e - n:e], g["band_pout"][e - n:e])
                  for n, e in zip(g["band_n"], ends)]  # bands[i]: the state decision i + 1 reads (i = 0: after the init branch)
    return g
