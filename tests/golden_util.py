"""Loader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def hash_noise(n, scale):
    """Deterministic float64 noise in [-scale/2, scale/2) from integer arithmetic only (bit-stable on every platform): breaks an
    intensity lattice into n distinct values without storing n doubles in the fixture."""
    x = np.arange(n, dtype=np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = (x ^ (x >> np.uint64(31))) >> np.uint64(11)  # 53 bits
    return (x.astype(np.float64) / float(1 << 53) - 0.5) * scale


def load_golden(name):
    f = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: f[k] for k in f.files}
    q = int(g["quantum"])
    if q == 0:  # continuous intensities are stored as they are
        k = g["data_f64"]
        g["data"] = k.astype(np.float64)
    elif "noise_scale" in g and float(g["noise_scale"]) > 0:  # lattice + hash noise: continuous data rebuilt bit for bit
        k = g["k"].astype(np.int64)
        g["data"] = k.astype(np.float64) / q + hash_noise(k.size, float(g["noise_scale"])).reshape(k.shape)
        g["quantum"] = np.int64(0)  # the tests treat it as a continuous case
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
