"""Golden fixtures for the mask-side operations, made with SciPy in the build container:

    python tests/golden/mask/make_golden_mask.py

The reference calls scipy.ndimage.distance_transform_edt (Code/generateVesselVolume.py:183,
Code/manualCorrectionGUI.py:248) and skimage.measure.label(connectivity=3) (Code/generateVesselVolume.py:126).  SciPy is
installed here (version recorded in every fixture); scikit-image is not, so the labelling fixtures come from
scipy.ndimage.label with the full 3x3x3 structure (the same 26-connectivity; both number components in raster order of
their first voxel).  Each ``<case>.npz`` holds: mask (bit-packed), shape, edt (float64, SciPy), labels (int32, SciPy),
and for the ``rule_*`` cases a vesselness volume (integer lattice k / quantum), the brain mask and the vessel mask that
the restated rule of generateVesselVolume.py:187-199 gives when SciPy does the EDT and the labelling.
"""
import os
import sys

import numpy as np
import scipy
from scipy import ndimage as ndi

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
FULL = np.ones((3, 3, 3), dtype=int)


def ball(shape, c, r):
    z, y, x = np.ogrid[: shape[0], : shape[1], : shape[2]]
    return (z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 <= r * r


def masks():
    rng = np.random.default_rng(7)
    out = {}
    m = np.zeros((9, 10, 11), dtype=bool); m[4, 5, 5] = True
    out["single_voxel"] = m
    m = np.ones((7, 8, 9), dtype=bool); m[0, 0, 0] = False
    out["one_zero_corner"] = m                      # every distance is to the same far corner
    out["ball"] = ball((24, 24, 24), (12, 11, 12), 9)
    out["random_dense"] = rng.random((12, 20, 37)) < 0.8   # odd X, many ties
    out["random_sparse"] = rng.random((16, 16, 40)) < 0.15
    m = np.zeros((20, 33, 70), dtype=bool)                  # tubes touching the array faces, X > 64
    m[8:12, 14:18, :] = True; m[:, 3:5, 60:63] = True; m[2:4, :, 30:33] = True
    out["tubes_to_faces"] = m
    m = np.ones((5, 6, 40), dtype=bool); m[:, :, 17] = False
    out["slab_zero_plane"] = m
    m = np.ones((3, 1, 50), dtype=bool); m[1, 0, 3] = False  # degenerate axis
    out["thin_axis"] = m
    return out


def rule_case(seed, shape, n_tubes):
    """A vesselness-like volume: bright tubes of different sizes + speckle, brain mask = a big ellipsoid."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    z, y, x = np.ogrid[:Z, :Y, :X]
    brain = ((z - Z / 2) / (Z / 2 - 1)) ** 2 + ((y - Y / 2) / (Y / 2 - 2)) ** 2 + ((x - X / 2) / (X / 2 - 2)) ** 2 <= 1.0
    k = np.rint(np.abs(rng.normal(0, 6, shape))).astype(np.int64)      # background speckle, quantum 1/64
    for _ in range(n_tubes):
        c = rng.integers(2, [Z - 2, Y - 2, X - 2])
        axis = rng.integers(0, 3)
        ln = int(rng.integers(3, max(shape)))
        r = int(rng.integers(1, 3))
        sl = [slice(max(0, c[i] - r), c[i] + r + 1) for i in range(3)]
        sl[axis] = slice(max(0, c[axis] - ln // 2), c[axis] + ln // 2)
        k[tuple(sl)] = rng.integers(46, 64)
    k[rng.random(shape) < 0.002] = 60                                   # isolated bright voxels (small components)
    k[0, 0, 0] = 0; k[-1, -1, -1] = 64                                    # pin the range
    return k, 64, brain


def vessel_mask_with_scipy(vesselness, brain, edge_distance=10, edge_fraction=0.8, fraction=0.7, min_size=150):
    v = vesselness.copy()
    edt = ndi.distance_transform_edt(brain)
    lo, hi = np.amin(v), np.amax(v)
    v[np.logical_and(edt <= edge_distance, v <= lo + edge_fraction * (hi - lo))] = 0
    v[v <= lo + fraction * (hi - lo)] = 0
    v[v != 0] = 1
    labeled, _ = ndi.label(v, structure=FULL)
    counts = np.bincount(labeled.ravel())
    for num in np.nonzero(counts)[0]:
        if counts[num] <= min_size:
            v[labeled == num] = 0
    return v.astype(np.uint8), edt, labeled.astype(np.int32)


def main():
    for name, m in masks().items():
        edt = ndi.distance_transform_edt(m)
        labeled, K = ndi.label(m, structure=FULL)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array(m.shape), mask=np.packbits(m),
                            edt=edt, labels=labeled.astype(np.int32), n_components=K, scipy_version=scipy.__version__)
        print(name, m.shape, "fg", int(m.sum()), "components", K, "max edt %.4f" % edt.max())
    for name, (seed, shape, nt, min_size) in {"rule_a": (1, (24, 40, 56), 14, 150), "rule_b": (2, (30, 30, 90), 20, 40),
                                              "rule_c": (3, (16, 64, 64), 10, 150)}.items():
        k, q, brain = rule_case(seed, shape, nt)
        vessel, edt, labeled = vessel_mask_with_scipy(k.astype(np.float64) / q, brain, min_size=min_size)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array(shape), k=k.astype(np.int16), quantum=q,
                            brain=np.packbits(brain), vessel_mask=np.packbits(vessel.astype(bool)), min_size=min_size,
                            brain_edt=edt, scipy_version=scipy.__version__)
        print(name, shape, "vessel voxels", int(vessel.sum()), "components before filter", int(labeled.max()))


if __name__ == "__main__":
    main()
