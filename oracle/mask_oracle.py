"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the steps on either side of the variational-region-growing path
(SURVEY.md section 8(f), rows N2 and N3), cited as GVV:line = ``/root/reference/Code/generateVesselVolume.py`` and
MCG:line = ``/root/reference/Code/manualCorrectionGUI.py``.  Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU
legs may import this; the product path (``arterynetwork_b200``) never does.

* ``edt_oracle``            scipy.ndimage.distance_transform_edt(mask), default arguments (GVV:183, MCG:248)
* ``label_oracle``          skimage.measure.label(volume, return_num=True, connectivity=3) + np.bincount (GVV:126-131)
* ``vessel_mask_oracle``    the vesselness -> vessel mask rule of GVV:187-199: two thresholds relative to the range of the
                            vesselness volume, the first applied only within 10 voxels of the brain-mask boundary, then
                            removal of the 26-connected components of at most 150 voxels

Parity status: PINNED for the first two against SciPy outputs computed in the build container
(``tests/golden/mask/*.npz`` made by ``tests/golden/mask/make_golden_mask.py``).  scikit-image and nibabel are not
installed, and GVV's ``main()`` is file-driven (NIfTI in, NIfTI out), so the rule itself cannot be executed from the
reference: its fixtures are this restatement's outputs with SciPy doing the EDT and the labelling -- stated as
"restated, pinned through SciPy" wherever they are used.

The arithmetic lives in ``oracle/mask_oracle.c`` (plain double loops / flood fill); this module is its ctypes front.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mask_oracle.c")
LIB = os.path.join(HERE, "_build", "libmask_oracle.so")
ORACLE_INF = 1 << 40

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        _lib.edt_sq_oracle.restype = ctypes.c_int
        _lib.edt_sq_oracle.argtypes = [p, i64, i64, i64, p]
        _lib.label26_oracle.restype = i64
        _lib.label26_oracle.argtypes = [p, i64, i64, i64, p, p, i64]
    return _lib


def _as3d(a):
    a = np.asarray(a)
    if a.ndim != 3:
        raise ValueError("3-D volumes only")
    return a


def edt_sq_oracle(mask) -> np.ndarray:
    """Squared distance (int64) of every non-zero voxel to the nearest zero voxel."""
    m = np.ascontiguousarray(_as3d(mask) != 0, dtype=np.uint8)
    out = np.empty(m.shape, dtype=np.int64)
    rc = _load().edt_sq_oracle(m.ctypes.data, *m.shape, out.ctypes.data)
    if rc != 0:
        raise MemoryError("edt oracle")
    return out


def edt_oracle(mask) -> np.ndarray:
    """distance_transform_edt(mask): float64 Euclidean distance to the nearest zero voxel (GVV:183, MCG:248)."""
    sq = edt_sq_oracle(mask)
    if (sq >= ORACLE_INF).any():
        raise ValueError("mask has no zero voxel")
    return np.sqrt(sq.astype(np.float64))


def label_oracle(volume):
    """GVV:126-131: (labeled int32 volume, [(label, size), ...] including the background entry (0, n0) when present)."""
    b = np.ascontiguousarray(_as3d(volume) != 0, dtype=np.uint8)
    labels = np.empty(b.shape, dtype=np.int32)
    K = int(_load().label26_oracle(b.ctypes.data, *b.shape, labels.ctypes.data, None, 0))
    if K < 0:
        raise MemoryError("label oracle")
    counts = np.bincount(labels.ravel())           # GVV:127
    loc = np.nonzero(counts)[0]                    # GVV:128
    return labels, list(zip(loc.tolist(), counts[loc].tolist()))  # GVV:129-131


def thresholds_oracle(vesselness, edge_fraction=0.8, fraction=0.7):
    """GVV:188-191: the two cut-offs, relative to the range of the vesselness volume (same float64 expression)."""
    lo, hi = np.amin(vesselness), np.amax(vesselness)
    return lo + edge_fraction * (hi - lo), lo + fraction * (hi - lo)


def vessel_mask_oracle(vesselness, brain_mask, edge_distance=10, edge_fraction=0.8, fraction=0.7, min_size=150,
                       brain_edt=None):
    """GVV:187-199 without the file I/O: uint8 vessel mask from a vesselness volume and the brain mask.

    ``brain_edt`` (optional) replaces the oracle's own EDT of ``brain_mask`` (e.g. SciPy's, when making fixtures).
    """
    v = np.array(vesselness, dtype=np.float64, copy=True)
    edt = edt_oracle(brain_mask) if brain_edt is None else np.asarray(brain_edt)
    t_edge, t_all = thresholds_oracle(v, edge_fraction, fraction)
    v[np.logical_and(edt <= edge_distance, v <= t_edge)] = 0   # GVV:189-190
    v[v <= t_all] = 0                                          # GVV:191-192
    v[v != 0] = 1                                              # GVV:195
    labels, result = label_oracle(v)                           # GVV:196
    sizes = np.zeros(max(num for num, _ in result) + 1, dtype=np.int64)
    for num, size in result:
        sizes[num] = size
    v[sizes[labels] <= min_size] = 0                           # GVV:198-200, all small labels at once (background: no-op)
    return v.astype(np.uint8)                                  # GVV:216 saves it as uint8
