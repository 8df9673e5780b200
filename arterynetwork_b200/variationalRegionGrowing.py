"""Drop-in for the reference module ``Code/variationalRegionGrowing.py``.

Same entry point, same arguments, same three return values, same stdout lines
(VRG:10, VRG:94-96) -- the iteration runs on a B200 through ``libvrg_b200.so``.

    from arterynetwork_b200.variationalRegionGrowing import variationalRegionGrowing
    segmented, segmentedMap, valueMap = variationalRegionGrowing(dataArray, valueMap)

Differences from the reference, all deliberate (DESIGN.md "Boundary"):

* the caller's ``valueMap`` is still mutated in place and returned as the same
  object (VRG:137-228), but it must start with labels 0 (seed), 3 (outside) and
  4 (excluded) only: labels 1/2 in the input put the reference's band lists and
  labels out of step, which has no order-free meaning;
* an empty seed set or a seed without boundary raises ``ValueError`` (the
  reference dies with ``IndexError`` at VRG:88);
* ``segmented`` rows come back in C order (the reference's row order is the
  history of its list appends); as a set it is ``argwhere(segmentedMap == 1)``;
* the 120 s wall-clock exit (VRG:97) is kept (``MAX_SECONDS``) but is checked
  between batches of iterations; set ``MAX_SECONDS = None`` for parity runs;
* intensities are processed as float64 whatever the input dtype (the reference
  silently drops to float32 sums for float32 input under NumPy 2);
* results are those of the order-free restatement of the band state machine:
  bit-identical to the reference wherever the reference's own result does not
  depend on its list order.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import _native as nat
from .engine import VRGEngine

A = (2 * np.pi) ** (-0.5)  # VRG:7
MAX_SECONDS = 120.0  # VRG:97
ITER_MAX = 200  # VRG:56
DEVICE = 0
INTENSITY = "index"  # how the sweep reads intensities: f64_dense | f64_band | index | continuous
CONTINUOUS_MAX_VOXELS = 1 << 24  # fall back to the brute-force Parzen mode (continuous data) up to this volume size

_EXIT_SUFFIX = {nat.EXIT_CONVERGED: "", nat.EXIT_MAX_TIME: " (Max time reached)",
                nat.EXIT_MAX_SEGMENT: " (Max segment size reached)"}


def _as_zyx(a):
    """View ``a`` as a C-contiguous (Z, Y, X) array; returns (view3d, transposed?)."""
    a = np.asarray(a)
    if a.ndim > 3 or a.ndim == 0:
        raise NotImplementedError("variationalRegionGrowing: 1-D to 3-D volumes only, got ndim=%d" % a.ndim)
    transposed = a.ndim > 1 and a.flags.f_contiguous and not a.flags.c_contiguous
    if transposed:
        a = a.T  # nibabel-style F-ordered volume: the slowest axis becomes z
    return a.reshape((1,) * (3 - a.ndim) + a.shape), transposed


def variationalRegionGrowing(dataArray, valueMap, H=2.25, maxSegmentSize=5000):
    """B200 implementation of VRG:10-121.  See the module docstring for the contract."""
    dataArray = np.asarray(dataArray)
    if not isinstance(valueMap, np.ndarray):
        raise TypeError("valueMap must be an ndarray (it is updated in place, as in the reference)")
    if dataArray.shape != valueMap.shape:
        raise ValueError("dataArray and valueMap must have the same shape")
    data3, transposed = _as_zyx(dataArray)
    vm_view = valueMap.T if transposed else valueMap
    vm3 = np.ascontiguousarray(vm_view).reshape(data3.shape)
    if vm3.size and (vm3.min() < 0 or vm3.max() > 255):
        raise ValueError("valueMap may only hold labels 0, 3 and 4")
    def run(mode):
        with VRGEngine(data3.shape, H=H, max_segment_size=maxSegmentSize, iter_max=ITER_MAX, device=DEVICE,
                       intensity=mode, max_seconds=MAX_SECONDS or 0.0) as eng:
            eng.upload(np.ascontiguousarray(data3, dtype=np.float64), vm3.astype(np.uint8))
            eng.init()
            res = eng.run()
            labels = eng.labels()
            if transposed or data3.shape != dataArray.shape:
                return res, labels, eng.segmented_map(), None
            return res, labels, None, eng.segmented()

    try:
        res, labels, seg_u8, segmented = run(INTENSITY)
    except nat.LevelsError:
        # more than 65536 distinct intensities: no level table.  Small volumes take the brute-force path, which is the
        # reference's own arithmetic (VRG:151-155, 232-255); large ones are as infeasible here as in the reference.
        if data3.size > CONTINUOUS_MAX_VOXELS:
            raise
        res, labels, seg_u8, segmented = run("continuous")
    lab_user = labels.reshape(vm_view.shape)
    lab_user = lab_user.T if transposed else lab_user
    valueMap[...] = lab_user  # in place, VRG:137-228
    if segmented is None:
        seg_user = seg_u8.reshape(vm_view.shape)
        seg_user = seg_user.T if transposed else seg_user
        segmentedMap = seg_user.astype(np.int64)  # np.full(shape, 0), VRG:45
        segmented = np.argwhere(segmentedMap == 1)
    else:
        segmentedMap = np.zeros(dataArray.shape, dtype=np.int64)
        segmentedMap[tuple(segmented.T)] = 1
    segmentedMap = np.ascontiguousarray(segmentedMap)
    total = '{}/{}'.format(segmented.shape[0], np.count_nonzero(dataArray))
    if res["exit_reason"] == nat.EXIT_MAX_ITER:  # VRG:118-120
        print('Segmented points are: \n', segmented)
        print('Max iteration reached! Finished at iteration {}'.format(res["iterations"]))
    else:  # VRG:94-104
        print('Finished at iteration {}{}'.format(res["iterations"], _EXIT_SUFFIX[res["exit_reason"]]))
    print('Total segmented voxels: ' + total)
    return segmented, segmentedMap, valueMap


def get_neighbours(p, exclude_p=True, shape=None):
    """In-bounds 3^ndim neighbourhood of ``p`` in lexicographic order, last axis fastest (VRG:263-282)."""
    p = np.asarray(p)
    offsets = np.array(list(itertools.product((-1, 0, 1), repeat=len(p))), dtype=np.int64)
    if exclude_p:
        offsets = offsets[np.any(offsets != 0, axis=1)]
    neighbours = p + offsets
    if shape is not None:
        ok = np.all((neighbours >= 0) & (neighbours < np.asarray(shape)), axis=1)
        neighbours = neighbours[ok]
    return neighbours
