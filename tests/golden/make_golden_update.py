"""Golden fixtures for the reference's ``update`` (VRG:124-261), made by calling the UNMODIFIED function.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_update.py

For each case the reference's own driver logic (VRG:44-52, 79-88) is replayed step by step around ``update`` and every call
is recorded: its inputs (segmentedMap, valueMap, flipedPoints) and its seven outputs -- valueMap and segmentedMap after
the call, the band lists (as C-sorted rows; the reference's row order is its append history) and the unnormalised Parzen
sums ``innerProb`` / ``outerProb`` at every band voxel.  ``tests/test_gpu_update.py`` feeds the same inputs to
``arterynetwork_b200.variationalRegionGrowing.update`` and compares.  A case is only written when the reference left no
stale label behind (valueMap == canonical labels after every call), i.e. where its result is order-free.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_golden import case_excl32, case_removal32, case_straight_line, case_tube_clean, case_tube_fat_seed  # noqa: E402
from oracle.ref_harness import canonical_labels, load_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "update")


def sort_rows(a):
    a = np.asarray(a, dtype=np.int64).reshape(-1, 3)
    return a[np.lexsort(a.T[::-1])]


def record(name, k, q, vm0, H, n_calls):
    ref = load_reference()
    data = k.astype(np.float64) / q if q > 1 else k
    vm = np.array(vm0, copy=True)
    segmented = np.array(np.where(vm == 0)).T  # VRG:44-46
    seg_map = np.full(data.shape, 0)
    seg_map[tuple(segmented.T)] = 1
    calls = {}
    state = None
    for c in range(n_calls):
        seg_in, vm_in = seg_map.copy(), vm.copy()
        if c == 0:
            out = ref.update(data, segmented, seg_map, vm, H)
            flips = np.zeros((0, 3), dtype=np.int64)
        else:
            segmented, seg_map, vm, innerBnd, outerBnd, innerProb, outerProb = state
            allBnd = np.concatenate((innerBnd, outerBnd))  # VRG:48,111
            n_in = np.count_nonzero((vm == 0) | (vm == 1))
            n_out = np.count_nonzero((vm == 2) | (vm == 3))
            pin = innerProb[tuple(allBnd.T)] / n_in  # VRG:79-82
            pout = outerProb[tuple(allBnd.T)] / n_out
            mask = np.logical_xor(seg_map[tuple(allBnd.T)], pin >= pout)  # VRG:87
            flips = allBnd[mask, :]
            if len(flips) == 0:
                break
            out = ref.update(data, segmented, seg_map, vm, H, flips, innerBnd, outerBnd, innerProb, outerProb)
        segmented, seg_map, vm, innerBnd, outerBnd, innerProb, outerProb = state = out
        canon = canonical_labels(seg_map == 1, vm == 4)
        if not np.array_equal(canon, vm):
            raise SystemExit("%s: the reference left stale labels after call %d -- not an order-free case" % (name, c))
        ib, ob = sort_rows(innerBnd), sort_rows(outerBnd)
        band = np.concatenate([ib, ob])
        calls.update({
            "c%d_seg_in" % c: np.packbits(seg_in.astype(bool)), "c%d_vm_in" % c: vm_in.astype(np.uint8),
            "c%d_flips" % c: np.asarray(flips, dtype=np.int64).reshape(-1, 3),
            "c%d_seg_out" % c: np.packbits(seg_map.astype(bool)), "c%d_vm_out" % c: vm.astype(np.uint8),
            "c%d_segmented" % c: sort_rows(segmented), "c%d_inner" % c: ib, "c%d_outer" % c: ob,
            "c%d_pin" % c: innerProb[tuple(band.T)].copy(), "c%d_pout" % c: outerProb[tuple(band.T)].copy(),
            "c%d_prob_nonzero" % c: np.array([np.count_nonzero(innerProb), np.count_nonzero(outerProb)]),
        })
        n_done = c + 1
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), k=k.astype(np.int16), quantum=q, data_is_int=(q == 1), H=H,
                        n_calls=n_done, numpy_version=np.__version__, **calls)
    print(name, "calls", n_done, "flips per call", [len(calls["c%d_flips" % c]) for c in range(n_done)])


def main():
    for name, case, n in (("tube_clean", case_tube_clean, 4), ("straight_line", case_straight_line, 3),
                          ("removal32", case_removal32, 4), ("excl32", case_excl32, 4), ("tube_fat_seed", case_tube_fat_seed, 4)):
        k, q, vm, kw = case()
        record(name, np.asarray(k), q, vm, kw["H"], n)


if __name__ == "__main__":
    main()
