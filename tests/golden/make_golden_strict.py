"""Golden fixtures for the strict (list-order) mode, SURVEY.md section 8(f) N4: outputs of the UNMODIFIED reference on
noisy inputs where its result DEPENDS on its list order, i.e. where it differs from the order-free restatement
(stale band labels in the returned valueMap, another iteration count / trace, another segmentation at the
``maxSegmentSize`` exit).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_strict.py        # ~2 min

Each ``tests/golden/strict/<case>.npz`` holds the input (integer lattice ``k``, ``data = k / quantum``; seeds), the
reference's final valueMap, ``segmented`` in the reference's ROW ORDER, stdout, the (n_flips, n_in, n_out) trace, and
for every iteration the band list in ``allBnd`` order (VRG:48,111) with innerProb/innerSize and outerProb/outerSize
at every band voxel (VRG:79-82) -- the running sums the reference actually holds, drift included.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from arterynetwork_b200.phantom import make_phantom  # noqa: E402
from oracle.ref_harness import run_reference  # noqa: E402
from oracle.vrg_oracle import vrg_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "strict")


def noisy_bar(seed, shape=(24, 24, 24), noise=0.3, q=16, cube=12, excl_below=None):
    """A 6x6 bar along x in noise, seeded with a cube wider than the bar: removals, cancelled additions, stale labels."""
    rng = np.random.default_rng(seed)
    d = np.zeros(shape)
    c = shape[0] // 2
    d[c - 3:c + 3, c - 3:c + 3, 2:shape[2] - 2] = 1.0
    k = np.round((d + rng.normal(0, noise, shape)) * q).astype(np.int64)
    vm = np.full(shape, 3)
    if excl_below is not None:
        vm[k <= excl_below] = 4
    lo = c - cube // 2
    vm[lo:lo + cube, lo:lo + cube, lo:lo + cube] = 0
    return k, q, vm


def noisy_forest(seed, sigma_k):
    data, vm, info = make_phantom((28, 28, 28), seed=seed, cell=(28, 28, 28), margin=3, depth=2, root_r2=4, min_len=7,
                                  max_len=11, quantum=16, sigma_k=sigma_k)
    return np.rint(data * info["quantum"]).astype(np.int64), info["quantum"], vm.astype(np.int64)


CASES = {
    # name: (builder, kwargs of the run)
    "bar_stale": (lambda: noisy_bar(128, noise=0.4, q=8), dict(H=2.25, max_segment_size=None)),      # 60 stale labels, 21 vs 20 iterations
    "bar_iters": (lambda: noisy_bar(110, noise=0.4, q=8), dict(H=2.25, max_segment_size=None)),      # 19 vs 18 iterations
    "bar_q16": (lambda: noisy_bar(3, noise=0.3, q=16), dict(H=2.25, max_segment_size=None)),
    "bar_maxseg": (lambda: noisy_bar(116, noise=0.4, q=8), dict(H=2.25, max_segment_size=3000)),     # stops at VRG:101 on another state
    "bar_excl": (lambda: noisy_bar(7, noise=0.4, q=8, excl_below=-2), dict(H=2.25, max_segment_size=None)),  # label 4 + absorption
    "bar_h1": (lambda: noisy_bar(21, noise=0.3, q=16), dict(H=1.0, max_segment_size=None)),
    "forest_noisy": (lambda: noisy_forest(4, 5), dict(H=2.25, max_segment_size=None)),               # small seeds growing in heavy noise
}


def generate(name):
    build, kw = CASES[name]
    k, q, vm = build()
    data = k.astype(np.float64) / q
    ms = kw["max_segment_size"]
    t0 = time.time()
    r = run_reference(data, vm, H=kw["H"], max_segment_size=ms, check_drift=False)
    o = vrg_oracle(data, vm, H=kw["H"], max_segment_size=(data.size + 1) if ms is None else ms)
    differs = dict(value_map=int((r["value_map"] != o["labels"]).sum()), seg=int((r["segmented_map"].astype(bool) != o["seg"]).sum()),
                   iterations=(r["iterations"], o["iterations"]),
                   trace=not (r["trace"].shape == o["trace"].shape and (r["trace"] == o["trace"]).all()))
    band_n, band_idx, pin, pout = [], [], [], []
    for s in r["band_samples"]:
        band_n.append(len(s["coords"]))
        band_idx.append(np.ravel_multi_index(tuple(s["coords"].T), data.shape).astype(np.int32))
        pin.append(s["p_in"])
        pout.append(s["p_out"])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        k=k.astype(np.int16), quantum=np.int64(q), value_map_in=vm.astype(np.uint8), H=np.float64(kw["H"]),
        max_segment_size=np.int64(-1 if ms is None else ms),
        value_map=r["value_map"].astype(np.uint8), segmented=r["segmented"].astype(np.int16),
        iterations=np.int64(r["iterations"]), stdout=np.array(r["stdout"]), trace=r["trace"],
        band_n=np.asarray(band_n, dtype=np.int64), band_idx=np.concatenate(band_idx),
        band_pin=np.concatenate(pin), band_pout=np.concatenate(pout),
        q2_voxels=np.int64(r["quirks"]["Q2_voxels"]), q3_dropped=np.int64(r["quirks"]["Q3_dropped"]),
        orderfree_value_map_diff=np.int64(differs["value_map"]), orderfree_seg_diff=np.int64(differs["seg"]),
        orderfree_iterations=np.int64(o["iterations"]))
    print("%-14s %5.1fs  iterations %d (order-free %d)  valueMap differs at %d voxels, seg at %d, trace differs: %s, Q2 %d Q3 %d" % (
        name, time.time() - t0, r["iterations"], o["iterations"], differs["value_map"], differs["seg"], differs["trace"],
        r["quirks"]["Q2_voxels"], r["quirks"]["Q3_dropped"]))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        generate(n)
