"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of the reference's
variational region growing, ``/root/reference/Code/variationalRegionGrowing.py``
(cited below as VRG:line).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this; the
product path (``arterynetwork_b200``) never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this restatement
against fixtures produced by running the unmodified reference in the build
container (``tests/golden/make_golden.py``): the reference's own two self-tests
(VRG:284-314: 16 iterations 80/80, 11 iterations 4169/4169) plus phantoms with
removals, excluded (label 4) voxels, array-edge contact and the
``maxSegmentSize`` exit -- final labels, printed iteration count and the
per-iteration (n_flips, n_in, n_out) trace are bit-identical, and the
normalised Parzen sums at band voxels agree to 1e-12 relative.

The restatement is *dense and order-free* (SURVEY.md section 7.1):

* the flip decision at a band voxel (VRG:79-88) depends only on its label, its
  intensity and the two global Parzen sums, which depend on the regions only
  through their integer intensity histograms, so the decision is an L-entry
  table over the distinct intensity levels;
* the sequential band state machine (VRG:165-230) is restated as set rules:
  removals R, candidate additions A0, the cancel rule (an addition all of
  whose segmented neighbours leave is dropped; VRG:183-190 then VRG:198), the
  4->3 absorption (3x3x3 around every flip, 5x5x5 around every executed flip;
  VRG:167-168,177-179,207-208) and canonical re-classification of the bands;
* the incremental float bookkeeping (VRG:232-255) becomes exact integer
  histogram updates.

Where the reference's result depends on its list processing order (quirks
Q2/Q3 and the re-promoted cancelled addition "Q4") this
restatement counts the *potential* for it (``quirk_potential``); bit-identity
with the reference is asserted only on inputs where that count is zero.
"""
from __future__ import annotations

import numpy as np

A = (2 * np.pi) ** (-0.5)  # VRG:7
ITER_MAX = 200  # VRG:56

EXIT_CONVERGED = 0  # VRG:91
EXIT_MAX_SEGMENT = 2  # VRG:101
EXIT_MAX_ITER = 3  # VRG:118-121


def dil3(mask: np.ndarray) -> np.ndarray:
    """OR over the in-bounds 3x3x3 neighbourhood (centre included).

    Out-of-bounds neighbours are dropped, not wrapped (get_neighbours, VRG:263-282).
    Separable: three passes of a 3-tap OR.
    """
    m = np.asarray(mask, dtype=bool)
    for ax in range(m.ndim):
        out = m.copy()
        sl_lo = [slice(None)] * m.ndim
        sl_hi = [slice(None)] * m.ndim
        sl_lo[ax] = slice(0, -1)
        sl_hi[ax] = slice(1, None)
        out[tuple(sl_hi)] |= m[tuple(sl_lo)]
        out[tuple(sl_lo)] |= m[tuple(sl_hi)]
        m = out
    return m


def canonical_labels(seg: np.ndarray, excl: np.ndarray) -> np.ndarray:
    """Label alphabet of VRG:21 from the (seg, excl) state.

    0 inside, 1 inner band (segmented with an in-bounds unsegmented neighbour,
    VRG:139-142), 2 outer band (unsegmented with a segmented neighbour,
    VRG:143-145), 3 outside, 4 excluded.
    """
    seg = seg.astype(bool)
    excl = excl.astype(bool) & ~seg
    has_nonseg = dil3(~seg)
    has_seg = dil3(seg)
    lab = np.full(seg.shape, 3, dtype=np.uint8)
    lab[seg] = 0
    lab[seg & has_nonseg] = 1
    lab[~seg & has_seg] = 2
    lab[excl] = 4
    return lab


def build_levels(data: np.ndarray):
    """Distinct intensity levels (sorted, float64) and the per-voxel level index."""
    levels, idx = np.unique(np.asarray(data), return_inverse=True)
    return levels.astype(np.float64), idx.reshape(np.shape(data)).astype(np.int32)


def kernel_matrix(levels: np.ndarray, H: float) -> np.ndarray:
    """K[c, b] = A * exp(-0.5 * H * (level_c - level_b)**2)  (VRG:152-155)."""
    diff = levels[:, None] - levels[None, :]
    return A * np.exp(-0.5 * H * diff ** 2)


def decision_table(hist_in, hist_out, n_in, n_out, kmat):
    """Per-level normalised Parzen sums and the decision bit (VRG:79-87).

    d[b] is True when a voxel of level b belongs inside: in >= out, ties inside.
    """
    p_in = hist_in.astype(np.float64) @ kmat
    p_out = hist_out.astype(np.float64) @ kmat
    with np.errstate(all="ignore"):
        pin_n = p_in / n_in
        pout_n = p_out / n_out
        d = pin_n >= pout_n
    return d, pin_n, pout_n


def init_state(value_map: np.ndarray):
    """Init branch of ``update`` (VRG:44-46, 129-145), order-free.

    seeds are ``valueMap == 0``; every 4 in the 3x3x3 of a seed is absorbed to 3
    (VRG:137).  Labels 1/2 in the input are outside the parity-defined domain
    (the reference keeps them as stale labels that are in no band list).
    """
    vm = np.asarray(value_map)
    bad = ~np.isin(vm, (0, 3, 4))
    if bad.any():
        raise ValueError("oracle: initial valueMap may only hold labels 0, 3 and 4")
    seg = vm == 0
    excl = (vm == 4) & ~dil3(seg)
    return seg, excl


def step(seg, excl, lab, dbit_vox):
    """One order-free application of the band state machine (VRG:165-230).

    ``lab`` are the canonical labels of (seg, excl); ``dbit_vox`` the decision
    bit looked up at every voxel.  Returns the new state plus the flip sets.
    """
    R = (lab == 1) & ~dbit_vox  # inner-band voxel leaves iff in < out (VRG:87)
    A0 = (lab == 2) & dbit_vox  # outer-band voxel enters iff in >= out
    keep = seg & ~R
    Aex = A0 & dil3(keep)  # cancel rule: needs a segmented neighbour that stays
    seg2 = keep | Aex
    F = R | A0  # every listed flip absorbs its 3x3x3 (VRG:167-168)
    E = R | Aex  # executed flips absorb their 5x5x5 (VRG:177-179, 207-208)
    absorbed = excl & (dil3(F) | dil3(dil3(E)))
    excl2 = excl & ~absorbed
    return seg2, excl2, R, A0, Aex, absorbed


def vrg_oracle(data, value_map, H=2.25, max_segment_size=5000, iter_max=ITER_MAX,
               record_tables=False):
    """Dense, histogram-based restatement of ``variationalRegionGrowing`` (VRG:10-121).

    Returns a dict: ``labels`` (uint8 valueMap), ``seg`` (bool segmentedMap),
    ``iterations`` (the number the reference prints, VRG:94), ``exit`` code,
    ``trace`` (rows of n_flips, n_in, n_out; row 0 is the init state with
    n_flips = -1), ``quirk_potential`` and ``min_margin`` (smallest relative
    gap |in-out|/max(in,out) seen at a band voxel: ties are the only place the
    summation order of the Parzen sums could matter).
    """
    data = np.asarray(data)
    levels, idx = build_levels(data)
    L = len(levels)
    kmat = kernel_matrix(levels, float(H))
    seg, excl = init_state(value_map)
    n_in = int(seg.sum())
    if n_in == 0:
        raise ValueError("oracle: empty seed set (reference raises IndexError at VRG:88)")
    lab = canonical_labels(seg, excl)
    if not ((lab == 1) | (lab == 2)).any():
        raise ValueError("oracle: seed has no boundary (reference raises IndexError at VRG:88)")
    flat = idx.ravel()
    hist_in = np.bincount(flat[seg.ravel()], minlength=L).astype(np.int64)
    out_mask = (~seg & ~excl).ravel()
    hist_out = np.bincount(flat[out_mask], minlength=L).astype(np.int64)
    n_out = int(out_mask.sum())
    trace = [(-1, n_in, n_out)]
    tables = []
    quirk = {"add_to_inside": 0, "remove_to_outside": 0, "cancel_repromoted": 0, "cancelled": 0}
    min_margin = np.inf
    iter_num = 1
    exit_code = EXIT_MAX_ITER
    while iter_num <= iter_max:
        d, pin_n, pout_n = decision_table(hist_in, hist_out, n_in, n_out, kmat)
        if record_tables:
            tables.append((pin_n.copy(), pout_n.copy()))
        dvox = d[idx]
        band = (lab == 1) | (lab == 2)
        if band.any():
            lv = np.unique(idx[band])
            with np.errstate(all="ignore"):
                mg = np.abs(pin_n[lv] - pout_n[lv]) / np.maximum(pin_n[lv], pout_n[lv])
            min_margin = min(min_margin, float(np.nanmin(mg)))
        seg2, excl2, R, A0, Aex, absorbed = step(seg, excl, lab, dvox)
        n_flips = int(R.sum() + A0.sum())
        if n_flips == 0:
            exit_code = EXIT_CONVERGED
            break
        if n_in >= max_segment_size:  # tested before the flips are applied (VRG:101)
            exit_code = EXIT_MAX_SEGMENT
            break
        lab2 = canonical_labels(seg2, excl2)
        cancelled = A0 & ~Aex
        quirk["cancelled"] += int(cancelled.sum())
        quirk["add_to_inside"] += int((Aex & (lab2 == 0)).sum())
        quirk["remove_to_outside"] += int((R & (lab2 == 3)).sum())
        quirk["cancel_repromoted"] += int((cancelled & dil3(Aex)).sum())
        hist_in += np.bincount(flat[Aex.ravel()], minlength=L) - np.bincount(flat[R.ravel()], minlength=L)
        hist_out += (np.bincount(flat[R.ravel()], minlength=L) - np.bincount(flat[Aex.ravel()], minlength=L)
                     + np.bincount(flat[absorbed.ravel()], minlength=L))
        n_in += int(Aex.sum()) - int(R.sum())
        n_out += int(R.sum()) - int(Aex.sum()) + int(absorbed.sum())
        seg, excl, lab = seg2, excl2, lab2
        trace.append((n_flips, n_in, n_out))
        iter_num += 1
    return {
        "labels": lab,
        "seg": seg,
        "iterations": iter_num,
        "exit": exit_code,
        "trace": np.asarray(trace, dtype=np.int64),
        "levels": levels,
        "tables": tables,
        "quirk_potential": quirk,
        "min_margin": float(min_margin),
    }


def vrg_oracle_exact(data, value_map, H=2.25, max_segment_size=5000, iter_max=ITER_MAX, record_band=False):
    """Same order-free restatement for CONTINUOUS intensities (no level table): the two Parzen sums of every band
    voxel are evaluated exactly against the whole volume (VRG:151-155, 249-255), O(n_band * N) per iteration --
    small volumes only.  This is the oracle of the brute-force mode.

    ``record_band``: per decision, (flat voxel index, pin/n_in, pout/n_out) of every band voxel.
    """
    data = np.asarray(data, dtype=np.float64)
    seg, excl = init_state(value_map)
    n_in = int(seg.sum())
    if n_in == 0:
        raise ValueError("oracle: empty seed set (reference raises IndexError at VRG:88)")
    lab = canonical_labels(seg, excl)
    if not ((lab == 1) | (lab == 2)).any():
        raise ValueError("oracle: seed has no boundary (reference raises IndexError at VRG:88)")
    n_out = int((~seg & ~excl).sum())
    trace = [(-1, n_in, n_out)]
    bands = []
    quirk = {"add_to_inside": 0, "remove_to_outside": 0, "cancel_repromoted": 0, "cancelled": 0}
    min_margin = np.inf
    flat = data.ravel()
    iter_num, exit_code = 1, EXIT_MAX_ITER
    while iter_num <= iter_max:
        band = ((lab == 1) | (lab == 2)).ravel()
        bidx = np.flatnonzero(band)
        vin, vout = flat[seg.ravel()], flat[(~seg & ~excl).ravel()]
        pin = np.empty(len(bidx))
        pout = np.empty(len(bidx))
        for i0 in range(0, len(bidx), 256):  # blocked to bound memory
            v = flat[bidx[i0:i0 + 256]][:, None]
            pin[i0:i0 + 256] = np.sum(A * np.exp(-0.5 * H * (vin[None, :] - v) ** 2), axis=1)
            pout[i0:i0 + 256] = np.sum(A * np.exp(-0.5 * H * (vout[None, :] - v) ** 2), axis=1)
        pin_n, pout_n = pin / n_in, pout / n_out
        if record_band:
            bands.append((bidx.copy(), pin_n.copy(), pout_n.copy()))
        if len(bidx):
            min_margin = min(min_margin, float(np.min(np.abs(pin_n - pout_n) / np.maximum(pin_n, pout_n))))
        dvox = np.zeros(flat.shape, dtype=bool)
        dvox[bidx] = pin_n >= pout_n
        seg2, excl2, R, A0, Aex, absorbed = step(seg, excl, lab, dvox.reshape(seg.shape))
        n_flips = int(R.sum() + A0.sum())
        if n_flips == 0:
            exit_code = EXIT_CONVERGED
            break
        if n_in >= max_segment_size:
            exit_code = EXIT_MAX_SEGMENT
            break
        lab2 = canonical_labels(seg2, excl2)
        cancelled = A0 & ~Aex
        quirk["cancelled"] += int(cancelled.sum())
        quirk["add_to_inside"] += int((Aex & (lab2 == 0)).sum())
        quirk["remove_to_outside"] += int((R & (lab2 == 3)).sum())
        quirk["cancel_repromoted"] += int((cancelled & dil3(Aex)).sum())
        n_in += int(Aex.sum()) - int(R.sum())
        n_out += int(R.sum()) - int(Aex.sum()) + int(absorbed.sum())
        seg, excl, lab = seg2, excl2, lab2
        trace.append((n_flips, n_in, n_out))
        iter_num += 1
    return {"labels": lab, "seg": seg, "iterations": iter_num, "exit": exit_code,
            "trace": np.asarray(trace, dtype=np.int64), "bands": bands, "quirk_potential": quirk,
            "min_margin": float(min_margin)}
