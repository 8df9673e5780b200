"""update() (VRG:124-261) on the GPU against fixtures recorded from the unmodified reference function
(tests/golden/make_golden_update.py), the resumable valueMap, the reference's two self-tests as callables (VRG:284-314),
the order-dependence counters of vrg_result, and the label hash (B200 only)."""
import contextlib
import glob
import io
import os
import warnings

import numpy as np
import pytest

from golden_util import load_golden

pytestmark = pytest.mark.gpu
UPDATE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "update")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(UPDATE_DIR, "*.npz")))
RTOL = 1e-12


def _rows_sorted(a):
    a = np.asarray(a, dtype=np.int64).reshape(-1, 3)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("name", NAMES)
def test_update_matches_reference_call_by_call(name):
    from arterynetwork_b200 import variationalRegionGrowing as mod
    f = np.load(os.path.join(UPDATE_DIR, name + ".npz"))
    k = f["k"].astype(np.int64)
    data = k if bool(f["data_is_int"]) else k.astype(np.float64) / int(f["quantum"])
    H = float(f["H"])
    shape = k.shape
    assert int(f["n_calls"]) >= 3
    for c in range(int(f["n_calls"])):
        seg_in = np.unpackbits(f["c%d_seg_in" % c])[: k.size].reshape(shape).astype(np.int64)
        vm_in = f["c%d_vm_in" % c].astype(np.int64)
        segmented_in = np.argwhere(seg_in == 1)
        flips = f["c%d_flips" % c]
        if c == 0:
            out = mod.update(data, segmented_in, seg_in, vm_in, H)
        else:  # the band lists and sums are rebuilt from the state: pass what a caller of the reference would hold
            prev_in, prev_out = f["c%d_inner" % (c - 1)], f["c%d_outer" % (c - 1)]
            ip, op = np.zeros(shape), np.zeros(shape)
            out = mod.update(data, segmented_in, seg_in, vm_in, H, flips, prev_in, prev_out, ip, op)
        segmented, seg_map, vm, inner, outer, iprob, oprob = out
        assert vm is vm_in and seg_map is seg_in  # updated in place, same objects (VRG:137-228,173,201)
        assert np.array_equal(vm, f["c%d_vm_out" % c])
        assert np.array_equal(seg_map == 1, np.unpackbits(f["c%d_seg_out" % c])[: k.size].reshape(shape).astype(bool))
        assert np.array_equal(_rows_sorted(segmented), f["c%d_segmented" % c])
        assert np.array_equal(_rows_sorted(inner), f["c%d_inner" % c]) and np.array_equal(_rows_sorted(outer), f["c%d_outer" % c])
        band = np.concatenate([f["c%d_inner" % c], f["c%d_outer" % c]])
        np.testing.assert_allclose(iprob[tuple(band.T)], f["c%d_pin" % c], rtol=RTOL, atol=0)
        np.testing.assert_allclose(oprob[tuple(band.T)], f["c%d_pout" % c], rtol=RTOL, atol=0)
        assert [np.count_nonzero(iprob), np.count_nonzero(oprob)] == f["c%d_prob_nonzero" % c].tolist()
        if c:
            assert iprob is ip and oprob is op  # the caller's arrays, as in the reference


def test_update_driven_loop_reproduces_the_whole_run():
    """The reference's own driver loop (VRG:47-117) written around OUR update(): same labels and iteration count as the
    fixture of the whole run."""
    from arterynetwork_b200 import variationalRegionGrowing as mod
    g = load_golden("tube_clean")
    data, vm = g["data"], g["value_map_in"].copy()
    segmented = np.array(np.where(vm == 0)).T
    seg_map = np.full(data.shape, 0)
    seg_map[tuple(segmented.T)] = 1
    segmented, seg_map, vm, ib, ob, ip, op = mod.update(data, segmented, seg_map, vm, g["H"])
    it = 1
    while it <= 200:
        allb = np.concatenate((ib, ob))
        n_in = np.count_nonzero((vm == 0) | (vm == 1)); n_out = np.count_nonzero((vm == 2) | (vm == 3))
        mask = np.logical_xor(seg_map[tuple(allb.T)], ip[tuple(allb.T)] / n_in >= op[tuple(allb.T)] / n_out)
        flips = allb[mask, :]
        if len(flips) == 0:
            break
        segmented, seg_map, vm, ib, ob, ip, op = mod.update(data, segmented, seg_map, vm, g["H"], flips, ib, ob, ip, op)
        it += 1
    assert it == g["iterations"] and np.array_equal(vm, g["labels"]) and np.array_equal(seg_map == 1, g["seg_bool"])


def test_update_ignores_points_outside_the_bands_and_rejects_points_outside_the_volume():
    from arterynetwork_b200 import variationalRegionGrowing as mod
    g = load_golden("tube_clean")
    data, vm = g["data"], g["value_map_in"].copy()
    segmented = np.array(np.where(vm == 0)).T
    seg_map = np.full(data.shape, 0); seg_map[tuple(segmented.T)] = 1
    out0 = mod.update(data, segmented, seg_map, vm, g["H"])
    vm0 = vm.copy()
    far = np.array([[0, 0, 0], [1, 2, 3]])  # label 3, far from every seed
    out1 = mod.update(data, out0[0], seg_map, vm, g["H"], far, out0[3], out0[4], out0[5], out0[6])
    assert np.array_equal(vm, vm0) and np.array_equal(out1[3], out0[3])
    with pytest.raises(ValueError):
        mod.update(data, out0[0], seg_map, vm, g["H"], np.array([[0, 0, 99]]), out0[3], out0[4], out0[5], out0[6])


def test_valuemap_with_band_labels_resumes_the_run():
    """A run stopped by maxSegmentSize returns a map holding labels 1 and 2; feeding it back continues the growth and ends
    exactly where the uninterrupted run ends (the reference itself dies on such a map, see test_oracle_golden.py)."""
    from arterynetwork_b200 import variationalRegionGrowing as mod
    g = load_golden("forest40")
    vm = g["value_map_in"].copy()
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mod.variationalRegionGrowing(g["data"], vm, H=g["H"], maxSegmentSize=120)
    assert "(Max segment size reached)" in buf.getvalue() and set(np.unique(vm)) >= {0, 1, 2, 3}
    stopped = int(np.count_nonzero(vm <= 1))
    assert 120 <= stopped < int(g["seg_bool"].sum())
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        _, seg_map, out = mod.variationalRegionGrowing(g["data"], vm, H=g["H"], maxSegmentSize=10 ** 9)
    assert out is vm and np.array_equal(vm, g["labels"]) and np.array_equal(seg_map == 1, g["seg_bool"])
    # a converged map fed back in is a fixed point: one decision, no flips
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mod.variationalRegionGrowing(g["data"], vm, H=g["H"], maxSegmentSize=10 ** 9)
    assert buf.getvalue().startswith("Finished at iteration 1\n") and np.array_equal(vm, g["labels"])
    bad = g["value_map_in"].copy(); bad[0, 0, 0] = 5
    with pytest.raises(ValueError):
        mod.variationalRegionGrowing(g["data"], bad)
    badf = g["value_map_in"].astype(np.float64); badf[0, 0, 0] = 3.7  # not a label
    with pytest.raises(ValueError):
        mod.variationalRegionGrowing(g["data"], badf)


def test_reference_self_tests_as_callables():
    """VRG:284-314: same inputs, same printed lines (known answers 16 iterations 80/80, 11 iterations 4169/4169)."""
    from arterynetwork_b200 import variationalRegionGrowing as mod
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        assert mod.test_StraightLine() is True
    assert buf.getvalue() == str(load_golden("straight_line")["stdout"]) + "Straight line test passed!\n"
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        assert mod.test_Sphere() is True
    assert buf.getvalue() == "4169\n" + str(load_golden("sphere")["stdout"]) + "Sphere test passed!\n"


def _noisy_case(seed):
    """A phantom on which the reference's order-dependent patterns do fire (heavy noise, coarse lattice, big seed)."""
    from arterynetwork_b200.phantom import make_phantom
    data, vm, _ = make_phantom((28, 30, 34), seed=seed, cell=(28, 30, 34), margin=3, depth=3, root_r2=9, min_len=6, max_len=12,
                               quantum=16, sigma_k=5)
    vm[8:20, 8:22, 8:26] = 0
    return data, vm


@pytest.mark.parametrize("mode", ["f64_dense", "f64_band", "index"])
@pytest.mark.parametrize("name", ["removal32", "cancel32", "excl32", "excl32_b", "forest40", "tube_fat_seed", "noisy0", "noisy1", "noisy2", "isolated"])
def test_order_dependence_counters_match_the_oracle(name, mode):
    """vrg_result's q_* fields against the oracle's quirk_potential (the same sets, counted on the device)."""
    from arterynetwork_b200.engine import VRGEngine
    from oracle.vrg_oracle import vrg_oracle
    if name.startswith("noisy"):
        data, vm = _noisy_case(int(name[5:]))
        H, max_seg = 2.25, 10 ** 12
    elif name == "isolated":  # stray seed voxels in the background leave at once and have no segmented neighbour left
        g = load_golden("tube_clean")
        data, vm, H, max_seg = g["data"], g["value_map_in"].copy(), g["H"], 10 ** 12
        vm[2, 2, 2] = 0
        vm[12, 13, 3:5] = 0
    else:
        g = load_golden(name)
        data, vm, H, max_seg = g["data"], g["value_map_in"], g["H"], g["max_segment_size"]
    ref = vrg_oracle(data, vm, H=H, max_segment_size=max_seg)
    with VRGEngine(data.shape, H=H, max_segment_size=max_seg, intensity=mode) as eng:
        eng.upload(np.asarray(data, dtype=np.float64), np.asarray(vm, dtype=np.uint8))
        eng.init()
        res = eng.run()
        assert np.array_equal(eng.labels(), ref["labels"]) and res["iterations"] == ref["iterations"]
    q = ref["quirk_potential"]
    assert (res["q_cancelled"], res["q_add_to_inside"], res["q_remove_to_outside"], res["q_cancel_repromoted"]) == (
        q["cancelled"], q["add_to_inside"], q["remove_to_outside"], q["cancel_repromoted"])
    if name in ("cancel32",):
        assert res["q_cancelled"] > 0
    if name.startswith("noisy"):
        assert res["q_add_to_inside"] + res["q_remove_to_outside"] + res["q_cancel_repromoted"] > 0
    if name == "isolated":
        assert res["q_remove_to_outside"] > 0


def test_dropin_warns_when_the_run_leaves_the_order_free_domain():
    from arterynetwork_b200 import variationalRegionGrowing as mod
    data, vm = _noisy_case(0)
    with warnings.catch_warnings(record=True) as w, contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("always")
        mod.variationalRegionGrowing(data, vm.astype(np.int64), maxSegmentSize=10 ** 12)
    assert any(issubclass(x.category, mod.VRGOrderDependenceWarning) for x in w)
    assert mod.LAST_RUN["q_add_to_inside"] + mod.LAST_RUN["q_remove_to_outside"] + mod.LAST_RUN["q_cancel_repromoted"] > 0
    g = load_golden("tube_clean")
    with warnings.catch_warnings(record=True) as w, contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("always")
        mod.variationalRegionGrowing(g["data"], g["value_map_in"].copy())
    assert not any(issubclass(x.category, mod.VRGOrderDependenceWarning) for x in w)


@pytest.mark.parametrize("name", ["forest40", "excl32", "sphere"])
def test_labels_hash_equals_the_oracle_hash(name):
    from arterynetwork_b200.engine import VRGEngine
    from oracle.c_oracle import hash_labels
    g = load_golden(name)
    with VRGEngine(g["data"].shape, H=g["H"], max_segment_size=g["max_segment_size"]) as eng:
        eng.upload(np.asarray(g["data"], dtype=np.float64), g["value_map_in"].astype(np.uint8))
        eng.init()
        eng.run()
        assert eng.labels_hash() == hash_labels(g["labels"])
        seg64 = eng.segmented_map_i64()
    assert seg64.dtype == np.int64 and np.array_equal(seg64 == 1, g["seg_bool"]) and set(np.unique(seg64)) <= {0, 1}
