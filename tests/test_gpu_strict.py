"""Strict (list-order) mode on the GPU, SURVEY.md section 8(f) N4: csrc/vrg_strict.cu through the C-ABI against
(a) outputs of the UNMODIFIED reference on inputs where it is order-dependent (tests/golden/strict) and (b) the pinned
list-order oracle (oracle/vrg_strict_oracle.c) on larger noisy inputs.  B200 only."""
import contextlib
import io

import numpy as np
import pytest

from strict_golden_util import load_strict, strict_names

pytestmark = pytest.mark.gpu
SUM_RTOL = 1e-11  # level-wise sums on the GPU, voxel-wise np.sum in the reference (VRG:154,239-241)


def _engine(g, **kw):
    from arterynetwork_b200.strict import StrictEngine
    return StrictEngine(g["data"].shape, H=g["H"], max_segment_size=g["max_segment_size"], **kw)


@pytest.mark.parametrize("name", strict_names())
def test_strict_mode_reproduces_the_reference(name):
    g = load_strict(name)
    with _engine(g) as eng:
        eng.init(g["data"], g["value_map_in"])
        for i, (idx, pin, pout) in enumerate(g["bands"]):  # the state every decision of the reference read
            b, spin, spout = eng.band()
            assert np.array_equal(b, idx), "band list order differs before decision %d" % (i + 1)
            np.testing.assert_allclose(spin, pin, rtol=SUM_RTOL, atol=0)
            np.testing.assert_allclose(spout, pout, rtol=SUM_RTOL, atol=0)
            res = eng.step()
        res = eng.run()
        assert res["iterations"] == g["iterations"]
        assert np.array_equal(eng.trace(), g["trace"])
        vm = eng.value_map()
        assert np.array_equal(vm, g["value_map"])  # stale band labels included
        assert np.array_equal(eng.segmented(), g["segmented"])  # the reference's row order
        assert np.array_equal(eng.segmented_map(), (g["value_map"] <= 1).astype(np.uint8))
        assert res["dropped"] == int(g["q3_dropped"])


def _noisy(seed, shape, noise, q, cube, excl_below=None):
    rng = np.random.default_rng(seed)
    d = np.zeros(shape)
    c = [s // 2 for s in shape]
    d[c[0] - 3:c[0] + 3, c[1] - 3:c[1] + 3, 2:shape[2] - 2] = 1.0
    d[2:shape[0] - 2, c[1] - 2:c[1] + 2, c[2] - 2:c[2] + 2] = 1.0
    k = np.round((d + rng.normal(0, noise, shape)) * q).astype(np.int64)
    vm = np.full(shape, 3, dtype=np.uint8)
    if excl_below is not None:
        vm[k <= excl_below] = 4
    vm[c[0] - cube // 2:c[0] + cube // 2, c[1] - cube // 2:c[1] + cube // 2, c[2] - cube // 2:c[2] + cube // 2] = 0
    return k / q, vm


CASES = [
    dict(seed=1, shape=(40, 36, 44), noise=0.4, q=8, cube=12),
    dict(seed=2, shape=(33, 47, 38), noise=0.3, q=16, cube=14, excl_below=-3),
    dict(seed=3, shape=(48, 48, 48), noise=0.45, q=8, cube=10),
    dict(seed=4, shape=(1, 64, 70), noise=0.3, q=8, cube=2),           # a 2-D image: 8 neighbours
    dict(seed=5, shape=(30, 30, 30), noise=0.45, q=4, cube=20, excl_below=-1),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "seed%d" % c["seed"])
def test_strict_mode_equals_the_list_order_oracle(case):
    from arterynetwork_b200.strict import StrictEngine
    from oracle.strict_oracle import vrg_strict_oracle
    data, vm = _noisy(**case)
    ms = 2500 if case["seed"] == 3 else data.size + 1
    o = vrg_strict_oracle(data, vm, H=2.25, max_segment_size=ms)
    assert o["min_margin"] > 1e-9
    with StrictEngine(data.shape, H=2.25, max_segment_size=ms) as eng:
        eng.init(data, vm)
        res = eng.run()
        assert res["iterations"] == o["iterations"] and res["exit_reason"] == o["exit"]
        assert np.array_equal(eng.trace(), o["trace"])
        assert np.array_equal(eng.value_map(), o["value_map"])
        assert np.array_equal(eng.segmented(), o["segmented"])
        b, _, _ = eng.band()
        assert np.array_equal(b, np.concatenate([o["inner"], o["outer"]]))
        pin, pout = eng.sums()
        np.testing.assert_allclose(pin, o["pin"], rtol=SUM_RTOL, atol=1e-300)
        np.testing.assert_allclose(pout, o["pout"], rtol=SUM_RTOL, atol=1e-300)
        assert res["skipped"] == o["skipped"] and res["dropped"] == o["dropped"]
        assert res["skipped"] > 0 or case["shape"][0] == 1  # the 3-D cases do exercise the order dependence


def test_dropin_list_order_switch(monkeypatch):
    """LIST_ORDER = True: the reference's stdout, its valueMap (stale labels) and its row order of `segmented`."""
    from arterynetwork_b200 import variationalRegionGrowing as mod
    g = load_strict("bar_maxseg")
    monkeypatch.setattr(mod, "LIST_ORDER", True)
    monkeypatch.setattr(mod, "MAX_SECONDS", None)
    vm = g["value_map_in"].copy()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        segmented, seg_map, vm_out = mod.variationalRegionGrowing(g["data"], vm, H=g["H"], maxSegmentSize=g["max_segment_size"])
    assert buf.getvalue() == str(g["stdout"])
    assert vm_out is vm and np.array_equal(vm, g["value_map"])
    assert np.array_equal(segmented, g["segmented"])
    assert seg_map.dtype == np.int64 and np.array_equal(seg_map, (g["value_map"] <= 1).astype(np.int64))
    # and the default (order-free) result on the same input is a different one: that is what the switch is for
    monkeypatch.setattr(mod, "LIST_ORDER", False)
    vm2 = g["value_map_in"].copy()
    with contextlib.redirect_stdout(io.StringIO()), pytest.warns(mod.VRGOrderDependenceWarning):
        mod.variationalRegionGrowing(g["data"], vm2, H=g["H"], maxSegmentSize=g["max_segment_size"])
    assert int((vm2 != vm).sum()) == int(g["orderfree_value_map_diff"])


def test_strict_mode_errors():
    from arterynetwork_b200.strict import StrictEngine
    from arterynetwork_b200 import _native as nat
    data = np.zeros((8, 8, 8))
    with StrictEngine(data.shape) as eng:
        with pytest.raises(ValueError):
            eng.init(data, np.full(data.shape, 3, dtype=np.uint8))  # no seed
        with pytest.raises(ValueError):
            eng.init(data, np.zeros(data.shape, dtype=np.uint8))  # seed without boundary
        bad = np.full(data.shape, 3, dtype=np.uint8)
        bad[4, 4, 4] = 0
        bad[0, 0, 0] = 2
        with pytest.raises(ValueError):
            eng.init(data, bad)  # band labels in the input
        nan = data.copy()
        nan[1, 1, 1] = np.nan
        ok = np.full(data.shape, 3, dtype=np.uint8)
        ok[4, 4, 4] = 0
        with pytest.raises(ValueError):
            eng.init(nan, ok)
        with pytest.raises(nat.LevelsError):
            big = np.random.default_rng(0).normal(size=(48, 48, 48))
            with StrictEngine(big.shape) as e2:
                vm = np.full(big.shape, 3, dtype=np.uint8)
                vm[24, 24, 24] = 0
                e2.init(big, vm)
        eng.init(data + (np.arange(512).reshape(8, 8, 8) % 3 == 0), ok)  # and the handle is still usable
        assert eng.run()["exit_reason"] in (0, 2, 3)
