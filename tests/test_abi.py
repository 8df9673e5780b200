"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/vrg_b200.h declares (no compute calls -- those need a GPU)."""
import os
import re

import numpy as np
import pytest

from arterynetwork_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vrg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vrg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libvrg_b200.so does not export %s" % n
    assert sorted(nat.EXPORTS) == names  # the ctypes binding covers the whole header
    assert lib.vrg_version() >= 100


def test_struct_layouts_match_header():
    # vrg_config: 3+2 int64, 2 int32, double, 2 int64, double ; vrg_result: 12 int64
    import ctypes
    assert ctypes.sizeof(nat.Config) == 8 * 5 + 4 * 2 + 8 + 8 * 2 + 8
    assert ctypes.sizeof(nat.Result) == 104


def test_create_rejects_bad_arguments_without_gpu():
    lib = nat.load()
    import ctypes
    cfg = nat.Config()
    cfg.shape[:] = (0, 4, 4)
    h = nat.vp()
    rc = lib.vrg_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == nat.ERR_ARG
    assert b"shape" in lib.vrg_last_error()
    with pytest.raises(ValueError):
        nat.check(rc)


def test_get_neighbours_matches_reference_order():
    """VRG:263-282 / SURVEY.md appendix A7."""
    from arterynetwork_b200.variationalRegionGrowing import get_neighbours
    n = get_neighbours([5, 5, 5], shape=(10, 10, 10))
    assert n.shape == (26, 3)
    assert n[:4].tolist() == [[4, 4, 4], [4, 4, 5], [4, 4, 6], [4, 5, 4]] and n[-1].tolist() == [6, 6, 6]
    assert get_neighbours([0, 0, 0], shape=(4, 4, 4)).shape == (7, 3)
    assert get_neighbours([1, 1], exclude_p=False).shape == (9, 2)
