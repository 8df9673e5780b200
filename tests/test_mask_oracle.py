"""CPU: the mask-side oracle (oracle/mask_oracle.*) against the SciPy golden fixtures, and its own invariants."""
import numpy as np
import pytest

from mask_golden_util import load_mask_case, mask_names, rule_names
from oracle import mask_oracle as mo


@pytest.mark.parametrize("name", mask_names())
def test_edt_oracle_equals_scipy_fixture(name):
    g = load_mask_case(name)
    assert np.array_equal(mo.edt_oracle(g["mask"]), g["edt"])  # sqrt of exact integers: bit-identical


@pytest.mark.parametrize("name", mask_names())
def test_label_oracle_equals_scipy_fixture(name):
    g = load_mask_case(name)
    labels, result = mo.label_oracle(g["mask"])
    assert np.array_equal(labels, g["labels"])
    assert len([r for r in result if r[0] != 0]) == g["n_components"]
    assert sum(s for _, s in result) == g["mask"].size


@pytest.mark.parametrize("name", rule_names())
def test_vessel_mask_oracle_equals_fixture(name):
    g = load_mask_case(name)
    out = mo.vessel_mask_oracle(g["vesselness"], g["brain"], min_size=g["min_size"])
    assert np.array_equal(out, g["vessel_mask"])
    assert np.array_equal(mo.edt_oracle(g["brain"]), g["brain_edt"])
    assert 0 < out.sum() < out.size


def test_scipy_agrees_when_installed():
    """The fixtures are SciPy outputs; where SciPy is importable, check a fresh random case directly."""
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(11)
    m = rng.random((9, 14, 23)) < 0.6
    m[4, 7, 11] = False
    assert np.array_equal(mo.edt_oracle(m), ndi.distance_transform_edt(m))
    lab, n = ndi.label(m, structure=np.ones((3, 3, 3), dtype=int))
    assert np.array_equal(mo.label_oracle(m)[0], lab)


def test_edt_oracle_rejects_mask_without_background():
    with pytest.raises(ValueError):
        mo.edt_oracle(np.ones((3, 3, 3), dtype=bool))
