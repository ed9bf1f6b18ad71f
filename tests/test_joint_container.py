"""CPU: the 2-D joint-histogram container, restating the reference's unittests/moments_histogram_two_dim_joint.py
(entry setters and their size checks JT:27-135, enter/make JT:137-236, to_json/from_json JT:238-272).  The JSON
fixture of the reference (unittests/reference/joint_test.json) is restated as a literal."""
import json

import numpy as np
import pytest

NINF = -np.inf
JOINT_TEST_JSON = {"bounds_idx": [[1, 3], [0, 4]], "ln(PI)": [[NINF, 1.0, 2.0, 3.0, NINF], [0.0, 1.0, 2.0, 3.0, 4.0]],
                   "op_1": [1.0, 2.0], "op_2": [0.0, 1.0, 2.0, 3.0, 4.0],
                   "props": {"N2": [[0.0, 1.0, 4.0, 8.0, 0.0], [1.0, 1.0, 1.0, 1.0, 1.0]],
                             "U": [[0.0, 5.0, 5.0, 5.0, 0.0], [5.0, 5.0, 5.0, 5.0, 5.0]]}}


@pytest.fixture
def hist():
    import FHMCAnalysis.moments.histogram.two_dim.joint_hist as jH           # import path of JT:12
    return jH.joint_hist()


PROPS3 = {"U": np.array([5, 5, 5]), "N2": np.array([1, 4, 8])}


def test_entry_setters(hist):
    en = hist.entry()
    en.set_lnpi(np.array([1, 2, 3]), np.array([0, 1, 2]))
    en = hist.entry()
    for p in PROPS3:
        en.set_prop(p, PROPS3[p])
    en = hist.entry()
    en.set(np.array([1, 2, 3]), np.array([0, 1, 2]), PROPS3)


@pytest.mark.parametrize("lnpi,ntot,props", [
    ([1, 2, 3], [0, 1, 2], {"U": [5, 5, 5], "N2": [1, 4]}),          # JT:89 bad property length
    ([1, 2], [0, 1], {"U": [5, 5, 5], "N2": [1, 4, 8]}),              # JT:105 ln(PI) shorter than the properties
    ([1, 2, 3], [0, 1], {"U": [5, 5, 5], "N2": [1, 4, 8]}),           # JT:121 ln(PI) and order parameter differ
])
def test_bad_sets_raise(hist, lnpi, ntot, props):
    en = hist.entry()
    with pytest.raises(Exception):
        en.set(np.array(lnpi), np.array(ntot), {k: np.array(v) for k, v in props.items()})


def test_enter_and_make(hist):
    hist.enter(1, np.array([1, 2, 3]), np.array([0, 1, 2]), PROPS3)
    hist.make()
    assert np.all(hist.data["ln(PI)"] == [[1, 2, 3]])


def test_double_make_sorts_by_first_order_parameter(hist):
    lnpi, ntot = np.array([1, 2, 3]), np.array([0, 1, 2])
    hist.enter(2, lnpi, ntot, PROPS3)
    hist.enter(1, lnpi * 2, ntot, PROPS3)
    hist.make()
    assert np.all(hist.data["ln(PI)"] == [[2, 4, 6], [1, 2, 3]])


def test_make_with_ragged_entries(hist):
    hist.enter(1, np.array([1, 2, 3]), np.array([0, 1, 2]), PROPS3)
    hist.enter(2, np.array([1, 2, 3, 4]), np.array([0, 1, 2, 3]), {"U": np.array([5, 5, 5, 5]), "N2": np.array([1, 4, 8, 12])})
    hist.make()
    assert np.all(hist.data["ln(PI)"] == [[1, 2, 3, NINF], [1, 2, 3, 4]])


def _vary2(hist):
    hist.enter(1, np.array([1, 2, 3]), np.array([1, 2, 3]), PROPS3)
    hist.enter(2, np.array([0, 1, 2, 3, 4]), np.array([0, 1, 2, 3, 4]), {"U": np.array([5, 5, 5, 5, 5]), "N2": np.array([1, 1, 1, 1, 1])})
    hist.make()


def _check_vary2(hist):
    assert np.all(hist.data["ln(PI)"] == JOINT_TEST_JSON["ln(PI)"])
    assert np.all(hist.data["op_1"] == [1, 2]) and np.all(hist.data["op_2"] == [0, 1, 2, 3, 4])
    assert np.all(hist.data["bounds_idx"] == [[1, 3], [0, 4]])
    for p in ("U", "N2"):
        assert np.all(hist.data["props"][p] == JOINT_TEST_JSON["props"][p])


def test_make_with_offset_entries(hist):
    _vary2(hist)
    _check_vary2(hist)


def test_json_round_trip_and_reference_fixture(hist, tmp_path):
    _vary2(hist)
    p = str(tmp_path / "joint.json")
    hist.to_json(p)
    assert json.load(open(p)) == JOINT_TEST_JSON             # same document the reference wrote (JT:238-255)
    import FHMCAnalysis.moments.histogram.two_dim.joint_hist as jH
    h2 = jH.joint_hist()
    h2.from_json(p)
    _check_vary2(h2)
    # and the reference's own file layout (restated literal, -Infinity tokens included)
    p2 = str(tmp_path / "ref.json")
    with open(p2, "w") as fh:
        json.dump(JOINT_TEST_JSON, fh, indent=4, sort_keys=True)
    h3 = jH.joint_hist()
    h3.from_json(p2)
    _check_vary2(h3)


def test_combine_isopleth_grids_host_bookkeeping():
    """unittests/moments_histogram_one_dim_gc_ntot_isopleth.py:26-91: misaligned / unequal dmu2 raise, aligned grids
    are concatenated along mu1 with the shared column dropped."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_binary as gcB
    x1, y1 = np.meshgrid(np.linspace(-15, -10, 10), np.linspace(-5, -3, 5))
    z1 = x1 ** 2 + y1 ** 2
    for dmu2 in (np.linspace(-5, -4, 5), np.linspace(-5, -3, 6)):
        x2, y2 = np.meshgrid(np.linspace(-10, -5, 10), dmu2)
        with pytest.raises(Exception):
            gcB.combine_isopleth_grids([x2, x1], [y2, y1], [x2 ** 2 + y2 ** 2, z1])
    x2, y2 = np.meshgrid(np.linspace(-10, -5, 10), np.linspace(-5, -3, 5))
    x3, y3 = np.meshgrid(np.concatenate((np.linspace(-15, -10, 10), np.linspace(-10, -5, 10)[1:])), np.linspace(-5, -3, 5))
    Z, (X, Y) = gcB.combine_isopleth_grids([x2, x1], [y2, y1], [x2 ** 2 + y2 ** 2, z1])
    assert np.all(np.abs(X - x3) < 1e-9) and np.all(np.abs(Y - y3) < 1e-9) and np.all(np.abs(Z - (x3 ** 2 + y3 ** 2)) < 1e-9)
