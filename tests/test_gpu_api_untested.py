"""GPU: API the reference exposes but its own unit tests never drive -- ``histogram.coexisting`` (gc_hist.pyx:417-449), the
``collect=`` hook of ``thermo`` / ``find_phase_eq`` (gc_hist.pyx:485-486, 653, 662; collect.py:32-80 ``janus_collect``),
``isopleth.get_hist`` (gc_binary.pyx:292-353) and ``isopleth.make_grid`` (gc_binary.pyx:355-476) -- against vectors
produced by the COMPILED reference (tests/golden/make_golden_api.py -> api_vectors.npz/json)."""
import contextlib
import copy
import io
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def api():
    z = np.load(os.path.join(GOLDEN, "api_vectors.npz"))
    with open(os.path.join(GOLDEN, "api_vectors.json")) as fh:
        return {k: z[k] for k in z.files}, json.load(fh)


def _H():
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    return oneDH.histogram


def _check_record(h, vec, prefix, props=True):
    assert np.array_equal(np.asarray(h.data["ln(PI)_maxima_idx"]), vec[prefix + "/maxima"])
    assert np.array_equal(np.asarray(h.data["ln(PI)_minima_idx"]), vec[prefix + "/minima"])
    P = len(h.data["thermo"])
    assert P == len(vec[prefix + "/fe"])
    assert np.array_equal(np.array([h.data["thermo"][p]["bound_idx"] for p in range(P)]), vec[prefix + "/bounds"])
    fe = np.array([h.data["thermo"][p]["F.E./kT"] for p in range(P)])
    assert np.allclose(fe, vec[prefix + "/fe"], rtol=1e-10, atol=1e-12)
    assert np.allclose(h.data["ln(PI)"], vec[prefix + "/lnpi"], rtol=1e-10, atol=1e-11)
    assert bool(h.is_safe()) == bool(vec[prefix + "/safe"])
    if props:
        mom = np.array([h.data["thermo"][p]["mom"] for p in range(P)])
        assert np.allclose(mom, vec[prefix + "/mom"], rtol=1e-10, atol=1e-12)


def test_coexisting_matches_reference(api):
    """GH:417-449: index groups of phases whose free energies agree to rtol, at mu* and four offsets from it."""
    vec, meta = api
    m = meta["coex"]
    base = _H().from_arrays(vec["coex/lnpi"], vec["coex/mom"], m["beta_ref"], m["mu_ref"], m["smooth"])
    with pytest.raises(Exception) as ei:
        copy.deepcopy(base).coexisting()
    assert str(ei.value) == m["before_thermo"]
    for case in m["cases"]:
        h = copy.deepcopy(base)
        h.reweight(case["mu"])
        h.thermo()
        fe = [h.data["thermo"][p]["F.E./kT"] for p in range(len(h.data["thermo"]))]
        assert np.allclose(fe, case["fe"], rtol=1e-10)
        for rtol, want in case["rtol"].items():
            assert h.coexisting(float(rtol)) == want, (case["mu"], rtol)
    # one phase only: the reference returns [[]] (GH:437-438)
    h = copy.deepcopy(base)
    h.reweight(m["mu_star"] + 40.0)
    h.thermo()
    assert len(h.data["thermo"]) == 1 and h.coexisting() == [[]]


def test_thermo_collect_hook_matches_reference(api):
    """GH:485-486 with collect.py:32-80: three ln(PI) peaks merged into two macrophases; bounds, F.E. and all moment
    averages of the merged phases come from the device integrals over the rewritten bounds."""
    from FHMCAnalysis.moments.histogram.one_dim.ntot.collect import janus_collect
    vec, meta = api
    m = meta["collect"]
    base = _H().from_arrays(vec["collect/lnpi"], vec["collect/mom"], m["beta_ref"], m["mu_ref"], m["smooth"])
    merged = 0
    for k, mu in enumerate(m["mus"]):
        h = copy.deepcopy(base)
        h.reweight(mu)
        h.thermo()
        _check_record(h, vec, "collect/plain%d" % k)
        h = copy.deepcopy(base)
        h.reweight(mu)
        h.thermo(True, False, janus_collect)
        _check_record(h, vec, "collect/janus%d" % k)
        merged += len(vec["collect/plain%d/fe" % k]) > len(vec["collect/janus%d/fe" % k])
    assert merged >= 1  # the vectors do exercise the merge


def test_find_phase_eq_collect_hook_matches_reference(api):
    """GH:598-668 with a collect callback: the reference's Nelder-Mead over device-evaluated objectives.  mu* within the
    simplex' own x tolerance (1e-4), same integers at mu*, F.E. equal between the two macrophases."""
    from FHMCAnalysis.moments.histogram.one_dim.ntot.collect import janus_collect
    vec, meta = api
    m = meta["collect"]
    base = _H().from_arrays(vec["collect/lnpi"], vec["collect/mom"], m["beta_ref"], m["mu_ref"], m["smooth"])
    with contextlib.redirect_stdout(io.StringIO()):
        eq, err = base.find_phase_eq(1e-10, 0.0, 0.0, [], 1, 10.0, False, True, False, janus_collect)
    assert abs(float(eq.data["curr_mu"][0]) - m["eq_mu"]) <= 1e-4
    assert err <= max(10 * m["eq_err"], 1e-8)
    assert np.array_equal(np.asarray(eq.data["ln(PI)_maxima_idx"]), vec["collect/eq/maxima"])
    assert np.array_equal(np.asarray(eq.data["ln(PI)_minima_idx"]), vec["collect/eq/minima"])
    P = len(eq.data["thermo"])
    assert P == 2 and np.array_equal(np.array([eq.data["thermo"][p]["bound_idx"] for p in range(P)]), vec["collect/eq/bounds"])
    fe = np.array([eq.data["thermo"][p]["F.E./kT"] for p in range(P)])
    assert abs(fe[0] - fe[1]) < 1e-4
    assert np.allclose(fe, vec["collect/eq/fe"], atol=2e-3)  # F.E. moves with mu* inside the x tolerance (slope ~ <N>)


def test_janus_collect_two_maxima_is_a_no_op(api):
    """collect.py:56-58 says two peaks are left as they are; as written the reference then falls through to an unassigned
    name (collect.py:77, UnboundLocalError -- recorded in api_vectors.json).  The drop-in follows the stated intent."""
    from FHMCAnalysis.moments.histogram.one_dim.ntot.collect import janus_collect
    vec, meta = api
    assert meta["collect"]["two_maxima"] == "UnboundLocalError"
    m = meta["coex"]
    h = _H().from_arrays(vec["coex/lnpi"], vec["coex/mom"], m["beta_ref"], m["mu_ref"], m["smooth"])
    h.thermo()
    before = (list(h.data["ln(PI)_maxima_idx"]), list(h.data["ln(PI)_minima_idx"]))
    janus_collect(hist=h)
    assert (list(h.data["ln(PI)_maxima_idx"]), list(h.data["ln(PI)_minima_idx"])) == before
    h2 = _H().from_arrays(vec["coex/lnpi"], vec["coex/mom"], m["beta_ref"], m["mu_ref"], m["smooth"])
    with pytest.raises(Exception, match="not been segmented"):
        janus_collect(hist=h2)


def _iso(golden, golden_meta, order):
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_binary as gcB
    gm = golden_meta["iso"]
    hists = [oneDH.histogram.from_arrays(golden["iso/lnpi"][k], golden["c3/mom"], gm["beta_ref"],
                                         [gm["mu1_ref"], gm["mu1_ref"] + d2], gm["smooth"], volume=gm["volume"])
             for k, d2 in enumerate(gm["dmu2"])]
    with contextlib.redirect_stdout(io.StringIO()):
        return gcB.isopleth(hists, gm["beta_ref"], order)


def test_isopleth_get_hist_matches_reference(api, golden, golden_meta):
    """GB:292-353: reweight -> dmu2 extrapolation of the neighbours -> distance-weighted mix, between two stored
    histograms, beyond either end (single neighbour) and exactly on one."""
    vec, meta = api
    iso = _iso(golden, golden_meta, 1)
    for k, (mu1, d2) in enumerate(meta["iso"]["points"]):
        with contextlib.redirect_stdout(io.StringIO()):
            h = iso.get_hist(mu1, d2, meta["iso"]["m"])
        p = "iso/o1/get%d" % k
        assert np.allclose(h.data["ln(PI)"], vec[p + "/lnpi"], rtol=1e-10, atol=1e-11)
        assert np.allclose(h.data["mom"][0, 1, 0, 0, 0], vec[p + "/mom_n1"], rtol=1e-10, atol=1e-12)
        assert np.allclose(h.data["mom"][0, 0, 0, 0, 1], vec[p + "/mom_u"], rtol=1e-10, atol=1e-12)
        assert np.allclose(h.data["curr_mu"], vec[p + "/curr_mu"], rtol=0, atol=1e-14)


def test_isopleth_make_grid_matches_reference(api, golden, golden_meta):
    """GB:355-476: the cell-by-cell grid (get_hist -> thermo -> is_safe -> most stable phase)."""
    vec, meta = api
    g = meta["iso"]["grid"]
    iso = _iso(golden, golden_meta, 1)
    with contextlib.redirect_stdout(io.StringIO()):
        Z, (X, Y) = iso.make_grid(g["mu1_bounds"], g["dmu2_bounds"], g["delta"], meta["iso"]["m"])
    assert np.allclose(X, vec["iso/o1/grid/X"]) and np.allclose(Y, vec["iso/o1/grid/Y"])
    assert np.array_equal(Z == 0, vec["iso/o1/grid/x1"] == 0)  # same cells skipped
    assert np.allclose(Z, vec["iso/o1/grid/x1"], rtol=1e-9, atol=1e-12)
    assert np.allclose(iso.data["density"], vec["iso/o1/grid/density"], rtol=1e-9, atol=1e-12)
    assert np.allclose(iso.data["F.E./kT"], vec["iso/o1/grid/fe"], rtol=1e-9, atol=1e-10)
