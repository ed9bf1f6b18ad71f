"""GPU: the N_1-order-parameter drop-in class (moments/histogram/one_dim/n1/gc_hist.pyx) against vectors produced by
the compiled reference (tests/golden/make_golden_n1.py): reweight -> thermo -> is_safe, temp_mu_extrap (orders 1/2),
temp_mu_extrap_multi, find_phase_eq (also at another temperature / mu_2), mix, and the one-component case.
Integers bit-exact, fp64 at 1e-10 relative (coexistence per SURVEY 7.3: same integers, |d mu| <= 1e-4)."""
import json
import os

import numpy as np
import pytest

from fhmcanalysis_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "n1_vectors.npz")), json.load(open(os.path.join(HERE, "golden", "n1_vectors.json")))


def H():
    from fhmcanalysis_b200.moments.histogram.one_dim.n1.gc_hist import histogram
    return histogram


def make(g, mom=None, lnpi=None, mu_ref=None):
    v, meta = g
    s = meta["setup"]
    mom = synth.n1_two_comp_moments(s["n"], 3) if mom is None else mom
    return H().from_arrays(v["lnpi"] if lnpi is None else lnpi, mom, s["beta_ref"], s["mu_ref"] if mu_ref is None else mu_ref,
                           s["smooth"], s["volume"])


def sub(m, addr, phase=False):
    m = np.asarray(m)
    if phase:
        return np.stack([m[..., a[0] % m.shape[-5], a[1], a[2] % m.shape[-3], a[3], a[4]] for a in addr], axis=-1)
    return np.stack([m[..., a[0] % m.shape[-6], a[1], a[2] % m.shape[-4], a[3], a[4], :] for a in addr], axis=-2)


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.max(np.abs(b))))
    assert np.max(np.abs(a - b)) <= tol * scale, float(np.max(np.abs(a - b)) / scale)


def check_record(h, v, prefix, addr, tol=1e-10, fe_tol=None):
    P = len(h.data["thermo"])
    assert np.array_equal(h.data["ln(PI)_maxima_idx"], v[prefix + "/maxima"])
    assert np.array_equal(h.data["ln(PI)_minima_idx"], v[prefix + "/minima"])
    assert [tuple(h.data["thermo"][p]["bound_idx"]) for p in range(P)] == [tuple(b) for b in v[prefix + "/bounds"]]
    assert bool(h.is_safe()) == bool(v[prefix + "/safe"])
    close(h.data["ln(PI)"], v[prefix + "/lnpi"], fe_tol or tol)
    close([h.data["thermo"][p]["F.E./kT"] for p in range(P)], v[prefix + "/fe"], fe_tol or tol)
    if prefix + "/mom" in v:
        close(sub(np.array([h.data["thermo"][p]["mom"] for p in range(P)]), addr, True), v[prefix + "/mom"], fe_tol or tol)
        close([h.data["thermo"][p]["x1"] for p in range(P)], v[prefix + "/x1"], fe_tol or 1e-9)
        close([h.data["thermo"][p]["density"] for p in range(P)], v[prefix + "/density"], fe_tol or tol)


def test_reweight_thermo_is_safe(g):
    v, meta = g
    for k, mu in enumerate(v["A/mu"]):
        h = make(g)
        h.reweight(mu)
        h.thermo()
        check_record(h, v, "A/%d" % k, meta["addr"])
        assert np.allclose(h.data["curr_mu"], v["A/%d/curr_mu" % k])        # only mu_1 moves (N1:276)


def test_temp_mu_extrap(g):
    v, meta = g
    for k, (order, tb, tm, skip) in enumerate(meta["C"]):
        h = make(g)
        h.reweight(-0.25)
        e = h.temp_mu_extrap(tb, np.array(tm), order, 10.0, False, True, skip)
        close(e.data["ln(PI)"], v["C/%d/lnpi" % k])
        close(sub(e.data["mom"], meta["addr"]), v["C/%d/mom" % k])
        assert np.allclose(e.data["curr_mu"], v["C/%d/curr_mu" % k]) and e.data["curr_beta"] == tb
        assert h.data["curr_beta"] == 1.0                                   # clone=True leaves self alone
        e.thermo()
        check_record(e, v, "C/%d/thermo" % k, meta["addr"])
    h = make(g)
    with pytest.raises(Exception, match="No implementation"):
        h.temp_mu_extrap(1.01, np.array([-1.5]), 3, 10.0, False, True, True)
    e = h.temp_mu_extrap(1.01, np.array([-1.5]), 1)
    with pytest.raises(Exception, match="twice"):
        e.temp_mu_extrap(1.02, np.array([-1.5]), 1)


def test_temp_mu_extrap_multi(g):
    v, meta = g
    h = make(g)
    h.reweight(-0.25)
    for order in (1, 2):
        grid = h.temp_mu_extrap_multi(v["D/betas"], v["D/mus"], order, 10.0, False, False)
        close(np.array([[grid[i][j].data["ln(PI)"] for j in range(2)] for i in range(3)]), v["D/%d/lnpi" % order])
        close(sub(np.array([[grid[i][j].data["mom"] for j in range(2)] for i in range(3)]), meta["addr"]), v["D/%d/mom" % order])
        assert grid[2][1].data["curr_beta"] == v["D/betas"][2] and grid[2][1].data["curr_mu"][1] == v["D/mus"][1][0]


def test_find_phase_eq(g):
    v, meta = g
    for k, (tb, tm, order) in enumerate(meta["E"]):
        h = make(g)
        eq = h.find_phase_eq(1e-8, -0.2, tb, tm, order, 10.0, True)
        ref_mu = v["E/%d/curr_mu" % k]
        assert abs(eq.data["curr_mu"][0] - ref_mu[0]) <= 1e-4               # the reference's own fmin x-tolerance
        assert np.allclose(eq.data["curr_mu"][1:], ref_mu[1:])
        # same integers at mu*, free energies equal to each other far better than the reference's
        assert np.array_equal(eq.data["ln(PI)_maxima_idx"], v["E/%d/maxima" % k])
        assert [tuple(eq.data["thermo"][p]["bound_idx"]) for p in range(2)] == [tuple(b) for b in v["E/%d/bounds" % k]]
        fe = [eq.data["thermo"][p]["F.E./kT"] for p in range(2)]
        assert abs(fe[0] - fe[1]) < 1e-8
        close(fe, v["E/%d/fe" % k], 2e-3)
        assert h.data["curr_mu"][0] == -0.2                                  # self untouched


def test_find_phase_eq_min_width_is_smooth(g):
    """N1:1479 passes smooth (not 2*smooth) as the minimum phase width: a 7-bin wide phase counts at smooth=5."""
    from fhmcanalysis_b200 import engine
    x = np.full(60, -30.0)
    x[:8] = -((np.arange(8) - 3.0) ** 2)            # narrow phase: bins [0, 7)
    x[7:] = -((np.arange(7, 60) - 35.0) ** 2) / 40.0 - 1.0
    dh = engine.DeviceHistogram(x, np.arange(60), 1.0, 0.0, smooth=5, sel=["N"])
    wide = dh.find_phase_eq(np.array([0.0]), min_width=5).host()
    narrow = dh.find_phase_eq(np.array([0.0])).host()            # 2*smooth = 10 > 7 bins: no admissible pair
    assert int(wide["code"][0]) == 0 and abs(wide["dfe"][0]) < 1e-9
    assert int(narrow["code"][0]) != 0


def test_mix(g):
    v, meta = g
    a = make(g)
    mom = synth.n1_two_comp_moments(201, 3)
    b = make(g, mom=mom[..., :150] * 1.001, lnpi=v["F/lnpi_b"])
    m = a.mix(b, [0.3, 0.7])
    close(m.data["ln(PI)"], v["F/lnpi"])
    close(sub(m.data["mom"], meta["addr"]), v["F/mom"])
    assert "used_ke" not in a.metadata and "used_ke" not in m.metadata


def test_one_component(g):
    v, meta = g
    mom1 = synth.one_comp_moments(201, 3)
    for k, (mu, tb, order) in enumerate(meta["G"]):
        h = make(g, mom=mom1, mu_ref=[-0.2])
        h.reweight(mu)
        if order:
            h = h.temp_mu_extrap(tb, np.array([]), order, 10.0, False, True, False)
        h.thermo()
        check_record(h, v, "G/%d" % k, meta["addr"])


def test_batched_matches_scalar(g):
    """reweight_batch over (mu_1, beta, mu_2) equals the scalar reweight -> temp_mu_extrap -> thermo chain."""
    h = make(g)
    mu1 = np.array([-0.3, -0.25, -0.2])
    beta = np.array([1.02, 0.98, 1.0])
    mu2 = np.array([-1.45, -1.55, -1.5])
    r = h.reweight_batch(mu1, beta=beta, mu2=mu2, order=2, moments=("N", "U"))
    for s in range(3):
        e = make(g)
        e.reweight(mu1[s])
        e = e.temp_mu_extrap(beta[s], np.array([mu2[s]]), 2, 10.0, False, True, True)
        e.thermo(props=False)
        P = len(e.data["thermo"])
        assert r["code"][s] == 0 and r["nphase"][s] == P
        assert r["max_idx"][s, :P].tolist() == list(e.data["ln(PI)_maxima_idx"])
        close(r["fe"][s, :P], [e.data["thermo"][p]["F.E./kT"] for p in range(P)], 1e-9)
