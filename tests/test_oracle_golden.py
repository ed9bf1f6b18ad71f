"""CPU: pin the restatement oracle (oracle/fhmc_oracle.c) against golden vectors produced by the compiled
reference (tests/golden/make_golden.py) and against the reference's own unit-test known answers
(unittests/moments_histogram_one_dim_gc_ntot.py, cited per test)."""
import numpy as np
import pytest

from conftest import sel_rows

T1 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0], dtype=np.float64)


def test_reweight_known_answer(oracle, golden):
    """T1:100-147: cumulative reweighting 5 -> 0 -> -5, bit-identical to the reference."""
    lnpi, ntot = golden["testnc/lnpi"], golden["testnc/ntot"]
    x = oracle.reweight(lnpi, ntot, 0.0, 5.0, 1.0)
    assert np.array_equal(x, golden["testnc/rew0"])
    y = oracle.reweight(x, ntot, -5.0, 0.0, 1.0)
    assert np.array_equal(y, golden["testnc/rew0_m5"])
    ref = lnpi + np.arange(31) * (0.0 - 5.0)
    ref -= np.log(np.sum(np.exp(ref)))
    assert np.all(np.abs(x - ref) < 1e-12)


@pytest.mark.parametrize("k,maxima,minima", [(0, [2, 8], [0, 4]), (1, [2, 5], [0, 4]), (2, [2], [0, 4]), (3, [0, 3], [1, 5])])
def test_relextrema_integer_arrays(oracle, golden, k, maxima, minima):
    """T1:149-198 hard-coded index lists (smooth=1)."""
    st, M, m, _ = oracle.relextrema(golden["t1/relext%d/x" % k], 1)
    assert st == 0
    assert M.tolist() == maxima and m.tolist() == minima
    assert M.tolist() == golden["t1/relext%d/maxima" % k].tolist()
    assert m.tolist() == golden["t1/relext%d/minima" % k].tolist()


def test_thermo_known_answers(oracle, golden):
    """T1:200-237: maxima [10,25], split at 20, n1=9.99979018961, phase 2 n1 = 25."""
    n = 31
    sel = np.stack([np.arange(n, dtype=float), 2.0 * np.arange(n)])
    r = oracle.state_point(T1, np.arange(n), 1.0, 5.0, 5.0, 1, sel=sel)
    assert r["status"] == 0 and r["nphase"] == 2
    assert r["max_idx"].tolist() == [10, 25] and r["min_idx"].tolist() == [0, 20, 30]
    assert r["bounds"].tolist() == [[0, 20], [20, 31]]
    assert abs(r["avg"][0, 0] - 9.99979018961) < 1e-6 and abs(r["avg"][0, 1] - 19.9995803792) < 1e-6
    assert abs(r["avg"][1, 0] - 25.0) < 1e-6 and abs(r["avg"][1, 1] - 50.0) < 1e-6
    assert np.array_equal(r["fe"], golden["t1/thermo/fe"])
    assert np.array_equal(r["lnpi"], golden["t1/thermo/lnpi"])
    assert np.allclose(r["avg"][:, 0], golden["t1/thermo/mom"][:, 0, 1, 0, 0, 0], rtol=1e-13, atol=0)


def test_is_safe_quartet(oracle, golden):
    """T1:269-291: is_safe(10)=False, (5)=True, (10,complete)=True, (10.1,complete)=False."""
    x, _ = oracle.normalize(T1)
    x, _ = oracle.normalize(x)
    st, M, m, _ = oracle.relextrema(x, 1)
    L = oracle.lib()
    Mi = np.ascontiguousarray(M, dtype=np.int32)
    got = [bool(L.fo_is_safe(oracle._d(x), 31, oracle._i(Mi), len(Mi), c, comp)) for c, comp in ((10.0, 0), (5.0, 0), (10.0, 1), (10.1, 1))]
    assert got == [False, True, True, False] == golden["t1/is_safe"].tolist()


def test_sweep_matches_reference_bitwise(oracle, golden, golden_meta):
    """config-2 generator at N=301: reweight -> thermo -> is_safe per mu, all outputs identical to the reference."""
    lnpi, mom, mus = golden["c2/lnpi"], golden["c2/mom"], golden["c2/mu"]
    n = len(lnpi)
    sel = sel_rows(n, mom)
    for k, mu in enumerate(mus):
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu, golden_meta["c2"]["smooth"], sel=sel)
        assert r["status"] == 0
        assert np.array_equal(r["lnpi"], golden["c2/%d/lnpi" % k])
        assert r["max_idx"].tolist() == golden["c2/%d/maxima" % k].tolist()
        assert r["min_idx"].tolist() == golden["c2/%d/minima" % k].tolist()
        assert r["bounds"].tolist() == golden["c2/%d/bounds" % k].tolist()
        assert np.array_equal(r["fe"], golden["c2/%d/fe" % k])
        assert r["safe"] == bool(golden["c2/%d/safe" % k])
        gm = golden["c2/%d/mom" % k]
        assert np.allclose(r["avg"][:, 0], gm[:, 0, 1, 0, 0, 0], rtol=1e-12, atol=0)
        assert np.allclose(r["avg"][:, 1], gm[:, 0, 2, 0, 0, 0], rtol=1e-12, atol=0)
        assert np.allclose(r["avg"][:, 2], gm[:, 0, 0, 0, 0, 1], rtol=1e-12, atol=0)


def test_stress_cases_and_raises(oracle, golden, golden_meta):
    """smooth in {1,2,30,60}, noise in {0,5e-2}: same extrema, and an error status exactly where the reference raises."""
    n_raise = 0
    for key, smooth, noise, mu, outcome in golden_meta["stress"]:
        lnpi = golden[key + "/input"]
        n = len(lnpi)
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu, smooth)
        if outcome == "ok":
            assert r["status"] == 0, key
            assert r["max_idx"].tolist() == golden[key + "/maxima"].tolist(), key
            assert r["min_idx"].tolist() == golden[key + "/minima"].tolist(), key
            assert np.array_equal(r["fe"], golden[key + "/fe"]), key
        else:
            n_raise += 1
            assert r["status"] != 0, key
    assert n_raise >= 1


def test_monotone_branch(oracle, golden):
    """GH:382-386: monotone ln(PI) -> maxima [0], minima [last], one phase."""
    x = golden["mono/input"]
    r = oracle.state_point(x, np.arange(len(x)), 1.0, 0.0, 0.0, 5)
    assert r["status"] == 0
    assert r["max_idx"].tolist() == golden["mono/maxima"].tolist() == [0]
    assert r["min_idx"].tolist() == golden["mono/minima"].tolist() == [len(x) - 1]
    assert r["bounds"].tolist() == golden["mono/bounds"].tolist()


def test_square_well_notebook_answer(oracle, golden, golden_meta):
    """example/ntot/square_well/example.ipynb cell 14: beta*mu_coex = -4.47264655, F.E./kT = -9.28506932479 / -9.28546354084."""
    meta = golden_meta["sw"]
    lnpi = golden["sw/lnpi"]
    n = len(lnpi)
    beta = meta["beta_ref"]
    mu = float(golden["sw/phase_eq/mu"][0])
    assert abs(beta * mu - meta["notebook_beta_mu"]) < 5e-9
    r = oracle.state_point(lnpi, np.arange(n), beta, 0.0, mu, meta["smooth"])
    assert r["max_idx"].tolist() == [10, 506] and r["min_idx"].tolist() == [0, 253, 572]
    assert np.allclose(r["fe"], meta["notebook_fe"], rtol=0, atol=5e-10)
    assert np.array_equal(r["fe"], golden["sw/phase_eq/fe"])
    # the reference's own Nelder-Mead leaves |dF.E.| ~ 4e-4; the restated solver reproduces its mu
    mu_nm, err, _ = oracle.find_phase_eq_fmin(lnpi, np.arange(n), beta, 0.0, meta["smooth"], meta["lnZ_tol"], meta["mu_guess"])
    assert abs(mu_nm - mu) < 1e-12 and abs(err - float(golden["sw/phase_eq/err"])) < 1e-15
    # tightened oracle: signed root
    mu_t = oracle.find_phase_eq_tight(lnpi, np.arange(n), beta, 0.0, meta["smooth"], mu - 0.01, mu + 0.01)
    d, _ = oracle.signed_dfe(lnpi, np.arange(n), beta, 0.0, meta["smooth"], mu_t)
    assert abs(d) < 1e-9 and abs(mu_t - mu) < 1e-4


def test_t1_phase_eq(oracle, golden):
    """T1:293-308: find_phase_eq(0.001, 5.0) -> mu = 5.33435059 (compiled reference), |dF.E.| < 1e-3."""
    mu_nm, err, _ = oracle.find_phase_eq_fmin(T1, np.arange(31), 1.0, 5.0, 1, 0.001, 5.0)
    assert abs(mu_nm - float(golden["t1/phase_eq/mu"][0])) < 1e-12
    assert abs(golden["t1/phase_eq/fe"][0] - golden["t1/phase_eq/fe"][1]) < 1e-3


def test_taylor_closed_form_matches_reference(oracle, golden, golden_meta):
    """SURVEY 8(a) row 9 closed form vs temp_dmu_extrap_multi of the compiled reference (orders 1 and 2, 2 species)."""
    meta = golden_meta["c3"]
    lnpi, mom = golden["c3/lnpi"], golden["c3/mom"]
    n = len(lnpi)
    N = np.arange(n, dtype=float)
    beta_ref, mu1_ref, mu1 = meta["beta_ref"], meta["mu_ref"][0], meta["mu1"]
    d0 = meta["mu_ref"][1] - meta["mu_ref"][0]
    A = oracle.taylor_coefficients(mom, d0)
    for order in (1, 2):
        for a, beta in enumerate(golden["c3/betas"]):
            for b, dmu in enumerate(golden["c3/dmus"][:, 0]):
                xb, xd = beta - beta_ref, dmu - d0
                x = lnpi + beta_ref * (mu1 - mu1_ref) * N + xb * (mu1 * N + A["A_b"]) + xd * beta_ref * A["A_d"]
                if order == 2:
                    x = x + 0.5 * xb * xb * A["A_bb"] + xb * xd * (A["A_d"] + beta_ref * A["A_bd"]) + 0.5 * xd * xd * beta_ref ** 2 * A["A_dd"]
                x = x - np.log(np.sum(np.exp(x - x.max()))) - x.max()
                g = golden["c3/o%d/%d_%d/lnpi" % (order, a, b)]
                assert np.max(np.abs(x - g)) < 5e-12, (order, a, b)


def test_taylor_one_component_square_well(oracle, golden, golden_meta):
    """1-species closed form (d lnPI/d beta = mu1 N - U, d2 = f_UU) vs temp_extrap of the reference on real data."""
    meta = golden_meta["sw"]
    lnpi, mom = golden["sw/lnpi"], golden["sw/mom"]
    n = len(lnpi)
    N = np.arange(n, dtype=float)
    beta_ref, mu1 = meta["beta_ref"], -4.47
    A = oracle.taylor_coefficients(mom)
    xb = 1.0 / 0.92 - beta_ref
    for order in (1, 2):
        x = lnpi + beta_ref * (mu1 - 0.0) * N + xb * (mu1 * N + A["A_b"])
        if order == 2:
            x = x + 0.5 * xb * xb * A["A_bb"]
        x = x - np.log(np.sum(np.exp(x - x.max()))) - x.max()
        assert np.max(np.abs(x - golden["sw/textrap%d" % order])) < 2e-10 * max(1.0, np.max(np.abs(x)))


def test_mix(oracle, golden):
    lnpi, mom = golden["c3/lnpi"], golden["c3/mom"]
    got = oracle.mix(lnpi, lnpi[:150] * 1.01, [0.3, 0.9])
    assert np.allclose(got, golden["mix/lnpi"], rtol=1e-15, atol=0)
    gm = oracle.mix(mom[1, 1, 0, 1, 1], mom[1, 1, 0, 1, 1, :150] * 0.99, [0.3, 0.9])
    assert np.allclose(gm, golden["mix/mom_sample"], rtol=1e-15, atol=0)
