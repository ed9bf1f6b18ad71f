"""GPU: every Taylor-extrapolation entry point of the drop-in class (temp_extrap orders 1-3, dmu_extrap, temp_dmu_extrap with
and without first_order_mom, kinetic-energy variant) and find_phase_eq at another (beta, dMu), against the compiled reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hist(golden, golden_meta, ke):
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    meta = golden_meta["h"]
    h = histogram.from_arrays(golden["h/lnpi"], golden["h/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"], ke=ke)
    h.reweight(meta["mu1"])
    return h


def _pick(t, sample):
    return np.array([t[tuple(a)] for a in sample])


def _mom_close(a, b, rtol=1e-9):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * np.max(np.abs(b)))) < rtol


@pytest.mark.parametrize("ke", [False, True])
def test_extrapolation_entry_points(golden, golden_meta, ke):
    meta = golden_meta["h"]
    sample = meta["sample"]
    tag = "h/ke%d" % int(ke)
    h = _hist(golden, golden_meta, ke)
    for order in ((1, 2, 3) if not ke else (1, 2)):
        hn = h.temp_extrap(meta["beta"], order, 10.0, True, True, False)
        assert np.max(np.abs(hn.data["ln(PI)"] - golden[tag + "/temp%d/lnpi" % order])) < 1e-9, order
        assert _mom_close(_pick(hn.data["mom"], sample), golden[tag + "/temp%d/mom" % order]), order
        assert hn.data["curr_beta"] == meta["beta"] and h.data["curr_beta"] == meta["beta_ref"]
    if ke:
        with pytest.raises(Exception):
            h.temp_extrap(meta["beta"], 3, 10.0, True, True, False)
    for order in (1, 2):
        hn = h.dmu_extrap(np.array([meta["dmu"]]), order, 10.0, True, True, False)
        assert np.max(np.abs(hn.data["ln(PI)"] - golden[tag + "/dmu%d/lnpi" % order])) < 1e-9
        assert _mom_close(_pick(hn.data["mom"], sample), golden[tag + "/dmu%d/mom" % order])
        assert abs(hn.data["curr_mu"][1] - hn.data["curr_mu"][0] - meta["dmu"]) < 1e-14
        for fom in (False, True):
            hn = h.temp_dmu_extrap(meta["beta"], np.array([meta["dmu"]]), order, 10.0, True, True, False, fom)
            assert np.max(np.abs(hn.data["ln(PI)"] - golden[tag + "/tdmu%d_%d/lnpi" % (order, int(fom))])) < 1e-9
            assert _mom_close(_pick(hn.data["mom"], sample), golden[tag + "/tdmu%d_%d/mom" % (order, int(fom))])
    with pytest.raises(Exception, match="twice"):
        h.temp_extrap(meta["beta"], 1, 10.0, True, False, True).temp_extrap(1.1, 1, 10.0, True, True, True)


def test_batched_taylor_sweep_matches_dropin(golden, golden_meta):
    """reweight_batch(beta=, dmu=) (coefficient rows inside the kernel) == the drop-in temp_dmu_extrap + thermo per point."""
    import copy
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    meta = golden_meta["h"]
    h0 = histogram.from_arrays(golden["h/lnpi"], golden["h/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"])
    mus = np.array([-1.95, -1.9, -1.85])
    betas = np.array([0.98, 1.0, 1.03])
    dmus = np.array([0.35, 0.4, 0.5])
    for order in (1, 2):
        out = h0.reweight_batch(mus, betas, dmus, order=order, moments=("N1", "N2", "U"), pmax=8)
        for k in range(3):
            h = copy.deepcopy(h0)
            h.reweight(mus[k])
            hn = h.temp_dmu_extrap(betas[k], np.array([dmus[k]]), order, 10.0, True, True, False, True)
            hn.thermo()
            P = len(hn.data["thermo"])
            assert out["code"][k] == 0 and out["nphase"][k] == P
            assert out["max_idx"][k, :P].tolist() == hn.data["ln(PI)_maxima_idx"].tolist()
            fe = np.array([hn.data["thermo"][p]["F.E./kT"] for p in range(P)])
            assert np.allclose(out["fe"][k, :P], fe, rtol=1e-9, atol=1e-9)
            n1 = np.array([hn.data["thermo"][p]["n1"] for p in range(P)])
            uu = np.array([hn.data["thermo"][p]["u"] for p in range(P)])
            assert np.allclose(out["avg"][k, :P, 0], n1, rtol=1e-8, atol=1e-10)
            assert np.allclose(out["avg"][k, :P, 2], uu, rtol=1e-8, atol=1e-10)


def test_find_phase_eq_other_conditions(golden, golden_meta):
    """find_phase_eq(lnZ_tol, mu_guess, beta, dMu, order) vs the reference's Nelder-Mead result: same extrema at coexistence,
    mu within the reference solver's x-tolerance (1e-4), residual far below the reference's."""
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    meta = golden_meta["eq"]
    for key, order, beta, dmu, status in meta["cases"]:
        assert status == "ok"
        h = histogram.from_arrays(golden["eq/lnpi"], golden["eq/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"])
        eq, err = h.find_phase_eq(1e-8, -3.0, beta, [dmu], order, 10.0, True, True)
        assert abs(eq.data["curr_mu"][0] - golden[key + "/mu"][0]) < 2e-4, key
        assert eq.data["ln(PI)_maxima_idx"].tolist() == golden[key + "/maxima"].tolist(), key
        assert eq.data["ln(PI)_minima_idx"].tolist() == golden[key + "/minima"].tolist(), key
        fe = [eq.data["thermo"][p]["F.E./kT"] for p in range(len(eq.data["thermo"]))]
        assert abs(fe[0] - fe[1]) < 1e-8 and err < 1e-16
        assert np.allclose(fe, golden[key + "/fe"], rtol=0, atol=5e-2)      # the reference stops ~1e-4 away in mu
        assert h.data["curr_mu"][0] == meta["mu_ref"][0]                     # self untouched


def _check_against_tight_oracle(oracle, lnpi, mom, betas, res, idx, smooth):
    """mu_coex within 1e-10 of the tightened oracle (brentq on the signed dF.E. of the same pair) and the record
    at mu_coex equal to the oracle's state point there: integers bit-exact, F.E. / averages 1e-10."""
    n = len(lnpi)
    A = oracle.taylor_coefficients(mom)
    N = np.arange(n, dtype=float)
    sel = np.stack([N, N * N, mom[0, 0, 0, 0, 1]])
    dsel = np.stack([np.zeros(n), np.zeros(n), -(mom[0, 0, 0, 0, 2] - mom[0, 0, 0, 0, 1] ** 2)])   # d<U>(N)/d beta (1 species)
    worst = 0.0
    for k in idx:
        xb = betas[k] - 1.0

        def coef_fn(mu, xb=xb):
            return np.stack([N, A["A_b"], A["A_bb"]]), np.array([xb * mu, xb, 0.5 * xb * xb])
        mu = res["mu_coex"][k]
        mu_t = None
        for half in (1e-7, 1e-6, 1e-5, 1e-4, 1e-3):     # the root nearest to the solver's
            try:
                mu_t = oracle.find_phase_eq_tight(lnpi, N, 1.0, 0.0, smooth, mu - half, mu + half, coef_fn=coef_fn)
                break
            except (RuntimeError, ValueError):
                continue
        assert mu_t is not None, k
        assert abs(mu - mu_t) <= 1e-10 * max(1.0, abs(mu_t)), (k, mu, mu_t)
        worst = max(worst, abs(mu - mu_t))
        coef, xi = coef_fn(mu)
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu, smooth, sel=sel + xb * dsel, coef=coef, xi=xi)
        P = r["nphase"]
        assert res["code"][k] == r["status"] == 0 and res["nphase"][k] == P
        assert res["max_idx"][k, :P].tolist() == r["max_idx"].tolist()
        assert res["min_idx"][k, :res["nmin"][k]].tolist() == r["min_idx"].tolist()
        assert res["bounds"][k, :P].tolist() == r["bounds"].tolist()
        assert bool(res["safe"][k]) == r["safe"]
        assert np.allclose(res["fe"][k, :P], r["fe"], rtol=1e-10, atol=1e-12)
        assert np.allclose(res["avg"][k, :P], r["avg"], rtol=1e-10, atol=0)
    return worst


def test_config4_coexistence_curve_at_size(oracle):
    """BASELINE config 4 AT SIZE: N_max = 2000 (2001 bins), smooth 10, 10^4 temperatures, order-2 beta extrapolation, every
    solve from the same cold guess (k_solve_lean).  >= 100 sampled solves against the tightened oracle; solves that ended on
    a jump of dF.E. carry FHMC_ST_JUMP (not 'converged'); the staged-continuation mode lands on the same roots."""
    from fhmcanalysis_b200 import _lib, synth
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    n, T, smooth, tol = 2001, 10000, 10, 1e-10
    lnpi, mom = synth.two_peak_lnpi(n, scale=2.0), synth.one_comp_moments(n, max_order=3)
    h4 = histogram.from_arrays(lnpi, mom, 1.0, [0.0], smooth)
    betas = 1.0 / np.linspace(0.90, 1.06, T)
    cold = h4.find_phase_eq_batch(betas, 0.0, order=2, lnZ_tol=tol, continuation=False)
    assert _lib.last_kernel() == "k_solve_lean"
    conv = cold["converged"]
    jump = (cold["code"] == 0) & ~conv
    assert conv.mean() > 0.98 and np.all(np.abs(cold["dfe"][conv]) <= tol)
    assert np.all(np.abs(cold["dfe"][jump]) > tol) and np.all((cold["status"][jump] & _lib.ST_JUMP) != 0)
    assert np.mean((cold["status"][conv] & _lib.ST_LEAN) != 0) > 0.99      # final records written by the lean evaluator
    idx = np.where(conv)[0][::83]
    assert len(idx) >= 100
    _check_against_tight_oracle(oracle, lnpi, mom, betas, cold, idx, smooth)
    # continuation along the curve (seeds first, interpolated guesses for the other solves): same roots, fewer evaluations.
    # The list is ordered in beta, so this is ONE launch of fhmc_find_phase_eq_curve (seeds and waiting inside the kernel).
    st = h4.find_phase_eq_batch(betas, 0.0, order=2, lnZ_tol=tol)           # automatic for one cold guess
    assert _lib.last_kernel() == "k_solve_lean"
    both = conv & st["converged"]
    assert both.mean() > 0.98 and st["iters"].mean() < 0.6 * cold["iters"].mean()
    same = np.abs(st["mu_coex"][both] - cold["mu_coex"][both]) <= 1e-9
    assert same.mean() > 0.995      # (a noisy ln(PI) near the critical temperature can hold more than one root)
    _check_against_tight_oracle(oracle, lnpi, mom, betas, st, np.where(st["converged"])[0][::89], smooth)
    again = h4.find_phase_eq_batch(betas, 0.0, order=2, lnZ_tol=tol)
    assert np.array_equal(again["mu_coex"], st["mu_coex"]) and np.array_equal(again["iters"], st["iters"])     # deterministic
    # a list that is NOT ordered in beta takes the host-staged form of the same continuation (two launches): same roots
    perm = np.random.default_rng(5).permutation(T)[:3000]
    sh = h4.find_phase_eq_batch(betas[perm], 0.0, order=2, lnZ_tol=tol)
    b2 = sh["converged"] & st["converged"][perm]
    assert b2.mean() > 0.97 and np.mean(np.abs(sh["mu_coex"][b2] - st["mu_coex"][perm][b2]) <= 1e-9) > 0.99
    # seed stride that does not divide the list, a list shorter than the stride, a two-entry list
    dh = h4.device_histogram(beta=betas, order=2, moments=("N",))
    for sub, stride in ((slice(100, 1131), 7), (slice(4000, 4011), 32), (slice(5000, 5002), 2)):
        bsub = betas[sub]
        a1 = dh._find_phase_eq_once(np.zeros_like(bsub), bsub, None, tol, None, 200, 4, None, None, None, seed_stride=stride).host()
        a0 = dh._find_phase_eq_once(np.zeros_like(bsub), bsub, None, tol, None, 200, 4, None, None, None).host()
        ok = (a1["code"] == 0) & (a0["code"] == 0) & (np.abs(a1["dfe"]) <= tol) & (np.abs(a0["dfe"]) <= tol)
        assert ok.mean() > 0.9 and np.all(a1["iters"] > 0)
        assert np.mean(np.abs(a1["mu_coex"][ok] - a0["mu_coex"][ok]) <= 1e-9) > 0.98
