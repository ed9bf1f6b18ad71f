"""GPU: the drop-in ``histogram`` class against the reference's unit tests (unittests/moments_histogram_one_dim_gc_ntot.py,
cited as T1:<line>) and golden vectors produced by the compiled reference."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
T1 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0])


def _hist(golden, smooth=1):
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    return oneDH.histogram.from_arrays(golden["testnc/lnpi"], golden["testnc/mom"], 1.0, [5.0, 0.0], smooth, volume=729.0)


def _t1_hist(golden):
    hist = _hist(golden)
    hist.data["mom"] = np.ones((2, 3, 2, 3, 3, 31), dtype=np.float64)
    hist.data["ln(PI)"] = T1.copy()
    hist.data["mom"][0, 1, 0, 0, :] = np.arange(0, 31)
    hist.data["mom"][1, 1, 0, 0, :] = np.arange(0, 31) * 2
    return hist


def test_norm_and_rew(golden):
    """T1:83-147."""
    hist = _hist(golden)
    lnpi_1 = copy.copy(hist.data["ln(PI)"])
    hist.normalize()
    assert abs(np.sum(np.exp(hist.data["ln(PI)"])) - 1.0) < 1e-12
    hist = _hist(golden)
    hist.reweight(0.0)
    x = lnpi_1 + np.arange(0, 31) * 1.0 * (0.0 - 5.0)
    x -= np.log(np.sum(np.exp(x)))
    assert np.all(np.abs(hist.data["ln(PI)"] - x) < 1e-12)
    assert np.max(np.abs(hist.data["ln(PI)"] - golden["testnc/rew0"])) < 1e-12
    assert np.all(hist.data["curr_mu"] == [0.0, -5.0])
    hist.reweight(-5.0)  # cumulative, from the modified data
    x = lnpi_1 + np.arange(0, 31) * 1.0 * (-5.0 - 5.0)
    x -= np.log(np.sum(np.exp(x)))
    assert np.all(np.abs(hist.data["ln(PI)"] - x) < 1e-12)
    assert np.max(np.abs(hist.data["ln(PI)"] - golden["testnc/rew0_m5"])) < 1e-12


def test_relextrema(golden):
    """T1:149-198 (integer arrays assigned straight into data['ln(PI)'])."""
    hist = _hist(golden)
    for arr, M, m in (([1, 2, 3, 2, 1, 2, 3, 4, 5], [2, 8], [0, 4]), ([1, 2, 3, 2, 1, 2], [2, 5], [0, 4]),
                      ([1, 2, 3, 2, 1], [2], [0, 4]), ([2, 1, 2, 3, 2, 1], [0, 3], [1, 5])):
        hist.data["ln(PI)"] = np.array(arr)
        hist.relextrema()
        assert np.all(hist.data["ln(PI)_maxima_idx"] == M)
        assert np.all(hist.data["ln(PI)_minima_idx"] == m)


def test_thermo(golden):
    """T1:200-237 + golden moment tensor of every phase."""
    hist = _t1_hist(golden)
    hist.thermo()
    th = hist.data["thermo"]
    assert len(th) == 2
    assert np.all(hist.data["ln(PI)_maxima_idx"] == [10, 25])
    lp = hist.data["ln(PI)"]
    assert abs(th[0]["F.E./kT"] - -np.log(np.sum(np.exp(lp[:20] - lp[0])))) < 1e-6
    assert abs(th[1]["F.E./kT"] - -np.log(np.sum(np.exp(lp[20:] - lp[0])))) < 1e-6
    assert abs(th[0]["n1"] - 9.99979018961) < 1e-6 and abs(th[0]["n2"] - 19.9995803792) < 1e-6
    assert abs(th[0]["ntot"] - 29.9993705688) < 1e-6
    assert abs(th[0]["x1"] - 9.99979018961 / 29.9993705688) < 1e-6
    assert abs(th[1]["n1"] - 25.0) < 1e-6 and abs(th[1]["n2"] - 50.0) < 1e-6 and abs(th[1]["ntot"] - 75.0) < 1e-6
    assert th[0]["bound_idx"] == (0, 20) and th[1]["bound_idx"] == (20, 31)
    for p in range(2):
        assert np.allclose(th[p]["mom"], golden["t1/thermo/mom"][p], rtol=1e-10, atol=0)
        assert abs(th[p]["F.E./kT"] - golden["t1/thermo/fe"][p]) < 1e-10 * abs(golden["t1/thermo/fe"][p])
    assert np.max(np.abs(lp - golden["t1/thermo/lnpi"])) < 1e-12


def test_thermo_complete(golden):
    """T1:239-267."""
    hist = _t1_hist(golden)
    hist.thermo(True, True)
    th = hist.data["thermo"]
    assert len(th) == 1
    assert abs(th[0]["n1"] - 10.0998274444) < 1e-6 and abs(th[0]["n2"] - 20.1996548887) < 1e-6
    assert abs(th[0]["ntot"] - 30.2994823331) < 1e-6
    assert np.allclose(th[0]["mom"], golden["t1/complete/mom"], rtol=1e-10, atol=0)
    assert abs(th[0]["F.E./kT"] - float(golden["t1/complete/fe"])) < 1e-10 * abs(float(golden["t1/complete/fe"]))


def test_is_safe(golden):
    """T1:269-291."""
    hist = _t1_hist(golden)
    hist.thermo()
    assert not hist.is_safe(10.0)
    assert hist.is_safe(5.0)
    assert hist.is_safe(10.0, True)
    assert not hist.is_safe(10.1, True)


def test_phase_eq(golden):
    """T1:293-308; the device root-finder must land within the reference solver's own x-tolerance (1e-4)."""
    hist = _hist(golden)
    hist.data["ln(PI)"] = T1.copy()
    before = copy.deepcopy(hist.data["ln(PI)"])
    eq_hist, err = hist.find_phase_eq(0.001, 5.0, reterr=True)
    assert abs(eq_hist.data["thermo"][0]["F.E./kT"] - eq_hist.data["thermo"][1]["F.E./kT"]) < 0.001
    assert abs(eq_hist.data["curr_mu"][0] - golden["t1/phase_eq/mu"][0]) < 2e-4
    assert abs(eq_hist.data["thermo"][0]["F.E./kT"] - eq_hist.data["thermo"][1]["F.E./kT"]) < 1e-9
    assert np.array_equal(hist.data["ln(PI)"], before)  # self untouched (GH:636)


def test_temp_extrap_1(golden):
    """T1:310-359."""
    hist = _t1_hist(golden)
    N = np.arange(0, 31)
    for a in ((0, 1, 1, 0), (0, 0, 0, 1), (1, 0, 0, 1)):
        hist.data["mom"][a + (slice(None),)] = N
    for a in ((1, 1, 1, 0), (0, 0, 1, 1), (1, 0, 1, 1)):
        hist.data["mom"][a + (slice(None),)] = N * 2
    hist.data["mom"][:, 1, :, 1, :] = 1.234 * np.ones(31)
    beta = 2.0 * hist.data["curr_beta"]
    hist.normalize()
    lnpi_orig = copy.copy(hist.data["ln(PI)"])
    ave_n2, ave_ntot, ave_u = 20.1996548887, 30.2994823331, 1.0
    dlnpi = hist.data["curr_mu"][0] * (N - ave_ntot) + (hist.data["curr_mu"][1] - hist.data["curr_mu"][0]) * (N * 2 - ave_n2) - (np.ones(31) - ave_u)
    ans = lnpi_orig + dlnpi * (beta - hist.data["curr_beta"])
    ans -= np.log(np.sum(np.exp(ans)))
    new_hist = hist.temp_extrap(beta, 1, 10.0, True, True, True)
    assert np.all(np.abs(ans - new_hist.data["ln(PI)"]) < 1e-12)
    assert abs(beta - new_hist.data["curr_beta"]) < 1e-12
    with pytest.raises(Exception):  # T1:361-376: max_order too low for order 2 with moments
        _hist(golden).temp_extrap(beta, 2, 10.0, True, True)


def test_temp_dmu_extrap_multi_matches_reference(golden, golden_meta):
    """T1:986-1043 pattern on the config-3 generator: grid == individually extrapolated == compiled reference."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    meta = golden_meta["c3"]
    for order in (1, 2):
        h = oneDH.histogram.from_arrays(golden["c3/lnpi"], golden["c3/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"])
        h.reweight(meta["mu1"])
        hs = h.temp_dmu_extrap_multi(golden["c3/betas"], golden["c3/dmus"], order, 10.0, True, True)
        for a in range(3):
            for b in range(3):
                g = golden["c3/o%d/%d_%d/lnpi" % (order, a, b)]
                assert np.max(np.abs(hs[a][b].data["ln(PI)"] - g)) < 1e-10
                assert hs[a][b].data["curr_beta"] == golden["c3/betas"][a]
                assert abs(hs[a][b].data["curr_mu"][1] - (meta["mu1"] + golden["c3/dmus"][b, 0])) < 1e-14
                hs[a][b].thermo(props=False)
                assert hs[a][b].data["ln(PI)_maxima_idx"].tolist() == golden["c3/o%d/%d_%d/maxima" % (order, a, b)].tolist()
                assert hs[a][b].data["ln(PI)_minima_idx"].tolist() == golden["c3/o%d/%d_%d/minima" % (order, a, b)].tolist()
        one = h.temp_dmu_extrap(golden["c3/betas"][0], golden["c3/dmus"][2], order, 10.0, True, True, True)
        assert np.max(np.abs(one.data["ln(PI)"] - hs[0][2].data["ln(PI)"])) < 1e-9
        assert h.data["curr_beta"] == meta["beta_ref"]  # original untouched


def test_moment_extrapolation_and_mix(golden, golden_meta):
    """skip_mom=False first-order moment update and mix() against the compiled reference."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    meta = golden_meta["c3"]
    h = oneDH.histogram.from_arrays(golden["c3/lnpi"], golden["c3/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"])
    h.reweight(meta["mu1"])
    hn = h.temp_dmu_extrap(1.03, np.array([0.6]), 1, 10.0, True, True, False)
    assert np.max(np.abs(hn.data["ln(PI)"] - golden["c3/mom1/lnpi"])) < 1e-10
    g = golden["c3/mom1/mom"]
    assert np.max(np.abs(hn.data["mom"] - g) / np.maximum(1.0, np.abs(g))) < 1e-12
    hn.thermo()
    for p in range(len(hn.data["thermo"])):
        assert np.allclose(hn.data["thermo"][p]["mom"], golden["c3/mom1/thermo/mom"][p], rtol=1e-9, atol=1e-12)
    ha = oneDH.histogram.from_arrays(golden["c3/lnpi"], golden["c3/mom"], 1.0, meta["mu_ref"], 5)
    hb = oneDH.histogram.from_arrays(golden["c3/lnpi"][:150] * 1.01, golden["c3/mom"][..., :150] * 0.99, 1.0, meta["mu_ref"], 5)
    hm = ha.mix(hb, [0.3, 0.9])
    assert np.allclose(hm.data["ln(PI)"], golden["mix/lnpi"], rtol=1e-14, atol=0)
    assert np.allclose(hm.data["mom"][1, 1, 0, 1, 1], golden["mix/mom_sample"], rtol=1e-14, atol=0)


def test_square_well_find_phase_eq(golden, golden_meta):
    """example/ntot/square_well/example.ipynb cell 14."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    meta = golden_meta["sw"]
    h = oneDH.histogram.from_arrays(golden["sw/lnpi"], golden["sw/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"], volume=meta["volume"])
    eq = h.find_phase_eq(meta["lnZ_tol"], meta["mu_guess"], meta["beta_ref"])
    assert eq.data["ln(PI)_maxima_idx"].tolist() == [10, 506] and eq.data["ln(PI)_minima_idx"].tolist() == [0, 253, 572]
    assert abs(meta["beta_ref"] * eq.data["curr_mu"][0] - meta["notebook_beta_mu"]) < 1e-4 * meta["beta_ref"]
    assert abs(eq.data["thermo"][0]["F.E./kT"] - eq.data["thermo"][1]["F.E./kT"]) < 1e-9
    assert abs(eq.data["thermo"][0]["F.E./kT"] - meta["notebook_fe"][0]) < 1e-3
    for order in (1, 2):
        hh = copy.deepcopy(h)
        hh.reweight(-4.47)
        hn = hh.temp_extrap(1.0 / 0.92, order, 10.0, False, True, True)
        assert np.max(np.abs(hn.data["ln(PI)"] - golden["sw/textrap%d" % order])) < 1e-9


def test_batched_entry_points(golden, golden_meta, oracle):
    """reweight_batch / find_phase_eq_batch: same answers as the scalar drop-in path and the tightened oracle."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    meta = golden_meta["sw"]
    lnpi, mom = golden["sw/lnpi"], golden["sw/mom"]
    n = len(lnpi)
    h = oneDH.histogram.from_arrays(lnpi, mom, meta["beta_ref"], meta["mu_ref"], meta["smooth"], volume=meta["volume"])
    out = h.reweight_batch(golden["sw/mu"])
    for k in range(len(golden["sw/mu"])):
        P = out["nphase"][k]
        assert out["max_idx"][k, :P].tolist() == golden["sw/%d/maxima" % k].tolist()
        assert np.allclose(out["fe"][k, :P], golden["sw/%d/fe" % k], rtol=1e-10, atol=0)
        assert np.allclose(out["avg"][k, :P, 2], golden["sw/%d/mom" % k][:, 0, 0, 0, 0, 1], rtol=1e-10, atol=0)
    betas = meta["beta_ref"] * np.array([1.0, 0.995, 1.004])
    res = h.find_phase_eq_batch(betas, meta["mu_guess"], order=2)
    assert np.all(res["code"] == 0) and np.all(np.abs(res["dfe"]) < 1e-9)
    A = oracle.taylor_coefficients(mom)
    N = np.arange(n, dtype=float)
    for k, beta in enumerate(betas):
        xb = beta - meta["beta_ref"]

        def coef_fn(mu):
            return np.stack([N, A["A_b"], A["A_bb"]]), np.array([xb * mu, xb, 0.5 * xb * xb])
        mu_t = oracle.find_phase_eq_tight(lnpi, N, meta["beta_ref"], 0.0, meta["smooth"], res["mu_coex"][k] - 0.01,
                                          res["mu_coex"][k] + 0.01, coef_fn=coef_fn)
        assert abs(res["mu_coex"][k] - mu_t) < 1e-10 * max(1.0, abs(mu_t))


def test_scalar_path_cache_follows_the_host_arrays(golden):
    """The scalar calls keep ln(PI), N and the moment tensor resident on the device between calls (engine.ScalarPath),
    keyed on the CONTENT of the host arrays, which stay the source of truth (T1:155, 206-209 overwrite hist.data[...]).
    Every kind of host-side change must reach the next call: a new array object, an in-place write of one element, a
    change confined to the low-order moments (invisible to a floating-point checksum next to the N^2 U^2 rows, the bug
    this test pins), two histogram objects of the same length taking turns, and a shorter ln(PI)."""
    from fhmcanalysis_b200 import engine
    a, b = _hist(golden, smooth=3), _hist(golden, smooth=3)
    n = len(a.data["ln(PI)"])
    lnpi0, mom0 = a.data["ln(PI)"].copy(), a.data["mom"].copy()

    def fresh_thermo(lnpi, mom):
        """the same state point through the batched kernels with nothing cached (a new DeviceHistogram per call)"""
        dh = engine.DeviceHistogram(lnpi, np.arange(n), 1.0, 5.0, smooth=3)
        h = dh.sweep(np.array([5.0]), pmax=8, lanes=32).host()
        P = int(h["nphase"][0])
        row = dh.lnpi_rows(dh.sweep(np.array([5.0]), pmax=8, lanes=32))[0].cpu().numpy()
        avg, _ = engine.phase_moments(row, mom.reshape(-1, n), h["bounds"][0, :P])
        return P, h["fe"][0, :P], avg

    def check(hist):
        lnpi, mom = np.array(hist.data["ln(PI)"], dtype=np.float64), np.array(hist.data["mom"], dtype=np.float64)
        hist.thermo()
        P, fe, avg = fresh_thermo(lnpi, mom)
        assert len(hist.data["thermo"]) == P
        for p in range(P):
            assert np.isclose(hist.data["thermo"][p]["F.E./kT"], fe[p], rtol=1e-12, atol=1e-12)
            assert np.allclose(hist.data["thermo"][p]["mom"].reshape(-1), avg[p], rtol=1e-12, atol=0, equal_nan=True)

    check(a)
    a.data["mom"][0, 1, 0, 0, 0, 7] += 1.0e-3             # one low-order entry, in place
    check(a)
    a.data["mom"] = mom0 * 1.0                            # new object, old content
    check(a)
    a.data["mom"][1, 1, 0, 0, 0] *= 1.5                   # a whole low-order row; the high-order rows (1e10 larger) untouched
    check(a)
    b.data["ln(PI)"] = lnpi0 + 1.0e-3 * np.sin(np.arange(n))      # the other object takes a turn with other arrays
    check(b)
    check(a)
    a.data["ln(PI)"][n // 2] += 0.25                      # in-place write of one bin
    check(a)
    a.reweight(5.01)
    check(a)
    b.data["ln(PI)"] = T1.astype(np.float64)              # shorter array, as T1:155 does
    b.metadata["smooth"] = 1
    b.relextrema()
    assert b.data["ln(PI)_maxima_idx"].tolist() == [10, 25] and b.data["ln(PI)_minima_idx"].tolist() == [0, 20, 30]


def test_scalar_path_more_extrema_than_its_record_holds(golden, oracle):
    """A noisy ln(PI) with smooth = 1 has dozens of windowed extrema: the scalar path's record (capacity 8) reports
    FHMC_E_CAPACITY and the drop-in methods fall back to the growing-capacity batched call; lists, bounds, F.E. and the
    27 moment averages equal the oracle's."""
    hist = _hist(golden, smooth=1)
    n = len(hist.data["ln(PI)"])
    rng = np.random.default_rng(11)
    lnpi = -1.0e-4 * (np.arange(n) - 0.5 * n) ** 2 + 0.05 * rng.normal(size=n)
    hist.data["ln(PI)"] = lnpi.copy()
    mom = np.asarray(hist.data["mom"], dtype=np.float64).reshape(-1, n)
    r = oracle.state_point(lnpi, np.arange(n), 1.0, 5.0, 5.0, 1, sel=mom)
    assert r["status"] == 0 and r["nphase"] > 8
    hist.relextrema()
    assert hist.data["ln(PI)_maxima_idx"].tolist() == r["max_idx"].tolist()
    assert hist.data["ln(PI)_minima_idx"].tolist() == r["min_idx"].tolist()
    hist.data["ln(PI)"] = lnpi.copy()
    hist.thermo()
    th = hist.data["thermo"]
    assert len(th) == r["nphase"]
    bounds = np.asarray(r["bounds"]).reshape(-1, 2)
    for p in range(r["nphase"]):
        assert tuple(th[p]["bound_idx"]) == tuple(bounds[p])
        assert np.isclose(th[p]["F.E./kT"], r["fe"][p], rtol=1e-10, atol=1e-12)
        assert np.allclose(th[p]["mom"].reshape(-1), r["avg"][p], rtol=1e-10, atol=0)


def test_scalar_path_cache_is_bounded(golden):
    """One resident ScalarPath per histogram length, at most MAX_CACHED of them (least recently used evicted)."""
    from fhmcanalysis_b200 import engine
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    old = engine.ScalarPath.MAX_CACHED
    engine.ScalarPath.MAX_CACHED = 3
    try:
        for n in (40, 41, 42, 43, 44, 41):
            x = np.arange(n, dtype=np.float64)
            h = oneDH.histogram.from_arrays(-0.01 * (x - 0.5 * n) ** 2, np.ones((1, 2, 1, 2, 2, n)), 1.0, [0.0], 2)
            h.normalize()
            assert np.isclose(np.log(np.sum(np.exp(h.data["ln(PI)"]))), 0.0, atol=1e-12)
            assert len(engine.ScalarPath._cache) <= 3
        assert [k[1] for k in engine.ScalarPath._cache][-1] == 41
    finally:
        engine.ScalarPath.MAX_CACHED = old
