"""GPU parity: the CUDA sweep (through the C ABI) against the oracle and the reference's golden vectors.
Tolerances: integers (extrema, bounds, nphase, is_safe, status) bit-exact; fp64 outputs 1e-10 relative."""
import numpy as np
import pytest

from conftest import sel_rows

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _dh(lnpi, mom, beta_ref, mu1_ref, smooth, **kw):
    from fhmcanalysis_b200 import engine
    n = len(lnpi)
    s = sel_rows(n, mom)
    dh = engine.DeviceHistogram(lnpi, np.arange(n), beta_ref, mu1_ref, smooth=smooth, sel=["N", s[1], s[2]], **kw)
    dh.ensure_hull()  # lanes=1 -> one-pass fast kernel, lanes=-1 -> generic one-lane kernel
    return dh


def _check_record(h, k, r, nsel=3):
    P = r["nphase"]
    assert h["code"][k] == r["status"]
    if r["status"] != 0:
        return
    assert h["nphase"][k] == P and h["nmin"][k] == len(r["min_idx"])
    assert h["max_idx"][k, :P].tolist() == r["max_idx"].tolist()
    assert h["min_idx"][k, :h["nmin"][k]].tolist() == r["min_idx"].tolist()
    assert h["bounds"][k, :P].tolist() == r["bounds"].tolist()
    assert bool(h["safe"][k]) == r["safe"]
    assert np.allclose(h["fe"][k, :P], r["fe"], rtol=RTOL, atol=0)
    if nsel:
        assert np.allclose(h["avg"][k, :P, :nsel], r["avg"][:, :nsel], rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("lanes", [1, -1, 4, 32])
def test_golden_config2_sweep(golden, golden_meta, lanes):
    lnpi, mom, mus = golden["c2/lnpi"], golden["c2/mom"], golden["c2/mu"]
    dh = _dh(lnpi, mom, 1.0, 0.0, golden_meta["c2"]["smooth"])
    res = dh.sweep(mus, pmax=4, lanes=lanes)
    h = res.host()
    rows = dh.lnpi_rows(res).cpu().numpy()
    for k in range(len(mus)):
        g = {q: golden["c2/%d/%s" % (k, q)] for q in ("lnpi", "maxima", "minima", "fe", "bounds", "safe", "mom")}
        P = len(g["maxima"])
        assert h["code"][k] == 0 and h["nphase"][k] == P
        assert h["max_idx"][k, :P].tolist() == g["maxima"].tolist()
        assert h["min_idx"][k, :h["nmin"][k]].tolist() == g["minima"].tolist()
        assert h["bounds"][k, :P].tolist() == g["bounds"].tolist()
        assert bool(h["safe"][k]) == bool(g["safe"])
        assert np.allclose(h["fe"][k, :P], g["fe"], rtol=RTOL, atol=0)
        assert np.allclose(h["avg"][k, :P, 0], g["mom"][:, 0, 1, 0, 0, 0], rtol=RTOL, atol=0)
        assert np.allclose(h["avg"][k, :P, 1], g["mom"][:, 0, 2, 0, 0, 0], rtol=RTOL, atol=0)
        assert np.allclose(h["avg"][k, :P, 2], g["mom"][:, 0, 0, 0, 0, 1], rtol=RTOL, atol=0)
        assert np.max(np.abs(rows[k] - g["lnpi"])) < 1e-11


@pytest.mark.parametrize("lanes", [1, -1, 4, 32])
def test_stress_cases_match_reference_and_oracle(golden, golden_meta, oracle, lanes):
    from fhmcanalysis_b200 import engine
    for key, smooth, noise, mu, outcome in golden_meta["stress"]:
        lnpi = golden[key + "/input"]
        n = len(lnpi)
        dh = engine.DeviceHistogram(lnpi, np.arange(n), 1.0, 0.0, smooth=smooth)
        h = dh.sweep_auto(np.array([mu]), pmax=4, lanes=lanes).host()
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu, smooth)
        if outcome == "ok":
            assert h["code"][0] == 0, key
            P = len(golden[key + "/maxima"])
            assert h["max_idx"][0, :P].tolist() == golden[key + "/maxima"].tolist(), key
            assert h["min_idx"][0, :h["nmin"][0]].tolist() == golden[key + "/minima"].tolist(), key
            assert np.allclose(h["fe"][0, :P], golden[key + "/fe"], rtol=RTOL, atol=0), key
        else:
            assert h["code"][0] == r["status"] != 0, key


@pytest.mark.parametrize("lanes", [1, 32])
def test_monotone_and_t1_arrays(golden, oracle, lanes):
    from fhmcanalysis_b200 import engine
    x = golden["mono/input"]
    dh = engine.DeviceHistogram(x, np.arange(len(x)), 1.0, 0.0, smooth=5)
    h = dh.sweep(np.array([0.0]), pmax=4, lanes=lanes).host()
    assert h["code"][0] == 0 and h["max_idx"][0, 0] == 0 and h["min_idx"][0, 0] == len(x) - 1
    assert h["bounds"][0, 0].tolist() == [0, len(x)]
    assert np.allclose(h["fe"][0, 0], golden["mono/fe"][0], rtol=RTOL)
    # T1:155-198 integer arrays through relextrema() semantics (compare on the raw array)
    for k in range(4):
        arr = golden["t1/relext%d/x" % k]
        dh = engine.DeviceHistogram(arr, np.arange(len(arr)), 1.0, 5.0, smooth=1)
        h = dh.sweep(np.array([5.0]), pmax=8, lanes=lanes, compare_raw=True).host()
        P = h["nphase"][0]
        assert h["max_idx"][0, :P].tolist() == golden["t1/relext%d/maxima" % k].tolist()
        assert h["min_idx"][0, :h["nmin"][0]].tolist() == golden["t1/relext%d/minima" % k].tolist()


def test_t1_thermo_and_is_safe(golden):
    from fhmcanalysis_b200 import engine
    t1 = golden["t1/thermo/lnpi"]  # normalised 31-bin two-peak array (T1:207)
    n = len(t1)
    N = np.arange(n, dtype=float)
    dh = engine.DeviceHistogram(t1, N, 1.0, 5.0, smooth=1, sel=["N", 2 * N])
    for cutoff, expect in ((10.0, False), (5.0, True)):
        h = dh.sweep(np.array([5.0]), pmax=4, cutoff=cutoff).host()
        assert h["max_idx"][0, :2].tolist() == [10, 25] and h["min_idx"][0, :3].tolist() == [0, 20, 30]
        assert bool(h["safe"][0]) == expect
    assert abs(h["avg"][0, 0, 0] - 9.99979018961) < 1e-6 and abs(h["avg"][0, 1, 0] - 25.0) < 1e-6
    assert abs(h["avg"][0, 0, 1] - 19.9995803792) < 1e-6
    for cutoff, expect in ((10.0, True), (10.1, False)):
        h = dh.sweep(np.array([5.0]), pmax=1, cutoff=cutoff, complete=True).host()
        assert bool(h["safe"][0]) == expect and h["nphase"][0] == 1
    assert abs(h["avg"][0, 0, 0] - 10.0998274444) < 1e-6


def test_square_well_real_data(golden, golden_meta):
    meta = golden_meta["sw"]
    lnpi, mom = golden["sw/lnpi"], golden["sw/mom"]
    dh = _dh(lnpi, mom, meta["beta_ref"], 0.0, meta["smooth"])
    mus = np.concatenate([golden["sw/mu"], golden["sw/phase_eq/mu"]])
    for lanes in (1, 32):
        h = dh.sweep(mus, pmax=4, lanes=lanes).host()
        for k in range(len(mus)):
            pre = "sw/%d" % k if k < len(mus) - 1 else "sw/phase_eq"
            P = len(golden[pre + "/maxima"])
            assert h["code"][k] == 0
            assert h["max_idx"][k, :P].tolist() == golden[pre + "/maxima"].tolist()
            assert h["min_idx"][k, :h["nmin"][k]].tolist() == golden[pre + "/minima"].tolist()
            assert bool(h["safe"][k]) == bool(golden[pre + "/safe"])
            assert np.allclose(h["fe"][k, :P], golden[pre + "/fe"], rtol=RTOL, atol=0)
            assert np.allclose(h["avg"][k, :P, 0], golden[pre + "/mom"][:, 0, 1, 0, 0, 0], rtol=RTOL, atol=0)
            assert np.allclose(h["avg"][k, :P, 2], golden[pre + "/mom"][:, 0, 0, 0, 0, 1], rtol=RTOL, atol=0)


@pytest.mark.parametrize("lanes,S", [(1, 4096), (-1, 2048), (4, 1024), (32, 256)])
def test_seeded_sweep_vs_oracle(oracle, lanes, S):
    """BASELINE config 2 generator at full N=1001; every 13th state point checked against the oracle."""
    from fhmcanalysis_b200 import synth
    n = 1001
    lnpi, mom = synth.two_peak_lnpi(n), synth.one_comp_moments(n)
    dh = _dh(lnpi, mom, 1.0, 0.0, 10)
    mus = np.linspace(-0.03, 0.03, S)
    h = dh.sweep(mus, pmax=4, lanes=lanes).host()
    sel = sel_rows(n, mom)
    for k in range(0, S, 13):
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mus[k], 10, sel=sel)
        _check_record(h, k, r)


def test_wide_mu_range_unlikely_phases(oracle):
    """Large |beta dmu N|: phases with weight far below 1e-300 keep their free energy (rescue path)."""
    from fhmcanalysis_b200 import synth
    n = 1001
    lnpi, mom = synth.two_peak_lnpi(n, noise=0.0), synth.one_comp_moments(n)
    dh = _dh(lnpi, mom, 1.0, 0.0, 10)
    mus = np.array([-3.0, -1.5, -0.5, 0.5, 1.5, 3.0])
    h = dh.sweep_auto(mus, pmax=4).host()
    sel = sel_rows(n, mom)
    for k, mu in enumerate(mus):
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu, 10, sel=sel)
        P = r["nphase"]
        assert h["code"][k] == r["status"] == 0
        assert h["max_idx"][k, :P].tolist() == r["max_idx"].tolist()
        assert np.allclose(h["fe"][k, :P], r["fe"], rtol=RTOL, atol=0)


def test_capacity_status_and_auto_growth(oracle):
    from fhmcanalysis_b200 import engine, synth
    n = 301
    lnpi = synth.two_peak_lnpi(n, noise=5e-2, scale=0.3, seed=7)
    dh = engine.DeviceHistogram(lnpi, np.arange(n), 1.0, 0.0, smooth=1)
    h = dh.sweep(np.array([0.0]), pmax=2).host()
    assert h["code"][0] == 8
    h = dh.sweep_auto(np.array([0.0]), pmax=2).host()
    r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, 0.0, 1)
    assert h["code"][0] == r["status"]
    if r["status"] == 0:
        assert h["max_idx"][0, :r["nphase"]].tolist() == r["max_idx"].tolist()


def test_full_size_properties():
    """BASELINE config 2 at full size (1e6 state points): size-independent properties.
    (i) sum_p exp(-(fe_p) + ...) consistency: ln sum_p exp(-fe_p) == lnnorm - u_0 ; (ii) <N> monotone in mu;
    (iii) phase bounds tile [0,n); (iv) identical results from the 1-lane and 4-lane kernels."""
    import torch
    from fhmcanalysis_b200 import synth
    n = 1001
    lnpi, mom = synth.two_peak_lnpi(n), synth.one_comp_moments(n)
    dh = _dh(lnpi, mom, 1.0, 0.0, 10)
    S = 1000000
    mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
    r1 = dh.sweep(mu, pmax=4, lanes=1)
    h = r1.host()
    assert np.all(h["code"] == 0) and np.all(h["safe"])
    P = h["nphase"]
    assert P.min() >= 1 and P.max() <= 4
    idx = np.arange(S)
    assert np.all(h["bounds"][:, 0, 0] == 0) and np.all(h["bounds"][idx, P - 1, 1] == n)
    mask = np.arange(4)[None, :] < P[:, None]
    # (slots p >= nphase are never written: mask them BEFORE they reach exp)
    w = np.where(mask, np.exp(-np.where(mask, h["fe"] - h["fe"][:, :1], 0.0)), 0.0)
    lse = np.log(np.sum(w, axis=1)) - h["fe"][:, 0]
    u0 = lnpi[0]  # N_0 = 0: reweighting leaves bin 0 unchanged
    assert np.max(np.abs(lse - (h["lnnorm"] - u0))) < 1e-9
    ntot = np.sum(np.where(mask, h["avg"][:, :, 0], 0.0) * w, axis=1) / np.sum(w, axis=1)
    assert np.all(np.diff(ntot) > -1e-9)
    sub = slice(0, S, 97)
    h4 = dh.sweep(mu[sub].contiguous(), pmax=4, lanes=4).host()
    assert np.array_equal(h4["max_idx"][:, :1], h["max_idx"][sub, :1]) and np.array_equal(h4["nphase"], h["nphase"][sub])
    assert np.allclose(h4["fe"][:, 0], h["fe"][sub, 0], rtol=1e-12, atol=0)


def test_headline_kernel_prod2_pinned_directly(oracle):
    """The kernel bench.py times -- k_sweep_prod2<2,1>: BASELINE config 2 (N = 1001 bins, smooth 10, 10^6-point mu sweep in
    [-0.03, 0.03], averaged quantities exactly (N, N^2) as bench.py builds them) -- checked DIRECTLY against the C oracle
    (>= 2000 strided records) and against the compiled reference itself (>= 200 records: reweight -> thermo -> is_safe,
    GH:71-78, 317-415, 498-596).  Integers bit-exact, F.E./kT and averages 1e-10 relative."""
    import copy
    import torch
    from fhmcanalysis_b200 import _lib, synth
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    from oracle import ref
    n, smooth, S = 1001, 10, 1000000
    lnpi, mom = synth.two_peak_lnpi(n), synth.one_comp_moments(n)
    hist = histogram.from_arrays(lnpi, mom, 1.0, [0.0], smooth)
    dh = hist.device_histogram(moments=("N", "N2"), device="cuda:0")      # bench.py:run_gpu_arm builds exactly this
    mu = np.linspace(-0.03, 0.03, S)
    res = dh.sweep(torch.from_numpy(mu).to("cuda:0"), pmax=4)
    assert _lib.last_kernel() == "k_sweep_prod2" and dh.desc.mu_recurrence == 3 and dh.n_sel == 2
    h = res.host()
    assert np.mean((h["status"] & 0x1000) != 0) > 0.99          # records written by the product-form walk itself
    Nf = np.arange(n, dtype=np.float64)
    sel = np.stack([Nf, Nf * Nf])
    idx = np.unique(np.concatenate([np.arange(0, S, 499), [S - 1]]))
    assert len(idx) >= 2000
    for k in idx:
        assert h["status"][k] & 0x1000
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu[k], smooth, sel=sel)
        _check_record(h, int(k), r, nsel=2)
    assert ref.available(), "compiled reference (oracle/_ref) missing: run oracle/build_ref.py where /root/reference exists"
    base = ref.make_histogram(lnpi, mom, 1.0, [0.0], smooth)
    worst = 0.0
    for k in idx[::9][:230]:
        g = copy.deepcopy(base)
        g.reweight(float(mu[k]))
        g.thermo()
        safe = g.is_safe()
        th = g.data["thermo"]
        P = len(th)
        assert h["code"][k] == 0 and h["nphase"][k] == P and bool(h["safe"][k]) == bool(safe)
        assert h["max_idx"][k, :P].tolist() == [int(v) for v in g.data["ln(PI)_maxima_idx"]]
        assert h["min_idx"][k, :h["nmin"][k]].tolist() == [int(v) for v in g.data["ln(PI)_minima_idx"]]
        for p in range(P):
            assert tuple(h["bounds"][k, p]) == tuple(th[p]["bound_idx"])
            got = np.array([h["fe"][k, p], h["avg"][k, p, 0], h["avg"][k, p, 1]])
            want = np.array([th[p]["F.E./kT"], th[p]["mom"][0, 1, 0, 0, 0], th[p]["mom"][0, 2, 0, 0, 0]])
            assert np.allclose(got, want, rtol=RTOL, atol=0)
            worst = max(worst, float(np.max(np.abs(got - want) / np.abs(want))))
    assert worst < RTOL
