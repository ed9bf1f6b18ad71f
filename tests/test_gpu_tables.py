"""GPU: the table-driven mu sweep (k_sweep_tab2, csrc/fhmc_tab.cuh; fhmc_mu_tables_build) -- relextrema() + the phase bounds of
thermo() (gc_hist.pyx:317-415, 498-520) looked up per elementary tilt interval instead of being re-derived per state point --
against the general evaluator, the oracle and the compiled reference.  Integers bit-exact, fe / averages to 1e-10."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
S_MIN = 80000   # the fused compact kernels take sweeps of more than 2 x 256 state points per SM (148 SMs)


def _both(lnpi, N, mu, smooth, sel=None, beta=1.0, mu_ref=0.0, pmax=4, cells=False):
    """compact records through the table-driven kernel + plain records of the general one-lane kernel"""
    from fhmcanalysis_b200 import _lib, engine
    sel = ["N", N * N] if sel is None else sel
    dh = engine.DeviceHistogram(lnpi, N, beta, mu_ref, smooth=smooth, sel=sel)
    dh.use_mu_cells = cells     # False: the table walk itself (tests/test_gpu_cells.py runs the same cases on the tilt cells)
    dh.CELLS_MIN_STATES = 1     # (the engine builds cells from 2^17 state points on; the test sweeps are shorter)
    c = dh.sweep_compact(mu, pmax=pmax)
    kern = _lib.last_kernel()
    g = dh.sweep(mu, pmax=pmax, lanes=-1).host()
    return dh, c, g, kern


def _check(c, g, pmax, tol=1e-10):
    st = c["status"].cpu().numpy().astype(np.int64)
    code = st & 0xFF
    bad = np.nonzero(code != g["code"])[0]
    assert len(bad) == 0, (len(bad), bad[:5], code[bad[:5]], g["code"][bad[:5]], g["nphase"][bad[:5]])
    safe_c, safe_g = (st & 0x100) != 0, (g["status"].astype(np.int64) & 0x100) != 0
    ok = code == 0
    assert np.array_equal(safe_c[ok], safe_g[ok])
    P = g["nphase"]
    assert np.array_equal(c["nphase"].cpu().numpy()[ok], P[ok])
    fe, av, bd = (c[k].cpu().numpy() for k in ("fe", "avg", "bounds"))
    worst = 0.0
    for p in range(pmax):
        live = ok & (P > p)
        assert np.array_equal(bd[live, p], g["bounds"][live, p])
        if live.any():
            worst = max(worst, float(np.max(np.abs(fe[live, p] - g["fe"][live, p]) / np.maximum(1.0, np.abs(g["fe"][live, p])))))
            worst = max(worst, float(np.max(np.abs(av[live, p] - g["avg"][live, p]) / np.maximum(1.0, np.abs(g["avg"][live, p])))))
        assert np.all(np.isnan(fe[~live, p])) and np.all(bd[~live, p] == -1)
    assert worst <= tol, worst
    return float(np.mean((st & 0x1000) != 0))   # fraction written by the table walk itself (FHMC_ST_FAST)


def test_tables_headline_config_against_general_kernel_oracle_and_reference(oracle):
    """BASELINE config 2 (1001 bins, smooth 10, <N>, <N^2>): every record of a 2x10^5-point sweep against the general kernel;
    a strided sample against the C oracle and the compiled reference (reweight -> thermo -> is_safe, GH:268-289, 451-596)."""
    from fhmcanalysis_b200 import synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    S = 200000
    mu = np.linspace(-0.03, 0.03, S)
    dh, c, g, kern = _both(lnpi, N, mu, 10)
    assert kern == "k_sweep_tab2<compact>"
    fast = _check(c, g, 4)
    assert fast > 0.999, fast
    fe, av, bd, P = (c[k].cpu().numpy() for k in ("fe", "avg", "bounds", "nphase"))
    sel = np.stack([N, N * N])
    safe = (c["status"].cpu().numpy().astype(np.int64) & 0x100) != 0
    for k in range(0, S, 997):
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu[k], 10, sel=sel)
        assert r["status"] == 0 and r["nphase"] == P[k] and bool(r["safe"]) == bool(safe[k])
        assert bd[k, :P[k]].tolist() == r["bounds"].tolist()
        assert np.allclose(fe[k, :P[k]], r["fe"], rtol=1e-10, atol=0)
        assert np.allclose(av[k, :P[k]], r["avg"][:, :2], rtol=1e-10, atol=1e-300)
    from oracle import ref
    if ref.load() is not None:
        import copy
        base = ref.make_histogram(lnpi, synth.one_comp_moments(n), 1.0, [0.0], 10)
        for k in range(0, S, 19997):
            h = copy.deepcopy(base)
            h.reweight(float(mu[k]))
            h.thermo()
            th = h.data["thermo"]
            assert len(th) == P[k]
            for p in range(P[k]):
                assert tuple(th[p]["bound_idx"]) == tuple(int(x) for x in bd[k, p])
                assert abs(th[p]["F.E./kT"] - fe[k, p]) <= 1e-10 * max(1.0, abs(fe[k, p]))
                assert abs(th[p]["n1"] - av[k, p, 0]) <= 1e-10 * max(1.0, abs(av[k, p, 0]))
            assert bool(h.is_safe()) == bool(int(c["status"][k].item()) & 0x100)


def test_tables_wide_sweeps_one_phase_monotone_and_steep_tilts():
    """One, two and three phases, monotone tilts (no windowed extremum: GH:382-386), tilts beyond the product form's range and
    tilts where a phase underflows next to the global maximum (rescue) -- every record equals the general kernel's."""
    from fhmcanalysis_b200 import synth
    n = 1001
    N = np.arange(n, dtype=np.float64)
    S = S_MIN + 3001
    rng = np.random.default_rng(11)
    i = N
    three = np.logaddexp(np.logaddexp(-(i - 120.0) ** 2 / 900.0, -(i - 480.0) ** 2 / 2500.0 - 0.7), -(i - 850.0) ** 2 / 1600.0 - 1.1)
    # (the 0.05-noise histogram shows more extrema than pmax at smooth 3: every state point is a capacity error of the general evaluator)
    for lnpi, smooth, span, want in ((synth.two_peak_lnpi(n), 10, 0.5, 0.9), (three + 2e-3 * rng.standard_normal(n), 7, 0.2, 0.9),
                                     (synth.two_peak_lnpi(n, noise=0.05), 3, 0.1, 0.0), (synth.two_peak_lnpi(n, noise=4e-3), 4, 0.1, 0.3),
                                     (synth.two_peak_lnpi(n), 10, 8.0, 0.5)):
        mu = np.concatenate([np.linspace(-span, span, S - 1001), rng.uniform(-span, span, 1001)])
        dh, c, g, kern = _both(lnpi, N, mu, smooth)
        assert kern == "k_sweep_tab2<compact>"
        fast = _check(c, g, 4)
        assert fast >= want, (smooth, span, fast)


def test_tables_ties_and_plateaus_fall_back_to_the_general_evaluator():
    """Integer-valued ln(PI) (the reference's own unit-test style, T1:149-198): at tilt 0 every chord slope is hit exactly, so
    no state point may take a table record there; half-integer N spacing; N offset from zero."""
    N = np.arange(0, 401, dtype=np.float64)
    tri = np.abs((N % 100) - 50.0)          # piecewise linear: long runs of equal chord slopes
    S = S_MIN + 17
    mu = np.concatenate([np.zeros(S // 2), np.linspace(-1.5, 1.5, S - S // 2)])
    dh, c, g, kern = _both(tri, N, mu, 2)
    assert kern == "k_sweep_tab2<compact>"
    _check(c, g, 4)
    st = c["status"].cpu().numpy().astype(np.int64)
    assert not np.any(st[:S // 2] & 0x1000)                 # exact ties: all through the general evaluator
    # half-integer spacing, N offset
    from fhmcanalysis_b200 import synth
    n = 601
    N2 = 10.0 + 0.5 * np.arange(n)
    lnpi = synth.two_peak_lnpi(n, scale=0.6)
    mu2 = np.linspace(-0.2, 0.2, S)
    dh, c, g, kern = _both(lnpi, N2, mu2, 5, sel=[N2, N2 * N2], beta=0.8, mu_ref=-0.3)
    assert kern == "k_sweep_tab2<compact>"
    assert _check(c, g, 4) > 0.99


def test_tables_host_pipeline_and_capacity():
    """The host-buffer pipeline (fhmc_sweep_host_compact16) picks the tables up through the descriptor; a sweep whose phase
    count exceeds pmax reports FHMC_E_CAPACITY exactly like the general kernel."""
    from fhmcanalysis_b200 import _lib, engine, synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    S = 1 << 18
    mu = np.linspace(-0.03, 0.03, S)
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    dh.use_mu_cells = False
    h = dh.sweep_host_compact(mu, pmax=4)
    assert _lib.last_kernel() == "k_sweep_tab2<compact>" and dh.desc.mu_tables
    g = dh.sweep(mu, pmax=4, lanes=-1).host()
    assert np.array_equal(h["status"].numpy().astype(np.int64) & 0xFF, g["code"])
    assert np.array_equal(h["nphase"].numpy(), g["nphase"])
    for p in range(2):
        live = g["nphase"] > p
        assert np.array_equal(h["bounds"].numpy()[live, p], g["bounds"][live, p])
        assert np.allclose(h["fe"].numpy()[live, p], g["fe"][live, p], rtol=1e-10, atol=1e-10)
        assert np.allclose(h["avg"].numpy()[live, p], g["avg"][live, p], rtol=1e-10, atol=1e-10)
    # pmax = 1 on a two-phase range: capacity code from the general evaluator, never a table record
    dh1 = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    c1 = dh1.sweep_compact(mu[:S_MIN + 5], pmax=1)
    g1 = dh1.sweep(mu[:S_MIN + 5], pmax=1, lanes=-1).host()
    assert np.array_equal(c1["status"].cpu().numpy().astype(np.int64) & 0xFF, g1["code"])
    assert np.any(g1["code"] == 8)
