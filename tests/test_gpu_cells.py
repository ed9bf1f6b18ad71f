"""GPU: dense mu sweeps on tilt cells (k_sweep_cell, csrc/fhmc_cell.cu; fhmc_mu_cells_build) -- the per-phase sums of reweight() +
thermo() (gc_hist.pyx:71-78, 498-554) as degree-7 moment expansions about cell centres, integers from the interval records of the
table walk -- against the general evaluator, the table walk, the oracle and the compiled reference.  Integers bit-exact, fe /
averages to 1e-10 (the expansion itself is cut at 1e-15)."""
import numpy as np
import pytest

from test_gpu_tables import S_MIN, _both, _check

pytestmark = pytest.mark.gpu


def _cell_fraction(c):
    return float(c["path"].cpu().numpy().astype(np.int64).mean())


def test_cells_headline_config_against_general_kernel_table_walk_oracle_and_reference(oracle):
    """BASELINE config 2 (1001 bins, smooth 10, <N>, <N^2>): every record of a 2x10^5-point sweep against the general kernel and the
    table walk; a strided sample against the C oracle and the compiled reference (reweight -> thermo -> is_safe)."""
    from fhmcanalysis_b200 import synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    S = 200000
    mu = np.linspace(-0.03, 0.03, S)
    dh, c, g, kern = _both(lnpi, N, mu, 10, cells=True)
    assert kern == "k_sweep_cell<compact>" and dh.desc.mu_cells
    assert _check(c, g, 4) > 0.999
    assert _cell_fraction(c) > 0.995, _cell_fraction(c)
    # the table walk on the same state points: same integers, values to the joint rounding
    dh.use_mu_cells = False          # (the descriptor keeps the cells; the switch alone must take them out of the call)
    w = dh.sweep_compact(mu, pmax=4)
    from fhmcanalysis_b200 import _lib
    assert _lib.last_kernel() == "k_sweep_tab2<compact>" and float(w["path"].double().mean()) == 0.0
    assert np.array_equal(w["nphase"].cpu().numpy(), c["nphase"].cpu().numpy())
    assert np.array_equal(w["bounds"].cpu().numpy(), c["bounds"].cpu().numpy())
    assert np.array_equal(w["status"].cpu().numpy() & 0x1FF, c["status"].cpu().numpy() & 0x1FF)
    fw, fc = w["fe"].cpu().numpy(), c["fe"].cpu().numpy()
    live = ~np.isnan(fw)
    assert np.array_equal(live, ~np.isnan(fc))
    assert np.max(np.abs(fw[live] - fc[live]) / np.maximum(1.0, np.abs(fw[live]))) < 1e-12
    aw, ac = w["avg"].cpu().numpy(), c["avg"].cpu().numpy()
    assert np.max(np.abs(aw[live] - ac[live]) / np.maximum(1.0, np.abs(aw[live]))) < 1e-12
    fe, av, bd, P = (c[k].cpu().numpy() for k in ("fe", "avg", "bounds", "nphase"))
    sel = np.stack([N, N * N])
    safe = (c["status"].cpu().numpy().astype(np.int64) & 0x100) != 0
    for k in range(0, S, 997):
        r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu[k], 10, sel=sel)
        assert r["status"] == 0 and r["nphase"] == P[k] and bool(r["safe"]) == bool(safe[k])
        assert bd[k, :P[k]].tolist() == r["bounds"].tolist()
        assert np.allclose(fe[k, :P[k]], r["fe"], rtol=1e-10, atol=1e-13)
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


def test_cells_wide_sweeps_one_phase_monotone_steep_tilts_and_random_order():
    """The cases of the table-walk test (one / two / three phases, monotone tilts, tilts beyond the product form, underflowing
    phases, noisy ln(PI) with capacity errors) through the cells: every record equals the general kernel's; ranges too wide for
    the cell budget are covered in part and finished by the table walk."""
    from fhmcanalysis_b200 import synth
    n = 1001
    N = np.arange(n, dtype=np.float64)
    S = S_MIN + 3001
    rng = np.random.default_rng(11)
    i = N
    three = np.logaddexp(np.logaddexp(-(i - 120.0) ** 2 / 900.0, -(i - 480.0) ** 2 / 2500.0 - 0.7), -(i - 850.0) ** 2 / 1600.0 - 1.1)
    for lnpi, smooth, span, want in ((synth.two_peak_lnpi(n), 10, 0.5, 0.9), (three + 2e-3 * rng.standard_normal(n), 7, 0.2, 0.9),
                                     (synth.two_peak_lnpi(n, noise=0.05), 3, 0.1, 0.0), (synth.two_peak_lnpi(n, noise=4e-3), 4, 0.1, 0.3),
                                     (synth.two_peak_lnpi(n), 10, 8.0, 0.5)):
        mu = np.concatenate([np.linspace(-span, span, S - 1001), rng.uniform(-span, span, 1001)])
        rng.shuffle(mu)
        dh, c, g, kern = _both(lnpi, N, mu, smooth, cells=True)
        assert kern == "k_sweep_cell<compact>"
        fast = _check(c, g, 4)
        assert fast >= want, (smooth, span, fast)
        if want >= 0.9 and span <= 0.5:
            assert _cell_fraction(c) > 0.8, (smooth, span, _cell_fraction(c))


def test_cells_ties_spacing_offsets_and_ranges():
    """Integer-valued ln(PI) (exact ties at tilt 0: no table record, hence no cell, may be used there); half-integer N spacing with
    N offset from zero, beta != 1 and a reference mu; a second sweep outside the first one's range rebuilds the cells for the
    union; a sweep with all state points equal; one averaged quantity that is not N; no averaged quantity."""
    N = np.arange(0, 401, dtype=np.float64)
    tri = np.abs((N % 100) - 50.0)
    S = S_MIN + 17
    mu = np.concatenate([np.zeros(S // 2), np.linspace(-1.5, 1.5, S - S // 2)])
    dh, c, g, kern = _both(tri, N, mu, 2, cells=True)
    assert kern == "k_sweep_cell<compact>"
    _check(c, g, 4)
    assert not np.any(c["path"].cpu().numpy()[:S // 2])
    from fhmcanalysis_b200 import _lib, engine, synth
    n = 601
    N2 = 10.0 + 0.5 * np.arange(n)
    lnpi = synth.two_peak_lnpi(n, scale=0.6)
    mu2 = np.linspace(-0.2, 0.2, S)
    dh, c, g, kern = _both(lnpi, N2, mu2, 5, sel=[N2, N2 * N2], beta=0.8, mu_ref=-0.3, cells=True)
    assert kern == "k_sweep_cell<compact>"
    assert _check(c, g, 4) > 0.99 and _cell_fraction(c) > 0.95
    # a second range, partly outside the first: the cells follow the sweep
    mu3 = np.linspace(0.1, 0.45, S)
    c3 = dh.sweep_compact(mu3, pmax=4)
    g3 = dh.sweep(mu3, pmax=4, lanes=-1).host()
    assert _check(c3, g3, 4) > 0.99 and _cell_fraction(c3) > 0.95
    # every state point the same
    mu4 = np.full(S, 0.0123)
    c4 = dh.sweep_compact(mu4, pmax=4)
    g4 = dh.sweep(mu4, pmax=4, lanes=-1).host()
    _check(c4, g4, 4)
    assert _cell_fraction(c4) > 0.99
    # device tensor input: cached by (storage, version); an in-place change of the tensor is seen; cells built for another sweep are
    # still correct for this one (state points outside them take the table walk)
    import torch
    mu_d = torch.from_numpy(mu2).cuda()
    c5 = dh.sweep_compact(mu_d, pmax=4)
    key = dh._cells_key
    c5 = dh.sweep_compact(mu_d, pmax=4)
    assert dh._cells_key == key and _cell_fraction(c5) > 0.95
    mu_d.mul_(3.0)
    c6 = dh.sweep_compact(mu_d, pmax=4)
    g6 = dh.sweep(mu_d, pmax=4, lanes=-1).host()
    assert dh._cells_key != key
    assert _check(c6, g6, 4) > 0.99 and _cell_fraction(c6) > 0.95
    stale = dh._cells_key
    mu_e = torch.from_numpy(np.linspace(-0.9, 0.9, S)).cuda()
    dh._cells_key = (mu_e.data_ptr(), mu_e.numel(), mu_e._version)     # pretend the cells (range [-0.6, 0.6]) belong to this sweep
    c7 = dh.sweep_compact(mu_e, pmax=4)
    g7 = dh.sweep(mu_e, pmax=4, lanes=-1).host()
    assert _check(c7, g7, 4) > 0.99 and 0.5 < _cell_fraction(c7) < 0.75
    # one quantity that is not N; no quantity
    for sel in ([N2 * N2], []):
        dhq = engine.DeviceHistogram(lnpi, N2, 0.8, -0.3, smooth=5, sel=sel)
        dhq.CELLS_MIN_STATES = 1
        cq = dhq.sweep_compact(mu2, pmax=4)
        assert _lib.last_kernel() == "k_sweep_cell<compact>"
        gq = dhq.sweep(mu2, pmax=4, lanes=-1).host()
        st = cq["status"].cpu().numpy().astype(np.int64)
        assert np.array_equal(st & 0xFF, gq["code"]) and np.array_equal(cq["nphase"].cpu().numpy(), gq["nphase"])
        P = gq["nphase"]
        for p in range(4):
            live = (gq["code"] == 0) & (P > p)
            assert np.array_equal(cq["bounds"].cpu().numpy()[live, p], gq["bounds"][live, p])
            if live.any():
                assert np.max(np.abs(cq["fe"].cpu().numpy()[live, p] - gq["fe"][live, p]) / np.maximum(1.0, np.abs(gq["fe"][live, p]))) < 1e-10
                if sel:
                    a, b = cq["avg"].cpu().numpy()[live, p, 0], gq["avg"][live, p, 0]
                    assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) < 1e-10
        assert _cell_fraction(cq) > 0.95


def test_cells_capacity_and_small_pmax():
    """pmax below the phase count: the capacity code of the general evaluator, never a cell record; pmax = 8."""
    from fhmcanalysis_b200 import engine, synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    mu = np.linspace(-0.03, 0.03, S_MIN + 5)
    for pmax in (1, 8):
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
        dh.CELLS_MIN_STATES = 1
        c = dh.sweep_compact(mu, pmax=pmax)
        g = dh.sweep(mu, pmax=pmax, lanes=-1).host()
        _check(c, g, pmax)
        if pmax == 1:
            assert np.any(g["code"] == 8)
            assert not np.any(c["path"].cpu().numpy()[g["code"] == 8])


def test_compact_sweep_below_the_fused_threshold_sizes_its_workspace():
    """Sweeps of 8192 < S <= 2 x 256 state points per SM take the general path (a scratch record per state point): the workspace
    query must say so (it used to answer with the fused path's size and the call failed with 'workspace too small')."""
    from fhmcanalysis_b200 import engine, synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    for S in (9000, 30000, 70000):
        mu = np.linspace(-0.03, 0.03, S)
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
        c = dh.sweep_compact(mu, pmax=4)
        g = dh.sweep(mu, pmax=4, lanes=-1).host()
        _check(c, g, 4)


def test_cells_in_the_host_pipeline():
    """fhmc_sweep_host_compact16 with cells built for the range of the host array (once per buffer): every chunk runs k_sweep_cell;
    a second call with other content in the same buffer is still correct (what the cells miss takes the table walk)."""
    import torch
    from fhmcanalysis_b200 import _lib, engine, synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    S = 1 << 18
    mu_h = torch.from_numpy(np.linspace(-0.03, 0.03, S)).pin_memory()
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    out = None
    for rnd in range(2):
        out = dh.sweep_host_compact(mu_h, pmax=4, out=out)
        assert _lib.last_kernel() == "k_sweep_cell<compact>" and dh.desc.mu_cells
        g = dh.sweep(mu_h.numpy(), pmax=4, lanes=-1).host()
        assert np.array_equal(out["status"].numpy().astype(np.int64) & 0xFF, g["code"])
        assert np.array_equal(out["nphase"].numpy(), g["nphase"])
        for p in range(3):
            live = (g["code"] == 0) & (g["nphase"] > p)
            assert np.array_equal(out["bounds"].numpy()[live, p], g["bounds"][live, p])
            if live.any():
                assert np.max(np.abs(out["fe"].numpy()[live, p] - g["fe"][live, p]) / np.maximum(1.0, np.abs(g["fe"][live, p]))) < 1e-10
                assert np.max(np.abs(out["avg"].numpy()[live, p] - g["avg"][live, p]) / np.maximum(1.0, np.abs(g["avg"][live, p]))) < 1e-10
        mu_h.mul_(2.5)     # same buffer, twice the range: the cells still cover the middle of it


def test_cells_two_destinations_take_the_transposing_kernel():
    """Two destination buffers (what a sweep fused with its gather hands the kernel: this GPU's buffer and its NVLink peers') run
    k_sweep_cell_t -- warp-uniform phase loop, fields transposed through shared memory into 16-byte stores.  Both buffers must hold
    the records of the general kernel; a first-record offset and an odd record count exercise the unaligned per-lane path, a shuffled
    sweep the warps whose lanes disagree about the phase count."""
    import torch
    from fhmcanalysis_b200 import _lib, engine, synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    rng = np.random.default_rng(5)
    L = _lib.load()
    for S, first, shuffle, sel in ((S_MIN + 64, 0, False, ["N", N * N]), (S_MIN + 33, 7, True, ["N", N * N]), (S_MIN + 1, 3, True, [N * N]),
                                   (S_MIN + 2, 0, False, [])):
        mu = np.linspace(-0.2, 0.2, S)
        if shuffle:
            rng.shuffle(mu)
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=sel)
        dh.CELLS_MIN_STATES = 1
        n_total = S + first + 5
        nbytes = int(L.fhmc_pack_soa16_bytes(n_total, 4, len(sel)))
        bufs = [torch.full((nbytes,), 0xFF, dtype=torch.uint8, device="cuda") for _ in range(2)]
        dh.sweep_compact(mu, pmax=4, dst=[b.data_ptr() for b in bufs], n_total=n_total, first=first, fill_dead=False)
        assert _lib.last_kernel() == "k_sweep_cell<compact>"
        g = dh.sweep(mu, pmax=4, lanes=-1).host()
        for b in bufs:
            v = engine.soa16_views(b, n_total, 4, len(sel))
            st = v["status"].cpu().numpy().astype(np.int64)[first:first + S]
            assert np.array_equal(st & 0xFF, g["code"]) and np.array_equal(v["nphase"].cpu().numpy()[first:first + S], g["nphase"])
            assert float(v["path"][first:first + S].double().mean()) > 0.95
            P = g["nphase"]
            for p in range(4):
                live = (g["code"] == 0) & (P > p)
                fe = v["fe"].cpu().numpy()[first:first + S, p]
                assert np.array_equal(v["bounds"].cpu().numpy()[first:first + S, p][live], g["bounds"][live, p])
                assert np.all(np.isnan(fe[~live]))          # untouched 0xFF bytes
                if live.any():
                    assert np.max(np.abs(fe[live] - g["fe"][live, p]) / np.maximum(1.0, np.abs(g["fe"][live, p]))) < 1e-10
                    if sel:
                        a = v["avg"].cpu().numpy()[first:first + S, p][live]
                        assert np.max(np.abs(a - g["avg"][live, p]) / np.maximum(1.0, np.abs(g["avg"][live, p]))) < 1e-10
            # records outside [first, first + S) stay untouched
            assert np.all(v["status"].cpu().numpy()[:first] == -1) and np.all(v["status"].cpu().numpy()[first + S:] == -1)
