"""GPU: isopleth grid, joint (2-D) histogram reweighting, sharded sweep helper, compare fast vs generic kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_isopleth_make_grid_multi_matches_reference(golden, golden_meta):
    """gc_binary.isopleth.make_grid_multi (GB:173-290; untested upstream) against the compiled reference."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_binary as gcB
    meta = golden_meta["iso"]
    assert meta["status"] == "ok"
    hists = []
    for k, d2 in enumerate(meta["dmu2"]):
        hists.append(oneDH.histogram.from_arrays(golden["iso/lnpi"][k], golden["c3/mom"], meta["beta_ref"],
                                                 [meta["mu1_ref"], meta["mu1_ref"] + d2], meta["smooth"], volume=meta["volume"]))
    iso = gcB.isopleth(hists, meta["beta_ref"], meta["order"])
    Z, (X, Y) = iso.make_grid_multi(meta["mu1_bounds"], meta["dmu2_bounds"], meta["delta"], meta["m"])
    assert np.allclose(X, golden["iso/X"]) and np.allclose(Y, golden["iso/Y"])
    assert np.array_equal(Z == 0, golden["iso/x1"] == 0)   # same cells skipped
    assert np.allclose(Z, golden["iso/x1"], rtol=1e-9, atol=1e-12)
    assert np.allclose(iso.data["density"], golden["iso/density"], rtol=1e-9, atol=1e-12)
    assert np.allclose(iso.data["F.E./kT"], golden["iso/fe"], rtol=1e-9, atol=1e-12)


def test_combine_isopleth_grids():
    """unittests/moments_histogram_one_dim_gc_ntot_isopleth.py:27-91 (pure host bookkeeping)."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_binary as gcB
    dmu2 = np.linspace(-5, -3, 5)
    x1, y1 = np.meshgrid(np.linspace(-15, -10, 10), dmu2)
    x2, y2 = np.meshgrid(np.linspace(-10, -5, 10), dmu2)
    z1, z2 = x1 ** 2 + y1 ** 2, x2 ** 2 + y2 ** 2
    x3, y3 = np.meshgrid(np.concatenate((np.linspace(-15, -10, 10), np.linspace(-10, -5, 10)[1:])), dmu2)
    Z, (X, Y) = gcB.combine_isopleth_grids([x2, x1], [y2, y1], [z2, z1])
    assert np.all(np.abs(X - x3) < 1e-9) and np.all(np.abs(Y - y3) < 1e-9) and np.all(np.abs(Z - (x3 ** 2 + y3 ** 2)) < 1e-9)
    xb, yb = np.meshgrid(np.linspace(-10, -5, 10), np.linspace(-5, -4, 5))
    with pytest.raises(Exception):
        gcB.combine_isopleth_grids([xb, x1], [yb, y1], [xb ** 2, z1])


def test_joint_hist_container_and_reweight(golden, golden_meta, oracle):
    """unittests/moments_histogram_two_dim_joint.py:238-244 layout (-inf padding, inclusive bounds) + K5 vs the oracle."""
    import FHMCAnalysis.moments.histogram.two_dim.joint_hist as jh
    J = jh.joint_hist()
    for op1, dct in golden_meta["joint"]["entries"]:
        J.enter(op1, np.array(dct["lnPI"]), np.array(dct["op2"]), {k: np.array(v) for k, v in dct["props"].items()})
    J.make()
    assert np.array_equal(J.data["ln(PI)"], golden["joint/lnpi"])
    assert J.data["bounds_idx"].tolist() == golden["joint/bounds"].tolist() == [[1, 3], [0, 4]]
    assert np.array_equal(J.data["op_1"], golden["joint/op1"]) and np.array_equal(J.data["op_2"], golden["joint/op2"])
    assert np.array_equal(J.data["props"]["e"], golden["joint/prop_e"])
    a1, a2 = np.array([0.0, 0.3, -1.0]), np.array([0.0, -0.2, 0.5])
    out = J.reweight_batch(a1, a2, props=("e",))
    b = J.data["bounds_idx"].astype(np.int32).copy()
    b[:, 1] += 1
    for s in range(3):
        ref = oracle.reweight_2d(J.data["ln(PI)"], b, J.data["op_1"], J.data["op_2"], a1[s], a2[s], J.data["props"]["e"][None])
        assert np.allclose(out[s], ref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("n1,n2,nprop", [(64, 48, 0), (129, 77, 1), (129, 77, 2), (512, 512, 2)])   # (129, 77, 2): odd rows * n2 with two staged property chunks
def test_reweight_2d_vs_oracle(oracle, n1, n2, nprop):
    """BASELINE config 5 generator (two anisotropic Gaussians, triangular -inf region) at three sizes."""
    from fhmcanalysis_b200 import engine, synth
    lnpi, bounds = synth.joint_2d(n1, n2, cut=int(0.625 * (n1 + n2)))
    op1, op2 = np.arange(n1, dtype=float), np.arange(n2, dtype=float)
    rng = np.random.default_rng(5)
    props = rng.random((nprop, n1, n2)) if nprop else None
    S = 300
    a1, a2 = rng.uniform(-0.05, 0.05, S), rng.uniform(-0.05, 0.05, S)
    for product in (False, True):   # exp-per-bin kernel and product-form kernel
        out = engine.reweight_2d(lnpi, bounds, op1, op2, a1, a2, props, product=product)
        for s in range(0, S, 37):
            ref = oracle.reweight_2d(lnpi, bounds, op1, op2, a1[s], a2[s], props)
            assert np.allclose(out[s], ref, rtol=1e-10, atol=0), (product, s)
    # an odd number of state points, stronger tilts, op2 starting away from zero with spacing 0.5, -inf holes inside the support
    lnpi2 = lnpi.copy()
    lnpi2[n1 // 3, bounds[n1 // 3, 0] + 2] = -np.inf
    op2b = 7.0 + 0.5 * np.arange(n2)
    a1b, a2b = rng.uniform(-0.3, 0.3, 77), rng.uniform(-0.5, 0.5, 77)
    outp = engine.reweight_2d(lnpi2, bounds, op1, op2b, a1b, a2b, props, product=True)
    oute = engine.reweight_2d(lnpi2, bounds, op1, op2b, a1b, a2b, props, product=False)
    assert np.allclose(outp, oute, rtol=1e-10, atol=0)
    ref = oracle.reweight_2d(lnpi2, bounds, op1, op2b, a1b[5], a2b[5], props)
    assert np.allclose(outp[5], ref, rtol=1e-10, atol=0)
    # linearity property at full batch: <op1> is non-decreasing in a1 at fixed a2
    a1s = np.linspace(-0.05, 0.05, 257)
    o = engine.reweight_2d(lnpi, bounds, op1, op2, a1s, np.zeros_like(a1s))
    assert np.all(np.diff(o[:, 1]) > -1e-9)


def test_fast_and_generic_kernels_agree():
    """The one-pass mu-sweep kernel (product form, chain recurrence, true exps) and the generic two-pass kernel must
    give identical integers and fp64 to 1e-10."""
    from fhmcanalysis_b200 import engine, synth
    for n, smooth, noise in ((1001, 10, 1e-3), (573, 3, 5e-2), (2001, 30, 0.0)):
        lnpi = synth.two_peak_lnpi(n, noise=noise, scale=n / 1001.0)
        N = np.arange(n, dtype=float)
        for rec in (3, 2, 1, 0):   # product form (two points / one point per thread), multiplicative chains, true exps
            dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=smooth, sel=["N", N * N])
            dh.use_recurrence = rec
            dh.ensure_hull()
            assert dh.desc.mu_recurrence == rec
            # mostly the coexistence region, plus strong tilts (product form: anchors underflow / chain switched off)
            mus = np.concatenate([np.linspace(-0.05, 0.05, 4600), np.linspace(-6.0, 6.0, 400)])
            if rec == 3:   # the two-point kernel only runs from 4 * 256 state points per SM on
                mus = np.concatenate([np.linspace(-0.05, 0.05, 158000), np.linspace(-6.0, 6.0, 4001)])
            a = dh.sweep_auto(mus, pmax=4, lanes=1).host()      # one-pass fast kernel
            if n == 1001:
                assert np.mean((a["status"] & 0x1000) != 0) > 0.99   # really produced by the fast kernel
            b = dh.sweep_auto(mus, pmax=a["fe"].shape[1], lanes=-1).host()     # generic one-lane kernel
            for k in ("code", "nphase", "nmin", "safe"):
                assert np.array_equal(a[k], b[k]), (n, k)
            ok = a["code"] == 0                                   # records of raised state points carry no thermo
            for k in ("bounds", "max_idx", "min_idx"):
                # entries beyond the list lengths (nphase maxima / bounds, nmin minima) are never written
                pm = np.arange(a[k].shape[1])[None, :] < (a["nmin"] if k == "min_idx" else a["nphase"])[:, None]
                m2 = (pm & ok[:, None]) if a[k].ndim == 2 else (pm & ok[:, None])[:, :, None].repeat(2, 2)
                assert np.array_equal(a[k][m2], b[k][m2]), (n, k)
            mask = (np.arange(a["fe"].shape[1])[None, :] < a["nphase"][:, None]) & ok[:, None]
            assert np.allclose(a["fe"][mask], b["fe"][mask], rtol=1e-10, atol=1e-11)
            assert np.allclose(a["avg"][mask], b["avg"][mask], rtol=1e-10, atol=0)
            assert np.allclose(a["lnnorm"], b["lnnorm"], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("nsel,sel0n", [(0, False), (1, True), (1, False), (2, False), (3, True), (4, False)])
def test_product_form_for_every_selection_shape(nsel, sel0n):
    """Every instantiation of the product-form kernels (0..4 summed quantities, first one N or not; one and two state points
    per thread) against the generic kernel: the tabulated P*X rows of quantities other than N, N-spacing other than 1 and a
    non-zero first bin are all exercised."""
    from fhmcanalysis_b200 import engine, synth
    n = 801
    rng = np.random.default_rng(100 * nsel + sel0n)
    lnpi = synth.two_peak_lnpi(n, noise=2e-3, scale=n / 1001.0)
    N = 3.0 + 0.5 * np.arange(n)                              # uniform spacing 0.5, first bin at 3
    extra = [N * N, -2.0 * N - 0.002 * N * N + rng.normal(size=n), np.sin(N / 40.0), 1.0 / (1.0 + N)]
    sel = (["N"] if sel0n else []) + extra[:nsel - (1 if sel0n else 0)]
    for rec, S in ((3, 170000), (2, 6000)):
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=7, sel=sel)
        dh.use_recurrence = rec
        dh.ensure_hull()
        assert dh.desc.mu_recurrence == rec
        mus = np.concatenate([np.linspace(-0.12, 0.12, S - 500), np.linspace(-9.0, 9.0, 500)])
        a = dh.sweep_auto(mus, pmax=4, lanes=1).host()
        b = dh.sweep_auto(mus, pmax=a["fe"].shape[1], lanes=-1).host()
        assert np.mean((a["status"] & 0x1000) != 0) > 0.9
        for k in ("code", "nphase", "nmin", "safe"):
            assert np.array_equal(a[k], b[k]), (rec, k)
        ok = a["code"] == 0
        mask = (np.arange(a["fe"].shape[1])[None, :] < a["nphase"][:, None]) & ok[:, None]
        assert np.array_equal(a["bounds"][mask], b["bounds"][mask])
        assert np.allclose(a["fe"][mask], b["fe"][mask], rtol=1e-10, atol=1e-10)
        if nsel:
            assert np.allclose(a["avg"][mask], b["avg"][mask], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("n", [9, 10, 11, 12, 130, 131, 257])
def test_product_form_small_and_odd_histograms(n):
    """Few bins (a single partial segment, every tail length n mod 4) and segment-boundary sizes through both product-form
    kernels, against the generic kernel; smooth = 1 makes nearly every block hold an extremum."""
    from fhmcanalysis_b200 import engine
    rng = np.random.default_rng(n)
    lnpi = np.cumsum(rng.normal(0.0, 0.6, size=n))
    N = np.arange(n, dtype=float)
    for rec, S in ((3, 152001), (2, 5003)):
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=1, sel=["N", N * N])
        dh.use_recurrence = rec
        dh.ensure_hull()
        assert dh.desc.mu_recurrence == rec
        mus = np.linspace(-1.5, 1.5, S)
        a = dh.sweep_auto(mus, pmax=8, lanes=1).host()
        b = dh.sweep_auto(mus, pmax=a["fe"].shape[1], lanes=-1).host()
        _agree(a, b, (n, rec))


@pytest.mark.parametrize("case", ["steep_large_tilt", "many_negligible_phases"])
def test_product_form_exponent_range_cases(case):
    """Cases scripts/soak_prod_parity.py found: (i) a steep ln(PI) (hundreds of units per 128-bin segment) under tilts of
    several units per bin -- exp(lnPI_i - A_seg) or the running factor leave the fp64 exponent range, and an anchor that
    underflowed comes back clamped, not zero; (ii) smooth = 1 on a random walk -- more than 32 phases, some of them
    negligible (rescued about their own maximum).  Both product-form kernels against the generic evaluator."""
    from fhmcanalysis_b200 import engine
    rng = np.random.default_rng(7)
    if case == "steep_large_tilt":
        n, smooth, span = 1578, 9, 8.0
        x = np.arange(n, dtype=float)
        lnpi = -1.45 * x + 40.0 * np.sin(x / 57.0) + 1e-3 * rng.normal(size=n)
        N = 0.5 * x
    else:
        n, smooth, span = 700, 1, 2.0
        lnpi = np.cumsum(rng.normal(0.0, 0.9, size=n))
        N = 2.0 + np.arange(n, dtype=float)
    for rec, S in ((3, 155000), (2, 6000)):
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=smooth, sel=["N", N * N])
        dh.use_recurrence = rec
        dh.ensure_hull()
        assert dh.desc.mu_recurrence == rec
        mus = rng.uniform(-span, span, size=S)
        a = dh.sweep_auto(mus, pmax=8, lanes=1).host()
        b = dh.sweep_auto(mus, pmax=a["fe"].shape[1], lanes=-1).host()
        if case == "many_negligible_phases":
            assert a["nphase"].max() > 32 and np.any(b["status"] & 0x800)     # phases beyond the rescue mask, rescued ones
        _agree(a, b, (case, rec))


def test_sharded_sweep_single_process(golden, golden_meta):
    """parallel.sweep_sharded without an initialised process group == plain sweep."""
    from fhmcanalysis_b200 import engine, parallel
    lnpi, mus = golden["c2/lnpi"], golden["c2/mu"]
    n = len(lnpi)
    N = np.arange(n, dtype=float)

    def mk():
        return engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=golden_meta["c2"]["smooth"], sel=["N", N * N])
    out = parallel.sweep_sharded(mk, mus, pmax=4, gather=True)
    ref = mk().sweep(mus, pmax=4).host()
    for k in ("nphase", "max_idx", "min_idx", "bounds"):
        assert np.array_equal(out[k], ref[k])
    assert np.allclose(out["fe"][:, :2], ref["fe"][:, :2], rtol=0, atol=0)


def test_histogram_larger_than_shared_memory(oracle):
    """N_max = 40000: the blob (3 rows x 320 KB) cannot be staged in 227 KB of shared memory; the generic kernels read the
    rows through L1/L2 instead.  Same outputs as the oracle."""
    from fhmcanalysis_b200 import engine, synth
    n = 40001
    lnpi = synth.two_peak_lnpi(n, noise=0.0, scale=40.0)
    N = np.arange(n, dtype=float)
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=50, sel=["N"])
    mus = np.array([-1e-3, 0.0, 2e-4])
    for lanes in (32, 4, -1):
        h = dh.sweep(mus, pmax=4, lanes=lanes).host()
        for k, mu in enumerate(mus):
            r = oracle.state_point(lnpi, np.arange(n), 1.0, 0.0, mu, 50, sel=N[None])
            P = r["nphase"]
            assert h["code"][k] == r["status"] == 0 and h["nphase"][k] == P
            assert h["max_idx"][k, :P].tolist() == r["max_idx"].tolist()
            assert h["min_idx"][k, :h["nmin"][k]].tolist() == r["min_idx"].tolist()
            assert np.allclose(h["fe"][k, :P], r["fe"], rtol=1e-10, atol=0)
            assert np.allclose(h["avg"][k, :P, 0], r["avg"][:, 0], rtol=1e-10, atol=0)
    res = dh.find_phase_eq(np.array([0.0]), lnz_tol=1e-10).host()
    assert res["code"][0] == 0 and abs(res["dfe"][0]) < 1e-9


def _agree(a, b, tag):
    for k in ("code", "nphase", "nmin", "safe"):
        assert np.array_equal(a[k], b[k]), (tag, k)
    ok = a["code"] == 0
    P = a["nphase"]
    for k in ("max_idx", "min_idx"):
        pm = np.arange(a[k].shape[1])[None, :] < (a["nmin"] if k == "min_idx" else P)[:, None]   # written entries only
        assert np.array_equal(a[k][pm & ok[:, None]], b[k][pm & ok[:, None]]), (tag, k)
    mask = (np.arange(a["fe"].shape[1])[None, :] < P[:, None]) & ok[:, None]
    assert np.allclose(a["fe"][mask], b["fe"][mask], rtol=1e-10, atol=1e-11), tag
    if a["avg"] is not None:
        assert np.allclose(a["avg"][mask], b["avg"][mask], rtol=1e-10, atol=1e-12), tag
    assert np.allclose(a["lnnorm"], b["lnnorm"], rtol=1e-13, atol=1e-12), tag


def test_fast_taylor_kernel_agrees_with_generic(golden, golden_meta):
    """One-thread-per-point Taylor kernel (NC > 0 instantiations) vs the generic evaluator on (beta x dmu) grids and on a
    1-species beta sweep of the real square-well data."""
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    meta = golden_meta["c3"]
    h = histogram.from_arrays(golden["c3/lnpi"], golden["c3/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"])
    h.reweight(meta["mu1"])
    betas, dmus = np.linspace(0.96, 1.04, 48), np.linspace(0.25, 0.75, 40)
    for order, moments in ((2, ()), (1, ("N1", "N2", "U")), (1, ())):
        dh = h.device_histogram(beta=betas, dmu=dmus, order=order, moments=moments)
        st = dh.make_states(np.array([meta["mu1"]]), betas, dmus, grid=True)
        a = dh.sweep(None, states=st, pmax=8, lanes=1).host()
        b = dh.sweep(None, states=st, pmax=8, lanes=-1).host()
        assert np.mean(a["code"] == 0) > 0.9
        _agree(a, b, ("c3", order, moments))
    sw = golden_meta["sw"]
    h1 = histogram.from_arrays(golden["sw/lnpi"], golden["sw/mom"], sw["beta_ref"], sw["mu_ref"], sw["smooth"])
    b1 = sw["beta_ref"] * np.linspace(0.99, 1.01, 300)
    for order, moments in ((1, ("N", "N2", "U")), (2, ())):
        dh = h1.device_histogram(beta=b1, order=order, moments=moments)
        mu = np.full(300, -4.02)
        a = dh.sweep(mu, b1, pmax=4, lanes=1).host()
        b = dh.sweep(mu, b1, pmax=4, lanes=-1).host()
        _agree(a, b, ("sw", order, moments))


@pytest.mark.gpu
def test_row_combined_taylor_grid_kernel_agrees_with_generic(golden, golden_meta):
    """k_sweep_rowc (coefficient rows combined once per (mu_1, beta) row of a beta x dmu_2 grid, two state points per thread,
    in-warp fallback on the combined rows) against the general evaluator on the flat rows: wide temperature range (monotone
    and one-phase cells), a run length that is not a multiple of the 64-point warp tile, orders 1 and 2, a small pmax."""
    from fhmcanalysis_b200 import _lib
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    meta = golden_meta["c3"]
    h = histogram.from_arrays(golden["c3/lnpi"], golden["c3/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"])
    h.reweight(meta["mu1"])
    for order, nb, nd, pmax, brange in ((2, 24, 600, 8, (0.90, 1.10)), (1, 16, 1111, 8, (0.97, 1.03)), (2, 8, 520, 2, (0.98, 1.02))):
        betas, dmus = np.linspace(brange[0], brange[1], nb), np.linspace(0.1, 0.9, nd)
        dh = h.device_histogram(beta=betas, dmu=dmus, order=order, moments=())
        st = dh.make_states(np.array([meta["mu1"]]), betas, dmus, grid=True)
        a = dh.sweep(None, states=st, pmax=pmax, lanes=1).host()
        assert _lib.last_kernel() == "k_sweep_rowc"
        b = dh.sweep(None, states=st, pmax=pmax, lanes=-1).host()
        assert _lib.last_kernel() == "k_sweep_1d<1>"
        if pmax == 8:
            assert np.mean(a["code"] == 0) > 0.9 and np.mean((a["status"] & 0x1000) != 0) > 0.5
        _agree(a, b, ("rowc", order, nb, nd, pmax))


def torch_int16():
    import torch
    return torch.int16


def torch_int32():
    import torch
    return torch.int32


def test_compact_host_sweep_matches_plain_host_sweep():
    """sweep_host_compact (phase-major repack, only live phase blocks cross PCIe) returns what sweep_host returns."""
    from fhmcanalysis_b200 import engine, synth
    n = 301
    lnpi = synth.two_peak_lnpi(n, noise=1e-3, scale=0.3)
    N = np.arange(n, dtype=np.float64)
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=5, sel=["N", N * N])
    for S, chunk, narrow in ((5000, 2048, True), (5000, 2048, False), (4097, 1 << 18, None), (1, 1 << 18, True)):
        mu = np.linspace(-0.12, 0.10, S)
        a = dh.sweep_host(mu, pmax=4, chunk=chunk)
        c = dh.sweep_host_compact(mu, pmax=4, chunk=chunk, narrow=narrow)
        st = a["status"].numpy().view(np.uint32)
        assert c["status"].dtype == (torch_int32() if narrow is False else torch_int16())   # narrow records by default
        assert np.array_equal(c["status"].numpy().astype(np.int64), st.astype(np.int64))
        P = a["nphase"].numpy()
        assert np.array_equal(c["nphase"].numpy(), P) and c["max_nphase"] == P.max()
        fe_a, fe_c = a["fe"].numpy(), c["fe"].numpy()
        av_a, av_c = a["avg"].numpy(), c["avg"].numpy()
        b_a, b_c = a["bounds"].numpy(), c["bounds"].numpy()
        assert fe_c.shape == fe_a.shape and av_c.shape == av_a.shape and b_c.shape == b_a.shape
        for p in range(4):
            live = ((st & 0xFF) == 0) & (P > p)
            assert np.array_equal(fe_c[live, p], fe_a[live, p])
            assert np.array_equal(av_c[live, p], av_a[live, p])
            assert np.array_equal(b_c[live, p], b_a[live, p])
            assert np.all(np.isnan(fe_c[~live, p])) and np.all(b_c[~live, p] == -1)
        assert c["d2h_bytes"] <= 8 * S + 4 * S * (16 + 16) + 64
        if narrow is not False and S == 5000:
            assert c["d2h_bytes"] <= 4 * S + c["max_nphase"] * S * 28 + 64
    # reuse of the result buffers: a later call with fewer phases must not leave stale phase blocks behind
    mu2 = np.full(5000, -3.0)            # far from coexistence: one phase everywhere
    c2 = dh.sweep_host_compact(mu2, pmax=4, chunk=2048, out=dh.sweep_host_compact(np.linspace(-0.12, 0.10, 5000), pmax=4, chunk=2048))
    assert c2["max_nphase"] == 1 and np.all(np.isnan(c2["fe"].numpy()[:, 1:]))


@pytest.mark.parametrize("narrow", [True, False])
def test_compact_host_sweep_first_call_ascending_phase_count(narrow):
    """First call on a FRESH result buffer (speculative copy guess = 1), ascending mu: the first two chunks are one-phase
    everywhere, the last one is two-phase.  Phase block 1 of the early chunks never crosses PCIe and must still read
    NaN / -1 (the header's contract: only blocks >= max_nphase are undefined)."""
    from fhmcanalysis_b200 import engine, synth
    n = 301
    lnpi = synth.two_peak_lnpi(n, noise=1e-3, scale=0.3)
    N = np.arange(n, dtype=np.float64)
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=5, sel=["N", N * N])
    mu = np.linspace(-3.0, 0.05, 5000)
    a = dh.sweep_host(mu, pmax=4, chunk=2048)
    P = a["nphase"].numpy()
    assert P[:4096].max() == 1 and P[4096:].max() == 2
    c = dh.sweep_host_compact(mu, pmax=4, chunk=2048, narrow=narrow)     # out=None: fresh pinned memory
    assert c["max_nphase"] == 2 and np.array_equal(c["nphase"].numpy(), P)
    fe_c, b_c, av_c = c["fe"].numpy(), c["bounds"].numpy(), c["avg"].numpy()
    for p in range(2):
        live = P > p
        assert np.array_equal(fe_c[live, p], a["fe"].numpy()[live, p])
        assert np.array_equal(b_c[live, p], a["bounds"].numpy()[live, p])
        assert np.all(np.isnan(fe_c[~live, p])) and np.all(np.isnan(av_c[~live, p])) and np.all(b_c[~live, p] == -1)
    assert np.all(np.isnan(fe_c[:, 2:])) and np.all(b_c[:, 2:] == -1)


def test_lean_solver_kernel_matches_group_kernel(monkeypatch):
    """K4: the default warp-per-solve kernel on the lean evaluator (k_solve_lean) and the general PointEval group kernel run
    the same iteration on bit-identical evaluations of u: same mu_coex, same integers, same evaluation counts."""
    from fhmcanalysis_b200 import _lib, synth
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    h = histogram.from_arrays(synth.two_peak_lnpi(801, scale=0.8), synth.one_comp_moments(801, max_order=3), 1.0, [0.0], 10)
    betas = 1.0 / np.linspace(0.92, 1.05, 700)
    for order, moments in ((2, ("N", "N2", "U")), (1, ("N",)), (3, ("N", "N2", "U"))):
        dh = h.device_histogram(beta=betas, order=order, moments=moments)
        monkeypatch.setenv("FHMC_SOLVER_LANES", "32")
        a = dh.find_phase_eq(np.zeros_like(betas), beta=betas, lnz_tol=1e-10, pmax=4, continuation=False).host()
        assert _lib.last_kernel() == "k_find_phase_eq"
        monkeypatch.delenv("FHMC_SOLVER_LANES")
        b = dh.find_phase_eq(np.zeros_like(betas), beta=betas, lnz_tol=1e-10, pmax=4, continuation=False).host()
        assert _lib.last_kernel() == "k_solve_lean"
        assert np.array_equal(a["code"], b["code"]) and (a["code"] == 0).mean() > 0.9
        assert np.array_equal(a["status"] & 0x2000, b["status"] & 0x2000) and np.array_equal(a["iters"], b["iters"])
        ok = a["code"] == 0
        assert np.allclose(a["mu_coex"][ok], b["mu_coex"][ok], rtol=0, atol=1e-13)
        assert np.array_equal(a["nphase"][ok], b["nphase"][ok]) and np.array_equal(a["bounds"][ok][:, :2], b["bounds"][ok][:, :2])
        assert np.array_equal(a["max_idx"][ok][:, :2], b["max_idx"][ok][:, :2])
        assert np.allclose(a["fe"][ok, :2], b["fe"][ok, :2], rtol=1e-11, atol=1e-12)
        assert np.allclose(a["avg"][ok, :2], b["avg"][ok, :2], rtol=1e-11, atol=0)
        assert (b["status"][ok] & 0x4000).mean() > 0.9          # the final records come from the lean evaluator
    # pure mu solves (no extrapolation)
    N = np.arange(801.0)
    from fhmcanalysis_b200 import engine
    dh = engine.DeviceHistogram(synth.two_peak_lnpi(801, scale=0.8), N, 1.0, 0.0, smooth=10, sel=["N"])
    g = np.linspace(-0.05, 0.05, 300)
    monkeypatch.setenv("FHMC_SOLVER_LANES", "32")
    a = dh.find_phase_eq(g).host()
    monkeypatch.delenv("FHMC_SOLVER_LANES")
    b = dh.find_phase_eq(g).host()
    assert _lib.last_kernel() == "k_solve_lean"
    assert np.all(a["code"] == 0) and np.all(b["code"] == 0)
    assert np.allclose(a["mu_coex"], b["mu_coex"], rtol=0, atol=1e-12) and np.ptp(b["mu_coex"]) < 1e-9


def _soa_equal_to_records(c, h, lo, S, pmax, status_mask=0xFFFF):
    """compact views c (records lo .. lo+S-1) against the host() dict of a plain sweep of the same state points"""
    st = c["status"].cpu().numpy()[lo:lo + S].astype(np.int64) & status_mask
    assert np.array_equal(st, h["status"].astype(np.int64) & status_mask)
    P = h["nphase"]
    assert np.array_equal(c["nphase"].cpu().numpy()[lo:lo + S], P)
    fe, av, bd = (c[k].cpu().numpy()[lo:lo + S] for k in ("fe", "avg", "bounds"))
    for p in range(pmax):
        live = (h["code"] == 0) & (P > p)
        assert np.array_equal(fe[live, p], h["fe"][live, p]) and np.array_equal(av[live, p], h["avg"][live, p])
        assert np.array_equal(bd[live, p], h["bounds"][live, p])
        assert np.all(np.isnan(fe[~live, p])) and np.all(np.isnan(av[~live, p])) and np.all(bd[~live, p] == -1)


def test_compact_records_written_by_the_sweep_kernel():
    """fhmc_sweep_1d_compact: the headline kernel writes the phase-major narrow records itself (k_sweep_prod2<compact>) -- same
    bits as a plain sweep followed by the repack; several destinations (as for a gather fused over NVLink peers), records at
    an offset inside a larger layout, dead slots written by the kernel or left to a 0xFF-filled buffer; the general path
    (small sweeps, three averaged quantities) through the same entry point."""
    import torch
    from fhmcanalysis_b200 import _lib, engine, synth
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=np.float64)
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    dh.use_mu_tables = False      # the table-driven kernel has its own test (tests/test_gpu_tables.py)
    S = 200001
    mu = np.concatenate([np.linspace(-0.03, 0.03, S - 4001), np.linspace(-6.0, 6.0, 4001)])   # strong tilts: queue / general evaluator
    h = dh.sweep_auto(mu, pmax=4).host()
    assert h["fe"].shape[1] == 4
    top = torch.zeros(1, dtype=torch.int32, device="cuda")
    c = dh.sweep_compact(mu, pmax=4, max_nphase=top)
    assert _lib.last_kernel() == "k_sweep_prod2<compact>"
    _soa_equal_to_records(c, h, 0, S, 4)
    assert int(top.item()) == int(h["nphase"][h["code"] == 0].max())
    # two destinations, records at an offset, dead slots left to the 0xFF fill
    n_total, first = S + 5000, 1234
    nbytes = int(_lib.load().fhmc_pack_soa16_bytes(n_total, 4, 2))
    d0 = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device="cuda")
    d1 = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device="cuda")
    dh.sweep_compact(mu, pmax=4, dst=[d0.data_ptr(), d1.data_ptr()], n_total=n_total, first=first, fill_dead=False)
    torch.cuda.synchronize()
    assert torch.equal(d0, d1)
    v = engine.soa16_views(d0, n_total, 4, 2)
    _soa_equal_to_records(v, h, first, S, 4)
    fe_all = v["fe"].cpu().numpy()
    assert np.all(np.isnan(fe_all[:first])) and np.all(np.isnan(fe_all[first + S:]))      # nothing outside the launch's range
    # general path: a small sweep, and three averaged quantities (no two-point product-form instantiation)
    c = dh.sweep_compact(mu[:3000], pmax=4)
    assert _lib.last_kernel() != "k_sweep_prod2<compact>"
    _soa_equal_to_records(c, dh.sweep(mu[:3000], pmax=4).host(), 0, 3000, 4)      # the same kernel choice inside: same bits
    dh3 = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N, -2.0 * N])
    h3 = dh3.sweep_auto(mu, pmax=4).host()
    c3 = dh3.sweep_compact(mu, pmax=4)
    _soa_equal_to_records(c3, h3, 0, S, 4)
