"""GPU: the reference's unit tests for the N_tot histogram that test_gpu_dropin.py does not already restate
(unittests/moments_histogram_one_dim_gc_ntot.py, "T1": test_init/load/clear T1:28-98, test_temp_extrap_2 T1:361,
test_dmu2_extrap_1/2 T1:377-438, test_temp_dmu2_extrap_1/2 T1:440-527, the *_ke variants T1:529-878 and
test_mix_symmetric/asymmetric T1:880-984), driven through the same constructor on a composite.nc that the built-in
writer makes from the fixture arrays (tests/golden: testnc/*; /root/reference does not exist on the GPU box).
Where T1 writes ``np.all(a - b) < tol`` (a no-op comparison) the intended ``|a - b| < tol`` is asserted."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BETA, MU, SMOOTH = 1.0, [5.0, 0.0], 1
T1_LNPI = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0], dtype=np.float64)


@pytest.fixture(scope="module")
def fname(tmp_path_factory, golden, golden_meta):
    from fhmcanalysis_b200.io.hdf5_min import write_composite
    p = str(tmp_path_factory.mktemp("t1") / "test.nc")
    m = golden_meta["testnc"]
    write_composite(p, golden["testnc/lnpi"], golden["testnc/ntot"], golden["testnc/mom"], m["volume"], m["nspec"], m["max_order"],
                    history=m["history"])
    return p


def H():
    from FHMCAnalysis.moments.histogram.one_dim.ntot import gc_hist as oneDH      # the import path T1:10 uses
    return oneDH.histogram


def lse(x):
    m = np.max(x)
    return m + np.log(np.sum(np.exp(x - m)))


def gc(h, x):
    w = np.exp(h.data["ln(PI)"] - np.max(h.data["ln(PI)"]))
    return np.sum(w * x) / np.sum(w)


def t1_moments():
    mom = np.ones((2, 3, 2, 3, 3, 31), dtype=np.float64)
    n = np.arange(0, 31)
    for idx in ((0, 1, 0, 0), (0, 1, 1, 0), (0, 0, 0, 1), (1, 0, 0, 1)):
        mom[idx] = n
    for idx in ((1, 1, 0, 0), (1, 1, 1, 0), (0, 0, 1, 1), (1, 0, 1, 1)):
        mom[idx] = n * 2
    mom[:, 1, :, 1, :] = 1.234
    return mom


def test_init_load_clear(fname, golden, golden_meta):
    hist = H()(fname, BETA, MU, SMOOTH)
    assert hist.metadata["beta_ref"] == BETA and np.all(hist.metadata["mu_ref"] == MU)
    assert hist.metadata["smooth"] == SMOOTH and hist.metadata["fname"] == fname and hist.metadata["used_ke"] is False
    assert hist.metadata["file_history"] == golden_meta["testnc"]["history"]
    assert np.array_equal(hist.data["ln(PI)"], golden["testnc/lnpi"]) and hist.data["mom"].shape == (2, 3, 2, 3, 3, 31)
    assert hist.data["lb"] == 0 and hist.data["ub"] == 30 and hist.data["volume"] == golden_meta["testnc"]["volume"]
    assert np.all(hist.data["curr_mu"] == MU) and hist.data["curr_beta"] == BETA
    hist.clear()
    assert hist.data == {} and hist.metadata["fname"] == fname
    hist.reload()
    assert np.array_equal(hist.data["ln(PI)"], golden["testnc/lnpi"])
    with pytest.raises(Exception):
        H()("does/not/exist.nc", BETA, MU, SMOOTH)
    with pytest.raises(AssertionError):
        H()(fname, -1.0, MU, SMOOTH)
    with pytest.raises(AssertionError):
        H()(fname, BETA, [5.0], SMOOTH)        # nspec mismatch, GH:149


@pytest.mark.parametrize("ke", [False, True])
def test_temp_extrap_2_needs_higher_order(fname, ke):
    hist = H()(fname, BETA, MU, SMOOTH, ke)
    assert hist.metadata["used_ke"] == ke
    with pytest.raises(Exception, match="Maximum order"):
        hist.temp_extrap(2.0 * BETA, 2, 10.0, True, True)


@pytest.mark.parametrize("ke", [False, True])
def test_temp_extrap_1_known_answer(fname, ke):
    """T1:309-359 / 529-581: identical with and without KE, dlnPI/dB has the same structure."""
    hist = H()(fname, BETA, MU, SMOOTH, ke)
    hist.data["mom"] = t1_moments()
    hist.data["ln(PI)"] = T1_LNPI.copy()
    beta = 2.0 * hist.data["curr_beta"]
    hist.normalize()
    lnpi0 = copy.copy(hist.data["ln(PI)"])
    ave_n2, ave_ntot, ave_u = 20.1996548887, 30.2994823331, 1.0
    n = np.arange(0, 31)
    dlnpi = hist.data["curr_mu"][0] * (n - ave_ntot) + (hist.data["curr_mu"][1] - hist.data["curr_mu"][0]) * (n * 2 - ave_n2) - (np.ones(31) - ave_u)
    ans = lnpi0 + dlnpi * (beta - hist.data["curr_beta"])
    ans -= lse(ans)
    new = hist.temp_extrap(beta, 1, 10.0, True, True, True)
    assert np.all(np.abs(ans - new.data["ln(PI)"]) < 1e-9)      # the rounded <N> constants of T1 cancel on renormalisation
    assert abs(beta - new.data["curr_beta"]) < 1e-12


def test_dmu2_extrap_1_and_2(fname):
    hist = H()(fname, BETA, MU, SMOOTH)
    n2 = hist.data["mom"][1, 1, 0, 0, 0]
    for order in (1, 2):
        newh = hist.dmu_extrap(np.array([-4.0]), order, 10.0, True, True, order == 2)
        assert np.all(newh.data["curr_mu"] == [5.0, 1.0]) and newh.data["curr_beta"] == BETA
        base = copy.deepcopy(hist)
        base.normalize()                                       # dmu_extrap normalises before extrapolating (GH:789)
        check = base.data["ln(PI)"] + base.data["curr_beta"] * (n2 - gc(base, n2)) * 1.0
        if order == 2:
            f_tilde = BETA * BETA * (hist.data["mom"][1, 2, 0, 0, 0] - n2 * n2)
            f_hat = BETA * BETA * (gc(base, hist.data["mom"][1, 2, 0, 0, 0]) - gc(base, n2) ** 2)
            check = check + 0.5 * (f_tilde - f_hat)
        check -= lse(check)
        newh.normalize()
        assert np.all(np.abs(newh.data["ln(PI)"] - check) < 1e-10)


@pytest.mark.parametrize("ke", [False, True])
def test_temp_dmu2_extrap_1_and_2(fname, ke):
    """T1:440-527 and 603-657: the Hessian is assembled 'VERY manually' from the private builders, like T1 does."""
    hist = H()(fname, BETA, MU, SMOOTH, ke)
    tb = 2.0 * hist.data["curr_beta"]
    n2, U, ntot = hist.data["mom"][1, 1, 0, 0, 0], hist.data["mom"][0, 0, 0, 0, 1], hist.data["ntot"]
    for order in (1, 2):
        newh = hist.temp_dmu_extrap(tb, np.array([-4.0]), order, 10.0, True, True, True)
        assert np.all(newh.data["curr_mu"] == [5.0, 1.0]) and newh.data["curr_beta"] == tb
        base = copy.deepcopy(hist)
        base.normalize()
        mu = base.data["curr_mu"]
        check = base.data["ln(PI)"] + base.data["curr_beta"] * (n2 - gc(base, n2)) * 1.0
        dlnpi = mu[0] * (ntot - gc(base, ntot)) + (mu[1] - mu[0]) * (n2 - gc(base, n2)) - (U - gc(base, U))
        check = check + dlnpi * (tb - base.data["curr_beta"])
        if order == 2:
            xi = np.array([tb - base.data["curr_beta"], 1.0])
            Hs = np.zeros((2, 2, 31))
            Hs[0, 0] = (-mu[0] * base._gc_dX_dB([0, 0, 0, 0, 0], 1)
                        + (mu[1] - mu[0]) * (base._sg_dX_dB([1, 1, 0, 0, 0], 0) - base._gc_dX_dB([1, 1, 0, 0, 0], 0))
                        - (base._sg_dX_dB([0, 0, 0, 0, 1], 0) - base._gc_dX_dB([0, 0, 0, 0, 1], 0)))
            Hs[0, 1] = (n2 - gc(base, n2)) + base.data["curr_beta"] * (base._sg_dX_dB([1, 1, 0, 0, 0], 0) - base._gc_dX_dB([1, 1, 0, 0, 0], 0))
            Hs[1, 0] = Hs[0, 1]
            Hs[1, 1] = base.data["curr_beta"] ** 2 * ((hist.data["mom"][1, 2, 0, 0, 0] - n2 ** 2) - base._gc_fluct_ii([1, 1, 0, 0, 0], [1, 1, 0, 0, 0]))
            for i in range(31):
                check[i] += 0.5 * np.sum(np.dot(xi, Hs[:, :, i]) * xi)
        check -= lse(check)
        newh.normalize()
        assert np.all(np.abs(newh.data["ln(PI)"] - check) < 1e-9 * max(1.0, np.max(np.abs(check))))


def test_ke_differences_of_the_private_builders(fname):
    """T1:583-878 (test_dlnpi_1_ke, test_dlnpi_2_ke, test_sg_dx_ke, test_gc_dx_ke): where KE terms enter."""
    ke, pe = H()(fname, BETA, MU, SMOOTH, True), H()(fname, BETA, MU, SMOOTH, False)
    for h in (ke, pe):
        h.normalize()
    ntot = pe.data["ntot"].astype(float)
    assert np.all(np.abs(ke._dB()[0] - pe._dB()[0]) < 1e-12)
    ave_ntot = gc(pe, ntot)
    assert np.all(np.abs((ke._dB2(True)[0] - pe._dB2(True)[0]) - 1.5 / BETA / BETA * (ntot - ave_ntot)) < 1e-9)
    for x, extra in (([0, 0, 0, 0, 1], 1.0), ([0, 1, 0, 0, 1], pe.data["mom"][0, 1, 0, 0, 0]), ([0, 1, 0, 1, 1], pe.data["mom"][0, 1, 0, 1, 0])):
        d = pe._sg_dX_dB(x, 0) - ke._sg_dX_dB(x, 0)
        assert np.all(np.abs(d - 1.5 / BETA / BETA * ntot * extra) < 1e-9)
    # quantities without an energy factor do not change
    assert np.all(ke._sg_dX_dB([0, 1, 0, 0, 0], 0) == pe._sg_dX_dB([0, 1, 0, 0, 0], 0))
    d = pe._gc_dX_dB([0, 0, 0, 0, 1], 0) - ke._gc_dX_dB([0, 0, 0, 0, 1], 0)
    assert abs(d - 1.5 / BETA / BETA * ave_ntot) < 1e-9


def _mix_pair(fname):
    h1 = H()(fname, BETA, MU, SMOOTH)
    h1.data["mom"] = t1_moments()
    h1.data["ln(PI)"] = T1_LNPI.copy()
    h2 = H()(fname, BETA, MU, SMOOTH)
    h2.data["mom"] = h1.data["mom"] * 2
    h2.data["ln(PI)"] = h1.data["ln(PI)"] * 2
    return h1, h2


@pytest.mark.parametrize("w", [[1.0, 1.0], [1.0, 0.1234]])
def test_mix_symmetric(fname, w):
    h1, h2 = _mix_pair(fname)
    mixed = h1.mix(h2, w)
    assert np.all(np.abs(mixed.data["ln(PI)"] - (T1_LNPI * w[0] + 2.0 * T1_LNPI * w[1]) / (w[0] + w[1])) < 1e-9)
    assert np.all(np.abs(mixed.data["mom"] - (h1.data["mom"] * w[0] + h2.data["mom"] * w[1]) / (w[0] + w[1])) < 1e-9)


@pytest.mark.parametrize("w", [[1.0, 1.0], [1.0, 0.1234]])
def test_mix_asymmetric(fname, w):
    h1, h2 = _mix_pair(fname)
    h2.data["mom"] = np.ascontiguousarray(h2.data["mom"][..., :29])       # trim the last two bins (T1:952-956)
    h2.data["ln(PI)"] = h2.data["ln(PI)"][:29].copy()
    h2.data["ntot"] = h2.data["ntot"][:29].copy()
    h2.data["ub"] = 28
    mixed = h1.mix(h2, w)
    assert len(mixed.data["ln(PI)"]) == 31
    assert np.all(np.abs(mixed.data["ln(PI)"][:29] - (w[0] + 2.0 * w[1]) / (w[0] + w[1]) * T1_LNPI[:29]) < 1e-9)
    assert np.all(mixed.data["ln(PI)"][29:] == T1_LNPI[29:])
    assert np.all(np.abs(mixed.data["mom"][..., :29] - (h1.data["mom"][..., :29] * w[0] + h2.data["mom"] * w[1]) / (w[0] + w[1])) < 1e-9)
    assert np.all(mixed.data["mom"][..., 29:] == h1.data["mom"][..., 29:])
    with pytest.raises(Exception, match="Requires 2 weights"):
        h1.mix(h2, [1.0])
