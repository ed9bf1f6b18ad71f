"""CPU: the host-side Taylor coefficient builders (fhmcanalysis_b200/.../_taylor.py) against the compiled reference's own
_dB/_dB2/_dB3/_dMU/_dMU2/_dBMU2/_gc_* (gc_hist.pyx:1241-2563) on a max_order-4, two-species tensor, with and without the
kinetic-energy terms.  (The reference's unit tests call these privates directly, T1:505-509, 659-878.)"""
import numpy as np
import pytest


def _hist(golden, golden_meta, oracle, ke):
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    meta = golden_meta["h"]
    h = histogram.from_arrays(golden["h/lnpi"], golden["h/mom"], meta["beta_ref"], meta["mu_ref"], meta["smooth"], ke=ke)
    # the reference state: reweight(mu1) (done with the oracle here: no GPU in this test)
    h.data["ln(PI)"] = oracle.reweight(h.data["ln(PI)"], h.data["ntot"], meta["mu1"], h.data["curr_mu"][0], h.data["curr_beta"])
    h.data["curr_mu"] = h.data["curr_mu"] + (meta["mu1"] - h.data["curr_mu"][0])
    return h


def _pick(t, sample, lead=()):
    return np.array([t[tuple(lead) + tuple(a)] for a in sample])


def _close(a, b, rtol=1e-11):
    scale = np.maximum(np.abs(b), 1e-9 * np.max(np.abs(b)) + 1e-300)
    return np.max(np.abs(a - b) / scale) < rtol


@pytest.mark.parametrize("ke", [False, True])
def test_derivative_builders_match_reference(golden, golden_meta, oracle, ke):
    sample = golden_meta["h"]["sample"]
    tag = "h/ke%d" % int(ke)
    h = _hist(golden, golden_meta, oracle, ke)
    d1, dm1 = h._dB(False)
    assert _close(d1, golden[tag + "/dB"]) and _close(_pick(dm1, sample), golden[tag + "/dB_mom"])
    assert abs(np.sum(np.abs(dm1)) / float(golden[tag + "/dB_sum"]) - 1) < 1e-12
    d2, dm2 = h._dB2(False)
    assert _close(d2, golden[tag + "/dB2"], 1e-9) and _close(_pick(dm2, sample), golden[tag + "/dB2_mom"], 1e-9)
    assert abs(np.sum(np.abs(dm2)) / float(golden[tag + "/dB2_sum"]) - 1) < 1e-11
    dmu, dmm = h._dMU(False)
    assert _close(dmu, golden[tag + "/dMU"]) and _close(_pick(dmm, sample, (0,)), golden[tag + "/dMU_mom"])
    H, Hm = h._dMU2(False)
    assert _close(H, golden[tag + "/dMU2"], 1e-9) and _close(_pick(Hm, sample, (0, 0)), golden[tag + "/dMU2_mom"], 1e-9)
    Hl, Hm = h._dBMU2(False)
    assert _close(Hl, golden[tag + "/dBMU2"], 1e-9) and _close(_pick(Hm, sample, (0, 1)), golden[tag + "/dBMU2_mom01"], 1e-9)
    assert abs(np.sum(np.abs(Hm)) / float(golden[tag + "/dBMU2_sum"]) - 1) < 1e-11
    gc = np.array([h._gc_dX_dB([0, 1, 0, 0, 0], 0), h._gc_dX_dB([0, 0, 0, 0, 1], 1), h._gc_d2X_dB2([1, 1, 0, 0, 0], 0),
                   h._gc_df_dB_ii(([0, 1, 0, 0, 0], 0), ([0, 0, 0, 0, 1], 0)), h._gc_df_dB_in(([1, 1, 0, 0, 0], 0), 1),
                   h._gc_fluct_ii([0, 1, 0, 0, 0], [1, 1, 0, 0, 0])])
    assert np.allclose(gc, golden[tag + "/gc"], rtol=1e-9, atol=1e-9 * np.max(np.abs(golden[tag + "/gc"])))
    if not ke:
        d3, dm3 = h._dB3(False)
        assert _close(d3, golden[tag + "/dB3"], 1e-8) and _close(_pick(dm3, sample), golden[tag + "/dB3_mom"], 1e-8)
    else:
        with pytest.raises(Exception):
            h._dB3(False)


def test_mom_prod_rules_and_errors(golden, golden_meta, oracle):
    h = _hist(golden, golden_meta, oracle, False)
    assert h._mom_prod([0, 1, 0, 0, 0], [1, 1, 0, 0, 0]).tolist() == [0, 1, 1, 1, 0]     # N1 * N2
    assert h._mom_prod([1, 1, 0, 0, 0], [1, 1, 0, 0, 1]).tolist() == [0, 0, 1, 2, 1]     # N2 * N2 U
    assert h._mom_prod([0, 2, 0, 1, 0], [0, 0, 0, 0, 2]).tolist() == [0, 3, 0, 0, 2]     # N1^3 * U^2
    assert h._mom_prod([0, 3, 0, 0, 0], [0, 2, 0, 0, 0]).tolist() == [0, 4, 0, 1, 0]     # overflow spills to slot 2
    with pytest.raises(AssertionError):
        h._mom_prod([0, 0, 0, 0, 3], [0, 0, 0, 0, 2])                                    # U^5 > max_order 4
    with pytest.raises(Exception):
        h._sg_dX_dB([0, 4, 0, 0, 0], 0)                                                  # max_order too low
    assert np.all(h._sg_dX_dB([0, 0, 1, 0, 0], 0) == 0)                                  # constant moment


def test_taylor_rows_closed_form(golden, golden_meta, oracle):
    """The coefficient rows handed to the kernels == N-dependent part of the reference's _dBMU/_dBMU2 lnPI derivatives."""
    from fhmcanalysis_b200 import _lib
    h = _hist(golden, golden_meta, oracle, False)
    rows = dict((k, r) for k, r in h.taylor_rows(2) if not isinstance(r, str))
    N = h.data["ntot"].astype(float)
    mu1 = h.data["curr_mu"][0]
    d1, _ = h._dBMU(True)
    H, _ = h._dBMU2(True)

    def same_up_to_const(a, b, tol=1e-9):
        d = a - b
        return np.max(np.abs(d - d.mean())) < tol * max(1.0, np.max(np.abs(b)))
    assert same_up_to_const(mu1 * N + rows[_lib.M_DB], d1[0])
    assert same_up_to_const(rows[_lib.M_DD], d1[1])
    assert same_up_to_const(rows[_lib.M_DB2], H[0, 0])
    assert same_up_to_const(rows[_lib.M_DBDD], H[0, 1])
    assert same_up_to_const(rows[_lib.M_DD2], H[1, 1])
