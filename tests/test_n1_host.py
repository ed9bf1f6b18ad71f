"""CPU: the host-side derivative builders of the N_1 histogram class against vectors from the compiled reference
(n1/gc_hist.pyx _dBMU/_dBMU2/_dB2/_sg_*/_gc_*, tests/golden/make_golden_n1.py section B) and the closed-form
coefficient rows the kernels consume (the N-dependent part of the gradient/Hessian)."""
import json
import os

import numpy as np
import pytest

from fhmcanalysis_b200 import _lib, synth
from fhmcanalysis_b200.moments.histogram.one_dim.n1.gc_hist import histogram

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "n1_vectors.npz")), json.load(open(os.path.join(HERE, "golden", "n1_vectors.json")))


def sub(m, addr):
    m = np.asarray(m)
    return np.stack([m[..., a[0] % m.shape[-6], a[1], a[2] % m.shape[-4], a[3], a[4], :] for a in addr], axis=-2)


def _hist(g):
    v, meta = g
    s = meta["setup"]
    mom = synth.n1_two_comp_moments(s["n"], 3)
    assert np.allclose([mom.sum(), np.abs(mom).max()], v["mom_checksum"], rtol=1e-13)
    lnpi = v["lnpi"] - np.logaddexp.reduce(v["lnpi"])       # what normalize() leaves (host restatement for this test)
    return histogram.from_arrays(lnpi, mom, s["beta_ref"], s["mu_ref"], s["smooth"], s["volume"]), meta["addr"]


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = max(1.0, float(np.max(np.abs(b))))
    assert np.max(np.abs(a - b)) <= tol * scale, float(np.max(np.abs(a - b)) / scale)


def test_n1_metadata_has_no_ke_key(g):
    h, _ = _hist(g)
    assert "used_ke" not in h.metadata and "n1" in h.data and "ntot" not in h.data


def test_gradient_and_hessian_builders(g):
    v, _ = g
    h, addr = _hist(g)
    d, dm = h._dBMU(False)
    close(d, v["B/dBMU/d"]); close(sub(dm, addr), v["B/dBMU/dm"])
    H, Hm = h._dBMU2(False)
    close(H, v["B/dBMU2/H"]); close(sub(Hm, addr), v["B/dBMU2/Hm"])
    d2, d2m = h._dB2(False)
    close(d2, v["B/dB2/d"]); close(sub(d2m, addr), v["B/dB2/dm"])


def test_private_pointwise_and_gc_derivatives(g):
    v, _ = g
    h, _ = _hist(g)
    close(h._sg_dX_dB([1, 1, 0, 0, 1]), v["B/sg_dX_dB"])
    close(h._sg_dX_dMU(0, [1, 1, 0, 0, 1]), v["B/sg_dX_dMU"])
    close(h._sg_d2X_dB2([1, 1, 0, 0, 0]), v["B/sg_d2X_dB2"])
    close(h._sg_d2X_dMU2(0, 0, [0, 0, 0, 0, 1]), v["B/sg_d2X_dMU2"])
    close(h._sg_df_dB([1, 1, 0, 0, 0], [0, 0, 0, 0, 1]), v["B/sg_df_dB"])
    close(h._sg_df_dMU(0, [1, 1, 0, 0, 0], [0, 0, 0, 0, 1]), v["B/sg_df_dMU"])
    close(h._gc_dX_dB([1, 1, 0, 0, 0]), v["B/gc_dX_dB"])
    close(h._gc_fluct_ii([1, 1, 0, 0, 0], [0, 0, 0, 0, 1]), v["B/gc_fluct_ii"])
    close(h._gc_fluct_vi(h.data["mom"][1, 1, 0, 0, 0], [0, 0, 0, 0, 1]), v["B/gc_fluct_vi"])


def test_kernel_coefficient_rows_are_the_n_dependent_part(g):
    """rows(N) must equal gradient/Hessian minus their N-independent constants (renormalisation removes those)."""
    v, _ = g
    h, _ = _hist(g)
    rows = dict((k, r) for k, r in h.taylor_rows(2))
    n1 = h.data["n1"].astype(float)
    d, H = v["B/dBMU/d"], v["B/dBMU2/H"]
    mu1 = h.data["curr_mu"][0]

    def same_up_to_constant(a, b):
        diff = np.asarray(a) - np.asarray(b)
        assert np.ptp(diff) <= 1e-9 * max(1.0, np.max(np.abs(b)))

    same_up_to_constant(mu1 * n1 + rows[_lib.M_DB], d[0])
    same_up_to_constant(rows[_lib.M_DD], d[1])
    same_up_to_constant(rows[_lib.M_DB2], H[0, 0])
    same_up_to_constant(rows[_lib.M_DBDD], H[0, 1])
    same_up_to_constant(rows[_lib.M_DD2], H[1, 1])
    assert rows[_lib.M_DB_MU1] == "N"


def test_dmu_family_is_absent(g):
    h, _ = _hist(g)
    for name in ("temp_extrap", "dmu_extrap", "temp_dmu_extrap", "temp_dmu_extrap_multi"):
        with pytest.raises(AttributeError):
            getattr(h, name)(1.0)


def test_n1_loader_requires_n1_variable(tmp_path):
    from fhmcanalysis_b200.io import hdf5_min as h5
    n = 12
    mom = synth.n1_two_comp_moments(n, 2)
    extra = {}
    for fam, shp in (("P_{N_i}(N_{1})", (2, n, 3)), ("P_{U}(N_{1})", (n, 3))):
        extra[fam] = np.zeros(shp)
        for sfx in ("lb", "ub", "bw"):
            extra[fam + "_{" + sfx + "}"] = np.zeros(shp[:-1])
    p = str(tmp_path / "n1.nc")
    h5.write_composite(p, -np.arange(n) * 0.1, np.arange(n), mom, 8.0, 2, 2, "t", extra, op_name="N_{1}")
    h = histogram(p, 1.0, [0.0, -1.0], 2)
    assert h.data["n1"].tolist() == list(range(n)) and h.data["pk_hist"]["hist"].shape == (2, n, 3)
    # a file written for the N_tot class has no N_{1}
    p2 = str(tmp_path / "ntot.nc")
    h5.write_composite(p2, -np.arange(n) * 0.1, np.arange(n), mom, 8.0, 2, 2, "t")
    with pytest.raises(KeyError):
        histogram(p2, 1.0, [0.0, -1.0], 2)
    # round trip through to_nc keeps the N_{1} naming
    p3 = str(tmp_path / "again.nc")
    h.to_nc(p3)
    assert np.array_equal(histogram(p3, 1.0, [0.0, -1.0], 2).data["mom"], mom)
