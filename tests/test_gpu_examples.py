"""The reference's own example composites (example/ntot/binary_ideal_gas, square_well) as regression data: vectors recorded
from the COMPILED reference by tests/golden/make_golden_examples.py.  Drop-in scalar API and the
batched entry point against them: integers bit-exact, fp64 to 1e-10 relative."""
import copy
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["ig_T1.00_m1.10", "ig_T1.00_0.00", "ig_T1.00_2.94", "ig_T1.00_1.10", "ig_T1.00_m2.94", "ig_T1.20_1.10", "ig_T1.20_m2.94",
        "sw_T1.10"]


@pytest.fixture(scope="module")
def ex():
    z = np.load(os.path.join(HERE, "golden", "examples_vectors.npz"))
    with open(os.path.join(HERE, "golden", "examples_vectors.json")) as fh:
        return {k: z[k] for k in z.files}, json.load(fh)


@pytest.mark.parametrize("key", KEYS)
def test_example_composite_sweep_and_extrapolation(ex, key):
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    g, meta = ex
    m = meta[key]
    two = len(m["mu_ref"]) == 2
    h0 = oneDH.histogram.from_arrays(g[key + "/lnpi"], g[key + "/mom"], m["beta_ref"], m["mu_ref"], m["smooth"], volume=m["volume"])
    # ---- scalar drop-in path: reweight -> thermo -> is_safe, exactly as the reference was driven ----------------------
    for k, mu in enumerate(m["mus"]):
        assert m["status"][k] == "ok"
        h = copy.deepcopy(h0)
        h.reweight(mu)
        h.thermo()
        pre = "%s/%d/" % (key, k)
        assert h.data["ln(PI)_maxima_idx"].tolist() == g[pre + "maxima"].tolist()
        assert h.data["ln(PI)_minima_idx"].tolist() == g[pre + "minima"].tolist()
        P = len(h.data["thermo"])
        assert [list(h.data["thermo"][p]["bound_idx"]) for p in range(P)] == g[pre + "bounds"].tolist()
        assert bool(h.is_safe()) == bool(g[pre + "safe"])
        assert np.allclose(h.data["ln(PI)"], g[pre + "lnpi"], rtol=1e-10, atol=1e-10)
        for name in ("fe", "n1", "u", "density") + (("n2", "x1") if two else ()):
            got = np.array([h.data["thermo"][p]["F.E./kT" if name == "fe" else name] for p in range(P)])
            assert np.allclose(got, g[pre + name], rtol=1e-10, atol=1e-12), (key, k, name)
    # ---- batched entry point on the same state points -------------------------------------------------------------------
    out = h0.reweight_batch(np.array(m["mus"]))
    for k in range(len(m["mus"])):
        pre = "%s/%d/" % (key, k)
        P = int(out["nphase"][k])
        assert out["max_idx"][k, :P].tolist() == g[pre + "maxima"].tolist()
        assert out["bounds"][k, :P].tolist() == g[pre + "bounds"].tolist()
        assert np.allclose(out["fe"][k, :P], g[pre + "fe"], rtol=1e-10, atol=1e-12)
    # ---- Taylor extrapolation of ln(PI), orders 1 and 2 (skip_mom=True) ------------------------------------------------
    for order in (1, 2):
        h = copy.deepcopy(h0)
        h.reweight(m["mus"][1])
        if two:
            hn = h.temp_dmu_extrap(m["extrap"]["beta"], np.array(m["extrap"]["dmu"]), order, 10.0, True, True, True, False)
        else:
            hn = h.temp_extrap(m["extrap"]["beta"], order, 10.0, True, True, True)
        ref = g["%s/extrap%d" % (key, order)]
        assert np.max(np.abs(hn.data["ln(PI)"] - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))


def test_isopleth_over_the_ideal_gas_composites(ex):
    """isopleth.make_grid_multi (gc_binary.pyx:173-290) over the five T* = 1.00 binary ideal-gas composites of the reference's
    examples: same cells filled, same x1 / density / F.E. grids as the compiled reference."""
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist as oneDH
    import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_binary as gcB
    g, meta = ex
    mi = meta["iso"]
    hists = []
    for k in mi["keys"]:
        m = meta[k]
        hists.append(oneDH.histogram.from_arrays(g[k + "/lnpi"], g[k + "/mom"], m["beta_ref"], m["mu_ref"], m["smooth"], volume=m["volume"]))
    iso = gcB.isopleth(hists, mi["beta_ref"], mi["order"])
    Z, (X, Y) = iso.make_grid_multi(mi["mu1_bounds"], mi["dmu2_bounds"], mi["delta"], mi["m"])
    assert np.allclose(X, g["iso/X"]) and np.allclose(Y, g["iso/Y"])
    assert np.array_equal(Z == 0, g["iso/x1"] == 0)
    assert np.mean(Z != 0) > 0.5
    assert np.allclose(Z, g["iso/x1"], rtol=1e-9, atol=1e-12)
    assert np.allclose(iso.data["density"], g["iso/density"], rtol=1e-9, atol=1e-12)
    assert np.allclose(iso.data["F.E./kT"], g["iso/fe"], rtol=1e-9, atol=1e-12)
