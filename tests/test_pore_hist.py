"""pore_hist normalise / thermo (reference moments/histogram/two_dim/h_ntot/pore_hist.pyx).
CPU: the oracle restatement against vectors recorded from the COMPILED reference (tests/golden/make_golden_pore.py).
GPU: fhmc_masked_lse_2d through the drop-in class against the oracle and the same vectors."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pore_golden():
    return np.load(os.path.join(HERE, "golden", "pore_vectors.npz"))


def test_oracle_normalize_matches_compiled_reference(pore_golden):
    from oracle import fhmc_oracle as fo
    for name in ("small", "mid", "steep"):
        ref = pore_golden[name + "/normalized"]
        got = fo.pore_normalize(pore_golden[name + "/lnpi"], pore_golden[name + "/edge"])
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(got))
        assert np.array_equal(got[fin], ref[fin]), name     # same fold order, same libm: bit-identical


def test_oracle_thermo_restatement_properties():
    """thermo() cannot be pinned on the reference (it raises, pore_hist.pyx:170): check the restatement on cases with
    known answers instead."""
    from oracle import fhmc_oracle as fo
    lp = np.log(np.array([[1.0, 2.0, 1.0], [4.0, 1.0, 1.0]]))
    props = {"n": np.array([[0.0, 1.0, 2.0], [0.0, 1.0, 2.0]]), "one": np.ones((2, 3))}
    mask = np.array([[True, True, False], [False, False, True]])
    ave, peak = fo.pore_thermo(lp, mask, props)
    assert abs(ave["one"] - 1.0) < 1e-15
    assert abs(ave["n"] - (2.0 * 1 + 1.0 * 2) / 4.0) < 1e-15
    assert (peak[0].tolist(), peak[1].tolist()) == ([0], [1])


def _make_pore(pore_golden, name, n_prop=3, seed=5):
    from fhmcanalysis_b200.moments.histogram.two_dim.h_ntot import pore_hist as ph
    lp, edge = pore_golden[name + "/lnpi"], pore_golden[name + "/edge"]
    rng = np.random.default_rng(seed)
    obj = ph.pore_hist.__new__(ph.pore_hist)

    class _JH(object):
        pass
    jh = _JH()
    jh.data = {"props": {"p%d" % k: rng.normal(size=lp.shape) * (k + 1) for k in range(n_prop)}}
    obj.data = {"ln(PI)": lp.copy(), "edge_idx": edge.copy(), "hist": jh}
    return obj


@pytest.mark.gpu
def test_normalize_matches_compiled_reference(pore_golden):
    for name in ("small", "mid", "steep"):
        obj = _make_pore(pore_golden, name)
        obj.normalize()
        ref = pore_golden[name + "/normalized"]
        got = obj.data["ln(PI)"]
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(got))
        # the normalisation constant is a parallel log-sum-exp here and a sequential fold there: 1e-10 relative on it
        c_ref = (pore_golden[name + "/lnpi"][fin] - ref[fin])[0]
        c_got = (pore_golden[name + "/lnpi"][fin] - got[fin])[0]
        assert abs(c_got - c_ref) <= 1e-10 * max(1.0, abs(c_ref))
        assert np.allclose(got[fin], ref[fin], rtol=1e-10, atol=1e-10)
        # normalised: sum of probabilities over the support is one
        assert abs(np.sum(np.exp(got[fin])) - 1.0) < 1e-12


@pytest.mark.gpu
def test_thermo_matches_oracle(pore_golden):
    from oracle import fhmc_oracle as fo
    for name, n_prop in (("small", 2), ("mid", 3), ("steep", 11)):   # 11 > 8: more than one launch
        obj = _make_pore(pore_golden, name, n_prop=n_prop)
        obj.normalize()
        lp = obj.data["ln(PI)"]
        n1, n2 = lp.shape
        fin = np.isfinite(lp)
        for mask in (fin & (np.arange(n2)[None, :] < n2 // 2), fin & (np.arange(n2)[None, :] >= n2 // 2), fin):
            got = obj.thermo(mask)
            ave, peak = fo.pore_thermo(lp, mask, obj.data["hist"].data["props"])
            for p in ave:
                assert abs(got[p] - ave[p]) <= 1e-10 * max(1.0, abs(ave[p])), (name, p)
            assert got["peak_idx"][0].tolist() == peak[0].tolist() and got["peak_idx"][1].tolist() == peak[1].tolist()


@pytest.mark.gpu
def test_thermo_ties_and_empty_mask(pore_golden):
    from fhmcanalysis_b200 import engine
    lp = np.zeros((5, 6))
    lp[2, 3] = lp[4, 1] = 2.0          # two bins tie for the maximum: both are reported, in row-major order
    r = engine.masked_lse_2d(lp, mask=np.ones_like(lp, dtype=bool), props=np.ones((1, 5, 6)))
    assert (r["peak_idx"][0].tolist(), r["peak_idx"][1].tolist()) == ([2, 4], [3, 1])
    assert abs(r["lnsum"] - np.log(28 + 2 * np.exp(2.0))) < 1e-13 and abs(r["avg"][0] - 1.0) < 1e-15
    r = engine.masked_lse_2d(lp, mask=np.zeros_like(lp, dtype=bool))
    assert r["lnsum"] == -np.inf and len(r["peak_idx"][0]) == 0
    r = engine.masked_lse_2d(np.zeros((3, 200)), peak_cap=4)     # 600 ties > buffer: the call retries with room for all
    assert len(r["peak_idx"][0]) == 600


@pytest.mark.gpu
def test_pore_hist_constructor_from_joint_hist():
    """The whole constructor (PH:91-137) from a joint histogram: rows shifted to -beta (F(h) + p A h) at N = 0, then
    normalised over the ragged support; compared with the oracle's sequential fold."""
    from oracle import fhmc_oracle as fo
    from fhmcanalysis_b200.moments.histogram.two_dim.h_ntot.pore_hist import pore_hist
    rng = np.random.default_rng(11)
    n1, n2 = 6, 12
    hs = np.linspace(1.0, 2.0, n1)
    edge = np.array([5, 6, 8, 9, 10, 11])

    class _JH(object):
        def make(self):
            pass
    jh = _JH()
    lp = rng.normal(size=(n1, n2))
    for i in range(n1):
        lp[i, edge[i] + 1:] = -np.inf
    jh.data = {"ln(PI)": lp, "op_1": hs, "op_2": np.arange(n2), "bounds_idx": np.stack([np.zeros(n1, dtype=int), edge], axis=1),
               "props": {"N": np.tile(np.arange(n2, dtype=float), (n1, 1))}}
    fh = lambda h: 0.3 * h * h     # noqa: E731
    obj = pore_hist(jh, fh, 0.2, 1.5, 0.9)
    want = lp.copy()
    for i, h in enumerate(hs):
        want[i, :] += -0.9 * (fh(h) + 0.2 * 1.5 * h) - lp[i, 0]
    want = fo.pore_normalize(want, edge)
    fin = np.isfinite(want)
    assert np.array_equal(obj.data["mask"], fin)
    assert np.allclose(obj.data["ln(PI)"][fin], want[fin], rtol=1e-10, atol=1e-12)
    with pytest.raises(NotImplementedError):
        obj.phase_average()
