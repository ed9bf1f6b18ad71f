"""Window patching shift solve (reference moments/win_patch/fhmc_patch.pyx:640-709)."""
import numpy as np
import pytest


class _Win(object):
    def __init__(self, lb, ub, lnpi, offset=2):
        self.lb, self.ub, self.lnPI, self.offset = lb, ub, np.asarray(lnpi, dtype=float), offset


def _windows(rng, lb1, ub1, lb2, ub2, shift, noise):
    full = np.cumsum(rng.normal(0.3, 1.0, size=ub1 + 1))
    w2 = _Win(lb2, ub2, full[lb2:ub2 + 1] + 1e-3 * noise * rng.normal(size=ub2 - lb2 + 1))
    w1 = _Win(lb1, ub1, full[lb1:ub1 + 1] - shift + 1e-3 * noise * rng.normal(size=ub1 - lb1 + 1))
    return w1, w2


def test_overlap_slices_host_logic():
    from fhmcanalysis_b200.moments.win_patch.fhmc_patch import overlap_slices
    rng = np.random.default_rng(0)
    w1, w2 = _windows(rng, 40, 100, 0, 60, 7.5, 0.0)
    s1, s2 = overlap_slices(w1, w2)
    assert len(s1) == len(s2) == 21 - 4          # bins 40..60 overlap, two trimmed from each end (offset = 2)
    assert np.allclose(s2 - s1, 7.5, atol=1e-12)
    with pytest.raises(AssertionError):
        overlap_slices(w2, w1)
    with pytest.raises(AssertionError):
        overlap_slices(_Win(70, 100, np.zeros(31)), w2)


def test_oracle_fmin_finds_the_closed_form():
    from oracle import fhmc_oracle as fo
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=40), rng.normal(size=40) + 3.0
    x, e = fo.patch_window_pair_slices(a, b)
    assert abs(x - np.mean(b - a)) < 1e-4           # Nelder-Mead's own x tolerance
    assert abs(e - np.sum((a + np.mean(b - a) - b) ** 2) / 40) < 1e-6


@pytest.mark.gpu
def test_patch_shifts_against_oracle():
    from oracle import fhmc_oracle as fo
    from fhmcanalysis_b200.moments.win_patch import fhmc_patch as fp
    rng = np.random.default_rng(2)
    pairs, wins = [], []
    for k in range(37):
        lb1 = int(rng.integers(20, 60))
        w1, w2 = _windows(rng, lb1, lb1 + int(rng.integers(60, 300)), 0, lb1 + int(rng.integers(8, 50)), rng.normal() * 50, 1.0)
        wins.append((w1, w2))
        pairs.append(fp.overlap_slices(w1, w2))
    shift, err2 = fp.patch_shifts(pairs)
    for k, (s1, s2) in enumerate(pairs):
        exact = np.mean(s2 - s1)
        assert abs(shift[k] - exact) <= 1e-12 * max(1.0, abs(exact))
        assert abs(err2[k] - np.sum((s1 + exact - s2) ** 2) / len(s1)) <= 1e-10 * max(err2[k], 1e-300)
        x, e = fo.patch_window_pair_slices(s1, s2)
        assert abs(shift[k] - x) < 1e-4 and abs(err2[k] - e) <= 1e-6 * max(1.0, e)   # the reference's tolerances (xtol, ftol)
    sx, ex = fp.patch_window_pair(*wins[5])
    assert sx == shift[5] and ex == err2[5]
    assert fp.patch_shifts([])[0].shape == (0,)
    # full-size property: 10^4 pairs of 10^3 bins, shift recovered exactly for noiseless overlaps
    big = [(np.arange(1000.0) * 0.01 + k, np.arange(1000.0) * 0.01 + 2.5 * k) for k in range(10000)]
    s, e = fp.patch_shifts(big)
    assert np.allclose(s, 1.5 * np.arange(10000), rtol=0, atol=1e-9) and np.all(e < 1e-18)


# ---- pinned on the reference itself: oracle/_ref/fhmc_patch (compiled fhmc_patch.pyx) run on its own fixture windows
#      unittests/reference/test_sim/{1,2} and on synthetic window pairs (tests/golden/make_golden_patch.py) ----------------
def _golden_patch():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "patch_vectors.npz"))
    cases = []
    for name in z["names"]:
        name = str(name)
        w = []
        for k in ("w1", "w2"):
            lb, ub, off = (int(v) for v in z["%s/%s/meta" % (name, k)])
            w.append(_Win(lb, ub, z["%s/%s/lnpi" % (name, k)], off))
        cases.append((name, w[0], w[1], z))
    return cases


def test_overlap_and_objective_match_compiled_reference():
    """The slices overlap_slices() cuts and the oracle's restated objective reproduce ``window_patch_error`` of the compiled
    reference at the recorded trial shifts (1e-12 relative); the pair the reference refuses raises the same assertion."""
    from oracle import fhmc_oracle as fo
    from fhmcanalysis_b200.moments.win_patch.fhmc_patch import overlap_slices
    seen_sim = seen_raise = 0
    for name, w1, w2, z in _golden_patch():
        if name + "/raises" in z.files:
            with pytest.raises(AssertionError, match="no overlap"):
                overlap_slices(w1, w2)
            seen_raise += 1
            continue
        s1, s2 = overlap_slices(w1, w2)
        for x, want in zip(z[name + "/trial"], z[name + "/obj"]):
            got = fo.window_patch_error(float(x), s1, s2)
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (name, x)
        shift, err2 = z[name + "/ref"]
        # the reference's answer is the objective at ITS shift, per overlapping bin
        assert abs(fo.window_patch_error(shift, s1, s2) / len(s1) - err2) <= 1e-12 * max(1.0, err2)
        seen_sim += name.startswith("sim")
    assert seen_sim >= 1 and seen_raise >= 1


@pytest.mark.gpu
def test_patch_window_pair_against_compiled_reference():
    """fhmc_patch_shifts against patch_window_pair of the compiled reference: |shift - reference| <= the reference's own
    xtol (1e-4); the parabola gives err2_ref - err2 = (shift_ref - shift)^2 exactly, checked to 1e-10; never worse than the
    reference's minimum."""
    from fhmcanalysis_b200.moments.win_patch import fhmc_patch as fp
    n = 0
    for name, w1, w2, z in _golden_patch():
        if name + "/raises" in z.files:
            with pytest.raises(AssertionError):
                fp.patch_window_pair(w1, w2)
            continue
        shift_ref, err_ref = z[name + "/ref"]
        shift, err2 = fp.patch_window_pair(w1, w2)
        assert abs(shift - shift_ref) <= 1e-4, name
        assert err2 <= err_ref + 1e-12 * max(1.0, err_ref)
        assert abs((err_ref - err2) - (shift_ref - shift) ** 2) <= 1e-10 * max(1.0, err_ref), name
        n += 1
    assert n >= 5
