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
