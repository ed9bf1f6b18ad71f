"""Window patching: the ln(PI) shift solve between overlapping windows (reference: moments/win_patch/fhmc_patch.pyx
``patch_window_pair`` :668-709 with the objective ``window_patch_error`` :640-664), SURVEY.md 8(f) row 4.

Only the shift solve lives here (batched over any number of window pairs, one kernel launch: ``fhmc_patch_shifts``).
The rest of win_patch (parsing the simulation's text output, equilibration tests, composite assembly) is out of scope.

The reference minimises the sum of squared differences over the overlap with ``scipy.optimize.fmin`` (Nelder-Mead,
ftol = 1e-6, default xtol = 1e-4); the objective is a parabola, so its exact minimiser is mean(other - this) and that is what
the kernel returns: the reference's answer agrees with it to its own x tolerance.
"""
import ctypes

import numpy as np

from ... import _lib, engine


def overlap_slices(window_hist1, window_hist2):
    """The two aligned overlap slices patch_window_pair fits (fhmc_patch.pyx:686-697): window_hist1 is the upper window;
    ``offset`` bins are trimmed from both ends of the overlap because of edge effects."""
    assert window_hist1.lb > window_hist2.lb, 'Histograms out of order, cannot patch'
    assert window_hist1.ub > window_hist2.ub, 'Histograms out of order, cannot patch'
    assert window_hist1.lb < window_hist2.ub, 'Histograms do not overlap, cannot patch'
    index = int(window_hist2.ub - window_hist1.lb + 1)
    off = int(window_hist1.offset)
    l2 = len(window_hist2.lnPI)
    data_slice1 = np.asarray(window_hist1.lnPI[off:index - off], dtype=np.float64)
    data_slice2 = np.asarray(window_hist2.lnPI[l2 - index + off:l2 - off], dtype=np.float64)
    assert len(data_slice1) > 1, 'Error, unable to patch windown because there is no overlap'
    assert len(data_slice2) > 1, 'Error, unable to patch windows because there is no overlap'
    return data_slice1, data_slice2


def patch_shifts(pairs, device=None):
    """Shift and mean squared patching error for every (this_lnPI_slice, other_lnPI_slice) pair at once.
    Returns (shift[W], err2[W]): other ~= this + shift, err2 = sum((this + shift - other)^2) / len."""
    L = _lib.load()
    t = engine.torch()
    dev = engine.require_cuda(device)
    W = len(pairs)
    if W == 0:
        return np.zeros(0), np.zeros(0)
    lens = [len(p[0]) for p in pairs]
    for (x, y), n in zip(pairs, lens):
        if len(y) != n:
            raise ValueError("the two slices of a pair must have the same length")
    offsets = np.zeros(W + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(lens)
    a = np.concatenate([np.asarray(p[0], dtype=np.float64) for p in pairs]) if W else np.zeros(0)
    b = np.concatenate([np.asarray(p[1], dtype=np.float64) for p in pairs]) if W else np.zeros(0)
    a_d, b_d, o_d = (t.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (a, b, offsets))
    out = t.empty((2, max(W, 1)), dtype=t.float64, device=dev)
    with t.cuda.device(dev):
        rc = L.fhmc_patch_shifts(ctypes.c_void_p(a_d.data_ptr()), ctypes.c_void_p(b_d.data_ptr()), ctypes.c_void_p(o_d.data_ptr()), W,
                                 ctypes.c_void_p(out[0].data_ptr()), ctypes.c_void_p(out[1].data_ptr()),
                                 ctypes.c_void_p(t.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "fhmc_patch_shifts")
    h = out.cpu().numpy()
    return h[0, :W].copy(), h[1, :W].copy()


def patch_window_pair(window_hist1, window_hist2, ftol=0.000001):
    """Shift necessary for window_hist1 to match window_hist2, error^2 / number of overlapping points
    (fhmc_patch.pyx:668-709).  ``ftol`` is accepted for signature compatibility: the closed-form minimiser needs none."""
    s1, s2 = overlap_slices(window_hist1, window_hist2)
    shift, err2 = patch_shifts([(s1, s2)])
    return float(shift[0]), float(err2[0])
