"""``collect=`` hooks for thermo()/find_phase_eq() (reference moments/histogram/one_dim/ntot/collect.py): host callbacks
that merge several ln(PI) peaks into "macrophases" by rewriting data['ln(PI)_maxima_idx'/'ln(PI)_minima_idx'] after the
device phase split; the per-phase integrals are then evaluated on the device for the rewritten bounds."""
import numpy as np


def check_order_(hist):
    """Maxima and minima must still alternate (collect.py:10-30)."""
    M, m = hist.data["ln(PI)_maxima_idx"], hist.data["ln(PI)_minima_idx"]
    order = np.zeros(len(M) + len(m))
    if M[0] < m[0]:
        order[::2], order[1::2] = M, m
    else:
        order[::2], order[1::2] = m, M
    if not np.all([order[i] <= order[i + 1] for i in range(len(order) - 1)]):
        raise Exception("Local maxima and minima not sorted correctly after collection")


def janus_collect(hist, **kwargs):
    """Last maximum = one (isotropic liquid) phase, all others = one micellar gas (collect.py:32-80)."""
    if "ln(PI)_maxima_idx" not in hist.data or "ln(PI)_minima_idx" not in hist.data:
        raise Exception("Histogram has not been segmented yet")
    check_order_(hist)
    M, m = hist.data["ln(PI)_maxima_idx"], hist.data["ln(PI)_minima_idx"]
    if len(M) <= 2:
        return
    max_idx = [int(round(np.mean(M[:-1]))), int(M[-1])]
    min_idx = [] if m[0] > 0 else [0]
    last = int(m[-1])
    if max_idx[0] < last < max_idx[1]:
        min_idx.append(last)
    elif last > max_idx[1]:
        assert len(m) > 1
        min_idx.append(int(m[-2]))
        min_idx.append(int(m[-1]))
    check_order_(hist)
    hist.data["ln(PI)_maxima_idx"] = np.array(max_idx, dtype=np.int64)
    hist.data["ln(PI)_minima_idx"] = np.array(min_idx, dtype=np.int64)
