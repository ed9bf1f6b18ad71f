"""pore_hist -- joint histogram lnPI(h, N_tot) of a slit pore (reference: moments/histogram/two_dim/h_ntot/pore_hist.pyx,
``PH`` below), SURVEY.md 8(f) row 3: the surface construction, ``normalize`` and ``thermo(mask)`` run on the B200
(``fhmc_masked_lse_2d``, csrc/fhmc_masked2d.cu).  The watershed segmentation workflow on top of them
(``phase_average`` and its helpers, PH:186-477: scikit-image) is out of scope and raises.

Notes on the reference as it stands (probed with the compiled module):
* ``__init__`` reads ``self.data['ln(PI)']`` before assigning it (PH:129 vs PH:132) and therefore always raises
  KeyError; here the mask is taken from the joint histogram's surface, which is what the next line copies.
* ``thermo`` indexes with ``lp[not mask]`` (PH:170, 172), which raises ValueError for any mask with more than one
  element; the evident intent (everything outside the mask gets -inf) is what is implemented, and the parity of
  ``thermo`` is pinned on a NumPy restatement only (tests/test_pore_hist.py).  ``normalize`` is pinned on
  the compiled reference.
"""
import copy

import numpy as np

from ..... import engine


class pore_hist(object):
    """A joint histogram in terms of (h, Ntot) built from a general joint histogram object (PH:82-89)."""

    def __init__(self, joint_hist, fh, p_tot, A, beta):
        """joint_hist: joint histogram in (h, Ntot); fh: F(h) of the empty adsorbent; p_tot: total pressure;
        A: cross-sectional area; beta: 1/kT (PH:91-108)."""
        self.clear()
        self.data['F(h)'] = fh
        self.data['p'] = float(p_tot)
        self.data['hist'] = copy.deepcopy(joint_hist)
        self.data['A'] = float(A)
        self.data['beta'] = float(beta)
        try:
            self.data['hist'].make()
        except Exception as e:
            raise Exception('Could not construct joint histogram: ' + str(e))
        hd = self.data['hist'].data
        # 0 <= N <= Nmax, continuous (PH:122-123)
        assert np.all(hd['op_2'] == np.arange(len(hd['op_2']))), 'Must be 0 <= N <= N_max in a continuous fashion'
        # lower bound all at 0 across the board, the upper bound defines the ridgeline or 'edges' (PH:125-127)
        b = np.asarray(hd['bounds_idx'])
        assert np.all(b[:, 0] == 0), 'Lower bound for N must start from 0'
        self.data['edge_idx'] = np.array(b[:, 1], dtype=int)
        self.data['mask'] = np.asarray(hd['ln(PI)']) > -np.inf
        # lnPI surface: every row shifted to -beta (F(h) + p A h) at N = 0, then normalised (PH:131-137)
        lp = np.array(hd['ln(PI)'], dtype=np.float64, copy=True)
        for i, h in enumerate(hd['op_1']):
            shift = -self.data['beta'] * (self.data['F(h)'](h) + self.data['p'] * self.data['A'] * h) - lp[i, 0]
            lp[i, :] += shift
        self.data['ln(PI)'] = lp
        self.normalize()

    def clear(self):
        """Clear all data in histogram (PH:139-145)."""
        self.data = {}

    def normalize(self):
        """Normalize the ln(PI) surface over j <= edge_idx[i] of every row (PH:147-152 -> _cy_normalize PH:57-80)."""
        r = engine.masked_lse_2d(self.data['ln(PI)'], edge=self.data['edge_idx'], shifted=True)
        self.data['ln(PI)'] = r["shifted"]

    def thermo(self, mask):
        """Average extensive properties over the region of (h, N) space ``mask`` selects (PH:154-184).
        Returns {property_name: average, 'peak_idx': np.where(lp == max(lp))}."""
        mask = np.asarray(mask, dtype=bool)
        names = list(self.data['hist'].data['props'])
        props = np.stack([np.asarray(self.data['hist'].data['props'][p], dtype=np.float64) for p in names]) if names else None
        r = engine.masked_lse_2d(self.data['ln(PI)'], mask=mask, props=props)
        ave_props = {p: r["avg"][k] for k, p in enumerate(names)}
        ave_props['peak_idx'] = r["peak_idx"]
        return ave_props

    def phase_average(self, nnebr=1, max_peaks=10):
        raise NotImplementedError("pore_hist.phase_average (watershed segmentation with scikit-image, PH:186-477) is out "
                                  "of scope of the B200 hot path; build the masks yourself and call thermo(mask)")


if __name__ == "__main__":
    print("pore_hist (B200)")
