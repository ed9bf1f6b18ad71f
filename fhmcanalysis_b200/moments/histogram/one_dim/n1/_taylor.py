"""Derivative builders of the N_1-order-parameter histogram (reference moments/histogram/one_dim/n1/gc_hist.pyx, "N1").

N_1 is sharp at fixed order parameter, the chemical potentials of species 2..N are ABSOLUTE (not differences to mu_1)
and there is no kinetic-energy option, so compared with the N_tot builders (ntot/_taylor.py) only the weights change:
species 1 drops out of every pointwise ("semi-grand") derivative and species s >= 2 enters with curr_mu[s] instead of
curr_mu[s] - curr_mu[0].  Private names and plain-index signatures follow N1 (N1:691-1433)."""
import numpy as np

from ..ntot._taylor import TaylorMixin


def _plain(x):
    """N1's privates take plain [i,j,k,m,p]; the shared builders pass (index, N_tot power) tuples."""
    if len(x) == 2 and hasattr(x[0], "__len__"):
        assert int(x[1]) == 0, "N_1 histograms carry no order-parameter powers"
        return [int(v) for v in x[0]]
    return [int(v) for v in x]


class N1TaylorMixin(TaylorMixin):
    _first_species = 1          # species index 0 is the order parameter itself

    def _npow(self, n):
        assert not n, "N_1 histograms carry no order-parameter powers"
        return 1.0

    def _d(self, i):            # weight of N_i in d/d(beta): absolute chemical potential, N1:824-838
        return self.data["curr_mu"][i] if i >= 1 else 0.0

    # ---- grand-canonical scalars ------------------------------------------------------------------
    def _gc_dX_dB(self, x_idx, n=0):
        """N1:1336-1366."""
        assert len(x_idx) == 5, "Bad indices"
        X = self._m(x_idx)
        der = self.data["curr_mu"][0] * self._gc_fluct_vi(X, [0, 1, 0, 0, 0])
        der -= self._gc_fluct_vi(X, [0, 0, 0, 0, 1])
        for i in range(1, self.data["nspec"]):
            der += self.data["curr_mu"][i] * self._gc_fluct_vi(X, [i, 1, 0, 0, 0])
        return der

    # ---- pointwise derivatives: same rules, plain-index signatures ----------------------------------
    def _sg_dX_dB(self, x_idx, n=0):
        return TaylorMixin._sg_dX_dB(self, _plain(x_idx) if n == 0 else x_idx, 0)

    def _sg_df_dB(self, x_idx, y_idx):
        return TaylorMixin._sg_df_dB(self, (_plain(x_idx), 0), (_plain(y_idx), 0))

    def _sg_d2X_dB2(self, x_idx, n=0):
        return TaylorMixin._sg_d2X_dB2(self, _plain(x_idx), 0)

    # ---- assemblers -------------------------------------------------------------------------------
    def _dB(self, skip_mom=False):
        """N1:739-788."""
        ns = self.data["nspec"]
        ave_u = self._gc_ave_i([0, 0, 0, 0, 1])
        d = self._zeros()
        for i in range(ns):
            d = d + self.data["curr_mu"][i] * (self._m([i, 1, 0, 0, 0]) - self._gc_ave_i([i, 1, 0, 0, 0]))
        d = d - (self._m([0, 0, 0, 0, 1]) - ave_u)
        dm = self._mom_like()
        if not skip_mom:
            for a in self._each_address(1):
                try:
                    dm[a] = self._sg_dX_dB(list(a))
                except Exception as e:
                    raise Exception("Cannot compute first derivative: " + str(e))
        return d, dm

    def _dB2(self, skip_mom=False):
        """N1:1295-1334."""
        ns = self.data["nspec"]
        d2 = self._zeros()
        for i in range(1, ns):
            d2 = d2 + self.data["curr_mu"][i] * (self._sg_dX_dB([i, 1, 0, 0, 0]) - self._gc_dX_dB([i, 1, 0, 0, 0]))
        d2 = d2 + self.data["curr_mu"][0] * (-self._gc_dX_dB([0, 1, 0, 0, 0]))
        d2 = d2 - (self._sg_dX_dB([0, 0, 0, 0, 1]) - self._gc_dX_dB([0, 0, 0, 0, 1]))
        d2m = self._mom_like()
        if not skip_mom:
            for a in self._each_address(2):
                try:
                    d2m[a] = self._sg_d2X_dB2(list(a))
                except Exception as e:
                    raise Exception("Cannot compute second derivative: " + str(e))
        return d2, d2m

    def _dB3(self, skip_mom=False):
        raise Exception("No implementation for third order extrapolation of N_1 histograms")

    def taylor_rows(self, order, ke_ok=True):
        if order > 2:
            raise Exception("No implementation for temperature + mu extrapolation of order " + str(order))
        return TaylorMixin.taylor_rows(self, order, ke_ok)
