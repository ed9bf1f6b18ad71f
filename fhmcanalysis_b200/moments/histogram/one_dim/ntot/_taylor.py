"""Taylor-expansion coefficient builders for the 1-D N_tot histogram (host-side, one-time setup per
histogram; the per-state-point work that uses them runs on the GPU).

What is restated here (reference: moments/histogram/one_dim/ntot/gc_hist.pyx, "GH"):
  * the moment-address algebra  ``_order_mom_address`` / ``_mom_prod``            GH:1515-1658
  * pointwise ("semi-grand", fixed N_tot) derivatives of the moment arrays
      ``_sg_dX_dB`` GH:1660, ``_sg_dX_dMU`` GH:1724, ``_sg_d2X_dB2`` GH:1776, ``_sg_d2X_dMU2`` GH:1829,
      ``_sg_d3X_dB3`` GH:1870, ``_sg_df_dB`` GH:1914, ``_sg_df_dMU`` GH:1943, ``_sg_d2f_dB2`` GH:1968
  * whole-histogram (grand-canonical) averages/fluctuations and their beta derivatives
      ``_gc_fluct_*`` GH:1241-1336, ``_gc_ave_*`` GH:1338-1380, ``_gc_dX_dB`` GH:1382, ``_gc_d2X_dB2`` GH:1420,
      ``_gc_df_dB_ii`` GH:1461, ``_gc_df_dB_in`` GH:1488
  * the gradient / Hessian assemblers ``_dB`` GH:2114, ``_dB2`` GH:2167, ``_dB3`` GH:2208, ``_dMU`` GH:2342,
      ``_dMU2`` GH:2389, ``_dBMU`` GH:2436, ``_dBMU2`` GH:2484.
The unit tests of the reference call several of these privates directly (T1:505-509, 659-878), so they keep
their names, argument meaning and error behaviour.  The implementation is a single generic
"which higher moment is X*N_s" rule + product rule instead of the reference's per-function copies.
"""
import numpy as np


def order_mom_address(idx):
    """GH:1515-1544: put the lower species index first (energy power untouched)."""
    idx = [int(v) for v in idx]
    if idx[0] > idx[2]:
        return [idx[2], idx[3], idx[0], idx[1], idx[4]]
    return list(idx)


def _canon(idx):
    o = [int(v) for v in idx]
    if o[0] == o[2]:  # N_a^j N_a^m -> N_a^(j+m) N_0^0
        o[1] += o[3]
        o[3] = 0
        o[2] = 0
    return order_mom_address(o)


def mom_prod(x_idx, y_idx, max_order, nspec):
    """Address of the moment <X*Y> (GH:1546-1658; valid for at most two species).

    Only the species pairings the reference resolves are accepted; the others raise 'Bad logic'
    exactly as its (empty-condition) branches do."""
    assert nspec <= 2, "Ordering moment indices is only valid for 2 or less components"
    x, y = _canon(x_idx), _canon(y_idx)
    if x[0] == y[0] and x[2] == y[2]:
        z = [x[0], x[1] + y[1], x[2], x[3] + y[3], x[4] + y[4]]
    elif x[0] == 0 and x[2] == 0 and y[0] == 0 and y[2] == 1:
        z = [y[0], y[1] + x[1] + x[3], y[2], y[3], y[4] + x[4]]
    elif x[0] == 0 and x[2] == 1 and y[0] == 0 and y[2] == 0:
        z = [x[0], x[1] + y[1] + y[3], x[2], x[3], x[4] + y[4]]
    else:
        raise Exception("Bad logic")
    if z[0] == z[2]:  # use symmetry to avoid overflowing max_order
        if z[1] > max_order:
            z[3] = z[1] - max_order
            z[1] = max_order
        elif z[3] > max_order:
            z[1] = z[3] - max_order
            z[3] = max_order
    assert z[1] <= max_order, "Order out of range"
    assert z[3] <= max_order, "Order out of range"
    assert z[4] <= max_order, "Order out of range"
    return np.array(z, dtype=np.int64)


class TaylorMixin(object):
    """Private derivative builders of ``histogram`` (operate on self.data / self.metadata)."""

    _first_species = 0   # first species whose N_s enters the pointwise derivatives (1 for the N_1 class)

    def _ke(self):
        return bool(self.metadata.get("used_ke", False))

    # ---- helpers -------------------------------------------------------------------------------
    def _m(self, x):
        return self.data["mom"][int(x[0]), int(x[1]), int(x[2]), int(x[3]), int(x[4])]

    def _npow(self, n):
        return self.data["ntot"].astype(np.float64) ** n if n else 1.0

    def _d(self, i):
        return self.data["curr_mu"][i] - self.data["curr_mu"][0]

    def _order_mom_address(self, idx):
        return np.array(order_mom_address(idx), dtype=np.int64)

    def _mom_prod(self, x_idx, y_idx):
        return mom_prod(x_idx, y_idx, self.data["max_order"], self.data["nspec"])

    def _guard(self, x):
        """Common preamble of every _sg_* routine.  Returns True when the derivative is identically 0."""
        assert len(x) == 5, "Bad indices"
        if x[1] == 0 and x[3] == 0 and x[4] == 0:
            return True
        mo = self.data["max_order"]
        if x[4] >= mo or x[3] >= mo or x[1] >= mo:
            raise Exception("max_order too low to take this derivative")
        return False

    def _times_species(self, x, s):
        """Moment array of X*N_s, chosen by the first matching rule of GH:1689-1700 / 1752-1763."""
        mo = self.data["max_order"]
        mom = self.data["mom"]
        i, j, k, m, p = [int(v) for v in x]
        if i == s and j + 1 <= mo:
            return mom[i, j + 1, k, m, p]
        if k == s and m + 1 <= mo:
            return mom[i, j, k, m + 1, p]
        if j == 0:
            return mom[s, 1, k, m, p]
        if m == 0:
            return mom[i, j, s, 1, p]
        if i == k and j + m <= mo:
            return mom[i, j + m, s, 1, p]
        raise Exception("max_order too low to take this derivative")

    def _zeros(self):
        return np.zeros(int(self.data["ub"] - self.data["lb"] + 1), dtype=np.float64)

    def _prob(self):
        prob = np.exp(self.data["ln(PI)"])
        return prob, np.sum(prob)

    # ---- grand-canonical scalars (whole histogram) ---------------------------------------------
    def _gc_ave_v(self, a):
        assert len(a) == len(self.data["ln(PI)"]), "Bad quantity array"
        prob, sp = self._prob()
        return np.sum(a * prob) / sp

    def _gc_ave_i(self, x_idx):
        assert len(x_idx) == 5, "Bad indices"
        return self._gc_ave_v(self._m(x_idx))

    def _gc_fluct_vv(self, a, b):
        assert len(a) == len(self.data["ln(PI)"]), "Bad quantity array"
        assert len(b) == len(self.data["ln(PI)"]), "Bad quantity array"
        prob, sp = self._prob()
        return np.sum(a * b * prob) / sp - np.sum(a * prob) / sp * np.sum(b * prob) / sp

    def _gc_fluct_vi(self, a, y_idx):
        assert len(y_idx) == 5, "Bad indices"
        return self._gc_fluct_vv(a, self._m(y_idx))

    def _gc_fluct_iv(self, y_idx, a):
        return self._gc_fluct_vi(a, y_idx)

    def _gc_fluct_ii(self, x_idx, y_idx):
        assert len(x_idx) == 5 and len(y_idx) == 5, "Bad indices"
        prob, sp = self._prob()
        z = self._mom_prod(x_idx, y_idx)
        return np.sum(self._m(z) * prob) / sp - np.sum(self._m(x_idx) * prob) / sp * np.sum(self._m(y_idx) * prob) / sp

    def _gc_dX_dB(self, x_idx, n=0):
        """d<X N_tot^n>/d beta over the whole histogram, GH:1382-1418."""
        assert len(x_idx) == 5, "Bad indices"
        ntot = self.data["ntot"].astype(np.float64)
        X = self._m(x_idx) * self._npow(n)
        der = self.data["curr_mu"][0] * self._gc_fluct_vv(X, ntot)
        der -= self._gc_fluct_vi(X, [0, 0, 0, 0, 1])
        for i in range(self.data["nspec"]):
            der += self._d(i) * self._gc_fluct_vi(X, [i, 1, 0, 0, 0])
        if self._ke() and x_idx[4] > 0:
            y = list(x_idx)
            y[4] -= 1
            der -= 1.5 * x_idx[4] / self.data["curr_beta"] ** 2 * self._gc_ave_v(self._m(y) * self._npow(n + 1))
        return der

    def _gc_df_dB_ii(self, x_idx_t, y_idx_t):
        (x, nx), (y, ny) = x_idx_t, y_idx_t
        z = self._mom_prod(x, y)
        X = self._m(x) * self._npow(nx)
        Y = self._m(y) * self._npow(ny)
        return self._gc_dX_dB(z, nx + ny) - self._gc_ave_v(X) * self._gc_dX_dB(y, ny) - self._gc_ave_v(Y) * self._gc_dX_dB(x, nx)

    def _gc_df_dB_in(self, x_idx_t, n=0):
        x, nx = x_idx_t
        X = self._m(x) * self._npow(nx)
        Y = self._m([0, 0, 0, 0, 0]) * self._npow(n)
        return self._gc_dX_dB(x, n + nx) - self._gc_ave_v(X) * self._gc_dX_dB([0, 0, 0, 0, 0], n) - self._gc_ave_v(Y) * self._gc_dX_dB(x, nx)

    def _gc_d2X_dB2(self, x_idx, n=0):
        """GH:1420-1459."""
        assert len(x_idx) == 5, "Bad indices"
        der = self.data["curr_mu"][0] * self._gc_df_dB_in((x_idx, n), 1) - self._gc_df_dB_ii((x_idx, n), ([0, 0, 0, 0, 1], 0))
        for i in range(self.data["nspec"]):
            der += self._d(i) * self._gc_df_dB_ii((x_idx, n), ([i, 1, 0, 0, 0], 0))
        if self._ke() and x_idx[4] > 0:
            y = list(x_idx)
            y[4] -= 1
            beta = self.data["curr_beta"]
            ave_run = self._gc_ave_v(self._m(y) * self._npow(n + 1))
            der -= 1.5 * x_idx[4] / beta ** 2 * (-2.0 / beta * ave_run + self._gc_dX_dB(y, n + 1))
        return der

    # ---- semi-grand (pointwise in N_tot) derivatives -------------------------------------------
    def _sg_dX_dB(self, x_idx, n=0):
        """d(X N^n)/d beta at fixed N_tot, GH:1660-1722."""
        if self._guard(x_idx):
            return self._zeros()
        x = [int(v) for v in x_idx]
        npw = self._npow(n)
        X = self._m(x) * npw
        up = list(x)
        up[4] += 1
        der = -(self._m(up) * npw - X * self._m([0, 0, 0, 0, 1]))
        for s in range(self._first_species, self.data["nspec"]):
            f = self._times_species(x, s) * npw - X * self._m([s, 1, 0, 0, 0])
            der = der + self._d(s) * f
        if self._ke() and x[4] > 0:
            dn = list(x)
            dn[4] -= 1
            der = der - 1.5 * x[4] / self.data["curr_beta"] ** 2 * self.data["ntot"] * (self._m(dn) * npw)
        return der

    def _sg_dX_dMU(self, q, x_idx):
        """dX/d(dmu_{q+2}) at fixed N_tot, GH:1724-1774."""
        assert q >= 0 and q < self.data["nspec"] - 1, "Bad dMu index"
        if self._guard(x_idx):
            return self._zeros()
        s = q + 1
        return self.data["curr_beta"] * (self._times_species(x_idx, s) - self._m(x_idx) * self._m([s, 1, 0, 0, 0]))

    def _sg_df_dB(self, x_idx_t, y_idx_t):
        """Product rule on f = <XY> - <X><Y>, GH:1914-1941."""
        (x, nx), (y, ny) = x_idx_t, y_idx_t
        z = self._mom_prod(x, y)
        return (self._sg_dX_dB(z, nx + ny) - self._m(x) * self._npow(nx) * self._sg_dX_dB(y, ny)
                - self._m(y) * self._npow(ny) * self._sg_dX_dB(x, nx))

    def _sg_df_dMU(self, j, x_idx, y_idx):
        assert len(x_idx) == 5 and len(y_idx) == 5, "Bad indices"
        assert j >= 0 and j < self.data["nspec"] - 1, "Bad species index"
        z = self._mom_prod(x_idx, y_idx)
        return self._sg_dX_dMU(j, z) - self._m(x_idx) * self._sg_dX_dMU(j, y_idx) - self._m(y_idx) * self._sg_dX_dMU(j, x_idx)

    def _sg_d2X_dB2(self, x_idx, n=0):
        """GH:1776-1827."""
        if self._guard(x_idx):
            return self._zeros()
        x = [int(v) for v in x_idx]
        der = -self._sg_df_dB((x, n), ([0, 0, 0, 0, 1], 0))
        for s in range(self._first_species, self.data["nspec"]):
            der = der + self._d(s) * self._sg_df_dB((x, n), ([s, 1, 0, 0, 0], 0))
        if self._ke() and x[4] > 0:
            y = list(x)
            y[4] -= 1
            beta = self.data["curr_beta"]
            a = -2.0 / beta * (self._m(y) * self._npow(n))
            der = der - 1.5 * x[4] * self.data["ntot"] / beta ** 2 * (a + self._sg_dX_dB(y, n))
        return der

    def _sg_d2X_dMU2(self, q, r, x_idx):
        assert q >= 0 and q < self.data["nspec"] - 1, "Bad dMu index"
        assert r >= 0 and r < self.data["nspec"] - 1, "Bad dMu index"
        if self._guard(x_idx):
            return self._zeros()
        return self.data["curr_beta"] * self._sg_df_dMU(q, x_idx, [r + 1, 1, 0, 0, 0])

    def _sg_d2f_dB2(self, x_idx_t, y_idx_t):
        (x, nx), (y, ny) = x_idx_t, y_idx_t
        z = self._mom_prod(x, y)
        dx, dy = self._sg_dX_dB(x, nx), self._sg_dX_dB(y, ny)
        return (self._sg_d2X_dB2(z, nx + ny) - self._m(x) * self._npow(nx) * self._sg_d2X_dB2(y, ny)
                - self._m(y) * self._npow(ny) * self._sg_d2X_dB2(x, nx) - 2.0 * dx * dy)

    def _sg_d3X_dB3(self, x_idx, n=0):
        if self._guard(x_idx):
            return self._zeros()
        x = [int(v) for v in x_idx]
        der = -self._sg_d2f_dB2((x, n), ([0, 0, 0, 0, 1], 0))
        for s in range(self._first_species, self.data["nspec"]):
            der = der + self._d(s) * self._sg_d2f_dB2((x, n), ([s, 1, 0, 0, 0], 0))
        if self._ke():
            raise Exception("No KE correction implemented for _sg_d3X_dB3()")
        return der

    # ---- assemblers ------------------------------------------------------------------------------
    def _mom_like(self, lead=()):
        ns, mo = self.data["nspec"], self.data["max_order"]
        return np.zeros(tuple(lead) + (ns, mo + 1, ns, mo + 1, mo + 1, len(self.data["ln(PI)"])), dtype=np.float64)

    def _each_address(self, extra):
        ns, mo = self.data["nspec"], self.data["max_order"]
        for i in range(ns):
            for j in range(mo + 1):
                for k in range(ns):
                    for m in range(mo + 1):
                        for p in range(mo + 1):
                            if j + m + p + extra <= mo:
                                yield i, j, k, m, p

    def _dB(self, skip_mom=False):
        """First beta derivative of lnPI and the moments, GH:2114-2165."""
        ns = self.data["nspec"]
        ave_u = self._gc_ave_i([0, 0, 0, 0, 1])
        ave_n = [self._gc_ave_i([i, 1, 0, 0, 0]) for i in range(ns)]
        d = self._zeros()
        for i in range(ns):
            d = d + self._d(i) * (self._m([i, 1, 0, 0, 0]) - ave_n[i])
        d = d + self.data["curr_mu"][0] * (self.data["ntot"] - sum(ave_n))
        d = d - (self._m([0, 0, 0, 0, 1]) - ave_u)
        dm = self._mom_like()
        if not skip_mom:
            for a in self._each_address(1):
                try:
                    dm[a] = self._sg_dX_dB(list(a), 0)
                except Exception as e:
                    raise Exception("Cannot compute first derivative: " + str(e))
        return d, dm

    def _dB2(self, skip_mom=False):
        """GH:2167-2206."""
        ns = self.data["nspec"]
        d2 = self._zeros()
        for i in range(ns):
            d2 = d2 + self._d(i) * (self._sg_dX_dB([i, 1, 0, 0, 0], 0) - self._gc_dX_dB([i, 1, 0, 0, 0], 0))
        d2 = d2 + self.data["curr_mu"][0] * (-self._gc_dX_dB([0, 0, 0, 0, 0], 1))
        d2 = d2 - (self._sg_dX_dB([0, 0, 0, 0, 1], 0) - self._gc_dX_dB([0, 0, 0, 0, 1], 0))
        d2m = self._mom_like()
        if not skip_mom:
            for a in self._each_address(2):
                try:
                    d2m[a] = self._sg_d2X_dB2(list(a), 0)
                except Exception as e:
                    raise Exception("Cannot compute second derivative: " + str(e))
        return d2, d2m

    def _dB3(self, skip_mom=False):
        """GH:2208-2252."""
        if self._ke():
            raise Exception("KE corrections not implemented for 3rd order beta extrapolation")
        ns = self.data["nspec"]
        d3 = self._zeros()
        for i in range(ns):
            d3 = d3 + self._d(i) * (self._sg_d2X_dB2([i, 1, 0, 0, 0], 0) - self._gc_d2X_dB2([i, 1, 0, 0, 0], 0))
        d3 = d3 + self.data["curr_mu"][0] * (-self._gc_d2X_dB2([0, 0, 0, 0, 0], 1))
        d3 = d3 - (self._sg_d2X_dB2([0, 0, 0, 0, 1], 0) - self._gc_d2X_dB2([0, 0, 0, 0, 1], 0))
        d3m = self._mom_like()
        if not skip_mom:
            for a in self._each_address(3):
                try:
                    d3m[a] = self._sg_d3X_dB3(list(a), 0)
                except Exception as e:
                    raise Exception("Cannot compute third derivative: " + str(e))
        return d3, d3m

    def _dMU(self, skip_mom=False):
        """GH:2342-2387."""
        ns = self.data["nspec"]
        n = len(self.data["ln(PI)"])
        d = np.zeros((ns - 1, n), dtype=np.float64)
        for i in range(ns - 1):
            d[i] = self.data["curr_beta"] * (self._m([i + 1, 1, 0, 0, 0]) - self._gc_ave_i([i + 1, 1, 0, 0, 0]))
        dm = self._mom_like((ns - 1,))
        if not skip_mom:
            for q in range(ns - 1):
                for a in self._each_address(1):
                    try:
                        dm[(q,) + a] = self._sg_dX_dMU(q, list(a))
                    except Exception as e:
                        raise Exception("Cannot compute first derivative: " + str(e))
        return d, dm

    def _h_dmu_block(self, i, j):
        f = self._m([i + 1, 1, j + 1, 1, 0]) - self._m([i + 1, 1, j + 1, 0, 0]) * self._m([i + 1, 0, j + 1, 1, 0])
        return self.data["curr_beta"] ** 2 * (f - self._gc_fluct_ii([i + 1, 1, 0, 0, 0], [j + 1, 1, 0, 0, 0]))

    def _dMU2(self, skip_mom=False):
        """GH:2389-2434."""
        ns = self.data["nspec"]
        n = len(self.data["ln(PI)"])
        H = np.zeros((ns - 1, ns - 1, n), dtype=np.float64)
        for i in range(ns - 1):
            for j in range(ns - 1):
                H[i, j] = self._h_dmu_block(i, j)
        Hm = self._mom_like((ns - 1, ns - 1))
        if not skip_mom:
            for q in range(ns - 1):
                for r in range(ns - 1):
                    for a in self._each_address(2):
                        try:
                            Hm[(q, r) + a] = self._sg_d2X_dMU2(q, r, list(a))
                        except Exception as e:
                            raise Exception("Cannot compute second derivative: " + str(e))
        return H, Hm

    def _dBMU(self, skip_mom=False):
        """Gradient in xi = [beta, dmu_2, ...], GH:2436-2482."""
        ns = self.data["nspec"]
        n = len(self.data["ln(PI)"])
        d = np.zeros((ns, n), dtype=np.float64)
        dm = self._mom_like((ns,))
        d[0], dm[0] = self._dB(skip_mom)
        for i in range(1, ns):
            d[i] = self.data["curr_beta"] * (self._m([i, 1, 0, 0, 0]) - self._gc_ave_i([i, 1, 0, 0, 0]))
        if not skip_mom:
            for q in range(1, ns):
                for a in self._each_address(1):
                    try:
                        dm[(q,) + a] = self._sg_dX_dMU(q - 1, list(a))
                    except Exception as e:
                        raise Exception("Cannot compute first derivative: " + str(e))
        return d, dm

    def _dBMU2(self, skip_mom=False):
        """Hessian in xi = [beta, dmu_2, ...], GH:2484-2563."""
        ns = self.data["nspec"]
        n = len(self.data["ln(PI)"])
        H = np.zeros((ns, ns, n), dtype=np.float64)
        Hm = self._mom_like((ns, ns))
        for i in range(ns - 1):
            for j in range(ns - 1):
                H[i + 1, j + 1] = self._h_dmu_block(i, j)
        if not skip_mom:
            for q in range(ns - 1):
                for r in range(ns - 1):
                    for a in self._each_address(2):
                        try:
                            Hm[(q + 1, r + 1) + a] = self._sg_d2X_dMU2(q, r, list(a))
                        except Exception as e:
                            raise Exception("Cannot compute second derivative: " + str(e))
        H[0, 0], Hm[0, 0] = self._dB2(skip_mom)
        beta = self.data["curr_beta"]
        for q in range(1, ns):
            tmp = self._m([q, 1, 0, 0, 0]) - self._gc_ave_i([q, 1, 0, 0, 0])
            tmp = tmp + beta * (self._sg_dX_dB([q, 1, 0, 0, 0], 0) - self._gc_dX_dB([q, 1, 0, 0, 0], 0))
            H[q, 0] = tmp
            H[0, q] = tmp
        if not skip_mom:
            for q in range(1, ns):
                for a in self._each_address(2):
                    try:
                        z = self._mom_prod([q, 1, 0, 0, 0], list(a))
                        f = self._m(z) - self._m([q, 1, 0, 0, 0]) * self._m(a)
                        x = beta * self._sg_df_dB(([q, 1, 0, 0, 0], 0), (list(a), 0)) + f
                    except Exception as e:
                        raise Exception("Cannot compute second derivative: " + str(e))
                    Hm[(q, 0) + a] = x
                    Hm[(0, q) + a] = x
        return H, Hm

    # ---- closed-form coefficient rows for the batched GPU path -----------------------------------
    def taylor_rows(self, order, ke_ok=True):
        """Coefficient rows (kind, array) of  lnPI'(N) - lnPI(N) - beta_ref*(mu1-mu1_ref)*N  for the fused
        sweep kernel: the N-dependent parts of _dBMU/_dBMU2 (the N-independent _gc_ constants are removed by
        the renormalisation, GH:885/964).  At most two species (as _mom_prod)."""
        from fhmcanalysis_b200 import _lib as L
        ns = self.data["nspec"]
        if ns > 2:
            raise NotImplementedError("batched Taylor extrapolation supports at most two species")
        if order < 1 or order > 3 or (order == 3 and ns > 1):
            raise Exception("No implementation for this order of extrapolation")
        if self.data["max_order"] < order:
            raise Exception("Maximum order stored in simulation not high enough to calculate this order of extrapolation")
        U = self._m([0, 0, 0, 0, 1])
        rows = [(L.M_DB_MU1, "N")]
        a_b = -U
        for s in range(1, ns):
            a_b = a_b + self._d(s) * self._m([s, 1, 0, 0, 0])
        if self._ke():
            pass  # the first-order KE term enters through _sg_dX_dB at second order only (GH:1717-1720)
        rows.append((L.M_DB, a_b))
        beta = self.data["curr_beta"]
        if ns == 2:
            rows.append((L.M_DD, beta * self._m([1, 1, 0, 0, 0])))
        if order >= 2:
            a_bb = -self._sg_dX_dB([0, 0, 0, 0, 1], 0)
            for s in range(1, ns):
                a_bb = a_bb + self._d(s) * self._sg_dX_dB([s, 1, 0, 0, 0], 0)
            rows.append((L.M_DB2, a_bb))
            if ns == 2:
                rows.append((L.M_DBDD, self._m([1, 1, 0, 0, 0]) + beta * self._sg_dX_dB([1, 1, 0, 0, 0], 0)))
                f = self._m([1, 1, 1, 1, 0]) - self._m([1, 1, 1, 0, 0]) * self._m([1, 0, 1, 1, 0])
                rows.append((L.M_DD2, beta ** 2 * f))
        if order >= 3:
            a_b3 = -self._sg_d2X_dB2([0, 0, 0, 0, 1], 0)
            rows.append((L.M_DB3, a_b3))
        return rows
