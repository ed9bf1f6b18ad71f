"""Drop-in replacement for FHMCAnalysis.moments.histogram.one_dim.n1.gc_hist (reference file
moments/histogram/one_dim/n1/gc_hist.pyx, "N1" below): ln(PI)(N_1) histograms, where the order parameter is the
particle number of species 1 and the chemical potentials of species 2..N are absolute.

SURVEY.md 8(f) row 2.  Everything numerical runs on the same sm_100a kernels as the N_tot class
(../ntot/gc_hist.py): the kernels only see an order-parameter row, a per-state-point monomial in
(mu_1, d beta, d mu_2) and coefficient rows, so the N_1 class is the N_tot class with
  * ``data['n1']`` / ``N_{1}`` instead of ``data['ntot']`` / ``N_{tot}`` (N1:156, 77),
  * ``reweight`` moving only ``curr_mu[0]`` (N1:276),
  * the derivative builders of ``_taylor.N1TaylorMixin`` (absolute mu weights, no KE terms),
  * ``temp_mu_extrap`` / ``temp_mu_extrap_multi`` / ``find_phase_eq(..., mus=...)`` in place of the ``dmu`` family
    (N1:566, 1497, 1435), and a minimum phase width of ``smooth`` bins in the coexistence objective (N1:1479).
There is no CPU fallback.
"""
import copy
import sys

import numpy as np

from fhmcanalysis_b200 import _lib, engine
from ..ntot import gc_hist as _ntot
from ._taylor import N1TaylorMixin


def phase_eq_error(mu_guess, orig_hist, beta, mus, order, cutoff, override, min_width):
    """Objective of the reference's coexistence search (N1:1739-1797)."""
    mu_guess = float(np.atleast_1d(mu_guess)[0])
    hist = copy.deepcopy(orig_hist)
    hist.reweight(mu_guess)
    curr_mu = np.array(hist.data["curr_mu"][1:])
    if beta != orig_hist.data["curr_beta"] or not np.all(curr_mu == mus):
        hist.temp_mu_extrap(beta, mus, order, cutoff, override, False, True)
    hist.thermo(False)
    th = hist.data["thermo"]
    best = 100.0
    for i in range(len(th)):
        if th[i]["bound_idx"][1] - th[i]["bound_idx"][0] < min_width:
            continue
        for j in range(i + 1, len(th)):
            if th[j]["bound_idx"][1] - th[j]["bound_idx"][0] < min_width:
                continue
            best = min(best, (th[i]["F.E./kT"] - th[j]["F.E./kT"]) ** 2)
    return best


class histogram(N1TaylorMixin, _ntot.histogram):
    """1-D ln(PI)(N_1) histogram from grand-canonical flat-histogram simulations (N1:80-1732)."""

    _op_key = "n1"
    _op_var = "N_{1}"

    def __init__(self, fname, beta_ref, mu_ref, smooth=0, ke=False):
        _ntot.histogram.__init__(self, fname, beta_ref, mu_ref, smooth, ke)

    def reload(self):
        """N1:133-174: as the N_tot loader, but the particle-number / energy histograms are mandatory and the order
        parameter must agree with the first moment of species 1."""
        _ntot.histogram.reload(self)
        self.metadata.pop("used_ke", None)                      # N1:88-123 keeps no such key
        if "hist" not in self.data["pk_hist"] or "hist" not in self.data["e_hist"]:
            raise KeyError("P_{N_i}(N_{1}) / P_{U}(N_{1}) not present in " + str(self.metadata["fname"]))
        assert np.all((self.data["mom"][0, 1, 0, 0, 0] - self.data["n1"]) < 1.0e-9), \
            "N_{1} order parameter inconsistent with moments"

    @classmethod
    def from_arrays(cls, lnpi, mom, beta_ref, mu_ref, smooth=0, volume=1.0, n1=None, ke=False):
        h = super(histogram, cls).from_arrays(lnpi, mom, beta_ref, mu_ref, smooth, volume, n1, False)
        h.metadata.pop("used_ke", None)
        return h

    def _device_hist(self, sel=(), coef=(), sel_kinds=(), smooth=None, cutoff=10.0, device=None):
        """As the N_tot upload, but the second state variable is mu_2 itself (dmu_ref := curr_mu[1])."""
        dh = _ntot.histogram._device_hist(self, sel, coef, sel_kinds, smooth, cutoff, device)
        if self.data["nspec"] > 1:
            dh.desc.dmu_ref = float(self.data["curr_mu"][1])
        return dh

    def mix(self, other, weights):
        """N1:176-249 (no used_ke bookkeeping)."""
        self.metadata["used_ke"] = other.metadata["used_ke"] = False
        try:
            mixed = _ntot.histogram.mix(self, other, weights)
        finally:
            self.metadata.pop("used_ke", None)
            other.metadata.pop("used_ke", None)
        mixed.metadata.pop("used_ke", None)
        return mixed

    def reweight(self, mu1_target, print_screen=False):
        """N1:259-279: only the chemical potential of species 1 moves."""
        mu1_target = float(mu1_target)
        self._cy_reweight(mu1_target)
        self.data["curr_mu"][0] = mu1_target
        if print_screen:
            for i in range(len(self.data["ln(PI)"])):
                print(i, self.data["ln(PI)"][i] - self.data["ln(PI)"][0])

    def thermo(self, props=True, complete=False):
        """N1:438-526 (no ``collect`` hook)."""
        _ntot.histogram.thermo(self, props, complete, None)

    # ------------------------------------------------------------------------------------------
    def _check_not_extrapolated(self, check_beta=True, check_dmu=True):
        if check_beta and np.abs(self.metadata["beta_ref"] - self.data["curr_beta"]) > 1.0e-6:
            raise Exception("Cannot extrapolate the same histogram class twice")
        if check_dmu and np.any(np.abs(self.metadata["mu_ref"][1:] - self.data["curr_mu"][1:]) > 1.0e-6):
            raise Exception("Cannot extrapolate the same histogram class twice")

    def _xi(self, target_beta, target_mus):
        xi = np.zeros(self.data["nspec"], dtype=np.float64)
        xi[0] = target_beta - self.data["curr_beta"]
        xi[1:] = np.asarray(target_mus, dtype=np.float64) - self.data["curr_mu"][1:]
        return xi

    def _temp_mu_extrap_1(self, target_beta, target_mus, cutoff=10.0, override=False, skip_mom=False):
        self._temp_dmu_extrap_1(target_beta, target_mus, cutoff, override, skip_mom)

    def _temp_mu_extrap_2(self, target_beta, target_mus, cutoff=10.0, override=False, skip_mom=False):
        self._temp_dmu_extrap_2(target_beta, target_mus, cutoff, override, skip_mom, False)

    def temp_mu_extrap(self, target_beta, target_mus, order=1, cutoff=10.0, override=False, clone=True, skip_mom=False):
        """Simultaneous temperature and mu_2..mu_N extrapolation (N1:566-640)."""
        self._check_not_extrapolated(check_dmu=False)
        target_mus = np.asarray(target_mus, dtype=np.float64)
        assert len(target_mus) == self.data["nspec"] - 1, "Must specify mu values for all components 2-N"
        self._check_not_extrapolated(check_beta=False)
        self._check_order(order, skip_mom)
        tmp_hist = copy.deepcopy(self) if clone else self
        tmp_hist.normalize()
        if order == 1:
            try:
                tmp_hist._temp_mu_extrap_1(target_beta, target_mus, cutoff, override, skip_mom)
            except Exception as e:
                raise Exception("Unable to extrapolate : " + str(e))
        elif order == 2:
            try:
                tmp_hist._temp_mu_extrap_2(target_beta, target_mus, cutoff, override, skip_mom)
            except Exception as e:
                raise Exception("Unable to extrapolate : " + str(e))
        else:
            raise Exception("No implementation for temperature + mu extrapolation of order " + str(order))
        tmp_hist.data["curr_beta"] = target_beta
        tmp_hist.data["curr_mu"][1:] = copy.copy(target_mus)
        tmp_hist.normalize()
        return tmp_hist

    def temp_mu_extrap_multi(self, target_betas, target_mus, order=1, cutoff=10.0, override=False, skip_mom=False):
        """(beta x mu) grid of extrapolated histograms (N1:1497-1569, 1571-1732); derivatives built once."""
        self._check_not_extrapolated(check_dmu=False)
        target_betas = np.asarray(target_betas, dtype=np.float64)
        target_mus = [np.asarray(t, dtype=np.float64) for t in target_mus]
        for t in target_mus:
            assert len(t) == self.data["nspec"] - 1, "Must specify mu for all components 2-N"
        self._check_not_extrapolated(check_beta=False)
        self._check_order(order, skip_mom)
        if order not in (1, 2):
            raise Exception("No implementation for temperature + mu extrapolation of order " + str(order))
        try:
            self._edge_assert(cutoff, override)
            cc = copy.deepcopy(self)
            cc.normalize()
            try:
                grad = cc._dBMU(skip_mom)
                hess = cc._dBMU2(skip_mom) if order == 2 else None
            except Exception:
                raise Exception("Unable to compute first derivative" if order == 1 else "Unable to compute derivatives")
        except Exception as e:
            raise Exception("Unable to extrapolate : " + str(e))
        hists = []
        for tb in target_betas:
            row = []
            for tm in target_mus:
                try:
                    clone = copy.deepcopy(self)
                    clone._taylor_update(list(self._xi(tb, tm)), grad, hess, skip_mom, False)
                except Exception:
                    clone = None
                row.append(clone)
            hists.append(row)
        for i in range(len(target_betas)):        # N1:1562-1567: a failed cell (None) raises here, like the reference
            for j in range(len(target_mus)):
                hists[i][j].data["curr_beta"] = copy.copy(target_betas[i])
                hists[i][j].data["curr_mu"][1:] = copy.copy(target_mus[j])
                hists[i][j].normalize()
        return hists

    # the dmu family does not exist for N_1 histograms
    def temp_extrap(self, *a, **k):
        raise AttributeError("N_1 histograms extrapolate with temp_mu_extrap()")

    dmu_extrap = temp_dmu_extrap = temp_dmu_extrap_multi = temp_extrap

    # ------------------------------------------------------------------------------------------
    def find_phase_eq(self, lnZ_tol, mu_guess, beta=0.0, mus=[], extrap_order=1, cutoff=10.0, override=False):
        """Coexistence search (N1:1435-1495) on the batched device solver; ``self`` is not modified."""
        tmp_hist = copy.deepcopy(self)
        curr_mu = np.array(self.data["curr_mu"][1:], dtype=np.float64)
        if len(mus) == 0:
            new_mu = copy.copy(curr_mu)
        else:
            assert len(mus) == self.data["nspec"] - 1, "Need to specify mu for components 2-N"
            new_mu = np.array(mus, dtype=np.float64)
        if beta <= 0.0:
            beta = self.data["curr_beta"]
        extrap = (beta != self.data["curr_beta"]) or not np.all(new_mu == curr_mu)
        tmp_hist.normalize()
        coef = ()
        if extrap:
            if np.abs(self.metadata["beta_ref"] - self.data["curr_beta"]) > 1.0e-6:
                raise Exception("Cannot extrapolate the same histogram class twice")
            coef = tmp_hist.taylor_rows(extrap_order)
        dh = tmp_hist._device_hist(sel=["N"], coef=coef, cutoff=cutoff)
        res = dh.find_phase_eq(np.array([float(mu_guess)]), beta=np.array([float(beta)]) if extrap else None,
                               dmu=np.array([float(new_mu[0])]) if (extrap and len(new_mu)) else None,
                               lnz_tol=min(float(lnZ_tol), 1e-10), pmax=8,
                               min_width=max(int(tmp_hist.metadata["smooth"]), 1))
        h = res.host()
        if int(h["code"][0]) != 0:
            raise Exception("Error, unable to locate phase coexistence : " +
                            _lib.STATUS_TEXT.get(int(h["code"][0]), "solver status %d" % int(h["code"][0])))
        try:
            tmp_hist.reweight(float(h["mu_coex"][0]))
            if extrap:
                tmp_hist.temp_mu_extrap(beta, new_mu, extrap_order, cutoff, override, False)
            tmp_hist.thermo()
        except Exception as e:
            raise Exception("Found coexistence, but unable to compute properties afterwards: " + str(e))
        return tmp_hist

    # ------------------------------------------------------------------------------------------
    # batched entry points: same kernels, absolute mu_2 as the second state variable
    # ------------------------------------------------------------------------------------------
    def reweight_batch(self, mu1, beta=None, mu2=None, **kw):
        return _ntot.histogram.reweight_batch(self, mu1, beta=beta, dmu=mu2, **kw)

    def find_phase_eq_batch(self, betas, mu_guess, mu2=None, **kw):
        betas = np.atleast_1d(np.asarray(betas, dtype=np.float64))
        order = kw.pop("order", 1)
        moments = kw.pop("moments", ("N", "N2", "U"))
        moments = ["N"] + [m for m in moments if m != "N"]
        dh = self.device_histogram(betas, mu2, order, moments, kw.pop("cutoff", 10.0), kw.pop("device", None))
        guess = np.broadcast_to(np.asarray(mu_guess, dtype=np.float64), betas.shape).copy()
        res = dh.find_phase_eq(guess, beta=betas, dmu=mu2, lnz_tol=kw.pop("lnZ_tol", 1e-10),
                               max_iter=kw.pop("max_iter", 200), pmax=kw.pop("pmax", 4),
                               min_width=max(int(self.metadata["smooth"]), 1))
        return res if kw.pop("return_device", False) else res.host()


if __name__ == "__main__":
    print("gc_hist (N_1, B200)")
    sys.exit(0)
