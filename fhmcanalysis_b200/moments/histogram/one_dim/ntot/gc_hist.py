"""Drop-in replacement for FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist (reference file
moments/histogram/one_dim/ntot/gc_hist.pyx, "GH" below), running on B200 kernels.

Same class name, constructor, ``metadata`` / ``data`` dictionaries, method names, positional argument
order and exceptions as the reference.  The NumPy arrays in ``self.data`` stay
the source of truth (users and the reference's tests overwrite them directly); every numerical method
uploads what it needs, runs the sm_100a kernels through the C ABI (libfhmc_b200.so) and writes the
results back.  There is no CPU fallback: without a GPU the numerical methods raise.

New batched entry points (north star): ``reweight_batch``, ``temp_dmu_extrap_batch`` (alias through
``reweight_batch(beta=..., dmu=...)``) and ``find_phase_eq_batch``.
"""
import copy
import sys

import numpy as np

from fhmcanalysis_b200 import _lib, engine
from ._taylor import TaylorMixin

try:  # the reference reads composite.nc with netCDF4 (GH:17, 143); use it when it exists
    from netCDF4 import Dataset  # type: ignore
except Exception:  # pragma: no cover - netCDF4 is absent in the build image
    from fhmcanalysis_b200.io.hdf5_min import Dataset

# GH:29: overflow / invalid / divide raise, underflow is ignored
_ERRSTATE = dict(divide="raise", over="raise", invalid="raise", under="ignore")


def _status_exception(code):
    return Exception(_lib.STATUS_TEXT.get(int(code), "status %d" % int(code)))


def phase_eq_error(mu_guess, orig_hist, beta, dMu, order, cutoff, override, min_width, collect):
    """Objective of the reference's coexistence search (GH:2570-2630): reweight a copy to ``mu_guess``,
    extrapolate if (beta, dMu) differ, thermo(props=False), and return the smallest squared F.E./kT
    difference between two phases at least ``min_width`` bins wide (100.0 when there is no such pair)."""
    mu_guess = float(np.atleast_1d(mu_guess)[0])
    hist = copy.deepcopy(orig_hist)
    hist.reweight(mu_guess)
    curr_dMu = np.array([hist.data["curr_mu"][i] - hist.data["curr_mu"][0] for i in range(1, hist.data["nspec"])])
    if beta != orig_hist.data["curr_beta"] or not np.all(curr_dMu == dMu):
        hist.temp_dmu_extrap(beta, dMu, order, cutoff, override, False, True)
    hist.thermo(props=False, collect=collect)
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


class histogram(TaylorMixin):
    """1-D ln(PI)(N_tot) histogram from grand-canonical flat-histogram simulations (GH:80-2563)."""

    _op_key = "ntot"        # data[] key of the order-parameter vector (the N_1 subclass uses "n1")
    _op_var = "N_{tot}"     # its name inside composite.nc

    def __init__(self, fname, beta_ref, mu_ref, smooth=0, ke=False):
        self.metadata = {}
        self.metadata["beta_ref"] = beta_ref
        if isinstance(mu_ref, list):
            assert len(mu_ref) > 0, "Incomplete chemical potential information"
            self.metadata["mu_ref"] = np.array(mu_ref, dtype=np.float64)
        elif isinstance(mu_ref, (float, np.float64, int, np.int32, np.int64)):
            self.metadata["mu_ref"] = np.array([mu_ref], dtype=np.float64)
        else:
            raise Exception("Unrecognized type for mu_ref")
        self.metadata["nspec"] = len(self.metadata["mu_ref"])
        assert self.metadata["beta_ref"] > 0, "Illegal beta value"
        self.metadata["smooth"] = smooth
        assert self.metadata["smooth"] >= 0, "Illegal smooth value"
        assert isinstance(fname, str), "Expects filename as a string"
        self.metadata["fname"] = fname
        self.metadata["used_ke"] = ke
        self.reload()

    @classmethod
    def from_arrays(cls, lnpi, mom, beta_ref, mu_ref, smooth=0, volume=1.0, ntot=None, ke=False):
        """Build a histogram from in-memory arrays (fills exactly the keys ``reload`` fills, GH:137-182)."""
        h = cls.__new__(cls)
        mu = np.atleast_1d(np.array(mu_ref, dtype=np.float64))
        h.metadata = {"beta_ref": float(beta_ref), "mu_ref": mu.copy(), "nspec": len(mu), "smooth": int(smooth),
                      "fname": "", "used_ke": bool(ke), "file_history": "from_arrays"}
        lnpi = np.array(lnpi, dtype=np.float64)
        if ntot is None:
            ntot = np.arange(len(lnpi), dtype=np.int64)
        ntot = np.array(ntot, dtype=np.int64)
        mom = np.array(mom, dtype=np.float64)
        h.data = {"curr_mu": mu.copy(), "curr_beta": float(beta_ref), "nspec": len(mu), "ln(PI)": lnpi,
                  "max_order": mom.shape[1] - 1, "volume": float(volume), cls._op_key: ntot, "lb": ntot[0], "ub": ntot[-1],
                  "pk_hist": {}, "e_hist": {}, "mom": mom}
        return h

    def to_nc(self, fname):
        """Write the CURRENT state as a ``composite.nc`` (names and attributes of ``window.to_nc``,
        win_patch/fhmc_patch.pyx:551-634) with the built-in HDF5 writer, so that ``histogram(fname, ...)`` reloads it."""
        from fhmcanalysis_b200.io.hdf5_min import write_composite
        extra = {}
        for fam, key in (("P_{N_i}(" + self._op_var + ")", "pk_hist"), ("P_{U}(" + self._op_var + ")", "e_hist")):
            hst = self.data.get(key, {})
            if "hist" in hst:
                extra[fam] = hst["hist"]
                for sfx in ("lb", "ub", "bw"):
                    extra[fam + "_{" + sfx + "}"] = hst[sfx]
        write_composite(fname, self.data["ln(PI)"], self.data[self._op_key], self.data["mom"], self.data["volume"],
                        self.data["nspec"], self.data["max_order"],
                        history=str(self.metadata.get("file_history", "")), histograms=extra, op_name=self._op_var)

    def clear(self):
        self.data = {}

    def reload(self):
        """(Re)load data from the netCDF4 file (GH:131-182)."""
        self.clear()
        self.data["curr_mu"] = copy.copy(self.metadata["mu_ref"])
        self.data["curr_beta"] = copy.copy(self.metadata["beta_ref"])
        self.data["nspec"] = copy.copy(self.metadata["nspec"])
        try:
            dataset = Dataset(self.metadata["fname"], "r", format="NETCDF4")
        except Exception as e:
            raise Exception("Unable to load data from " + str(self.metadata["fname"]) + " : " + str(e))
        self.metadata["file_history"] = copy.copy(dataset.history)
        self.data["ln(PI)"] = np.array(dataset.variables["ln(PI)"][:], dtype=np.float64)
        assert dataset.nspec == self.metadata["nspec"], \
            "Different number of species in datafile from information initially specified"
        self.data["max_order"] = int(dataset.max_order)
        assert self.data["max_order"] > 0, "Error, max_order < 1"
        self.data["volume"] = float(dataset.volume)
        assert self.data["volume"] > 0, "Error, volume <= 0"
        op, ov = self._op_key, self._op_var
        self.data[op] = np.array(dataset.variables[ov][:], dtype=np.int64)
        self.data["lb"] = self.data[op][0]
        self.data["ub"] = self.data[op][len(self.data[op]) - 1]
        assert self.data["lb"] < self.data["ub"], "Error, bad bounds for N_tot"
        self.data["pk_hist"] = {}
        try:
            self.data["pk_hist"]["hist"] = np.array(dataset.variables["P_{N_i}(" + ov + ")"][:])
            self.data["pk_hist"]["lb"] = np.array(dataset.variables["P_{N_i}(" + ov + ")_{lb}"][:])
            self.data["pk_hist"]["ub"] = np.array(dataset.variables["P_{N_i}(" + ov + ")_{ub}"][:])
            self.data["pk_hist"]["bw"] = np.array(dataset.variables["P_{N_i}(" + ov + ")_{bw}"][:])
        except Exception:
            pass
        self.data["e_hist"] = {}
        try:
            self.data["e_hist"]["hist"] = np.array(dataset.variables["P_{U}(" + ov + ")"][:])
            self.data["e_hist"]["lb"] = np.array(dataset.variables["P_{U}(" + ov + ")_{lb}"][:])
            self.data["e_hist"]["ub"] = np.array(dataset.variables["P_{U}(" + ov + ")_{ub}"][:])
            self.data["e_hist"]["bw"] = np.array(dataset.variables["P_{U}(" + ov + ")_{bw}"][:])
        except Exception:
            pass
        self.data["mom"] = np.array(dataset.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:])
        mo, ns = self.data["max_order"], self.data["nspec"]
        assert self.data["mom"].shape == (ns, mo + 1, ns, mo + 1, mo + 1, len(self.data[op]))
        dataset.close()

    # ------------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------------
    def _device_hist(self, sel=(), coef=(), sel_kinds=(), smooth=None, cutoff=10.0, device=None):
        """Upload the CURRENT state (ln(PI), ntot, curr_beta, curr_mu) as a kernel blob."""
        dmu_ref = float(self.data["curr_mu"][1] - self.data["curr_mu"][0]) if self.data["nspec"] > 1 else 0.0
        sm = self.metadata["smooth"] if smooth is None else smooth
        lnpi = np.asarray(self.data["ln(PI)"], dtype=np.float64)
        ntot = self.data[self._op_key]
        if len(ntot) != len(lnpi):
            # the reference's tests assign shorter arrays to data['ln(PI)'] (T1:155-198); normalize/relextrema/thermo
            # never touch ntot (no reweighting shift: s = 0), so any N row of the right length will do
            ntot = np.arange(len(lnpi))
        return engine.DeviceHistogram(lnpi, ntot,
                                      self.data["curr_beta"], self.data["curr_mu"][0], dmu_ref, smooth=max(int(sm), 1),
                                      cutoff=cutoff, coef=coef, sel=sel, sel_kinds=sel_kinds, device=device)

    def _scalar_point(self, mu1_target, **kw):
        """One state point through the cached scalar path (engine.ScalarPath -> fhmc_scalar_point): the host arrays are the
        source of truth, the device copies are refreshed only when their content key changed."""
        lnpi = np.asarray(self.data["ln(PI)"], dtype=np.float64)
        ntot = self.data[self._op_key]
        if len(ntot) != len(lnpi):
            # the reference's tests assign shorter arrays to data['ln(PI)'] (T1:155-198); normalize/relextrema/thermo
            # never touch ntot (no reweighting shift: s = 0), so any N row of the right length will do
            ntot = np.arange(len(lnpi))
        sp = engine.ScalarPath.get(len(lnpi))
        return sp.point(lnpi, ntot, self.data["curr_beta"], self.data["curr_mu"][0], self.metadata["smooth"], float(mu1_target), **kw)

    def _renormalised(self, mu1_target):
        """ln(PI) reweighted to mu1_target and normalised (K1): returns (array, lnNorm)."""
        r = self._scalar_point(mu1_target, complete=True, want_row=True)
        return r["row"], r["lnnorm"]

    # ------------------------------------------------------------------------------------------
    def mix(self, other, weights):
        """Weighted blend of two histograms at identical conditions (GH:184-258); arithmetic on the device."""
        tol = 1.0e-9
        bad = Exception("Difference in conditions, cannot mix histograms")
        if self.metadata["nspec"] != other.metadata["nspec"]: raise bad
        if self.metadata["used_ke"] != other.metadata["used_ke"]: raise bad
        if self.data["nspec"] != other.data["nspec"]: raise bad
        if abs(self.data["curr_beta"] - other.data["curr_beta"]) > tol: raise bad
        if not np.all(np.abs(self.data["curr_mu"] - other.data["curr_mu"]) < tol): raise bad
        if abs(self.data["volume"] - other.data["volume"]) > tol: raise bad
        if self.data["max_order"] != other.data["max_order"]: raise bad
        if len(self.data["mom"]) != len(other.data["mom"]): raise bad
        if self.data["lb"] != other.data["lb"]: raise bad
        if not isinstance(weights, (np.ndarray, list, tuple)): raise Exception("Requires 2 weights, cannot mix histograms")
        if len(weights) != 2: raise Exception("Requires 2 weights, cannot mix histograms")
        if len(self.data["ln(PI)"]) >= len(other.data["ln(PI)"]):
            longer_one, max_idx = self, len(other.data["ln(PI)"])
        else:
            longer_one, max_idx = other, len(self.data["ln(PI)"])
        mixed = copy.deepcopy(longer_one)
        mixed.data["file_history"] = "this is a mixed histogram"
        mixed.metadata["fname"] = ""
        mixed.metadata["beta_ref"] = mixed.data["curr_beta"]
        mixed.metadata["mu_ref"] = mixed.data["curr_mu"]
        w0, w1 = float(weights[0]), float(weights[1])
        wsum = w0 + w1
        # (X_self*w0 + w1*X_other)/(w0+w1): one fused device pass over ln(PI) and the whole moment tensor
        stack_a = np.concatenate([np.asarray(self.data["ln(PI)"], dtype=np.float64)[None, :max_idx],
                                  np.asarray(self.data["mom"], dtype=np.float64)[..., :max_idx].reshape(-1, max_idx)])
        stack_b = np.concatenate([np.asarray(other.data["ln(PI)"], dtype=np.float64)[None, :max_idx],
                                  np.asarray(other.data["mom"], dtype=np.float64)[..., :max_idx].reshape(-1, max_idx)])
        out = engine.axpy_rows([stack_a, stack_b], [w0 / wsum, w1 / wsum])
        mixed.data["ln(PI)"] = mixed.data["ln(PI)"].astype(np.float64)
        mixed.data["ln(PI)"][:max_idx] = out[0]
        mixed.data["mom"] = mixed.data["mom"].astype(np.float64)
        mixed.data["mom"][..., :max_idx] = out[1:].reshape(mixed.data["mom"].shape[:-1] + (max_idx,))
        mixed.data["pk_hist"] = {}
        mixed.data["e_hist"] = {}
        return mixed

    def normalize(self):
        """ln(PI) <- ln(PI) - ln sum exp ln(PI)  (GH:260-266, 57-67)."""
        self._cy_normalize()

    def _cy_normalize(self):
        self.data["ln(PI)"], _ = self._renormalised(self.data["curr_mu"][0])

    def _cy_reweight(self, mu1_new):
        self.data["ln(PI)"], _ = self._renormalised(mu1_new)

    def reweight(self, mu1_target, print_screen=False):
        """Reweight to another chemical potential of species 1 and normalise (GH:268-289)."""
        mu1_target = float(mu1_target)
        dmu1 = mu1_target - self.data["curr_mu"][0]
        self._cy_reweight(mu1_target)
        self.data["curr_mu"] = self.data["curr_mu"] + dmu1
        if print_screen:
            for i in range(len(self.data["ln(PI)"])):
                print(i, self.data["ln(PI)"][i] - self.data["ln(PI)"][0])

    # ------------------------------------------------------------------------------------------
    def _split(self, compare_raw, want_norm, mom=None):
        """K3 (+K1 normalisation, + K2 over the rows of ``mom``): extrema lists / bounds / F.E. of the current ln(PI)."""
        if int(self.metadata["smooth"]) < 1:
            raise ValueError("Order must be an int >= 1")  # scipy.signal.argrelextrema (GH:329)
        r = self._scalar_point(self.data["curr_mu"][0], compare_raw=compare_raw, want_row=want_norm, mom=mom)
        if r["code"] == _lib.E_CAPACITY:
            # more extrema than the scalar path's record holds (very noisy ln(PI)): the growing-capacity batched call
            dh = self._device_hist()
            res = dh.sweep_auto(np.array([float(self.data["curr_mu"][0])]), pmax=32, lanes=32, compare_raw=compare_raw)
            h = res.host()
            P, nm = int(h["nphase"][0]), int(h["nmin"][0])
            out = {"code": int(h["code"][0]), "maxima": h["max_idx"][0, :P].astype(np.int64), "minima": h["min_idx"][0, :nm].astype(np.int64),
                   "bounds": h["bounds"][0, :P].astype(np.int64), "fe": h["fe"][0, :P].copy(), "lnnorm": float(h["lnnorm"][0])}
            if want_norm or mom is not None:
                out["lnpi"] = dh.lnpi_rows(res)[0].cpu().numpy()
            return out
        P, nm = r["nphase"], r["nmin"]
        out = {"code": r["code"], "maxima": r["max_idx"][:P].astype(np.int64), "minima": r["min_idx"][:nm].astype(np.int64),
               "bounds": r["bounds"][:P].astype(np.int64), "fe": r["fe"][:P].copy(), "lnnorm": r["lnnorm"]}
        if want_norm or mom is not None:
            out["lnpi"] = r["row"]
        if mom is not None:
            out["avg"], out["lnsum"] = r["avg"], r["lnsum"]
        return out

    def relextrema(self):
        """Locate the local extrema of ln(PI) (GH:317-415)."""
        last_idx = len(self.data["ln(PI)"]) - 1
        if last_idx <= 1:
            raise Exception("ln(PI) not long enough to analyze for relative extrema")
        s = self._split(compare_raw=True, want_norm=False)
        self.data["ln(PI)_maxima_idx"] = s["maxima"]
        self.data["ln(PI)_minima_idx"] = s["minima"]
        if s["code"] == 4:
            raise Exception("There are " + str(len(s["maxima"])) + " local maxima and " + str(len(s["minima"])) +
                            " local minima, so cannot be alternating, try adjusting the value of smooth")
        if s["code"] == 5:
            raise Exception("Local maxima and minima not sorted correctly, try adjusting the value of smooth (max,min) = " +
                            str(s["maxima"]) + ", " + str(s["minima"]))
        if s["code"] != 0:
            raise _status_exception(s["code"])

    def coexisting(self, rtol=1.0e-3):
        """Indices of phases with equal free energy (GH:417-449)."""
        if "thermo" not in self.data:
            raise Exception("Thermodynamic properties should be called first (self.thermo())")
        th = self.data["thermo"]
        if len(th) == 1:
            return [[]]
        eq = []
        for i in range(len(th)):
            x = [i]
            for j in range(i + 1, len(th)):
                if abs((th[i]["F.E./kT"] - th[j]["F.E./kT"]) / th[i]["F.E./kT"]) < rtol:
                    x.append(j)
            if len(x) > 1:
                eq.append(x)
        return eq

    @staticmethod
    def _bounds_from_lists(n, maxima, minima):
        """GH:498-520."""
        bounds, ctr = [], 0
        for p in range(len(maxima)):
            if maxima[p] > 0:
                left = int(minima[ctr])
                ctr += 1
            else:
                left = 0
            right = int(minima[ctr]) if maxima[p] < n - 1 else n
            if right == n - 1:
                right += 1
            bounds.append((left, right))
        return bounds

    def thermo(self, props=True, complete=False, collect=None):
        """Integrate ln(PI) per phase and average every moment array (GH:451-554)."""
        n = len(self.data["ln(PI)"])
        fused = None
        if not complete:
            use_mom = props and collect is None and np.shape(self.data["mom"])[-1] == n
            try:
                s = self._split(compare_raw=False, want_norm=True,
                                mom=np.asarray(self.data["mom"], dtype=np.float64).reshape(-1, n) if use_mom else None)
            except ValueError as e:
                raise Exception("Unable to find relative extrema : " + str(e))
            self.data["ln(PI)"] = s["lnpi"]  # GH:475: thermo leaves the normalised array behind
            self.data["ln(PI)_maxima_idx"] = s["maxima"]
            self.data["ln(PI)_minima_idx"] = s["minima"]
            if s["code"] != 0:
                if s["code"] == 4:
                    msg = ("There are " + str(len(s["maxima"])) + " local maxima and " + str(len(s["minima"])) +
                           " local minima, so cannot be alternating, try adjusting the value of smooth")
                elif s["code"] == 5:
                    msg = "Local maxima and minima not sorted correctly, try adjusting the value of smooth"
                else:
                    msg = str(_status_exception(s["code"]))
                if s["code"] == 6:
                    raise IndexError("index out of bounds")
                raise Exception("Unable to find relative extrema : " + msg)
            bounds, fe = [tuple(int(v) for v in b) for b in s["bounds"]], s["fe"]
            if "avg" in s:
                fused = s["avg"]       # averaged on the device in the same call, over the bounds of the record
            if collect is not None:
                collect(hist=self)
                bounds = self._bounds_from_lists(n, self.data["ln(PI)_maxima_idx"], self.data["ln(PI)_minima_idx"])
                fe = None
        else:
            self.data["ln(PI)"], _ = self._renormalised(self.data["curr_mu"][0])
            bounds, fe = [(0, n)], None
        nphases = len(bounds)
        phase = {}
        avg = None
        if fused is not None:
            avg = fused
        elif props or fe is None:
            mom = np.asarray(self.data["mom"], dtype=np.float64) if props else None
            avg, lnsum = engine.phase_moments(self.data["ln(PI)"], mom.reshape(-1, n) if props else None, np.array(bounds))
            if fe is None:
                fe = -(lnsum - self.data["ln(PI)"][0])
        ns = self.data["nspec"]
        for p in range(nphases):
            phase[p] = {"F.E./kT": float(fe[p]), "bound_idx": bounds[p]}
            if props:
                pm = avg[p].reshape(self.data["mom"].shape[:-1])
                phase[p]["mom"] = pm
                with np.errstate(**_ERRSTATE):
                    nsum = 0.0
                    for i in range(ns):
                        phase[p]["n" + str(i + 1)] = pm[i, 1, 0, 0, 0]
                        nsum += pm[i, 1, 0, 0, 0]
                    phase[p]["ntot"] = nsum
                    phase[p]["density"] = nsum / self.data["volume"]
                    phase[p]["u"] = pm[0, 0, 0, 0, 1]
                    for i in range(ns):
                        phase[p]["x" + str(i + 1)] = np.float64(pm[i, 1, 0, 0, 0]) / np.float64(nsum)
        self.data["thermo"] = phase

    def is_safe(self, cutoff=10.0, complete=False):
        """Edge test (GH:556-596).  Like the reference it uses the CACHED maxima indices when they exist
        (two array reads and one comparison on the host arrays that are the source of truth)."""
        lnpi = self.data["ln(PI)"]
        if not complete:
            if "ln(PI)_maxima_idx" not in self.data:
                try:
                    self.normalize()
                except Exception as e:
                    raise Exception("Unable to normalize ln(PI) : " + str(e))
                try:
                    self.relextrema()
                except Exception as e:
                    raise Exception("Unable to find relative extrema in ln(PI) : " + str(e))
                lnpi = self.data["ln(PI)"]
            maxima = lnpi[self.data["ln(PI)_maxima_idx"]]
            return not (maxima[len(maxima) - 1] - lnpi[len(lnpi) - 1] < cutoff)
        return not (np.max(lnpi) - lnpi[len(lnpi) - 1] < cutoff)

    # ------------------------------------------------------------------------------------------
    def find_phase_eq(self, lnZ_tol, mu_guess, beta=0.0, dMu=[], extrap_order=1, cutoff=10.0, override=False,
                      reterr=False, first_order_mom=False, collect=None):
        """Coexistence search (GH:598-668).  ``self`` is not modified.  The search itself is the batched
        device solver (K4) for one temperature; with a ``collect`` callback (arbitrary host code per
        evaluation) it is the reference's Nelder-Mead over device-evaluated objectives."""
        tmp_hist = copy.deepcopy(self)
        curr_dMu = np.array([self.data["curr_mu"][i] - self.data["curr_mu"][0] for i in range(1, self.data["nspec"])],
                            dtype=np.float64)
        if len(dMu) == 0:
            new_dMu = copy.copy(curr_dMu)
        else:
            assert len(dMu) == self.data["nspec"] - 1, "Need to specify dMu for components 2-N"
            new_dMu = np.array(dMu, dtype=np.float64)
        if beta <= 0.0:
            beta = self.data["curr_beta"]
        extrap = (beta != self.data["curr_beta"]) or not np.all(new_dMu == curr_dMu)
        tmp_hist.normalize()
        min_width = tmp_hist.metadata["smooth"] * 2
        coef, use_host_search = (), collect is not None
        if extrap and not use_host_search:
            if np.abs(self.metadata["beta_ref"] - self.data["curr_beta"]) > 1.0e-6:
                raise Exception("Cannot extrapolate the same histogram class twice")
            # (no edge assert here: the reference runs the whole search with override=True, GH:652, and applies the
            # caller's `override` only to the final extrapolation at mu*, GH:660)
            try:
                coef = tmp_hist.taylor_rows(extrap_order)
            except NotImplementedError:
                use_host_search = True      # > 2 species: no coefficient-row form; the reference's simplex on device evaluations
        if use_host_search:
            from scipy.optimize import fmin
            full_out = fmin(phase_eq_error, mu_guess, ftol=lnZ_tol,
                            args=(tmp_hist, beta, new_dMu, extrap_order, cutoff, True, min_width, collect),
                            maxfun=100000, maxiter=100000, full_output=True, disp=False)
            if full_out[4] != 0:
                raise Exception("Error, unable to locate phase coexistence : " + str(full_out))
            mu_star, err2 = float(full_out[0][0]), float(full_out[1])
        else:
            dh = tmp_hist._device_hist(sel=["N"], coef=coef, cutoff=cutoff)
            pmax = 8
            while True:   # a noisy ln(PI) can show more extrema than pmax at some intermediate mu: grow, like sweep_auto
                res = dh.find_phase_eq(np.array([float(mu_guess)]), beta=np.array([float(beta)]) if extrap else None,
                                       dmu=np.array([float(new_dMu[0])]) if (extrap and len(new_dMu)) else None,
                                       lnz_tol=min(float(lnZ_tol), 1e-10), pmax=pmax)
                h = res.host()
                if int(h["code"][0]) != _lib.E_CAPACITY or pmax > len(tmp_hist.data["ln(PI)"]):
                    break
                pmax = min(pmax * 4, len(tmp_hist.data["ln(PI)"]) + 1)
            if int(h["code"][0]) != 0:
                raise Exception("Error, unable to locate phase coexistence : " +
                                _lib.STATUS_TEXT.get(int(h["code"][0]), "solver status %d" % int(h["code"][0])))
            mu_star, err2 = float(h["mu_coex"][0]), float(h["dfe"][0]) ** 2
        try:
            tmp_hist.reweight(mu_star)
            if extrap:
                tmp_hist.temp_dmu_extrap(beta, new_dMu, extrap_order, cutoff, override, False, False, first_order_mom)
            tmp_hist.thermo(collect=collect)
        except Exception as e:
            raise Exception("Found coexistence, but unable to compute properties afterwards: " + str(e))
        if reterr:
            return tmp_hist, err2
        return tmp_hist

    # ------------------------------------------------------------------------------------------
    # Taylor extrapolation (GH:670-1239, 1995-2112, 2254-2340)
    # ------------------------------------------------------------------------------------------
    def _check_not_extrapolated(self, check_beta=True, check_dmu=True):
        if check_beta and np.abs(self.metadata["beta_ref"] - self.data["curr_beta"]) > 1.0e-6:
            raise Exception("Cannot extrapolate the same histogram class twice")
        if check_dmu:
            orig = self.metadata["mu_ref"][1:] - self.metadata["mu_ref"][0]
            curr = self.data["curr_mu"][1:] - self.data["curr_mu"][0]
            if np.any(np.abs(orig - curr) > 1.0e-6):
                raise Exception("Cannot extrapolate the same histogram class twice")

    def _check_order(self, order, skip_mom):
        if self.data["max_order"] < (order if skip_mom else order + 1):
            raise Exception("Maximum order stored in simulation not high enough to calculate this order of extrapolation")

    def _edge_assert(self, cutoff, override):
        if not override:
            lp = self.data["ln(PI)"]
            assert np.max(lp) - cutoff > lp[len(lp) - 1], "Error, histogram edge effect encountered in temperature extrapolation"

    @staticmethod
    def _apply_terms(base, terms):
        """base + sum_t w_t * array_t on the device (terms: list of (w, array)); chunks of <= 7 terms."""
        terms = [(float(w), a) for (w, a) in terms if w != 0.0]
        out = np.array(base, dtype=np.float64)
        step = _lib.MAX_TERMS - 1
        for k in range(0, len(terms), step):
            chunk = terms[k:k + step]
            out = engine.axpy_rows([out] + [a for _, a in chunk], [1.0] + [w for w, _ in chunk])
        return out

    def _taylor_update(self, xi, grad, hess, skip_mom, first_order_mom, third=None):
        """ln(PI) += xi.grad + 1/2 xi^T H xi (+ xi0^3/6 third); the same for every moment array
        (GH:1023-1034, 1157-1171, 2023-2031, 2062-2070, 2103-2112)."""
        dlnpi, dm = grad
        nx = len(xi)
        t_l = [(xi[q], dlnpi[q]) for q in range(nx)]
        t_m = [(xi[q], dm[q]) for q in range(nx)]
        if hess is not None:
            H, Hm = hess
            for q in range(nx):
                for r in range(nx):
                    t_l.append((0.5 * xi[q] * xi[r], H[q, r]))
                    if not first_order_mom:
                        t_m.append((0.5 * xi[q] * xi[r], Hm[q, r]))
        if third is not None:
            t_l.append((xi[0] ** 3 / 6.0, third[0]))
            t_m.append((xi[0] ** 3 / 6.0, third[1]))
        self.data["ln(PI)"] = self._apply_terms(self.data["ln(PI)"], t_l)
        if not skip_mom:
            self.data["mom"] = self._apply_terms(self.data["mom"], t_m)

    def _temp_extrap_1(self, target_beta, cutoff=10.0, override=False, skip_mom=False):
        self._edge_assert(cutoff, override)
        try:
            d, dm = self._dB(skip_mom)
        except Exception:
            raise Exception("Unable to compute first derivative")
        self._taylor_update([target_beta - self.data["curr_beta"]], (d[None], dm[None]), None, skip_mom, False)

    def _temp_extrap_2(self, target_beta, cutoff=10.0, override=False, skip_mom=False):
        self._edge_assert(cutoff, override)
        try:
            d, dm = self._dB(skip_mom)
            d2, d2m = self._dB2(skip_mom)
        except Exception:
            raise Exception("Unable to compute derivatives")
        self._taylor_update([target_beta - self.data["curr_beta"]], (d[None], dm[None]), (d2[None, None], d2m[None, None]),
                            skip_mom, False)

    def _temp_extrap_3(self, target_beta, cutoff=10.0, override=False, skip_mom=False):
        self._edge_assert(cutoff, override)
        try:
            d, dm = self._dB(skip_mom)
            d2, d2m = self._dB2(skip_mom)
            d3, d3m = self._dB3(skip_mom)
        except Exception:
            raise Exception("Unable to compute derivatives")
        self._taylor_update([target_beta - self.data["curr_beta"]], (d[None], dm[None]), (d2[None, None], d2m[None, None]),
                            skip_mom, False, third=(d3, d3m))

    def temp_extrap(self, target_beta, order=1, cutoff=10.0, override=False, clone=True, skip_mom=False):
        """Temperature extrapolation (GH:670-740)."""
        self._check_not_extrapolated(check_dmu=False)
        self._check_order(order, skip_mom)
        tmp_hist = copy.deepcopy(self) if clone else self
        tmp_hist.normalize()
        fn = {1: tmp_hist._temp_extrap_1, 2: tmp_hist._temp_extrap_2, 3: tmp_hist._temp_extrap_3}.get(order)
        if fn is None:
            raise Exception("No implementation for temperature extrapolation of order " + str(order))
        try:
            fn(target_beta, cutoff, override, skip_mom)
        except Exception as e:
            raise Exception("Unable to extrapolate in temperature: " + str(e))
        tmp_hist.data["curr_beta"] = target_beta
        tmp_hist.normalize()
        return tmp_hist

    def _dmu_extrap_1(self, target_dmu, cutoff=10.0, override=False, skip_mom=False):
        self._edge_assert(cutoff, override)
        try:
            d, dm = self._dMU(skip_mom)
        except Exception:
            raise Exception("Unable to compute first derivative")
        xi = np.asarray(target_dmu, dtype=np.float64) - (self.data["curr_mu"][1:] - self.data["curr_mu"][0])
        self._taylor_update(list(xi), (d, dm), None, skip_mom, False)

    def _dmu_extrap_2(self, target_dmu, cutoff=10.0, override=False, skip_mom=False):
        self._edge_assert(cutoff, override)
        try:
            d, dm = self._dMU(skip_mom)
            H, Hm = self._dMU2(skip_mom)
        except Exception:
            raise Exception("Unable to compute derivatives")
        xi = np.asarray(target_dmu, dtype=np.float64) - (self.data["curr_mu"][1:] - self.data["curr_mu"][0])
        self._taylor_update(list(xi), (d, dm), (H, Hm), skip_mom, False)

    def dmu_extrap(self, target_dmu, order=1, cutoff=10.0, override=False, clone=True, skip_mom=False):
        """Delta-mu extrapolation (GH:742-811)."""
        target_dmu = np.asarray(target_dmu, dtype=np.float64)
        assert len(target_dmu) == self.data["nspec"] - 1, "Must specify delta mu for all components 2-N"
        self._check_not_extrapolated(check_beta=False)
        self._check_order(order, skip_mom)
        tmp_hist = copy.deepcopy(self) if clone else self
        tmp_hist.normalize()
        fn = {1: tmp_hist._dmu_extrap_1, 2: tmp_hist._dmu_extrap_2}.get(order)
        if fn is None:
            raise Exception("No implementation for dMu extrapolation of order " + str(order))
        try:
            fn(target_dmu, cutoff, override, skip_mom)
        except Exception as e:
            raise Exception("Unable to extrapolate in dMu: " + str(e))
        tmp_hist.data["curr_mu"][1:] = tmp_hist.data["curr_mu"][0] + target_dmu
        tmp_hist.normalize()
        return tmp_hist

    def _xi(self, target_beta, target_dmu):
        xi = np.zeros(self.data["nspec"], dtype=np.float64)
        xi[0] = target_beta - self.data["curr_beta"]
        xi[1:] = np.asarray(target_dmu, dtype=np.float64) - (self.data["curr_mu"][1:] - self.data["curr_mu"][0])
        return xi

    def _temp_dmu_extrap_1(self, target_beta, target_dmu, cutoff=10.0, override=False, skip_mom=False):
        self._edge_assert(cutoff, override)
        try:
            grad = self._dBMU(skip_mom)
        except Exception:
            raise Exception("Unable to compute first derivative")
        self._taylor_update(list(self._xi(target_beta, target_dmu)), grad, None, skip_mom, False)

    def _temp_dmu_extrap_2(self, target_beta, target_dmu, cutoff=10.0, override=False, skip_mom=False, first_order_mom=False):
        self._edge_assert(cutoff, override)
        try:
            grad = self._dBMU(skip_mom)
            hess = self._dBMU2(skip_mom)
        except Exception:
            raise Exception("Unable to compure derivatives")
        self._taylor_update(list(self._xi(target_beta, target_dmu)), grad, hess, skip_mom, first_order_mom)

    def temp_dmu_extrap(self, target_beta, target_dmu, order=1, cutoff=10.0, override=False, clone=True, skip_mom=False,
                        first_order_mom=False):
        """Simultaneous temperature and delta-mu extrapolation (GH:889-966)."""
        self._check_not_extrapolated(check_dmu=False)
        target_dmu = np.asarray(target_dmu, dtype=np.float64)
        assert len(target_dmu) == self.data["nspec"] - 1, "Must specify delta mu for all components 2-N"
        self._check_not_extrapolated(check_beta=False)
        self._check_order(order, skip_mom)
        tmp_hist = copy.deepcopy(self) if clone else self
        tmp_hist.normalize()
        try:
            if order == 1:
                tmp_hist._temp_dmu_extrap_1(target_beta, target_dmu, cutoff, override, skip_mom)
            elif order == 2:
                tmp_hist._temp_dmu_extrap_2(target_beta, target_dmu, cutoff, override, skip_mom, first_order_mom)
            else:
                raise Exception("No implementation for temperature + dMu extrapolation of order " + str(order))
        except Exception as e:
            if str(e).startswith("No implementation"):
                raise
            raise Exception("Unable to extrapolate : " + str(e))
        tmp_hist.data["curr_beta"] = target_beta
        tmp_hist.data["curr_mu"][1:] = copy.copy(tmp_hist.data["curr_mu"][0] + target_dmu)
        tmp_hist.normalize()
        return tmp_hist

    def temp_dmu_extrap_multi(self, target_betas, target_dmus, order=1, cutoff=10.0, override=False, skip_mom=False,
                              first_order_mom=False):
        """(beta x dMu) grid of extrapolated histograms (GH:813-887, 968-1043, 1093-1180): derivatives are built
        once, every grid cell is one fused device update + normalisation; failed cells are None."""
        self._check_not_extrapolated(check_dmu=False)
        target_betas = np.asarray(target_betas, dtype=np.float64)
        target_dmus = [np.asarray(t, dtype=np.float64) for t in target_dmus]
        for t in target_dmus:
            assert len(t) == self.data["nspec"] - 1, "Must specify delta mu for all components 2-N"
        self._check_not_extrapolated(check_beta=False)
        self._check_order(order, skip_mom)
        if order not in (1, 2):
            raise Exception("No implementation for temperature + dMu extrapolation of order " + str(order))
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
            for td in target_dmus:
                try:
                    clone = copy.deepcopy(self)
                    clone._taylor_update(list(self._xi(tb, td)), grad, hess, skip_mom, first_order_mom)
                    clone.data["curr_beta"] = copy.copy(tb)
                    clone.data["curr_mu"][1:] = copy.copy(clone.data["curr_mu"][0] + td)
                    clone.normalize()
                except Exception:
                    clone = None
                row.append(clone)
            hists.append(row)
        return hists

    # ------------------------------------------------------------------------------------------
    # batched entry points (new)
    # ------------------------------------------------------------------------------------------
    _MOMENT_ADDR = {"N": None, "N2": None, "U": [0, 0, 0, 0, 1], "U2": [0, 0, 0, 0, 2]}

    def _sel_rows(self, moments, order, extrap):
        """Rows (and their Taylor terms) of the quantities averaged inside the fused sweep."""
        ntot = self.data[self._op_key].astype(np.float64)
        kinds = []
        if extrap:
            kinds = [_lib.M_DB] + ([_lib.M_DD] if self.data["nspec"] == 2 else [])
        sel = []
        for name in moments:
            if name == "N":
                base, addr = ntot, None
            elif name == "N2":
                base, addr = ntot * ntot, None
            elif name in ("U", "U2"):
                addr = self._MOMENT_ADDR[name]
                base = self._m(addr)
            elif name.startswith("N") and name[1:].isdigit():  # "N1", "N2_" style species averages: N<i>
                addr = [int(name[1:]) - 1, 1, 0, 0, 0]
                base = self._m(addr)
            else:
                addr = list(name)
                base = self._m(addr)
            if not extrap:
                sel.append(base if name != "N" else "N")
                continue
            terms = [base]
            if addr is None or self.data["max_order"] < sum(addr[1::2]) + addr[4] + 1:
                terms += [np.zeros_like(ntot) for _ in kinds]  # sharp in N_tot, or order not stored: no first-order term
            else:
                terms.append(self._sg_dX_dB(addr, 0))
                if self.data["nspec"] == 2:
                    terms.append(self._sg_dX_dMU(0, addr))
            sel.append(terms)
        return sel, kinds

    def device_histogram(self, beta=None, dmu=None, order=1, moments=("N", "N2", "U"), cutoff=10.0, device=None):
        """Device-resident blob of the current histogram for batched sweeps (Taylor rows included when
        ``beta``/``dmu`` state points will be supplied)."""
        extrap = beta is not None or dmu is not None
        coef = ()
        src = self
        if extrap:
            self._check_not_extrapolated()
            src = copy.deepcopy(self)
            coef = src.taylor_rows(order)
        sel, kinds = src._sel_rows(list(moments), order, extrap)
        return src._device_hist(sel=sel, coef=coef, sel_kinds=kinds, cutoff=cutoff, device=device)

    def reweight_batch(self, mu, beta=None, dmu=None, order=1, moments=("N", "N2", "U"), grid=False, pmax=4, cutoff=10.0,
                       lanes=0, device=None, return_device=False, dh=None):
        """Reweight (+ Taylor-extrapolate) + phase split + thermo + is_safe for MANY state points at once.

        mu : array of mu_1 targets;  beta, dmu : optional arrays of target 1/kT and mu_2-mu_1.
        grid=False: flat lists (length-1 arrays broadcast); grid=True: outer product (mu x beta x dmu).
        Every state point starts from the CURRENT state of this histogram (no cumulative mutation; self is
        unchanged).  Returns a dict of NumPy arrays: status/code/safe, nphase, nmin, lnnorm, fe[S,pmax],
        avg[S,pmax,len(moments)], bounds[S,pmax,2], max_idx, min_idx (or the device SweepResult)."""
        if dh is None:
            dh = self.device_histogram(beta, dmu, order, moments, cutoff, device)
        if beta is None and dmu is None and not grid and not return_device and isinstance(mu, np.ndarray) and mu.size >= (1 << 16):
            # large host-resident mu sweep: chunked, double-buffered H2D -> kernel -> D2H pipeline
            out = {k: v.numpy() for k, v in dh.sweep_host(mu, pmax=pmax, lanes=lanes).items()}
            out["status"] = out["status"].view(np.uint32)
            out["code"] = (out["status"] & _lib.ST_CODE_MASK).astype(np.int32)
            out["safe"] = (out["status"] & _lib.ST_SAFE) != 0
            return out
        res = dh.sweep(mu, beta, dmu, grid=grid, pmax=pmax, lanes=lanes)
        return res if return_device else res.host()

    def find_phase_eq_batch(self, betas, mu_guess, dmu=None, order=1, lnZ_tol=1e-10, moments=("N", "N2", "U"), pmax=4,
                            cutoff=10.0, max_iter=200, device=None, return_device=False, continuation=None):
        """One coexistence solve per temperature in ``betas`` (and optional ``dmu``), all concurrently (K4).
        ``mu_guess``: scalar or one guess per temperature.  Returns the thermo records at coexistence plus
        'mu_coex', 'dfe' (signed residual F.E._i - F.E._j), 'iters' and 'converged' (code 0 and |dfe| <= lnZ_tol; a solve
        that ended on a jump of the free-energy difference has code 0, status bit ST_JUMP and converged False).
        ``continuation``: see ``DeviceHistogram.find_phase_eq`` (default: automatic for one cold guess and many
        temperatures -- a coarse subset is solved first and the other guesses are interpolated from its roots)."""
        betas = np.atleast_1d(np.asarray(betas, dtype=np.float64))
        moments = ["N"] + [m for m in moments if m != "N"]
        dh = self.device_histogram(betas, dmu, order, moments, cutoff, device)
        guess = np.broadcast_to(np.asarray(mu_guess, dtype=np.float64), betas.shape).copy()
        res = dh.find_phase_eq(guess, beta=betas, dmu=dmu, lnz_tol=lnZ_tol, max_iter=max_iter, pmax=pmax, continuation=continuation)
        if return_device:
            return res
        out = res.host()
        out["converged"] = (out["code"] == 0) & ((out["status"] & _lib.ST_JUMP) == 0)
        return out


if __name__ == "__main__":
    print("gc_hist (B200)")
    sys.exit(0)
