"""Drop-in replacement for FHMCAnalysis.moments.histogram.one_dim.ntot.gc_binary (reference file
moments/histogram/one_dim/ntot/gc_binary.pyx, "GB"): isopleth grids of binary mixtures from a set of
(mu_1, dmu_2) histograms.

``isopleth.make_grid_multi`` is re-designed for the GPU.  The reference loops over mu_1, re-weights every stored
histogram, Taylor-extrapolates it to each dmu_2 row, mixes neighbours and calls thermo() per cell (GB:243-288,
~100 ms per cell).  Here, for every dmu_2 ROW the two neighbouring histograms are extrapolated in (beta, dmu_2) and
blended ONCE (the Taylor terms and the mixing weights do not depend on mu_1, and the mu_1 dependence of the blend is
an exact reweighting at the target beta), which leaves one pure mu_1 sweep per row for the fused one-pass kernel.
"""
import bisect
import copy
import json

import numpy as np

from fhmcanalysis_b200 import _lib, engine
from . import gc_hist as gch

np.seterr(divide="raise", over="raise", invalid="raise", under="ignore")  # GB:27


def _find_left_right(ordered_dmu2, val, bound=False):
    """Indices of the stored dmu_2 values bracketing ``val`` (GB:31-81)."""
    tol = 1.0e-9
    n = len(ordered_dmu2)
    if val <= np.min(ordered_dmu2):
        return (0, 0) if bound else (-1, -1)
    if val >= np.max(ordered_dmu2):
        return (n - 1, n - 1) if bound else (n, n)
    if np.any([np.isclose(val, x) for x in ordered_dmu2]):
        x = np.where(np.abs(ordered_dmu2 - val) < tol)[0]
        if len(x) != 1:
            raise Exception("dmu2 values repeat, " + str(x) + " , " + str(ordered_dmu2) + " , " + str(val))
        return int(x[0]), int(x[0])
    left = bisect.bisect(list(ordered_dmu2), val) - 1
    return left, left + 1


def _get_most_stable_phase(hist):
    """Index of the phase with the lowest F.E./kT after thermo() (GB:83-107)."""
    fe = {p: hist.data["thermo"][p]["F.E./kT"] for p in hist.data["thermo"]}
    return sorted(fe.items(), key=lambda kv: kv[1])[0][0]


class isopleth(object):
    """Isopleths from a series of (mu1, dMu2) histograms (GB:109-523)."""

    def __init__(self, histograms, beta_target, order=2):
        if not isinstance(histograms, (list, np.ndarray)):
            raise Exception("Expects an array of histograms to construct isopleths")
        for h in histograms:
            if not isinstance(h, gch.histogram):
                raise Exception("Expects a vector of histograms to construct isopleths")
        if beta_target <= 0:
            raise Exception("Illegal beta, cannot construct isopleths")
        if order < 1 or order > 2:
            raise Exception("Illegal order, cannot construct isopleths")
        self.meta = {"beta": beta_target, "tol": 1.0e-9, "order": order, "cutoff": 10.0}
        self.clear()
        for h in histograms:
            if h.data["nspec"] != 2:
                raise Exception("Component mismatch in isopleth generation")
        dummy = {}
        t_ = -1.0
        for h in histograms:
            if len(h.data["curr_mu"]) != 2:
                raise Exception("Only expects 2 chemical potentials, one for each component, cannot construct isopleth")
            dmu2 = float(h.data["curr_mu"][1] - h.data["curr_mu"][0])
            dummy[dmu2] = h
            if t_ > 0:
                if abs(h.metadata["beta_ref"] - t_) > self.meta["tol"]:
                    raise Exception("Expects all histograms to be performed at the same temperature")
            else:
                if h.metadata["beta_ref"] <= 0:
                    raise Exception("Illegal temperature in histograms")
                t_ = h.metadata["beta_ref"]
        srt = sorted(dummy.items(), key=lambda kv: kv[0])
        self.data["dmu2"] = np.array([x[0] for x in srt])
        self.data["histograms"] = [copy.deepcopy(x[1]) for x in srt]

    def clear(self):
        self.data = {}

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _check_grid_args(mu1_bounds, dmu2_bounds, delta):
        if not isinstance(mu1_bounds, (list, np.ndarray, tuple)): raise Exception("Expects an array of mu1 bounds to construct isopleths")
        if not isinstance(dmu2_bounds, (list, np.ndarray, tuple)): raise Exception("Expects an array of dmu2 bounds to construct isopleths")
        if not isinstance(delta, (list, np.ndarray, tuple)): raise Exception("Expects an array of delta mu values to construct isopleths")
        if len(mu1_bounds) != 2: raise Exception("mu1_bound error in constructing isopleths")
        if len(dmu2_bounds) != 2: raise Exception("dmu2_bound error in constructing isopleths")
        if len(delta) != 2: raise Exception("delta error in constructing isopleths")
        if mu1_bounds[1] <= mu1_bounds[0]: raise Exception("mu1_bound error in constructing isopleths")
        if dmu2_bounds[1] <= dmu2_bounds[0]: raise Exception("dmu2_bound error in constructing isopleths")
        if delta[0] <= 0: raise Exception("delta error in constructing isopleths")
        if delta[1] <= 0: raise Exception("delta error in constructing isopleths")

    def _alloc_grid(self, mu1_bounds, dmu2_bounds, delta):
        nx = int(np.ceil((mu1_bounds[1] - mu1_bounds[0]) / delta[0]) + 1)
        ny = int(np.ceil((dmu2_bounds[1] - dmu2_bounds[0]) / delta[1]) + 1)
        mu1_v = np.linspace(mu1_bounds[0], mu1_bounds[1], nx)
        dmu2_v = np.linspace(dmu2_bounds[0], dmu2_bounds[1], ny)
        self.data["X"], self.data["Y"] = np.meshgrid(mu1_v, dmu2_v)
        for k in ("Z", "density", "F.E./kT"):
            self.data[k] = np.zeros(self.data["X"].shape, dtype=np.float64)
        return mu1_v, dmu2_v

    def _row_source(self, h, dmu2, want):
        """Histogram ``h`` extrapolated to (beta_target, dmu2) with the mu_1 dependence factored out:
        returns rows R such that the extrapolated, UN-normalised lnPI at mu_1 is R['lnpi'] + beta*mu_1*N and the
        extrapolated <N_1>(N), <N_2>(N) arrays are R['n1'], R['n2'] (independent of mu_1)."""
        order = self.meta["order"]
        beta = self.meta["beta"]
        if h.data["max_order"] < order + 1:
            raise Exception("Maximum order stored in simulation not high enough to calculate this order of extrapolation")
        if np.abs(h.metadata["beta_ref"] - h.data["curr_beta"]) > 1.0e-6:
            raise Exception("Cannot extrapolate the same histogram class twice")
        beta_ref = h.data["curr_beta"]
        d0 = h.data["curr_mu"][1] - h.data["curr_mu"][0]
        dB, dD = beta - beta_ref, dmu2 - d0
        mono = {_lib.M_DB: dB, _lib.M_DD: dD, _lib.M_DB2: 0.5 * dB * dB, _lib.M_DBDD: dB * dD, _lib.M_DD2: 0.5 * dD * dD}
        N = h.data["ntot"].astype(np.float64)
        terms = [(1.0, np.asarray(h.data["ln(PI)"], dtype=np.float64))]
        for kind, row in h.taylor_rows(order):
            if kind == _lib.M_DB_MU1:
                continue  # dB*mu_1*N joins beta_ref*mu_1*N: the sweep reweights at the TARGET beta
            terms.append((mono[kind], row))
        terms.append((-beta_ref * h.data["curr_mu"][0], N))
        out = {"lnpi": engine.axpy_rows([t[1] for t in terms], [t[0] for t in terms]) if len(terms) <= _lib.MAX_TERMS
               else gch.histogram._apply_terms(terms[0][1], terms[1:])}
        for name, addr in want:
            t = [(1.0, h._m(addr)), (dB, h._sg_dX_dB(addr, 0)), (dD, h._sg_dX_dMU(0, addr))]
            if order == 2:
                t += [(0.5 * dB * dB, h._sg_d2X_dB2(addr, 0)), (0.5 * dD * dD, h._sg_d2X_dMU2(0, 0, addr))]
                z = h._mom_prod([1, 1, 0, 0, 0], addr)
                f = h._m(z) - h._m([1, 1, 0, 0, 0]) * h._m(addr)
                t.append((dB * dD, h.data["curr_beta"] * h._sg_df_dB(([1, 1, 0, 0, 0], 0), (addr, 0)) + f))
            out[name] = engine.axpy_rows([x[1] for x in t], [x[0] for x in t])
        return out

    def make_grid_multi(self, mu1_bounds, dmu2_bounds, delta, m=2.5, pmax=8):
        """x_1 of the most stable phase on a (mu_1, dmu_2) grid (GB:173-290).  Returns Z, (X, Y); density and
        F.E./kT grids are left in self.data like the reference.  Cells the reference would skip (unsafe edge,
        failed extrema) stay 0."""
        self._check_grid_args(mu1_bounds, dmu2_bounds, delta)
        mu1_v, dmu2_v = self._alloc_grid(mu1_bounds, dmu2_bounds, delta)
        hs = self.data["histograms"]
        beta = self.meta["beta"]
        want = (("n1", [0, 1, 0, 0, 0]), ("n2", [1, 1, 0, 0, 0]))
        # edge test of temp_dmu_extrap_multi (GH:1003/1133, override=False) per stored histogram and mu_1:
        # max(lnPI) - cutoff > lnPI[-1] on the re-weighted histogram
        edge_ok = np.zeros((len(hs), len(mu1_v)), dtype=bool)
        for j, h in enumerate(hs):
            try:
                r = h._device_hist(cutoff=self.meta["cutoff"]).sweep(mu1_v, pmax=1, complete=True).host()
                edge_ok[j] = r["safe"] & (r["code"] == 0)
            except Exception:
                edge_ok[j] = False
        for i, dmu2 in enumerate(dmu2_v):
            left, right = _find_left_right(self.data["dmu2"], dmu2, True)
            dl = abs(self.data["dmu2"][left] - dmu2) ** m
            dr = abs(self.data["dmu2"][right] - dmu2) ** m
            if dl + dr < 1.0e-9:
                assert left == right, "Unknown mixing distance error"
                wl = wr = 1.0
            else:
                wl, wr = dr / (dr + dl), dl / (dr + dl)
            try:
                L = self._row_source(hs[left], dmu2, want)
                R = L if right == left else self._row_source(hs[right], dmu2, want)
            except Exception as e:
                print("Error during extrapolation : " + str(e))
                continue
            # mix (GH:244-252): weighted blend over the common range, the longer histogram supplies the tail
            nl, nr = len(L["lnpi"]), len(R["lnpi"])
            longer, nmix = (L, nr) if nl >= nr else (R, nl)
            rows = {}
            for k in ("lnpi", "n1", "n2"):
                rows[k] = np.array(longer[k], dtype=np.float64)
                rows[k][:nmix] = (L[k][:nmix] * wl + wr * R[k][:nmix]) / (wl + wr)
            href = hs[left] if nl >= nr else hs[right]
            N = href.data["ntot"].astype(np.float64)
            dh = engine.DeviceHistogram(rows["lnpi"], N, beta, 0.0, 0.0, smooth=max(int(href.metadata["smooth"]), 1),
                                        cutoff=self.meta["cutoff"], sel=[rows["n1"], rows["n2"]])
            r = dh.sweep_auto(mu1_v, pmax=pmax).host()
            vol = href.data["volume"]
            # the whole row of cells at once: most stable phase = lowest F.E./kT among the nphase[jx] phases (GB:83-107);
            # cells whose x1 would divide by zero / be invalid stay 0 like the reference's FloatingPointError branch
            ok = edge_ok[left] & edge_ok[right] & (r["code"] == 0) & r["safe"].astype(bool)
            nph = np.maximum(r["nphase"].astype(np.int64), 1)
            fe = np.where(np.arange(r["fe"].shape[1])[None, :] < nph[:, None], r["fe"], np.inf)
            p = np.argmin(fe, axis=1)
            jj = np.arange(len(mu1_v))
            n1, n2 = r["avg"][jj, p, 0], r["avg"][jj, p, 1]
            with np.errstate(divide="ignore", invalid="ignore"):
                x1 = n1 / (n1 + n2)
            ok &= np.isfinite(x1) & ((n1 + n2) != 0.0)
            self.data["Z"][i, ok] = x1[ok]
            self.data["density"][i, ok] = (n1[ok] + n2[ok]) / vol
            self.data["F.E./kT"][i, ok] = r["fe"][jj, p][ok]
        return self.data["Z"], (self.data["X"], self.data["Y"])

    # ------------------------------------------------------------------------------------------
    def get_hist(self, mu1, dmu2, m=2.5):
        """Histogram at (mu1, dmu2) by reweighting, extrapolating and mixing neighbours (GB:292-353).  Like the
        reference this re-weights the stored histograms in place."""
        left, right = _find_left_right(self.data["dmu2"], dmu2, False)
        hs = self.data["histograms"]
        tgt = np.array([dmu2], dtype=np.float64)
        if left == right:
            h_l = hs[0] if left < 0 else (hs[-1] if left == len(self.data["dmu2"]) else hs[left])
            try:
                h_l.reweight(mu1)
                return h_l.temp_dmu_extrap(self.meta["beta"], tgt, self.meta["order"], self.meta["cutoff"], False, True, False)
            except Exception as e:
                raise Exception("Unable to get histogram : " + str(e))
        h_l, h_r = hs[left], hs[right]
        try:
            h_l.reweight(mu1)
            h_l = h_l.temp_dmu_extrap(self.meta["beta"], tgt, self.meta["order"], self.meta["cutoff"], False, True, False)
            h_r.reweight(mu1)
            h_r = h_r.temp_dmu_extrap(self.meta["beta"], tgt, self.meta["order"], self.meta["cutoff"], False, True, False)
        except Exception as e:
            raise Exception("Unable to get histogram : " + str(e))
        dl = abs(self.data["dmu2"][left] - dmu2) ** m
        dr = abs(self.data["dmu2"][right] - dmu2) ** m
        return h_l.mix(h_r, [dr / (dr + dl), dl / (dr + dl)])

    def make_grid(self, mu1_bounds, dmu2_bounds, delta, m=2.5):
        """Cell-by-cell variant of the grid (GB:355-476): every cell goes through get_hist -> thermo -> is_safe."""
        self._check_grid_args(mu1_bounds, dmu2_bounds, delta)
        self._alloc_grid(mu1_bounds, dmu2_bounds, delta)
        X, Y = self.data["X"], self.data["Y"]
        for i in range(X.shape[0]):
            for j in range(X.shape[1]):
                mu1, dmu2 = X[i, j], Y[i, j]
                try:
                    h = self.get_hist(mu1, dmu2, m)
                    h.thermo()
                    if not h.is_safe():
                        raise Exception("extrapolated ln(PI) in histogram is not safe to use")
                    p = _get_most_stable_phase(h)
                    self.data["Z"][i, j] = h.data["thermo"][p]["x1"]
                    self.data["density"][i, j] = h.data["thermo"][p]["density"]
                    self.data["F.E./kT"][i, j] = h.data["thermo"][p]["F.E./kT"]
                except Exception as e:
                    print("Error at (mu_1,dmu_2) = (" + str(mu1) + "," + str(dmu2) + ") : " + str(e) + ", continuing on...")
        return self.data["Z"], (X, Y)

    def dump(self, fname):
        """Grids to JSON (GB:478-497; same keys and layout)."""
        info = {"mu_1": self.data["X"].tolist(), "dmu_2": self.data["Y"].tolist(), "x_1": self.data["Z"].tolist(),
                "density": self.data["density"].tolist(), "F.E./kT": self.data["F.E./kT"].tolist()}
        with open(fname, "w") as f:
            json.dump(info, f, sort_keys=True, indent=4)

    def load(self, fname):
        """Grids from JSON (GB:499-523)."""
        with open(fname, "r") as f:
            info = json.load(f)
        self.data["X"] = np.array(info["mu_1"], dtype=np.float64)
        self.data["Y"] = np.array(info["dmu_2"], dtype=np.float64)
        self.data["Z"] = np.array(info["x_1"], dtype=np.float64)
        self.data["density"] = np.array(info["density"], dtype=np.float64)
        self.data["F.E./kT"] = np.array(info["F.E./kT"], dtype=np.float64)
        for k in ("Y", "Z", "density", "F.E./kT"):
            assert self.data["X"].shape == self.data[k].shape, "Shape mismatch in " + fname

    def zoom(self, factor, order=3, inplace=False):
        """Spline-resample the result grids (GB:525-564; scipy post-processing of results, host side)."""
        import scipy.ndimage
        z = {k: scipy.ndimage.zoom(self.data[k], factor, order=order) for k in ("X", "Y", "Z", "density", "F.E./kT")}
        if inplace:
            self.data.update(z)
        return z["Z"], (z["X"], z["Y"]), z["density"], z["F.E./kT"]


def combine_isopleth_grids(mu1_arrays, dmu2_arrays, x1_arrays, rho_arrays=None, fe_arrays=None):
    """Concatenate isopleth grids along mu_1 after trimming duplicated columns (GB:705-820; tested by the reference
    in unittests/moments_histogram_one_dim_gc_ntot_isopleth.py:27-91).  Pure array bookkeeping on result grids."""
    for name, arr in (("mu1_arrays", mu1_arrays), ("dmu2_arrays", dmu2_arrays), ("x1_arrays", x1_arrays)):
        if not isinstance(arr, (list, np.ndarray, tuple)):
            raise Exception("Expects an array of " + name + " to combine isopleths")
    if not (len(mu1_arrays) == len(dmu2_arrays) == len(x1_arrays)):
        raise Exception("Must specify one mu_1, dmu_2, and x_1 for each isopleth")
    extras = []
    for arr, what in ((rho_arrays, "density"), (fe_arrays, "free energy")):
        if arr is not None:
            if not isinstance(arr, (list, np.ndarray, tuple)):
                raise Exception("Expects an array to combine isopleths")
            if len(arr) != len(mu1_arrays):
                raise Exception("Must specify one " + what + " for each isopleth")
            extras.append(arr)
    stacks = [mu1_arrays, dmu2_arrays, x1_arrays] + extras
    for i in range(len(mu1_arrays)):
        for st in stacks[1:]:
            if mu1_arrays[i].shape != st[i].shape:
                raise Exception("Each set of isopleth grids must have the same size")
    for i in range(len(mu1_arrays) - 1):
        for st in stacks:
            if st[i].shape[0] != st[i + 1].shape[0]:
                raise Exception("dmu2 dimension not aligned")
    order = sorted(range(len(mu1_arrays)), key=lambda k: np.min(mu1_arrays[k]))
    out = [copy.copy(st[order[0]]) for st in stacks]
    dmu2_ref = dmu2_arrays[order[0]][:, 1]
    for a, b in zip(order[:-1], order[1:]):
        if not np.all(np.abs(dmu2_arrays[b][:, 0] - dmu2_ref) < 1.0e-9):
            raise Exception("dmu2 dimension not aligned")
        mu1_right = mu1_arrays[b][0, :]
        max_mu1_left = np.max(mu1_arrays[a][0, :])
        ncols = bisect.bisect_left(list(mu1_right), max_mu1_left)
        if mu1_right[ncols] == max_mu1_left:
            ncols += 1
        out = [np.concatenate((o, st[b][:, ncols:]), axis=1) for o, st in zip(out, stacks)]
    X, Y, Z = out[0], out[1], out[2]
    if len(out) == 3:
        return Z, (X, Y)
    if len(out) == 4:
        return Z, (X, Y), out[3]
    return Z, (X, Y), out[3], out[4]
