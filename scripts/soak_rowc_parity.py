#!/usr/bin/env python
"""Soak test: the row-combined Taylor-grid kernel (k_sweep_rowc) against the general evaluator on random two-species histograms --
sizes, extrema windows, noise levels, temperature / dmu ranges and extrapolation orders drawn at random.  The two paths form u in a
different order (rows combined per grid row vs eight separate terms), so a near-tie may legitimately order differently: mismatching
records are counted and the worst one is reported with its margin.  usage: soak_rowc_parity.py [trials] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import _lib, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
    t0 = time.time()
    tot = bad_int = bad_fe = 0
    for trial in range(trials):
        n = int(rng.integers(40, 1500))
        kind = int(rng.integers(0, 3))
        x = np.arange(n, dtype=float)
        if kind == 0:
            lnpi = synth.two_peak_lnpi(n, noise=float(rng.choice([0.0, 1e-4, 1e-3, 3e-2])), seed=int(rng.integers(1, 1000)))
        elif kind == 1:
            lnpi = np.cumsum(rng.normal(0.0, rng.uniform(0.05, 0.6), size=n))
        else:
            lnpi = -rng.uniform(0.2, 3.0) * x + 30.0 * np.sin(x / rng.uniform(20, 200)) + 1e-3 * rng.normal(size=n)
        mom = synth.two_comp_moments(n)
        smooth = int(rng.integers(1, 25))
        h = histogram.from_arrays(lnpi, mom, 1.0, [-3.0, -2.5], smooth)
        h.reweight(float(rng.uniform(-3.1, -2.8)))
        order = int(rng.integers(1, 3))
        wb, wd = float(rng.choice([0.01, 0.05, 0.2])), float(rng.choice([0.05, 0.3, 1.0]))
        nb, nd = int(rng.integers(2, 12)), int(rng.integers(512, 1400))
        betas, dmus = np.sort(rng.uniform(1 - wb, 1 + wb, nb)), np.sort(rng.uniform(0.5 - wd, 0.5 + wd, nd))
        pmax = int(rng.choice([2, 4, 8]))
        dh = h.device_histogram(beta=betas, dmu=dmus, order=order, moments=())
        st = dh.make_states(np.array([h.data["curr_mu"][0]]), betas, dmus, grid=True)
        a = dh.sweep(None, states=st, pmax=pmax, lanes=1).host()
        kern = _lib.last_kernel()
        b = dh.sweep(None, states=st, pmax=pmax, lanes=-1).host()
        assert kern == "k_sweep_rowc", kern
        S = st.n_states
        tot += S
        mism = np.zeros(S, dtype=bool)
        for k in ("code", "nphase", "nmin", "safe"):
            mism |= a[k] != b[k]
        ok = (a["code"] == 0) & ~mism
        P = a["nphase"]
        for k in ("max_idx", "min_idx", "bounds"):
            lim = a["nmin"] if k == "min_idx" else P
            pm = (np.arange(a[k].shape[1])[None, :] < lim[:, None]) & ok[:, None]
            d = (a[k] != b[k])
            d = d.reshape(S, a[k].shape[1], -1).any(-1) & pm
            mism |= d.any(1)
        ok &= ~mism
        mask = (np.arange(a["fe"].shape[1])[None, :] < P[:, None]) & ok[:, None]
        err = np.abs(a["fe"] - b["fe"]) / np.maximum(1.0, np.abs(b["fe"]))
        err[~mask] = 0.0
        nfe = int((err > 1e-10).sum())
        bad_int += int(mism.sum())
        bad_fe += nfe
        fastf = float(np.mean((a["status"] & 0x1000) != 0))
        print("trial %2d n=%4d kind=%d smooth=%2d order=%d grid %2dx%4d pmax=%d: ok %.3f fast %.3f  integer mismatches %d  fe>1e-10 %d  max fe err %.1e"
              % (trial, n, kind, smooth, order, nb, nd, pmax, float(np.mean(a["code"] == 0)), fastf, int(mism.sum()), nfe, float(err.max())), flush=True)
        if mism.any():
            k = int(np.where(mism)[0][0])
            print("   first mismatch: state", k, "codes", a["code"][k], b["code"][k], "nphase", a["nphase"][k], b["nphase"][k],
                  "max", a["max_idx"][k].tolist(), b["max_idx"][k].tolist(), "min", a["min_idx"][k].tolist(), b["min_idx"][k].tolist())
    print("state points %d  integer mismatches %d  fe mismatches %d  (%.0f s)" % (tot, bad_int, bad_fe, time.time() - t0))


if __name__ == "__main__":
    main()
