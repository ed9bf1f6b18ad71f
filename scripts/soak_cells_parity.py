#!/usr/bin/env python
"""Soak test: the tilt-cell kernel (k_sweep_cell + the indexed table walk behind it) against the general evaluator on random
histograms -- sizes, extrema windows, noise levels, N spacings, reference conditions, tilt ranges and orderings of mu drawn at random
(soak_prod_parity.py's generator).  Integers must agree exactly, fe / averages to 1e-10 relative to max(|x|, 1).
usage: soak_cells_parity.py [trials] [seed]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from fhmcanalysis_b200 import _lib, engine  # noqa: E402
from soak_prod_parity import random_hist  # noqa: E402


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4321)
    bad, points, by_cells, worst = 0, 0, 0, 0.0
    t0 = time.time()
    for trial in range(trials):
        lnpi, N, smooth = random_hist(rng)
        span = float(rng.choice([0.005, 0.02, 0.2, 2.0]))
        centre = float(rng.choice([0.0, 0.0, rng.uniform(-0.5, 0.5)]))
        beta, mu_ref = float(rng.choice([1.0, 0.7, 1.3])), float(rng.choice([0.0, -1.5]))
        S = int(rng.integers(70000, 160000))
        mus = mu_ref + centre + (np.sort(rng.uniform(-span, span, size=S)) if rng.integers(0, 2) else rng.uniform(-span, span, size=S))
        nsel = int(rng.integers(0, 3))
        sel = [["N", N * N], [np.sqrt(N + 1.0)], []][2 - nsel] if nsel < 2 else ["N", N * N]
        dh = engine.DeviceHistogram(lnpi, N, beta, mu_ref, smooth=smooth, sel=sel)
        dh.CELLS_MIN_STATES = 1
        pmax = int(rng.choice([4, 8]))
        c = dh.sweep_compact(mus, pmax=pmax)
        kern = _lib.last_kernel()
        g = dh.sweep(mus, pmax=pmax, lanes=-1).host()
        st = c["status"].cpu().numpy().astype(np.int64)
        msg = None
        if not np.array_equal(st & 0xFF, g["code"]):
            msg = "code"
        ok = g["code"] == 0
        if msg is None and not np.array_equal((st & 0x100) != 0, (g["status"].astype(np.int64) & 0x100) != 0):
            msg = "safe"
        P = g["nphase"]
        if msg is None and not np.array_equal(c["nphase"].cpu().numpy()[ok], P[ok]):
            msg = "nphase"
        fe, bd = c["fe"].cpu().numpy(), c["bounds"].cpu().numpy()
        av = c["avg"].cpu().numpy() if len(sel) else None
        w = 0.0
        for p in range(pmax):
            live = ok & (P > p)
            if msg is None and not np.array_equal(bd[live, p], g["bounds"][live, p]):
                msg = "bounds"
            if live.any():
                w = max(w, float(np.max(np.abs(fe[live, p] - g["fe"][live, p]) / np.maximum(1.0, np.abs(g["fe"][live, p])))))
                if av is not None:
                    w = max(w, float(np.max(np.abs(av[live, p] - g["avg"][live, p]) / np.maximum(1.0, np.abs(g["avg"][live, p])))))
        if msg is None and not w <= 1e-10:
            msg = "values %.3g" % w
        frac = float(c["path"].double().mean())
        points += S
        by_cells += int(round(frac * S))
        worst = max(worst, w)
        if msg is not None or os.environ.get("SOAK_VERBOSE"):
            bad += msg is not None
            print(json.dumps({"trial": trial, "n": len(lnpi), "smooth": smooth, "span": span, "centre": centre, "beta": beta, "mu_ref": mu_ref, "S": S,
                              "nsel": len(sel), "pmax": pmax, "kernel": kern, "cell_fraction": frac, "ok_fraction": float(ok.mean()), "worst": w,
                              "mismatch": msg}), flush=True)
    print(json.dumps({"trials": trials, "mismatching_runs": bad, "state_points": points, "evaluated_by_the_cells": by_cells, "worst_rel": worst,
                      "seconds": time.time() - t0}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
