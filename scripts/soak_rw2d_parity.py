#!/usr/bin/env python
"""Soak test of the product-form 2-D reweighting kernel against the exp-per-bin kernel on random surfaces (sizes, ragged
supports, -inf holes, steep rows, op2 spacings, tilts up to the documented limit).  usage: soak_rw2d_parity.py [trials] [seed]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine  # noqa: E402


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    worst = 0.0
    for trial in range(trials):
        n1, n2 = int(rng.integers(3, 300)), int(rng.integers(8, 700))
        i, j = np.arange(n1)[:, None], np.arange(n2)[None, :]
        steep = rng.choice([0.01, 0.3, 3.0])
        lnpi = -steep * np.abs(j - rng.uniform(0, n2)) - 0.02 * (i - rng.uniform(0, n1)) ** 2 + rng.normal(0, rng.choice([0, 1e-3, 0.5]), size=(n1, n2))
        lo = rng.integers(0, max(1, n2 // 4), size=n1)
        hi = rng.integers(n2 // 2, n2 + 1, size=n1)
        bounds = np.stack([lo, hi], axis=1).astype(np.int32)
        for r in range(n1):
            lnpi[r, :lo[r]] = -np.inf
            lnpi[r, hi[r]:] = -np.inf
        holes = rng.random((n1, n2)) < 0.01
        lnpi[holes] = -np.inf
        d2 = float(rng.choice([0.5, 1.0, 2.0]))
        op1, op2 = rng.uniform(0, 3) + np.arange(n1, dtype=float), rng.uniform(0, 5) + d2 * np.arange(n2)
        nprop = int(rng.integers(0, 3))
        props = rng.normal(size=(nprop, n1, n2)) if nprop else None
        S = int(rng.integers(64, 3000))
        amax = min(2.0, 299.0 / (op2[-1] - op2[0]))
        a1, a2 = rng.uniform(-1.0, 1.0, S), rng.uniform(-amax, amax, S)
        p = engine.reweight_2d(lnpi, bounds, op1, op2, a1, a2, props, product=True)
        e = engine.reweight_2d(lnpi, bounds, op1, op2, a1, a2, props, product=False)
        fin = np.isfinite(e).all(axis=1)
        err = np.max(np.abs(p[fin] - e[fin]) / np.maximum(1.0, np.abs(e[fin]))) if fin.any() else 0.0
        same_nan = np.array_equal(np.isfinite(p), np.isfinite(e))
        worst = max(worst, float(err))
        if err > 1e-10 or not same_nan:
            bad += 1
            print(json.dumps({"trial": trial, "n1": n1, "n2": n2, "steep": float(steep), "nprop": nprop, "err": float(err), "same_nan": bool(same_nan)}), flush=True)
    print(json.dumps({"trials": trials, "mismatching": bad, "worst_rel": worst}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
