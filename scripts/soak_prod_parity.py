#!/usr/bin/env python
"""Soak test: product-form sweep kernels (one and two state points per thread) against the generic evaluator on random
histograms -- sizes, extrema windows, noise levels, N spacings and tilt ranges drawn at random.  Integers must agree
exactly, fp64 to 1e-10.  usage: soak_prod_parity.py [trials] [seed]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine  # noqa: E402


def random_hist(rng):
    n = int(rng.integers(9, 1600))
    kind = rng.integers(0, 4)
    x = np.arange(n, dtype=float)
    if kind == 0:      # random walk
        lnpi = np.cumsum(rng.normal(0.0, rng.uniform(0.05, 1.0), size=n))
    elif kind == 1:    # two or three Gaussian peaks + noise
        lnpi = np.full(n, -np.inf)
        for _ in range(int(rng.integers(2, 4))):
            c, w, h = rng.uniform(0, n), rng.uniform(0.03, 0.2) * n, rng.uniform(-3, 0)
            lnpi = np.logaddexp(lnpi, -(x - c) ** 2 / (2 * w * w) + h)
        lnpi = lnpi + rng.choice([0.0, 1e-6, 1e-3, 5e-2]) * rng.normal(size=n)
    elif kind == 2:    # steep, square-well like
        lnpi = -rng.uniform(1.0, 5.0) * x + 40.0 * np.sin(x / rng.uniform(20, 200)) + 1e-3 * rng.normal(size=n)
    else:              # plateaus and exact ties (integers)
        lnpi = np.round(np.cumsum(rng.normal(0, 0.7, size=n))).astype(float)
    dN = float(rng.choice([0.5, 1.0, 2.0]))
    N = float(rng.integers(0, 5)) + dN * x
    return lnpi, N, int(rng.integers(1, 31))


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
    only = int(os.environ.get("SOAK_ONLY", "-1"))
    bad = 0
    t0 = time.time()
    for trial in range(trials):
        lnpi, N, smooth = random_hist(rng)
        span = rng.choice([0.02, 0.2, 2.0, 8.0])
        for rec, S in ((3, 155000), (2, 6000)):
            dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=smooth, sel=["N", N * N])
            dh.use_recurrence = rec
            dh.ensure_hull()
            mus = rng.uniform(-span, span, size=S)
            if only >= 0 and trial != only:
                if dh.desc.mu_recurrence != rec:
                    break
                continue
            a = dh.sweep_auto(mus, pmax=8, lanes=1).host()
            b = dh.sweep_auto(mus, pmax=a["fe"].shape[1], lanes=-1).host()
            msg = None
            for k in ("code", "nphase", "nmin", "safe"):
                if not np.array_equal(a[k], b[k]):
                    msg = k
            ok = a["code"] == 0
            P = a["nphase"]
            if msg is None:
                for k in ("max_idx", "min_idx", "bounds"):
                    lim = a["nmin"] if k == "min_idx" else P
                    pm = (np.arange(a[k].shape[1])[None, :] < lim[:, None]) & ok[:, None]
                    if not np.array_equal(a[k][pm], b[k][pm]):
                        msg = k
            if msg is None:
                mask = (np.arange(a["fe"].shape[1])[None, :] < P[:, None]) & ok[:, None]
                if not np.allclose(a["fe"][mask], b["fe"][mask], rtol=1e-10, atol=1e-10):
                    msg = "fe"
                elif not np.allclose(a["avg"][mask], b["avg"][mask], rtol=1e-10, atol=1e-10):
                    msg = "avg"
            if only >= 0:
                mask = (np.arange(a["fe"].shape[1])[None, :] < P[:, None]) & ok[:, None]
                err = np.abs(a["fe"] - b["fe"]) / np.maximum(1.0, np.abs(b["fe"]))
                err[~mask] = 0
                w = np.unravel_index(np.argmax(err), err.shape)
                print(json.dumps({"rec": rec, "worst_rel": float(err[w]), "mu": float(mus[w[0]]), "phase": int(w[1]), "nphase": int(P[w[0]]),
                                  "fe_fast": float(a["fe"][w]), "fe_generic": float(b["fe"][w]), "status_fast": int(a["status"][w[0]]),
                                  "status_generic": int(b["status"][w[0]]), "bounds": a["bounds"][w[0], :P[w[0]]].tolist(),
                                  "lnpi_range": [float(lnpi.min()), float(lnpi.max())], "dN": float(N[1] - N[0]),
                                  "n_bad": int(np.sum(err > 1e-10))}), flush=True)
            if msg is not None:
                bad += 1
                print(json.dumps({"trial": trial, "rec": rec, "n": len(lnpi), "smooth": smooth, "span": float(span), "mismatch": msg,
                                  "dmu": None if dh.desc.mu_recurrence == rec else "recurrence not enabled"}), flush=True)
            if dh.desc.mu_recurrence != rec:
                break
    print(json.dumps({"trials": trials, "mismatching_runs": bad, "seconds": time.time() - t0}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
