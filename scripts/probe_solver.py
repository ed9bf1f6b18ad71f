#!/usr/bin/env python
"""K4 diagnostics on config 4: distribution of evaluations per solve, cost of the non-converging tail, lane width."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
    T = np.linspace(0.90, 1.06, 10000)
    betas = 1.0 / T
    for moments in (("N", "N2", "U"), ("N",)):
        dh = h4.device_histogram(beta=betas, order=2, moments=moments)
        g = np.zeros_like(betas)
        r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4)
        hr = r.host()
        it = hr["iters"]
        ok = hr["code"] == 0
        row = {"moments": moments, "rows": int(dh.desc.n_rows), "ok": float(ok.mean()), "codes": {int(c): int((hr["code"] == c).sum()) for c in np.unique(hr["code"])},
               "iters_pct_ok": [float(np.percentile(it[ok], q)) for q in (50, 90, 99, 100)],
               "iters_pct_fail": [float(np.percentile(it[~ok], q)) for q in (50, 90, 99, 100)] if (~ok).any() else None,
               "ms_all": timed(lambda: dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4))}
        bo = betas[ok]
        go = np.zeros_like(bo)
        row["ms_converging_only"] = timed(lambda: dh.find_phase_eq(go, beta=bo, lnz_tol=1e-10, pmax=4))
        row["n_converging"] = int(ok.sum())
        for lanes in ("4", "1"):
            os.environ["FHMC_SOLVER_LANES"] = lanes
            row["ms_all_lanes" + lanes] = timed(lambda: dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4))
        os.environ.pop("FHMC_SOLVER_LANES")
        # single evaluation cost: a plain sweep of the same 10^4 state points (generic kernel, same lanes rule)
        mu = hr["mu_coex"]
        for lanes in (32, 4, -1, 0):
            row["ms_sweep_lanes%d" % lanes] = timed(lambda: dh.sweep(mu, beta=betas, pmax=4, lanes=lanes))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
