#!/usr/bin/env python
"""Config 4 (10^4 temperatures, N_max = 2000, order-2 beta extrapolation, every guess 0): every solve from its own cold guess
against the in-kernel continuation of fhmc_find_phase_eq_curve at several seed strides (one launch each)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import _lib, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

n4 = 2001
h4 = histogram.from_arrays(synth.two_peak_lnpi(n4, scale=2.0), synth.one_comp_moments(n4, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.90, 1.06, 10000)
dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
g = np.zeros_like(betas)


def run(stride):
    hold = {}

    def f():
        if stride:
            hold["r"] = dh._find_phase_eq_once(g, betas, None, 1e-10, None, 200, 4, None, None, None, seed_stride=stride)
        else:
            hold["r"] = dh._find_phase_eq_once(g, betas, None, 1e-10, None, 200, 4, None, None, None)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / 3, hold["r"].host()


ms0, h0 = run(0)
conv0 = (h0["code"] == 0) & ((h0["status"].view(np.uint32) & _lib.ST_JUMP) == 0)
print("cold: %.3f ms  %.3g solves/s  evals %.2f  converged %.4f" % (ms0, 1e4 / ms0 * 1e3, h0["iters"].mean(), conv0.mean()))
for stride in (32, 64, 128, 256, 512, 1024):
    ms, h = run(stride)
    conv = (h["code"] == 0) & ((h["status"].view(np.uint32) & _lib.ST_JUMP) == 0)
    both = conv & conv0
    dmu = np.abs(h["mu_coex"][both] - h0["mu_coex"][both])
    same_int = all(np.array_equal(h[k][both], h0[k][both]) for k in ("nphase",)) and np.array_equal(h["bounds"][both][:, :2], h0["bounds"][both][:, :2])
    print("stride %2d: %.3f ms  %.3g solves/s  evals %.2f  converged %.4f  (both %d)  max|dmu| %.2e  max|dfe| %.2e  same bounds %s  code-diff %d" % (
        stride, ms, 1e4 / ms * 1e3, h["iters"].mean(), conv.mean(), both.sum(), dmu.max(), np.abs(h["dfe"][conv]).max(), same_int,
        int(np.sum(h["code"] != h0["code"]))))
ms, h = run(32)
conv = (h["code"] == 0) & ((h["status"].view(np.uint32) & _lib.ST_JUMP) == 0)
both = conv & conv0
dm = np.abs(h["mu_coex"] - h0["mu_coex"])
for thr in (1e-10, 1e-8, 1e-6, 1e-4):
    print("both converged and |dmu| > %g: %d" % (thr, int(np.sum(both & (dm > thr)))))
idx = np.where(both & (dm > 1e-8))[0]
print("indices (T):", idx[:20], (1.0 / betas[idx[:20]]).round(4), "first/last", idx.min() if len(idx) else None, idx.max() if len(idx) else None)
k = idx[0] if len(idx) else 0
for name, hh in (("cold", h0), ("curve", h)):
    print(name, "mu", hh["mu_coex"][k], "dfe", hh["dfe"][k], "nphase", hh["nphase"][k], "bounds", hh["bounds"][k].tolist(), "iters", hh["iters"][k])
only0, only1 = conv0 & ~conv, conv & ~conv0
print("converged only cold:", int(only0.sum()), "only curve:", int(only1.sum()), "T range of non-converged (curve):", (1.0 / betas[~conv]).min().round(4) if (~conv).any() else None)
