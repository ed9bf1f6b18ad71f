#!/usr/bin/env python
"""Two batched coexistence solves on config 4 (cold guesses) -- run under ncu with -k regex:k_find_phase_eq."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.90, 1.06, 10000)
dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
g = np.zeros_like(betas)
for _ in range(2):
    r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4)
torch.cuda.synchronize()
hr = r.host()
print("ok", float((hr["code"] == 0).mean()), "evals", float(hr["iters"].mean()))
