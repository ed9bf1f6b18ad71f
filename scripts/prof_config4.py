#!/usr/bin/env python
"""One generic (32 lanes per state point) Taylor sweep over the config-4 temperatures: the per-evaluation work of K4.
Meant to be run under ncu (-k regex:k_sweep_1d)."""
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
mu = np.full_like(betas, -0.003)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for _ in range(3):
    r = dh.sweep(mu, beta=betas, pmax=4, lanes=lanes)
torch.cuda.synchronize()
print("ok", float((r.host()["code"] == 0).mean()))
