#!/usr/bin/env python
"""Config-3 Taylor grid at 1024 x 1024 state points (order-2 lnPI) -- run under ncu with -k regex:k_sweep_fast."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

n = 1001
h = histogram.from_arrays(synth.two_peak_lnpi(n), synth.two_comp_moments(n), 1.0, [-3.0, -2.5], 10)
h.reweight(-2.9)
betas, dmus = np.linspace(0.95, 1.05, 1024), np.linspace(0.2, 0.8, 1024)
dh = h.device_histogram(beta=betas, dmu=dmus, order=2, moments=())
st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
res = engine.SweepResult(st.n_states, 8, dh.n_sel, dh.device)
for _ in range(3):
    dh.sweep(None, states=st, out=res, pmax=8)
torch.cuda.synchronize()
print("ok", float(((res.status & 0xFF) == 0).double().mean().item()), "fast", float(((res.status & 0x1000) != 0).double().mean().item()))
