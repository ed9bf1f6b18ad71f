#!/usr/bin/env python
"""Device timeline of the host-buffer sweep pipeline (FHMC_PIPE_TRACE=1): where the time between the D2H floor and the
measured e2e goes."""
import os
import sys
import time

import numpy as np
import torch

os.environ["FHMC_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine, synth  # noqa: E402

S = 1000000
lnpi = synth.two_peak_lnpi(1001)
N = np.arange(1001.0)
dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
mu = torch.from_numpy(np.linspace(-0.03, 0.03, S)).pin_memory()
o = None
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = dh.sweep_host_compact(mu, pmax=4, out=o)
    print("call %d wall %.3f ms" % (rep, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
