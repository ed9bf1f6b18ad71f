"""One 10^6-point compact sweep of config 2 (k_sweep_prod2<compact>, what bench.py times) for ncu."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
h = histogram.from_arrays(synth.two_peak_lnpi(1001), synth.one_comp_moments(1001), 1.0, [0.0], 10)
dh = h.device_histogram(moments=("N", "N2"))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
for _ in range(2):
    r = dh.sweep_compact(mu, pmax=4, fill_dead=False)
torch.cuda.synchronize()
print(_lib.last_kernel(), float((r["status"].to(torch.int32) & 0xFF).eq(0).double().mean()))
