"""Kernel time of the compact mu sweep (k_sweep_tab2) against the number of state points: fixed cost, cost per full round of
warp tiles (296 CTAs x 8 warps x 64 points = 151552 points on a B200) and the cost of a partly filled last round."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
h = histogram.from_arrays(synth.two_peak_lnpi(1001), synth.one_comp_moments(1001), 1.0, [0.0], 10)
dh = h.device_histogram(moments=("N", "N2"))
R = 296 * 512
sizes = [100000, R, 2 * R, 3 * R, 6 * R, 1000000, 7 * R, 2000000, 4000000]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S in sizes:
    mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
    buf = None
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = dh.sweep_compact(mu, pmax=4, fill_dead=False, dst=buf)
        e1.record()
        torch.cuda.synchronize()
        buf = r["buf"]
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    print("S %8d  rounds %.2f  median %.1f us  min %.1f us  -> %.3e points/s   %s" % (S, S / R, 1e3 * ts[len(ts) // 2], 1e3 * ts[0], S / (1e-3 * ts[len(ts) // 2]), _lib.last_kernel()))
