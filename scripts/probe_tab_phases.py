"""Probe build (-DFHMC_TAB_PROFILE): average cycles a warp of k_sweep_tab2 spends per warp tile in init / walk / finish."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
h = histogram.from_arrays(synth.two_peak_lnpi(1001), synth.one_comp_moments(1001), 1.0, [0.0], 10)
dh = h.device_histogram(moments=("N", "N2"))
L = _lib.load()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
r = dh.sweep_compact(mu, pmax=4, fill_dead=False)
buf = (ctypes.c_ulonglong * 8)()
L.fhmc_tab_profile(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = dh.sweep_compact(mu, pmax=4, fill_dead=False, dst=r["buf"]); e1.record()
L.fhmc_tab_profile(buf, 0)
v = list(buf)
t = max(v[4], 1)
print("grid %s S %d  %.1f us  per warp tile: init %.0f  walk %.0f  finish %.0f cycles (sum %.0f), tiles %d" % (
    os.environ.get("FHMC_TAB_GRID", "full"), S, 1e3 * e0.elapsed_time(e1), v[0] / t, v[1] / t, v[2] / t, (v[0] + v[1] + v[2]) / t, t))
