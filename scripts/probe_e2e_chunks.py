"""e2e (host buffers) of the 10^6-point sweep against the chunk size of the host pipeline, with and without tilt cells."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fhmcanalysis_b200 import _lib, engine, synth
n = 1001
N = np.arange(n, dtype=np.float64)
S = 1000000
mu_h = torch.from_numpy(np.linspace(-0.03, 0.03, S)).pin_memory()
for cells in (False, True):
    dh = engine.DeviceHistogram(synth.two_peak_lnpi(n), N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    dh.use_mu_cells = cells
    if cells:
        dh.sweep_compact(mu_h.cuda(), pmax=4)     # builds the cells for this range
        keep = mu_h.cuda(); dh.sweep_compact(keep, pmax=4)
    for lg in (15, 16, 17, 18, 19, 20):
        out = None
        for _ in range(3):
            out = dh.sweep_host_compact(mu_h, pmax=4, chunk=1 << lg, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            t0 = time.perf_counter()
            out = dh.sweep_host_compact(mu_h, pmax=4, chunk=1 << lg, out=out)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        print("cells %d chunk 2^%d: median %.3f ms min %.3f ms -> %.3e state points/s  (%s)" % (cells, lg, 1e3 * ts[7], 1e3 * ts[0], S / ts[7], _lib.last_kernel()), flush=True)
