"""A/B of the two cell kernels (FHMC_CELL_T=0/1 in the environment): timing at 10^6 and 4x10^6 state points."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fhmcanalysis_b200 import _lib, engine, synth
n = 1001
lnpi = synth.two_peak_lnpi(n)
N = np.arange(n, dtype=np.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S in (1000000, 4000000):
    mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    st = dh.make_states(mu)
    buf = torch.empty(int(_lib.load().fhmc_pack_soa16_bytes(S, 4, 2)), dtype=torch.uint8, device="cuda")
    fn = lambda: dh.sweep_compact(None, pmax=4, dst=buf, states=st, fill_dead=False)
    r = fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("FHMC_CELL_T=%s S %d median %.1f us min %.1f us  cell fraction %.4f" % (os.environ.get("FHMC_CELL_T"), S, 1e3 * ts[15], 1e3 * ts[0], float(r["path"].double().mean())), flush=True)
