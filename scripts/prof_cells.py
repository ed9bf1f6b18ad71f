"""One warm + a few profiled compact mu sweeps on tilt cells (run under ncu; config 2 of bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fhmcanalysis_b200 import _lib, engine, synth

n = 1001
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
N = np.arange(n, dtype=np.float64)
dh = engine.DeviceHistogram(synth.two_peak_lnpi(n), N, 1.0, 0.0, smooth=10, sel=["N", N * N])
mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
st = dh.make_states(mu)
buf = torch.empty(int(_lib.load().fhmc_pack_soa16_bytes(S, 4, 2)), dtype=torch.uint8, device="cuda")
for _ in range(3):
    r = dh.sweep_compact(None, pmax=4, dst=buf, states=st, fill_dead=False)
torch.cuda.synchronize()
print(_lib.last_kernel(), float(r["path"].double().mean()))
