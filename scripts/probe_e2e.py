#!/usr/bin/env python
"""End-to-end host sweep: time vs chunk size, and the pieces (H2D, kernels, D2H) alone."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine, synth  # noqa: E402

S = 1000000
lnpi = synth.two_peak_lnpi(1001)
N = np.arange(1001.0)
dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
mu = torch.from_numpy(np.linspace(-0.03, 0.03, S)).pin_memory()


def wall(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
    st = {"o": None, "p": None}

    def f():
        st["o"] = dh.sweep_host_compact(mu, pmax=4, chunk=chunk, out=st["o"])

    def fpy():
        st["p"] = dh.sweep_host_compact_py(mu, pmax=4, chunk=chunk, out=st["p"])
    print(json.dumps({"chunk": chunk, "compact_c_ms": wall(f), "compact_py_ms": wall(fpy), "d2h_bytes": st["o"]["d2h_bytes"]}), flush=True)
# pieces
dev = torch.device("cuda:0")
buf_d = torch.empty(72000000, dtype=torch.uint8, device=dev)
buf_h = torch.empty(72000000, dtype=torch.uint8).pin_memory()
print(json.dumps({"d2h_72MB_ms": wall(lambda: buf_h.copy_(buf_d, non_blocking=True))}))
mu_d = torch.empty(S, dtype=torch.float64, device=dev)
print(json.dumps({"h2d_8MB_ms": wall(lambda: mu_d.copy_(mu, non_blocking=True))}))
dh.ensure_hull()
res = engine.SweepResult(S, 4, dh.n_sel, dev)
print(json.dumps({"kernel_ms": wall(lambda: dh.sweep(mu_d, pmax=4, out=res))}))
