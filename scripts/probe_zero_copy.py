#!/usr/bin/env python
"""Can SM stores stream the packed records straight into pinned host memory at PCIe speed?  Runs the repack kernel with
its output pointer in (UVA-mapped) pinned host memory and compares with repack-to-device + cudaMemcpy D2H."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import _lib, engine, synth  # noqa: E402

S = 1000000
L = _lib.load()
lnpi = synth.two_peak_lnpi(1001)
N = np.arange(1001.0)
dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
dh.ensure_hull()
dev = dh.device
mu = dh._dev_array(np.linspace(-0.03, 0.03, S))
res = dh.sweep(mu, pmax=4)
torch.cuda.synchronize()
nb = int(L.fhmc_pack_bytes(S, 2, 2))   # pmax = 2 view of the records is not possible; pack all 4 and look at bytes
nb4 = int(L.fhmc_pack_bytes(S, 4, 2))
host = torch.empty(nb4, dtype=torch.uint8).pin_memory()
devbuf = torch.empty(nb4, dtype=torch.uint8, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
cs = res.c_struct()
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def run(dst_ptr):
    _lib.check(L.fhmc_pack_phase_major(ctypes.byref(cs), S, 4, 2, ctypes.c_void_p(dst_ptr), ctypes.c_void_p(flag.data_ptr()), sp), "pack")


for name, ptr in (("pack_to_device", devbuf.data_ptr()), ("pack_to_pinned_host", host.data_ptr())):
    for _ in range(3):
        run(ptr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        run(ptr)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print(json.dumps({"case": name, "ms": ms, "bytes": nb4, "GBps": nb4 / ms / 1e6}), flush=True)
chk = devbuf.cpu()
print(json.dumps({"host_copy_equals_device": bool(torch.equal(chk.view(torch.int32)[:2 * S], host.view(torch.int32)[:2 * S]))}))
