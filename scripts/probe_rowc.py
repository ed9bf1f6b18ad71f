#!/usr/bin/env python
"""Config-3 Taylor grid through the row-combined kernel (k_sweep_rowc) and, with FHMC_ROWC=0 in a child process, through the
flat Taylor kernel: timing of both and a full-grid comparison of the records (integers identical?  fe to 1e-10?)."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import _lib, engine, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

NB = int(os.environ.get("NB", 4096))
ND = int(os.environ.get("ND", 4096))
ORDER = int(os.environ.get("ORDER", 2))
n = 1001
h = histogram.from_arrays(synth.two_peak_lnpi(n), synth.two_comp_moments(n), 1.0, [-3.0, -2.5], 10)
h.reweight(-2.9)
betas, dmus = np.linspace(0.95, 1.05, NB), np.linspace(0.2, 0.8, ND)
dh = h.device_histogram(beta=betas, dmu=dmus, order=ORDER, moments=())
st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
res = engine.SweepResult(st.n_states, 8, dh.n_sel, dh.device)
dh.sweep(None, states=st, out=res, pmax=8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    dh.sweep(None, states=st, out=res, pmax=8)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 3
tag = "rowc" if os.environ.get("FHMC_ROWC", "1") != "0" else "flat"
print(tag, _lib.last_kernel(), "ms", ms, "points/s %.4g" % (st.n_states / ms * 1e3),
      "ok", float(((res.status & 0xFF) == 0).double().mean().item()), "fast", float(((res.status & 0x1000) != 0).double().mean().item()))
if os.environ.get("COMPARE", "1") != "1":
    sys.exit(0)
out = os.path.join("/tmp", "rowc_cmp_%s.pt" % tag)
P = res.nphase.clamp(0, 8)
m = torch.arange(8, device=res.fe.device)[None, :] < P[:, None]
torch.save({"status": res.status.cpu(), "nphase": res.nphase.cpu(), "nmin": res.nmin.cpu(), "lnnorm": res.lnnorm.cpu(),
            "fe": torch.where(m, res.fe, torch.zeros_like(res.fe)).cpu(),
            "bounds": torch.where(m[:, :, None], res.bounds.view(-1, 8, 2), torch.zeros_like(res.bounds.view(-1, 8, 2))).cpu(),
            "max_idx": torch.where(m, res.max_idx, torch.zeros_like(res.max_idx)).cpu()}, out)
if tag == "rowc":
    env = dict(os.environ, FHMC_ROWC="0")
    subprocess.check_call([sys.executable, os.path.abspath(__file__)], env=env)
    a, b = torch.load(out), torch.load(out.replace("rowc.pt", "flat.pt"))
    code_a, code_b = a["status"] & 0xFF, b["status"] & 0xFF
    print("code differs", int((code_a != code_b).sum()), "safe differs", int(((a["status"] ^ b["status"]) & 0x100).ne(0).sum()),
          "nphase differs", int((a["nphase"] != b["nphase"]).sum()), "nmin differs", int((a["nmin"] != b["nmin"]).sum()))
    okm = (code_a == 0) & (code_b == 0) & (a["nphase"] == b["nphase"])
    print("bounds differ", int((a["bounds"][okm] != b["bounds"][okm]).any(-1).any(-1).sum()),
          "max_idx differ", int((a["max_idx"][okm] != b["max_idx"][okm]).any(-1).sum()))
    d = (a["fe"][okm] - b["fe"][okm]).abs() / b["fe"][okm].abs().clamp(min=1.0)
    print("max rel fe", float(d.max()), "max rel lnnorm", float(((a["lnnorm"] - b["lnnorm"]).abs() / b["lnnorm"].abs().clamp(min=1.0))[okm].max()))
    os.remove(out)
    os.remove(out.replace("rowc.pt", "flat.pt"))
