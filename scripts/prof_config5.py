#!/usr/bin/env python
"""Config 5 (512x512 joint histogram, 10^5 (mu1, mu2) pairs) through the product-form K5 kernels -- the command profiled for
profiles/r01b_rw2d_prod_ncu_summary.txt."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine, synth  # noqa: E402

n1 = n2 = 512
lnpi2d, bounds = synth.joint_2d(n1, n2, 640)
g1, g2 = np.meshgrid(np.linspace(-0.02, 0.02, 316), np.linspace(-0.02, 0.02, 317), indexing="ij")
a1, a2 = g1.ravel()[:100000].copy(), g2.ravel()[:100000].copy()
dev = engine.require_cuda()
tl, tb = torch.from_numpy(lnpi2d).to(dev), torch.from_numpy(bounds).to(dev)
to1, to2 = torch.arange(n1, dtype=torch.float64, device=dev), torch.arange(n2, dtype=torch.float64, device=dev)
ta1, ta2 = torch.from_numpy(a1).to(dev), torch.from_numpy(a2).to(dev)
for _ in range(4):
    out = engine.reweight_2d(tl, tb, to1, to2, ta1, ta2, None, return_device=True, product=True)
torch.cuda.synchronize()
print("ok", out[0].tolist())
