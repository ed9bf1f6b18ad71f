#!/usr/bin/env python
"""Golden vectors for pore_hist.normalize from the COMPILED reference (oracle/_ref/pore_hist, built by
oracle/build_ref.py from moments/histogram/two_dim/h_ntot/pore_hist.pyx).  Run in the build container (needs
/root/reference to have been compiled); writes tests/golden/pore_vectors.npz.

pore_hist.__init__ of the reference always raises (it reads data['ln(PI)'] before assigning it, pore_hist.pyx:129), so
objects are made with __new__ and a hand-filled data dict; thermo() raises for every mask with more than one element
(pore_hist.pyx:170), so only normalize() can be recorded."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import pore_hist as ph  # noqa: E402


def surface(n1, n2, seed, steep=1.0):
    """two ridges in (h, N) with a ragged right edge (shorter rows at small h), -inf beyond the edge"""
    rng = np.random.default_rng(seed)
    h = np.linspace(0.0, 1.0, n1)[:, None]
    N = np.arange(n2)[None, :]
    lp = np.logaddexp(-((N - 0.25 * n2) ** 2) / (2 * (0.08 * n2) ** 2) - 6 * (h - 0.3) ** 2,
                      -((N - 0.7 * n2) ** 2) / (2 * (0.1 * n2) ** 2) - 4 * (h - 0.8) ** 2 - 1.0) * steep
    lp = lp + 1e-3 * rng.normal(size=lp.shape)
    edge = np.minimum(n2 - 1, (0.45 * n2 + 0.55 * n2 * np.linspace(0, 1, n1) ** 0.7).astype(np.int32)).astype(np.int32)
    for i in range(n1):
        lp[i, edge[i] + 1:] = -np.inf
    return lp, edge


def main():
    out = {}
    for name, (n1, n2, seed, steep) in {"small": (7, 9, 1, 1.0), "mid": (40, 130, 2, 1.0), "steep": (33, 257, 3, 40.0)}.items():
        lp, edge = surface(n1, n2, seed, steep)
        obj = ph.pore_hist.__new__(ph.pore_hist)
        obj.data = {"ln(PI)": lp.copy(), "edge_idx": edge.copy()}
        obj.normalize()
        out[name + "/lnpi"] = lp
        out[name + "/edge"] = edge
        out[name + "/normalized"] = np.asarray(obj.data["ln(PI)"])
    np.savez_compressed(os.path.join(HERE, "pore_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
