#!/usr/bin/env python
"""Generate tests/golden/n1_vectors.npz from the COMPILED REFERENCE module moments/histogram/one_dim/n1/gc_hist.pyx
(oracle/_ref/gc_hist_n1, built by oracle/build_ref.py).  Build container only; the GPU box uses the committed file.
Driven like the N_tot vectors: fresh histogram -> reweight(mu) -> [temp_mu_extrap] -> thermo() -> is_safe();
find_phase_eq; mix; temp_mu_extrap_multi; the private derivative builders."""
import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from fhmcanalysis_b200 import synth  # noqa: E402


def make(H, lnpi, mom, beta_ref, mu_ref, smooth, volume=1.0):
    h = H.__new__(H)
    mu_ref = np.atleast_1d(np.array(mu_ref, dtype=np.float64))
    h.metadata = {"beta_ref": float(beta_ref), "mu_ref": mu_ref.copy(), "nspec": len(mu_ref), "smooth": int(smooth),
                  "fname": "", "file_history": "synthetic"}
    n = len(lnpi)
    h.data = {"curr_mu": mu_ref.copy(), "curr_beta": float(beta_ref), "nspec": len(mu_ref),
              "ln(PI)": np.array(lnpi, dtype=np.float64), "max_order": mom.shape[1] - 1, "volume": float(volume),
              "n1": np.arange(n, dtype=np.int64), "lb": 0, "ub": n - 1, "pk_hist": {}, "e_hist": {},
              "mom": np.array(mom, dtype=np.float64)}
    return h


# the (2,4,2,4,4,n) tensors are large: keep the entries the kernels/tests read plus mixed high-order ones
ADDR = [(0, 1, 0, 0, 0), (1, 1, 0, 0, 0), (0, 0, 0, 0, 1), (1, 1, 1, 1, 0), (1, 1, 0, 0, 1), (1, 2, 0, 0, 0), (0, 0, 0, 0, 2),
        (0, 1, 1, 1, 0), (0, 1, 0, 0, 1), (1, 0, 1, 1, 1)]


def sub(m):
    """Selected moment entries of a tensor whose LAST 6 axes are (i,j,k,m,p,n) (or 5 axes without n)."""
    m = np.asarray(m)
    if m.shape[-1] in (3, 4) and m.ndim >= 5 and m.shape[-5:] in ((2, 4, 2, 4, 4), (1, 4, 1, 4, 4)):   # phase averages
        return np.stack([m[..., a[0] % m.shape[-5], a[1], a[2] % m.shape[-3], a[3], a[4]] for a in ADDR], axis=-1)
    return np.stack([m[..., a[0] % m.shape[-6], a[1], a[2] % m.shape[-4], a[3], a[4], :] for a in ADDR], axis=-2)


def record(h, out, prefix):
    P = len(h.data["thermo"])
    out[prefix + "/lnpi"] = np.array(h.data["ln(PI)"])
    out[prefix + "/maxima"] = np.array(h.data["ln(PI)_maxima_idx"], dtype=np.int64)
    out[prefix + "/minima"] = np.array(h.data["ln(PI)_minima_idx"], dtype=np.int64)
    out[prefix + "/fe"] = np.array([h.data["thermo"][p]["F.E./kT"] for p in range(P)])
    out[prefix + "/bounds"] = np.array([h.data["thermo"][p]["bound_idx"] for p in range(P)], dtype=np.int64)
    out[prefix + "/safe"] = np.array(bool(h.is_safe()))
    if "mom" in h.data["thermo"][0]:
        out[prefix + "/mom"] = sub(np.array([h.data["thermo"][p]["mom"] for p in range(P)]))
        out[prefix + "/x1"] = np.array([h.data["thermo"][p]["x1"] for p in range(P)])
        out[prefix + "/density"] = np.array([h.data["thermo"][p]["density"] for p in range(P)])
    out[prefix + "/curr_mu"] = np.array(h.data["curr_mu"])


def main():
    ns = ref.load()
    if ns is None:
        raise SystemExit("compiled reference unavailable: %s" % ref._cache.get("error"))
    H = ns.gc_hist_n1.histogram
    out, meta = {}, {}
    n = 201
    lnpi = synth.two_peak_lnpi(n, noise=1e-3, scale=0.2)
    mom = synth.n1_two_comp_moments(n, 3)
    beta_ref, mu_ref, smooth = 1.0, [-0.2, -1.5], 5
    meta["setup"] = {"n": n, "beta_ref": beta_ref, "mu_ref": mu_ref, "smooth": smooth, "volume": 512.0,
                     "lnpi": "synth.two_peak_lnpi(201, 1e-3, 0.2)", "mom": "synth.n1_two_comp_moments(201, 3)"}
    out["lnpi"], out["mom_checksum"] = lnpi, np.array([mom.sum(), np.abs(mom).max()])

    # A. reweight -> thermo -> is_safe
    mus = np.linspace(-0.45, 0.05, 11)
    out["A/mu"] = mus
    for k, mu in enumerate(mus):
        h = make(H, lnpi, mom, beta_ref, mu_ref, smooth, 512.0)
        h.reweight(mu)
        h.thermo()
        record(h, out, "A/%d" % k)

    # B. private derivative builders at the reference state (normalised first, as temp_mu_extrap does)
    h = make(H, lnpi, mom, beta_ref, mu_ref, smooth, 512.0)
    h.normalize()
    d, dm = h._dBMU(False)
    out["B/dBMU/d"], out["B/dBMU/dm"] = d, sub(dm)
    Hl, Hm = h._dBMU2(False)
    out["B/dBMU2/H"], out["B/dBMU2/Hm"] = Hl, sub(Hm)
    d2, d2m = h._dB2(False)
    out["B/dB2/d"], out["B/dB2/dm"] = d2, sub(d2m)
    out["B/sg_dX_dB"] = h._sg_dX_dB([1, 1, 0, 0, 1])
    out["B/sg_dX_dMU"] = h._sg_dX_dMU(0, [1, 1, 0, 0, 1])
    out["B/sg_d2X_dB2"] = h._sg_d2X_dB2([1, 1, 0, 0, 0])
    out["B/sg_d2X_dMU2"] = h._sg_d2X_dMU2(0, 0, [0, 0, 0, 0, 1])
    out["B/sg_df_dB"] = h._sg_df_dB([1, 1, 0, 0, 0], [0, 0, 0, 0, 1])
    out["B/sg_df_dMU"] = h._sg_df_dMU(0, [1, 1, 0, 0, 0], [0, 0, 0, 0, 1])
    out["B/gc_dX_dB"] = np.array(h._gc_dX_dB([1, 1, 0, 0, 0]))
    out["B/gc_fluct_ii"] = np.array(h._gc_fluct_ii([1, 1, 0, 0, 0], [0, 0, 0, 0, 1]))
    out["B/gc_fluct_vi"] = np.array(h._gc_fluct_vi(h.data["mom"][1, 1, 0, 0, 0], [0, 0, 0, 0, 1]))

    # C. temp_mu_extrap, orders 1 and 2, with and without the moments, after a reweight
    cases = [(1, 1.03, [-1.45], False), (2, 1.03, [-1.45], False), (2, 0.96, [-1.6], True), (1, 1.0, [-1.3], True),
             (2, 1.05, [-1.5], False)]
    meta["C"] = [[o, b, m, s] for o, b, m, s in cases]
    for k, (order, tb, tm, skip) in enumerate(cases):
        h = make(H, lnpi, mom, beta_ref, mu_ref, smooth, 512.0)
        h.reweight(-0.25)
        e = h.temp_mu_extrap(tb, np.array(tm), order, 10.0, False, True, skip)
        out["C/%d/lnpi" % k] = np.array(e.data["ln(PI)"])
        out["C/%d/mom" % k] = sub(e.data["mom"])
        out["C/%d/curr_mu" % k] = np.array(e.data["curr_mu"])
        e.thermo()
        record(e, out, "C/%d/thermo" % k)

    # D. temp_mu_extrap_multi grid
    h = make(H, lnpi, mom, beta_ref, mu_ref, smooth, 512.0)
    h.reweight(-0.25)
    tbs, tms = np.array([0.98, 1.0, 1.04]), np.array([[-1.55], [-1.4]])
    out["D/betas"], out["D/mus"] = tbs, tms
    for order in (1, 2):
        g = h.temp_mu_extrap_multi(tbs, tms, order, 10.0, False, False)
        out["D/%d/lnpi" % order] = np.array([[g[i][j].data["ln(PI)"] for j in range(2)] for i in range(3)])
        out["D/%d/mom" % order] = sub(np.array([[g[i][j].data["mom"] for j in range(2)] for i in range(3)]))

    # E. find_phase_eq: same conditions, other temperature + mu_2 (order 1 and 2)
    eq_cases = [(0.0, [], 1), (1.02, [-1.45], 1), (0.97, [-1.55], 2)]
    meta["E"] = [[b, m, o] for b, m, o in eq_cases]
    for k, (tb, tm, order) in enumerate(eq_cases):
        h = make(H, lnpi, mom, beta_ref, mu_ref, smooth, 512.0)
        with redirect_stdout(io.StringIO()):
            eq = h.find_phase_eq(1e-8, -0.2, tb, tm, order, 10.0, True)
        record(eq, out, "E/%d" % k)

    # F. mix two histograms of different length
    a = make(H, lnpi, mom, beta_ref, mu_ref, smooth, 512.0)
    b = make(H, lnpi[:150] + 0.01 * np.cos(np.arange(150) / 9.0), mom[..., :150] * 1.001, beta_ref, mu_ref, smooth, 512.0)
    a.metadata["used_ke"] = b.metadata["used_ke"] = False   # n1 mix never reads it; harmless
    m = a.mix(b, [0.3, 0.7])
    out["F/lnpi_b"] = b.data["ln(PI)"].copy()
    out["F/lnpi"], out["F/mom"] = np.array(m.data["ln(PI)"]), sub(m.data["mom"])

    # G. one-component N_1 histogram (N_1 = N_tot): reweight/thermo + temperature-only extrapolation
    mom1 = synth.one_comp_moments(n, 3)
    for k, (mu, tb, order) in enumerate([(-0.3, 1.0, 0), (-0.25, 1.04, 1), (-0.25, 0.95, 2)]):
        h = make(H, lnpi, mom1, 1.0, [-0.2], smooth, 512.0)
        h.reweight(mu)
        if order:
            h = h.temp_mu_extrap(tb, np.array([]), order, 10.0, False, True, False)
        h.thermo()
        record(h, out, "G/%d" % k)
    meta["addr"] = [list(a) for a in ADDR]
    meta["G"] = [[-0.3, 1.0, 0], [-0.25, 1.04, 1], [-0.25, 0.95, 2]]

    np.savez_compressed(os.path.join(HERE, "n1_vectors.npz"), **out)
    with open(os.path.join(HERE, "n1_vectors.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("wrote %d arrays, %.1f kB" % (len(out), os.path.getsize(os.path.join(HERE, "n1_vectors.npz")) / 1e3))


if __name__ == "__main__":
    main()
