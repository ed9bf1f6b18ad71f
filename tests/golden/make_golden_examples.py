#!/usr/bin/env python
"""Golden vectors from the reference's own example composites (example/ntot/*/composite.nc; SURVEY 8(f) row 1: "running the
surviving example composites as regression data"), produced by the COMPILED reference (oracle/_ref).  Run in the build
container only.  For every composite: a short mu_1 sweep (reweight -> thermo -> is_safe) and first/second order
temperature (and dmu_2) extrapolations of ln(PI).  The fixture keeps ln(PI) and the moment sub-tensor up to order 2 (all
the recorded operations read), not the whole files.  Writes tests/golden/examples_vectors.npz + .json."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref  # noqa: E402
from make_golden import thermo_record, pack  # noqa: E402
from fhmcanalysis_b200.io.hdf5_min import Dataset  # noqa: E402

REF = os.environ.get("FHMC_REFERENCE", "/root/reference")
CASES = [   # (key, path below example/ntot, T, dMu2 or None, smooth)
    ("ig_T1.00_m1.10", "binary_ideal_gas/T_1.00/dMu2_-1.10", 1.00, -1.10, 5),
    ("ig_T1.00_0.00", "binary_ideal_gas/T_1.00/dMu2_0.00", 1.00, 0.00, 5),
    ("ig_T1.00_2.94", "binary_ideal_gas/T_1.00/dMu2_2.94", 1.00, 2.94, 5),
    ("ig_T1.00_1.10", "binary_ideal_gas/T_1.00/dMu2_1.10", 1.00, 1.10, 5),
    ("ig_T1.00_m2.94", "binary_ideal_gas/T_1.00/dMu2_-2.94", 1.00, -2.94, 5),
    ("ig_T1.20_1.10", "binary_ideal_gas/T_1.20/dMu2_1.10", 1.20, 1.10, 5),
    ("ig_T1.20_m2.94", "binary_ideal_gas/T_1.20/dMu2_-2.94", 1.20, -2.94, 5),
    ("sw_T1.10", "square_well/T_1.10", 1.10, None, 10),
]


def main():
    ns = ref.load()
    if ns is None:
        raise SystemExit("compiled reference unavailable: %s" % ref._cache.get("error"))
    H = ns.gc_hist.histogram
    out, meta = {}, {}
    for key, rel, T, dmu2, smooth in CASES:
        path = os.path.join(REF, "example/ntot", rel, "composite.nc")
        d = Dataset(path)
        lnpi = np.array(d.variables["ln(PI)"][:], dtype=np.float64)
        mom = np.array(d.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:], dtype=np.float64)
        beta = 1.0 / T
        mu_ref = [0.0] if dmu2 is None else [0.0, dmu2]
        out[key + "/lnpi"] = lnpi
        out[key + "/mom"] = mom[:, :3, :, :3, :3, :].copy()       # everything orders <= 2 read
        # where the sweep is interesting: tilts that move the maximum of ln(PI) through the N range
        n = len(lnpi)
        slope = np.gradient(lnpi)
        mus = [float(-slope[int(f * (n - 1))] / beta) for f in (0.15, 0.4, 0.6, 0.85)] + [0.0]
        meta[key] = {"beta_ref": beta, "mu_ref": mu_ref, "smooth": smooth, "volume": float(d.volume), "mus": mus, "status": []}
        for k, mu in enumerate(mus):
            h = H(path, beta, mu_ref, smooth)
            try:
                with redirect_stdout(io.StringIO()):
                    h.reweight(mu)
                    h.thermo()
                rec = thermo_record(h)
                P = len(h.data["thermo"])
                rec["n1"] = np.array([h.data["thermo"][p]["n1"] for p in range(P)])
                rec["u"] = np.array([h.data["thermo"][p]["u"] for p in range(P)])
                rec["density"] = np.array([h.data["thermo"][p]["density"] for p in range(P)])
                if dmu2 is not None:
                    rec["n2"] = np.array([h.data["thermo"][p]["n2"] for p in range(P)])
                    rec["x1"] = np.array([h.data["thermo"][p]["x1"] for p in range(P)])
                rec.pop("mom", None)
                pack("%s/%d" % (key, k), rec, out)
                meta[key]["status"].append("ok")
            except Exception as e:
                meta[key]["status"].append("raise: " + str(e)[:80])
        # Taylor extrapolation of ln(PI) (skip_mom=True) from the state reweighted to mus[1]
        tb = 1.0 / (T * 1.02)
        for order in (1, 2):
            h = H(path, beta, mu_ref, smooth)
            h.reweight(mus[1])
            with redirect_stdout(io.StringIO()):
                if dmu2 is None:
                    hn = h.temp_extrap(tb, order, 10.0, True, True, True)
                else:
                    hn = h.temp_dmu_extrap(tb, np.array([dmu2 + 0.05]), order, 10.0, True, True, True, False)
            out["%s/extrap%d" % (key, order)] = np.array(hn.data["ln(PI)"], dtype=np.float64)
        meta[key]["extrap"] = {"beta": tb, "dmu": None if dmu2 is None else [dmu2 + 0.05]}
    # ---- isopleth.make_grid_multi over the five T* = 1.00 ideal-gas composites (gc_binary.pyx:173-290) ----------------
    iso_keys = ["ig_T1.00_m2.94", "ig_T1.00_m1.10", "ig_T1.00_0.00", "ig_T1.00_1.10", "ig_T1.00_2.94"]
    hs = []
    for k in iso_keys:
        rel = [c[1] for c in CASES if c[0] == k][0]
        hs.append(H(os.path.join(REF, "example/ntot", rel, "composite.nc"), 1.0, meta[k]["mu_ref"], 5))
    meta["iso"] = {"keys": iso_keys, "beta_ref": 1.0, "order": 1, "mu1_bounds": [-4.5, -3.0], "dmu2_bounds": [-2.5, 2.5],
                   "delta": [0.25, 0.5], "m": 2.5}
    with redirect_stdout(io.StringIO()):
        iso = ns.gc_binary.isopleth(hs, 1.0, 1)
        Z, (X, Y) = iso.make_grid_multi([-4.5, -3.0], [-2.5, 2.5], [0.25, 0.5], 2.5)
    out["iso/x1"], out["iso/density"], out["iso/fe"] = np.array(Z), np.array(iso.data["density"]), np.array(iso.data["F.E./kT"])
    out["iso/X"], out["iso/Y"] = np.array(X), np.array(Y)
    np.savez_compressed(os.path.join(HERE, "examples_vectors.npz"), **out)
    with open(os.path.join(HERE, "examples_vectors.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote %d arrays, %.0f KB" % (len(out), os.path.getsize(os.path.join(HERE, "examples_vectors.npz")) / 1024.0))
    for k, v in meta.items():
        if "status" in v:
            print(k, v["status"], [round(m, 4) for m in v["mus"]])
    print("iso filled fraction", float(np.mean(out["iso/x1"] != 0)))


if __name__ == "__main__":
    main()
