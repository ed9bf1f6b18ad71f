#!/usr/bin/env python
"""Golden vectors for the API the reference exposes but its own unit tests never drive, produced by the COMPILED reference
(oracle/_ref): ``histogram.coexisting`` (gc_hist.pyx:417-449), the ``collect=`` hook of ``thermo`` / ``find_phase_eq``
(gc_hist.pyx:485-486, 653, 662 with collect.py:32-80 ``janus_collect``), ``isopleth.get_hist`` (gc_binary.pyx:292-353) and
``isopleth.make_grid`` (gc_binary.pyx:355-476).  Run in the build container only.

collect.py is Python-2 source with one tab/space mix (collect.py:63) and ``xrange`` (collect.py:29): it is read from
/root/reference, those two porting edits are applied in memory and the result is exec'd -- no copy is written anywhere.
Writes tests/golden/api_vectors.npz + api_vectors.json."""
import copy
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref  # noqa: E402
from fhmcanalysis_b200 import synth  # noqa: E402
from make_golden import thermo_record, pack  # noqa: E402

REF = os.environ.get("FHMC_REFERENCE", "/root/reference")


def reference_collect_module():
    src = open(os.path.join(REF, "moments/histogram/one_dim/ntot/collect.py")).read()
    src = src.replace("xrange", "range").replace("                        min_idx = [0]", "\t\t\tmin_idx = [0]")
    mod = types.ModuleType("reference_collect")
    exec(compile(src, "collect.py", "exec"), mod.__dict__)
    return mod


def three_peak_lnpi(n=301):
    i = np.arange(n, dtype=np.float64)
    rng = np.random.default_rng(4242)
    g = [-(i - 30.0) ** 2 / (2 * 9.0 ** 2), -(i - 95.0) ** 2 / (2 * 12.0 ** 2) - 0.5, -(i - 230.0) ** 2 / (2 * 25.0 ** 2) - 1.0]
    return np.logaddexp(np.logaddexp(g[0], g[1]), g[2]) + 1e-3 * rng.normal(size=n)


def main():
    ns = ref.load()
    assert ns is not None, ref._cache.get("error")
    coll = reference_collect_module()
    out, meta = {}, {}
    quiet = io.StringIO()

    # ---- A. coexisting -----------------------------------------------------------------------------------------
    n = 201
    lnpi, mom = synth.two_peak_lnpi(n, scale=0.2), synth.one_comp_moments(n)
    out["coex/lnpi"], out["coex/mom"] = lnpi, mom
    base = ref.make_histogram(lnpi, mom, 1.0, [0.0], 5)
    with redirect_stdout(quiet):
        eq = base.find_phase_eq(1e-10, 0.0)
    mu_star = float(eq.data["curr_mu"][0])
    cases = []
    for dmu in (0.0, 1e-5, 2e-3, 5e-2, 0.6):
        h = copy.deepcopy(base)
        h.reweight(mu_star + dmu)
        h.thermo()
        fe = [float(h.data["thermo"][p]["F.E./kT"]) for p in range(len(h.data["thermo"]))]
        cases.append({"mu": mu_star + dmu, "fe": fe, "rtol": {str(r): h.coexisting(r) for r in (1e-3, 1e-6, 0.5)}})
    meta["coex"] = {"beta_ref": 1.0, "mu_ref": [0.0], "smooth": 5, "mu_star": mu_star, "cases": cases}
    h = copy.deepcopy(base)
    try:
        h.coexisting()
        meta["coex"]["before_thermo"] = "no error"
    except Exception as e:
        meta["coex"]["before_thermo"] = str(e)

    # ---- B. collect= hook ----------------------------------------------------------------------------------------
    n3 = 301
    l3, m3 = three_peak_lnpi(n3), synth.one_comp_moments(n3)
    out["collect/lnpi"], out["collect/mom"] = l3, m3
    b3 = ref.make_histogram(l3, m3, 1.0, [0.0], 5)
    mus = [0.0, 0.01, -0.02]
    meta["collect"] = {"beta_ref": 1.0, "mu_ref": [0.0], "smooth": 5, "mus": mus}
    for k, mu in enumerate(mus):
        h = copy.deepcopy(b3)
        h.reweight(mu)
        h.thermo()
        pack("collect/plain%d" % k, thermo_record(h), out)
        h = copy.deepcopy(b3)
        h.reweight(mu)
        h.thermo(True, False, coll.janus_collect)
        pack("collect/janus%d" % k, thermo_record(h), out)
    with redirect_stdout(quiet):
        eq, err = b3.find_phase_eq(1e-10, 0.0, 0.0, [], 1, 10.0, False, True, False, coll.janus_collect)
    pack("collect/eq", thermo_record(eq), out)
    meta["collect"]["eq_mu"], meta["collect"]["eq_err"] = float(eq.data["curr_mu"][0]), float(err)
    # two maxima only: the reference's janus_collect falls through to an unassigned name (collect.py:77)
    h = copy.deepcopy(base)
    h.thermo()
    try:
        coll.janus_collect(hist=h)
        meta["collect"]["two_maxima"] = "no error"
    except Exception as e:
        meta["collect"]["two_maxima"] = type(e).__name__

    # ---- C. isopleth.get_hist / make_grid (same three histograms as section F of make_golden.py) -------------------
    g = np.load(os.path.join(HERE, "reference_vectors.npz"))
    gm = json.load(open(os.path.join(HERE, "reference_vectors.json")))["iso"]
    mom2 = g["c3/mom"]

    def hists():
        return [ref.make_histogram(g["iso/lnpi"][k], mom2, gm["beta_ref"], [gm["mu1_ref"], gm["mu1_ref"] + d2], gm["smooth"], volume=gm["volume"])
                for k, d2 in enumerate(gm["dmu2"])]
    pts = [(-3.0, 0.3), (-2.95, 0.5), (-3.05, -0.5), (-3.0, 1.5), (-2.9, 0.9)]
    meta["iso"] = {"points": pts, "order": 1, "m": 2.5, "grid": {"mu1_bounds": [-3.1, -2.9], "dmu2_bounds": [0.0, 1.0], "delta": [0.1, 0.25]}}
    for order in (1,):  # the synthetic moment tensor is max_order 2: second-order extrapolation refuses (gc_hist.pyx:937)
        with redirect_stdout(quiet):
            iso = ns.gc_binary.isopleth(hists(), gm["beta_ref"], order)
        for k, (mu1, d2) in enumerate(pts):
            with redirect_stdout(quiet):
                hm = iso.get_hist(mu1, d2, 2.5)
            out["iso/o%d/get%d/lnpi" % (order, k)] = np.array(hm.data["ln(PI)"])
            out["iso/o%d/get%d/mom_n1" % (order, k)] = np.array(hm.data["mom"][0, 1, 0, 0, 0])
            out["iso/o%d/get%d/mom_u" % (order, k)] = np.array(hm.data["mom"][0, 0, 0, 0, 1])
            out["iso/o%d/get%d/curr_mu" % (order, k)] = np.array(hm.data["curr_mu"])
        with redirect_stdout(quiet):
            iso = ns.gc_binary.isopleth(hists(), gm["beta_ref"], order)
            Z, (X, Y) = iso.make_grid([-3.1, -2.9], [0.0, 1.0], [0.1, 0.25], 2.5)
        out["iso/o%d/grid/x1" % order], out["iso/o%d/grid/density" % order] = np.array(Z), np.array(iso.data["density"])
        out["iso/o%d/grid/fe" % order] = np.array(iso.data["F.E./kT"])
        out["iso/o%d/grid/X" % order], out["iso/o%d/grid/Y" % order] = np.array(X), np.array(Y)
        print("order", order, "grid", Z.shape, "cells filled", int(np.sum(Z != 0)))
    np.savez_compressed(os.path.join(HERE, "api_vectors.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "api_vectors.json"), "w"), indent=1)
    print(json.dumps(meta["coex"], indent=None)[:600])
    print(json.dumps(meta["collect"]))


if __name__ == "__main__":
    main()
