#!/usr/bin/env python
"""Golden vectors for the window-patching shift solve from the COMPILED reference
(oracle/_ref/fhmc_patch, built from moments/win_patch/fhmc_patch.pyx by oracle/build_ref.py: ``patch_window_pair`` and
``window_patch_error`` turned from cdef into def, nothing else).  Run in the build container only.

Inputs: the reference's own fixture windows unittests/reference/test_sim/{1,2,3} (``get_patch_sequence`` ->
``window(...)`` exactly as unittests/moments_win_patch_fhmc.py:289-312, 520-528 drives them), every adjacent pair patched
with offsets 1 and 2, plus synthetic window pairs (long overlaps, large shifts) built with ``window.__new__``.
Recorded: the windows' lb / ub / offset / lnPI arrays, the reference's (shift, err2) and its objective evaluated at a few
trial shifts.  Writes tests/golden/patch_vectors.npz."""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REF = os.environ.get("FHMC_REFERENCE", "/root/reference")


def main():
    assert ref.load() is not None, ref._cache.get("error")
    import importlib
    fp = importlib.import_module("fhmc_patch")
    out = {}
    cases = []
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "unittests"))
    try:
        seq = list(fp.get_patch_sequence("reference/test_sim/"))
        for offset in (1, 2):
            wins = [fp.window(s[0], s[1], s[2], s[3], offset, False) for s in seq]
            wins.sort(key=lambda w: w.lb)
            for k in range(len(wins) - 1):
                cases.append(("sim_off%d_%d" % (offset, k), wins[k + 1], wins[k]))
    finally:
        os.chdir(cwd)
    rng = np.random.default_rng(77)
    for k, (lb1, ub1, ub2, shift, noise, offset) in enumerate([(40, 160, 70, 12.5, 1e-3, 2), (100, 900, 400, -733.0, 5e-2, 2),
                                                              (5, 60, 12, 0.25, 0.0, 1), (300, 2000, 1500, 3.0e4, 1.0, 3)]):
        full = np.cumsum(rng.normal(0.3, 1.0, size=ub1 + 1))
        w1, w2 = fp.window.__new__(fp.window), fp.window.__new__(fp.window)
        w2.lb, w2.ub, w2.offset = 0, ub2, offset
        w2.lnPI = full[:ub2 + 1] + noise * rng.normal(size=ub2 + 1)
        w1.lb, w1.ub, w1.offset = lb1, ub1, offset
        w1.lnPI = full[lb1:ub1 + 1] - shift + noise * rng.normal(size=ub1 - lb1 + 1)
        cases.append(("syn_%d" % k, w1, w2))
    names = []
    for name, w1, w2 in cases:
        try:
            with redirect_stdout(io.StringIO()):
                shift, err2 = fp.patch_window_pair(w1, w2)
        except AssertionError as e:       # (a pair the reference refuses: kept as an error-behaviour case)
            names.append(name)
            for k, w in (("w1", w1), ("w2", w2)):
                out["%s/%s/lnpi" % (name, k)] = np.asarray(w.lnPI, dtype=np.float64)
                out["%s/%s/meta" % (name, k)] = np.array([w.lb, w.ub, w.offset], dtype=np.int64)
            out[name + "/raises"] = np.array([str(e)])
            print(name, "lb/ub", w1.lb, w1.ub, w2.lb, w2.ub, "AssertionError:", e)
            continue
        index = w2.ub - w1.lb + 1
        s1 = np.asarray(w1.lnPI[w1.offset:index - w1.offset], dtype=np.float64)
        s2 = np.asarray(w2.lnPI[len(w2.lnPI) - index + w1.offset:len(w2.lnPI) - w1.offset], dtype=np.float64)
        trial = np.array([shift, shift + 0.5, 0.0, -3.25])
        obj = np.array([fp.window_patch_error(float(x), s1, s2) for x in trial])
        names.append(name)
        for k, w in (("w1", w1), ("w2", w2)):
            out["%s/%s/lnpi" % (name, k)] = np.asarray(w.lnPI, dtype=np.float64)
            out["%s/%s/meta" % (name, k)] = np.array([w.lb, w.ub, w.offset], dtype=np.int64)
        out[name + "/ref"] = np.array([shift, err2])
        out[name + "/trial"], out[name + "/obj"] = trial, obj
        print(name, "lb/ub", w1.lb, w1.ub, w2.lb, w2.ub, "overlap", len(s1), "shift %.9g err2 %.6g" % (shift, err2))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "patch_vectors.npz"), **out)


if __name__ == "__main__":
    main()
