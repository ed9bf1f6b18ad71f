"""Loader for the compiled reference oracle in oracle/_ref (TEST INFRASTRUCTURE ONLY).

``load()`` returns a namespace with the reference's own modules ``gc_hist``, ``gc_binary`` and
``joint_hist`` (jeetain/FHMCAnalysis, compiled by oracle/build_ref.py), or ``None`` when they
are not available (e.g. /root/reference absent and nothing prebuilt).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def available():
    return load() is not None


def load(build=True):
    if "ns" in _cache:
        return _cache["ns"]
    ns = None
    try:
        if build:
            sys.path.insert(0, HERE)
            try:
                import build_ref
                build_ref.build()
            finally:
                sys.path.remove(HERE)
        ref_dir = os.path.join(HERE, "_ref")
        stubs = os.path.join(HERE, "stubs")
        added = []
        for mod, path in (("netCDF4", stubs), ("statsmodels", stubs), ("matplotlib", stubs)):
            try:
                importlib.import_module(mod)
            except Exception:
                if path not in sys.path:
                    sys.path.append(path)
                    added.append(path)
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        ns = types.SimpleNamespace(
            gc_hist=importlib.import_module("gc_hist"),
            gc_binary=importlib.import_module("gc_binary"),
            joint_hist=importlib.import_module("joint_hist"),
            gc_hist_n1=importlib.import_module("gc_hist_n1"),
        )
    except Exception as e:  # pragma: no cover - depends on environment
        _cache["error"] = repr(e)
        ns = None
    _cache["ns"] = ns
    return ns


def make_histogram(lnpi, mom, beta_ref, mu_ref, smooth, volume=1.0, ntot=None, ke=False):
    """Construct a reference ``histogram`` from arrays without going through a file
    (fills the same keys ``reload`` does, gc_hist.pyx:104-182)."""
    import copy
    import numpy as np
    ns = load()
    h = ns.gc_hist.histogram.__new__(ns.gc_hist.histogram)
    mu_ref = np.atleast_1d(np.array(mu_ref, dtype=np.float64))
    h.metadata = {"beta_ref": float(beta_ref), "mu_ref": mu_ref.copy(), "nspec": len(mu_ref),
                  "smooth": int(smooth), "fname": "", "used_ke": bool(ke), "file_history": "synthetic"}
    lnpi = np.array(lnpi, dtype=np.float64)
    n = len(lnpi)
    if ntot is None:
        ntot = np.arange(n, dtype=np.int64)
    mom = np.array(mom, dtype=np.float64)
    h.data = {"curr_mu": mu_ref.copy(), "curr_beta": float(beta_ref), "nspec": len(mu_ref),
              "ln(PI)": lnpi, "max_order": mom.shape[1] - 1, "volume": float(volume),
              "ntot": np.array(ntot, dtype=np.int64), "lb": int(ntot[0]), "ub": int(ntot[-1]),
              "pk_hist": {}, "e_hist": {}, "mom": mom}
    return h
