"""Import-path shim: lets code written for the reference (``import FHMCAnalysis.moments.histogram.one_dim.ntot.gc_hist``,
unittests/moments_histogram_one_dim_gc_ntot.py:10-12) run unchanged on the B200 implementation."""
import importlib
import sys

_MODULES = [
    "moments", "moments.histogram", "moments.histogram.one_dim", "moments.histogram.one_dim.ntot",
    "moments.histogram.one_dim.ntot.gc_hist", "moments.histogram.one_dim.ntot.gc_binary",
    "moments.histogram.one_dim.ntot.collect", "moments.histogram.one_dim.n1", "moments.histogram.one_dim.n1.gc_hist",
    "moments.histogram.two_dim", "moments.histogram.two_dim.joint_hist",
    "moments.histogram.two_dim.h_ntot", "moments.histogram.two_dim.h_ntot.pore_hist",
    "moments.win_patch", "moments.win_patch.fhmc_patch",
]
for _m in _MODULES:
    try:
        sys.modules[__name__ + "." + _m] = importlib.import_module("fhmcanalysis_b200." + _m)
    except ImportError:  # optional sub-module not present
        pass
moments = sys.modules[__name__ + ".moments"]
