"""Stub: the reference imports lowess (gc_hist.pyx:14) but only dead code (_lowess_smooth) calls it."""


def lowess(*args, **kwargs):
    raise NotImplementedError("statsmodels is not installed; lowess is not on the hot path")
