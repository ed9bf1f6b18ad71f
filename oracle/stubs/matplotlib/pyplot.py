"""Stub: gc_binary.pyx:10 imports pyplot; only the out-of-scope plotting helpers use it."""


def __getattr__(name):
    raise NotImplementedError("matplotlib is not installed (stub used by the reference oracle)")
