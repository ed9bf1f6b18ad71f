"""Import-time stand-in for scikit-image (absent from this image); the compiled reference pore_hist only needs the
names to exist.  Every function raises: the watershed workflow (pore_hist.pyx:186-477) is out of scope."""
