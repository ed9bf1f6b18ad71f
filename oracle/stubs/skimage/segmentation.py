def find_boundaries(*a, **k):
    raise RuntimeError("scikit-image stub: find_boundaries is not available")
