def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError("scikit-image is not installed (stub)")
    return _missing
