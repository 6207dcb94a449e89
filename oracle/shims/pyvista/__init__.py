"""Inert stand-in for pyvista (visualisation is out of scope; only import-time names)."""


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


def __getattr__(name):
    return _Inert
