"""Inert stand-in for vtk (thermca/plot/primitives.py:25)."""


def __getattr__(name):
    raise AttributeError(name)
