"""Import environment for the unmodified reference (TEST INFRASTRUCTURE ONLY).

`/root/reference` is importable only in the build container (it does not exist on
the GPU box).  ``activate()`` puts the stand-in third-party modules in
``oracle/shims`` and the reference tree on ``sys.path`` and returns the imported
``thermca`` package.  Nothing under ``thermca_b200/`` may import this module.
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "thermca"))


def activate():
    if not available():
        raise RuntimeError("reference tree not present (only available in the build container)")
    for p in (REFERENCE_ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import thermca  # noqa: F401
    return thermca
