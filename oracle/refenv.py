"""Import environment for the unmodified reference (TEST INFRASTRUCTURE ONLY).

The reference is importable from ``/root/reference`` in the build container and from the staged
copy ``oracle/_ref`` (made by ``oracle/build_ref.py``, git-ignored, shipped with the snapshot) on
the GPU box.  ``activate()`` puts the stand-in third-party modules of ``oracle/shims`` and the
reference tree on ``sys.path`` and returns the imported ``thermca`` package.  Nothing under
``thermca_b200/`` may import this module.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
_CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))


def _find_root():
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "thermca", "solver.py")):
            return root
    return _CANDIDATES[0]


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "thermca", "solver.py"))


def is_staged_copy():
    return available() and REFERENCE_ROOT != _CANDIDATES[0]


def activate():
    if not available():
        raise RuntimeError("reference tree not present (neither /root/reference nor oracle/_ref)")
    for p in (REFERENCE_ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import thermca  # noqa: F401
    return thermca
