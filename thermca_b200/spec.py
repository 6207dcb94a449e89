"""Flat, serialisable description of a built network ("lowered network").

A lowered network is a nested ``dict`` of numpy arrays / ints / lists.  It holds
exactly what the transient solve reads off a ``thermca.Network``
(thermca/solver.py:96-191 and the attribute list in SURVEY.md §8b), with the Python
callables replaced by device programs (:mod:`thermca_b200.trace`).  It is what the
C-ABI consumes, what the oracle restates the right-hand side on, and what the golden
fixtures under ``tests/golden`` store, because the reference itself cannot travel to
the GPU box.

``save_spec`` / ``load_spec`` flatten the nesting into ``.npz`` keys like
``ffe/0/A_data``; lists become ``name/<i>/...`` plus ``name/__len__``.
"""
from __future__ import annotations

import numpy as np


def _flatten(obj, prefix, out):
    if isinstance(obj, dict):
        out[prefix + "__dict__"] = np.int64(1)
        for k, v in obj.items():
            _flatten(v, f"{prefix}{k}/", out)
    elif isinstance(obj, (list, tuple)):
        out[prefix + "__len__"] = np.int64(len(obj))
        for i, v in enumerate(obj):
            _flatten(v, f"{prefix}{i}/", out)
    elif obj is None:
        out[prefix + "__none__"] = np.int64(1)
    elif isinstance(obj, str):
        out[prefix + "__str__"] = np.array(obj)
    else:
        out[prefix + "__val__"] = np.asarray(obj)


def _unflatten(keys, data, prefix):
    if prefix + "__val__" in data:
        v = data[prefix + "__val__"]
        return v[()] if v.ndim == 0 else v
    if prefix + "__none__" in data:
        return None
    if prefix + "__str__" in data:
        return str(data[prefix + "__str__"])
    if prefix + "__len__" in data:
        n = int(data[prefix + "__len__"])
        return [_unflatten(keys, data, f"{prefix}{i}/") for i in range(n)]
    if prefix + "__dict__" in data:
        names = []
        plen = len(prefix)
        for k in keys:
            if k.startswith(prefix) and k != prefix + "__dict__":
                name = k[plen:].split("/", 1)[0]
                if name not in names:
                    names.append(name)
        return {name: _unflatten(keys, data, f"{prefix}{name}/") for name in names}
    raise KeyError(prefix)


def save_spec(path, spec, **extra):
    out = {}
    _flatten(spec, "spec/", out)
    for k, v in extra.items():
        _flatten(v, f"{k}/", out)
    np.savez_compressed(path, **out)


def load_spec(path):
    """Returns ``(spec, extras)``; extras are the additional top-level trees."""
    with np.load(path, allow_pickle=False) as z:
        data = {k: z[k] for k in z.files}
    keys = list(data.keys())
    tops = []
    for k in keys:
        t = k.split("/", 1)[0]
        if t not in tops:
            tops.append(t)
    trees = {t: _unflatten(keys, data, t + "/") for t in tops}
    spec = trees.pop("spec")
    return spec, trees


def state_length(spec):
    n = int(spec["u"])
    n += sum(int(p["dof"]) for p in spec["lp"])
    n += sum(int(p["dof"]) for p in spec["ffe"])
    n += sum(int(p["S"]) * int(p["r"]) for p in spec["mor"])
    return n


def state_offsets(spec):
    """Offsets of [net | lp... | ffe... | mor...] blocks in the stacked state
    (thermca/solver.py:151-175)."""
    off = int(spec["u"])
    lp, ffe, mor = [], [], []
    for p in spec["lp"]:
        lp.append(off)
        off += int(p["dof"])
    for p in spec["ffe"]:
        ffe.append(off)
        off += int(p["dof"])
    for p in spec["mor"]:
        mor.append(off)
        off += int(p["S"]) * int(p["r"])
    return lp, ffe, mor, off
