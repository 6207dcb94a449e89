"""CPU ORACLE — independent evaluator of lowered link programs and of the stacked-state layout.

TEST INFRASTRUCTURE ONLY (see ``rhs_oracle.py``).  The lowered network (``thermca_b200/spec.py``) stores
every film / conductance / heat / flux / bound-temperature callable (thermca/network.py:773-796, library in
thermca/lib/*) as a straight-line fp64 program.  The product evaluates such programs in ``csrc/vm.cuh`` (device)
and ``thermca_b200.trace.run_program`` (lowering-time validation).  This module is a SECOND evaluator that
shares nothing with them but the storage format: the opcode names below are the format's vocabulary
(``tests/test_host_logic.py`` asserts the numbering equals the product's), each name is bound to the numpy
function the reference closures themselves call, and evaluation is a table lookup instead of the product's
``if`` chain.  ``tests/test_link_programs.py`` pins it against values of the UNMODIFIED reference closures.

``state_offsets`` / ``initial_state`` restate thermca/solver.py:132-175 (stacked vector
``[T_u | lp… | ffe… | mor…]``).
"""
from __future__ import annotations

import numpy as np

OPCODES = (
    "ARG CONST INPUT "
    "ADD SUB MUL DIV POW MIN MAX LT LE GT GE EQ NE AND OR "
    "NEG ABS SQRT EXP LOG LOG10 SIN COS TAN TANH ARCTAN NOT SQUARE RECIP SIGN CBRT LOG2 EXP2 FLOOR CEIL ISNAN "
    "WHERE INTERP ARCTAN2 FMOD"
).split()

_as_f = lambda m: m.astype(np.float64)
_truth = lambda x: x != 0

_BINARY = {
    "ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.true_divide, "POW": np.power,
    "MIN": np.minimum, "MAX": np.maximum, "ARCTAN2": np.arctan2, "FMOD": np.fmod,
    "LT": lambda x, y: _as_f(np.less(x, y)), "LE": lambda x, y: _as_f(np.less_equal(x, y)),
    "GT": lambda x, y: _as_f(np.greater(x, y)), "GE": lambda x, y: _as_f(np.greater_equal(x, y)),
    "EQ": lambda x, y: _as_f(np.equal(x, y)), "NE": lambda x, y: _as_f(np.not_equal(x, y)),
    "AND": lambda x, y: _as_f(np.logical_and(_truth(x), _truth(y))),
    "OR": lambda x, y: _as_f(np.logical_or(_truth(x), _truth(y))),
}
_UNARY = {
    "NEG": np.negative, "ABS": np.absolute, "SQRT": np.sqrt, "EXP": np.exp, "LOG": np.log, "LOG10": np.log10,
    "SIN": np.sin, "COS": np.cos, "TAN": np.tan, "TANH": np.tanh, "ARCTAN": np.arctan,
    "NOT": lambda x: _as_f(np.logical_not(_truth(x))), "SQUARE": lambda x: np.multiply(x, x),
    "RECIP": lambda x: np.true_divide(1.0, x), "SIGN": np.sign, "CBRT": np.cbrt, "LOG2": np.log2,
    "EXP2": np.exp2, "FLOOR": np.floor, "CEIL": np.ceil, "ISNAN": lambda x: _as_f(np.isnan(x)),
}


def eval_program(code, consts, result, args, inputs, tables):
    """Value of instruction ``result`` of the program ``code`` (rows ``(op, a, b, c)``) for the argument
    vectors ``args``; ``inputs`` are the current ``Input.value``s (thermca/input.py:127-146), ``tables`` the
    material tables ``(x, y)`` evaluated with ``numpy.interp`` as thermca/materials.py:120-130 does."""
    width = np.atleast_1d(np.asarray(args[0])).shape[0] if len(args) else 1
    leaf = lambda x: np.broadcast_to(np.asarray(x, dtype=np.float64), (width,))
    slots = []
    with np.errstate(all="ignore"):
        for op, a, b, c in code:
            name = OPCODES[int(op)]
            if name in _BINARY:
                out = _BINARY[name](slots[a], slots[b])
            elif name in _UNARY:
                out = _UNARY[name](slots[a])
            elif name == "ARG":
                out = leaf(args[a])
            elif name == "CONST":
                out = leaf(consts[a])
            elif name == "INPUT":
                out = leaf(inputs[a])
            elif name == "WHERE":
                out = np.where(_truth(slots[a]), slots[b], slots[c])
            elif name == "INTERP":
                xs, ys = tables[a]
                out = np.interp(slots[b], xs, ys)
            else:  # pragma: no cover - the vocabulary above is complete
                raise ValueError(f"unknown opcode {op}")
            slots.append(out)
    return np.array(slots[int(result)], dtype=np.float64)


def state_offsets(spec):
    """Start of every part block in the stacked state and its total length (thermca/solver.py:151-175):
    unknown network nodes first, then LP, full-order FE and MOR parts, each in model order."""
    pos = int(spec["u"])
    starts = {"lp": [], "ffe": [], "mor": []}
    for kind in ("lp", "ffe", "mor"):
        for part in spec[kind]:
            starts[kind].append(pos)
            pos += int(part["S"]) * int(part["r"]) if kind == "mor" else int(part["dof"])
    return starts["lp"], starts["ffe"], starts["mor"], pos


def initial_state(spec):
    """Stacked initial vector (thermca/solver.py:132-151): initial temperatures of the unknown nodes,
    uniform part temperatures, and the reference's own projected MOR start vector ``x0``."""
    u = int(spec["u"])
    blocks = [np.array(spec["init_temp"][:u], dtype=np.float64)]
    blocks += [np.full(int(p["dof"]), float(p["init_temp"])) for p in spec["lp"]]
    blocks += [np.full(int(p["dof"]), float(p["init_temp"])) for p in spec["ffe"]]
    blocks += [np.array(p["x0"], dtype=np.float64).reshape(-1) for p in spec["mor"]]
    return np.concatenate(blocks)
