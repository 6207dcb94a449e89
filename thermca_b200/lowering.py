"""Lower a built ``thermca.Network`` to the flat description the device solver consumes.

Everything is read from the *built* objects (never re-derived from the model), so node
order, CSR pattern order and scatter targets are the reference's own
(thermca/network.py:56-100,706-745,870-905; thermca/_utils/sparse.py:38-103).  The
attribute list follows SURVEY.md §8b.  Callables are traced into device programs
(:mod:`thermca_b200.trace`).

Duck-typed: no ``import thermca`` here, so the module also works where the reference
is not installed (it only needs an object with the same attributes).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .trace import TableRegistry, trace, UntraceableError  # noqa: F401


# ---- index maps restated from thermca/_utils/sparse.py ---------------------------
def csr_data_idxs(indptr, indices, rows, cols):
    """Position in ``data`` of each dense (row, col): FIRST match in the row
    (thermca/_utils/sparse.py:38-60).  Raises IndexError like the reference."""
    rows = np.asarray(rows, dtype=np.int64).ravel()
    cols = np.asarray(cols, dtype=np.int64).ravel()
    out = np.empty(rows.size, dtype=np.int32)
    for k in range(rows.size):
        lo, hi = int(indptr[rows[k]]), int(indptr[rows[k] + 1])
        hit = np.nonzero(indices[lo:hi] == cols[k])[0]
        if hit.size == 0:
            raise IndexError("No value stored for given index!")
        out[k] = lo + int(hit[0])
    return out


def csr_sub_matrix_idxs(indptr, indices, row_start, row_stop, col_start, col_stop):
    """Sub-matrix (pattern kept, zeros included) and back-map into the source ``data``
    (thermca/_utils/sparse.py:63-103).  Returns (sub_indptr, sub_indices, orig_data_idxs)."""
    sub_indptr, sub_indices, orig = [0], [], []
    for r in range(row_start, row_stop):
        for k in range(int(indptr[r]), int(indptr[r + 1])):
            c = int(indices[k])
            if col_start <= c < col_stop:
                orig.append(k)
                sub_indices.append(c - col_start)
        sub_indptr.append(len(orig))
    return (np.asarray(sub_indptr, dtype=np.int32), np.asarray(sub_indices, dtype=np.int32),
            np.asarray(orig, dtype=np.int32))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _csr(m):
    m = sp.csr_matrix(m)
    return {"indptr": np.asarray(m.indptr, dtype=np.int32),
            "indices": np.asarray(m.indices, dtype=np.int32),
            "data": _f64(m.data), "shape": np.asarray(m.shape, dtype=np.int64)}


def _sparse_columns(dense):
    """Per column: (row indices of non-zeros, values) of a dense (n x S) array."""
    idx, val = [], []
    for col in np.asarray(dense).T:
        nz = np.nonzero(col)[0]
        idx.append(nz.astype(np.int64))
        val.append(_f64(col[nz]))
    return idx, val


class _Lowering:
    def __init__(self, net):
        self.net = net
        self.tables = TableRegistry()
        self.inputs = []          # Input objects, order of net.input_set_methods
        self.input_arrays = []
        self.programs = []
        self._prog_ix = {}

    def program(self, func, n_args, extra=()):
        key = (id(func), n_args) + tuple(id(e) for e in extra)
        if key not in self._prog_ix:
            p = trace(func, n_args, self.input_arrays, self.tables, extra, self.inputs)
            code, consts, result = p.arrays()
            self.programs.append({"code": code, "consts": consts, "result": result,
                                  "n_args": np.int32(n_args)})
            self._prog_ix[key] = len(self.programs) - 1
        return np.int32(self._prog_ix[key])

    def table(self, xp, fp):
        return np.int32(self.tables.index(xp, fp))


def lower_network(net):
    """``thermca.Network`` -> lowered network dict (see :mod:`thermca_b200.spec`)."""
    lo = _Lowering(net)
    N, u = int(net._node_count), int(net._free_temp_lgth)
    rc, crc = net.run_cond, net.clean_run_cond
    if not (np.array_equal(rc.indptr, crc.indptr) and np.array_equal(rc.indices, crc.indices)):
        raise ValueError("run_cond and clean_run_cond patterns differ")
    indptr = np.asarray(rc.indptr, dtype=np.int32)
    indices = np.asarray(rc.indices, dtype=np.int32)

    # -- inputs (thermca/input.py:135-146) --------------------------------------
    for meth in net.input_set_methods:
        inp = meth.__self__
        lo.inputs.append(inp)
        lo.input_arrays.append(inp._value)
    inputs = [{"t": _f64(i._time_table[:, 0]), "v": _f64(i._time_table[:, 1])} for i in lo.inputs]

    spec = {
        "N": np.int64(N), "u": np.int64(u),
        "init_temp": _f64(net.init_temp), "heat_src": _f64(net.heat_src),
        "capy": _f64(net.capy), "vol": _f64(net._vol), "vol_capy": _f64(net._vol_capy),
        "condy": _f64(net.condy), "posn": _f64(net.posn),
        "rc_indptr": indptr, "rc_indices": indices,
        "rc_data": _f64(rc.data), "rc_clean_data": _f64(crc.data),
        "stat_idx": _i64(net.stationary_elem_idxs),
        "inputs": inputs,
    }
    # maps computed inside transient_solution (thermca/solver.py:116-122)
    diag = np.arange(u)
    spec["diag_didx"] = csr_data_idxs(indptr, indices, diag, diag)
    _, _, spec["uu_didx"] = csr_sub_matrix_idxs(indptr, indices, 0, u, 0, u)
    _, _, spec["uk_didx"] = csr_sub_matrix_idxs(indptr, indices, 0, u, u, N)

    # -- temperature dependent MatlNodes (thermca/network.py:761-768) -------------
    groups = []
    for matl, idx in net._all_temp_dependent_matl_idxs.items():
        groups.append({
            "idx": _i64(idx),
            "vol_capy_tab": lo.table(matl._vol_capy_temp, matl._vol_capy),
            "condy_tab": lo.table(matl._condy[0], matl._condy[1]),
            "limits": _f64(matl.limits),
        })
    spec["matl_groups"] = groups
    spec["capy_idxs"] = _i64(net._all_temp_dependent_capy_idxs)

    # -- flow links (thermca/network.py:769-779) -----------------------------------
    didx, i0, i1, vf = net._flow_link_data
    spec["flow_const"] = {"didx": _i64(didx), "i0": _i64(i0), "i1": _i64(i1), "vol_flow": _f64(vf)}
    flow_funcs = []
    for func, per_fluid in net._flow_link_func_data.items():
        for fluid, (didx, i0, i1, vf) in per_fluid.items():
            flow_funcs.append({"prog": lo.program(func, 2, (fluid,)), "didx": _i64(didx),
                               "i0": _i64(i0), "i1": _i64(i1)})
    spec["flow_funcs"] = flow_funcs

    # -- film / conductance functions (thermca/network.py:781-785) -----------------
    film_funcs = []
    for func, per_matl in net._film_link_func_data.items():
        for matl, (fwd, bwd, i0, i1, areas) in per_matl.items():
            film_funcs.append({"prog": lo.program(func, 2, (matl,)), "fwd": _i64(fwd),
                               "bwd": _i64(bwd), "i0": _i64(i0), "i1": _i64(i1),
                               "areas": _f64(areas)})
    spec["film_funcs"] = film_funcs

    # -- heat / flux / bound-temperature functions (thermca/network.py:787-796) ----
    heat_funcs = []
    for func, (idxs_arr, counts) in net._heat_srce_func_data.items():
        idxs_arr = np.asarray(idxs_arr, dtype=np.int64)
        if idxs_arr.ndim != 2:
            raise NotImplementedError("ragged HeatSource index groups")
        heat_funcs.append({"prog": lo.program(func, 1), "idxs": np.ascontiguousarray(idxs_arr),
                           "counts": _f64(counts)})
    spec["heat_funcs"] = heat_funcs
    spec["flux_funcs"] = [{"prog": lo.program(func, 1), "idxs": _i64(idxs), "areas": _f64(areas)}
                          for func, (idxs, areas) in net._flux_srce_func_data.items()]
    spec["bound_funcs"] = [{"prog": lo.program(func, 0), "idxs": _i64(idxs)}
                           for func, idxs in net._bound_temp_func_idxs.items()]

    # -- parts ------------------------------------------------------------------------
    spec["lp"] = [_lower_lp(lo, net, k, io) for k, io in enumerate(net._lp_ios)]
    spec["ffe"] = [_lower_ffe(lo, net, k, io) for k, io in enumerate(net._ffe_ios)]
    spec["mor"] = [_lower_mor(lo, net, k, io) for k, io in enumerate(net._mor_ios)]

    # two linked LP parts (thermca/network.py:831-852)
    lp2lp = []
    for elems, surf_idx, idx_idx in zip(net.run_lp_to_lp_link_elems, net.run_lp_to_lp_surf_idxs,
                                        net.run_lp_to_lp_link_idxs_idxs):
        lp2lp.append({"p0": np.int32(net._lp_parts.index(elems[0])),
                      "p1": np.int32(net._lp_parts.index(elems[1])),
                      "s0": np.int32(surf_idx[0]), "s1": np.int32(surf_idx[1]),
                      "f0": np.int32(idx_idx[0]), "f1": np.int32(idx_idx[1])})
    spec["lp_to_lp"] = lp2lp

    spec["programs"] = lo.programs
    spec["tables"] = [{"x": x, "y": y} for x, y in lo.tables.tables]
    return spec


def _find(lst, obj):
    for k, o in enumerate(lst):
        if o is obj:
            return k
    return -1


def _lower_lp(lo, net, k, io):
    sysm = io.lp_sys
    dof = int(sysm.dof)
    F = len(io.films)
    p = {
        "dof": np.int64(dof), "init_temp": np.float64(sysm.init_temp),
        "A_body": _csr(io.A_body), "A_data": _f64(io.A.data),
        "A_diag_didx": _i64(io.A_diag_data_idxs), "M_inv": _f64(io.M_inv.data),
        "Qs": _f64(sysm.Qs), "C_marg": _f64(io.C_marg), "surf_areas": _f64(sysm.surf_areas),
        "films": _f64(io.films), "film_surf_idxs": _i64(io.film_surf_idxs),
        "surf_net_idxs": _i64(io.surf_net_idxs), "link_elem_net_idxs": _i64(io.link_elem_net_idxs),
        "L_film_margs": _f64(io.L_film_margs), "L_data": _f64(sysm.L.data),
        "face_areas": [_f64(sysm.surf_face_areas[fi]) for fi in io.film_surf_idxs],
        "smx_didx": [_i64(sysm.surf_to_marg_data_idxs[fi]) for fi in io.film_surf_idxs],
        "marg_idxs": [_i64(sysm.marg_dof_idxs[fi]) for fi in io.film_surf_idxs],
        "dock_areas": _f64(sysm.surf_areas[io.film_surf_idxs]) if F else _f64([]),
    }
    j = _find(net._var_film_lp_ios, io)
    p["var_film_rc_idx"] = _i64(net._var_lp_run_cond_film_idxs[j]) if j >= 0 else None
    j = _find(net._var_lp_ios, io)
    p["var"] = np.int32(j >= 0)
    p["leak_fwd"] = _i64(net._var_lp_run_cond_fwd_idxs[j]) if j >= 0 else None
    p["leak_bwd"] = _i64(net._var_lp_run_cond_bwd_idxs[j]) if j >= 0 else None
    if _find(net._var_body_lp_ios, io) >= 0:
        # LPSystem.update_material_properties (thermca/lpm/lp_system.py:106-118)
        matl = sysm.matl
        coo = sysm.Lx_base_coo
        L = sp.csr_matrix(sysm.L)
        lookup = {}
        for r in range(L.shape[0]):
            for k in range(L.indptr[r], L.indptr[r + 1]):
                lookup.setdefault((r, int(L.indices[k])), k)
        ab = sp.csr_matrix(io.A_body)
        ab_rows = np.repeat(np.arange(ab.shape[0]), np.diff(ab.indptr))
        from_L = np.array([lookup[(int(r), int(c))] for r, c in zip(ab_rows, ab.indices)], dtype=np.int64)
        p["tdep"] = {
            "Lx_base_data": _f64(sysm.Lx_base.data), "row": _i64(coo.row), "col": _i64(coo.col),
            "L_indptr": _i64(L.indptr), "n_rows": np.int64(L.shape[0]),
            "diag_didx": _i64(sysm.L_diag_data_idxs), "Mx_diag": _f64(sysm.Mx_diag),
            "condy_tab": lo.table(matl._condy[0], matl._condy[1]),
            "vol_capy_tab": lo.table(matl._vol_capy_temp, matl._vol_capy),
            "marg_all": [_i64(m) for m in sysm.marg_dof_idxs], "surf_all": [_i64(x) for x in sysm.surf_dof_idxs],
            "abody_from_L": from_L, "abody_rows": _i64(ab_rows),
        }
    else:
        p["tdep"] = None
    return p


def _lower_ffe(lo, net, k, io):
    sysm = io.fe_sys
    n = int(sysm.dof)
    F = len(io.films)
    B_idx, B_val = _sparse_columns(io.B)
    C_idx, C_val = _sparse_columns(io.C_surf)
    p = {
        "dof": np.int64(n), "init_temp": np.float64(sysm.init_temp),
        "A_body": _csr(io.A_body), "A_data": _f64(io.A.data), "M_inv": _f64(io.M_inv.data),
        "B_idx": B_idx, "B_val": B_val, "C_idx": C_idx, "C_val": C_val,
        "surf_areas": _f64(sysm.surf_areas), "S": np.int64(len(sysm.surf_areas)),
        "films": _f64(io.films), "film_surf_idxs": _i64(io.film_surf_idxs),
        "surf_net_idxs": _i64(io.surf_net_idxs), "link_elem_net_idxs": _i64(io.link_elem_net_idxs),
        "film_didx": [_i64(ix) for ix in io.A_data_films_idxs] if F else [],
        "film_adata": [_f64(Af.data) for Af in io.A_films] if F else [],
        "dock_areas": _f64(sysm.surf_areas[io.film_surf_idxs]) if F else _f64([]),
    }
    j = _find(net._var_film_ffe_io, io)
    p["var_film_rc_idx"] = _i64(net._var_ffe_run_cond_idxs[j]) if j >= 0 else None
    if _find(net._var_body_ffe_ios, io) >= 0:
        matl = sysm.matl
        coo = sysm.Lx_base_coo
        p["tdep"] = {
            "Lx_base_data": _f64(sysm.Lx_base.data), "row": _i64(coo.row), "col": _i64(coo.col),
            "diag_didx": _i64(sysm.L_diag_data_idxs), "Mx_diag": _f64(sysm.Mx_diag),
            "Qs_idx": _sparse_columns(sysm.Qs)[0], "Qs_val": _sparse_columns(sysm.Qs)[1],
            "condy_tab": lo.table(matl._condy[0], matl._condy[1]),
            "vol_capy_tab": lo.table(matl._vol_capy_temp, matl._vol_capy),
            "limits": _f64(matl.limits),
        }
    else:
        p["tdep"] = None
    return p


def _lower_mor(lo, net, k, io):
    sysm = io.fe_sys
    S = len(sysm.surf_names)
    r = int(io.mor_dof)
    x0 = np.vstack(io.VT0) @ (np.full(sysm.dof, sysm.init_temp - io.ref_temp) / S)
    F = len(io.films)
    p = {
        "S": np.int64(S), "r": np.int64(r), "F": np.int64(F),
        "A_body": _f64(io.A_body), "A_films": _f64(io.A_films) if F else np.zeros((0, S, r, r)),
        "B_T": _f64(io.B_T), "C_surfs": _f64(io.C_surfs), "ref_temp": np.float64(io.ref_temp),
        "films": _f64(io.films), "film_surf_idxs": _i64(io.film_surf_idxs),
        "surf_net_idxs": _i64(io.surf_net_idxs), "link_elem_net_idxs": _i64(io.link_elem_net_idxs),
        "surf_areas": _f64(sysm.surf_areas), "x0": _f64(x0),
        "dock_areas": _f64(sysm.surf_areas[io.film_surf_idxs]) if F else _f64([]),
    }
    j = _find(net._var_film_mor_ios, io)
    p["var_film_rc_idx"] = _i64(net._var_mor_run_cond_idxs[j]) if j >= 0 else None
    return p


def initial_state(spec):
    """Stacked initial vector [T_u, lp.., ffe.., mor..] (thermca/solver.py:132-151)."""
    parts = [np.asarray(spec["init_temp"][: int(spec["u"])], dtype=np.float64)]
    for p in spec["lp"]:
        parts.append(np.full(int(p["dof"]), float(p["init_temp"])))
    for p in spec["ffe"]:
        parts.append(np.full(int(p["dof"]), float(p["init_temp"])))
    for p in spec["mor"]:
        parts.append(np.asarray(p["x0"], dtype=np.float64).ravel())
    return np.hstack(parts)
