"""CPU ORACLE — numpy restatement of the reference's transient right-hand side.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / reference arm, never by the product path under
``thermca_b200/``.  It restates, on a lowered network (``thermca_b200/spec.py``):

* ``update_time_derivative``                       thermca/solver.py:193-302
* ``Network._set_time_dependent_properties``      thermca/network.py:747-868
* ``solve_stationary_nodes``                       thermca/solver.py:30-42
* ``cond_to_differential_equation_representation`` thermca/solver.py:45-58
* ``FFEIO.update_state_matrix``                    thermca/fem/ffe_io.py:76-82
* ``FESystem.update_material_properties``          thermca/fem/fe_system.py:185-193
* ``MORIO.update_state_matrix``                    thermca/fem/mor_io.py:90-92
* ``LPIO.update_state_matrix`` / ``film_marg_series`` thermca/lpm/lp_io.py:68-103
* ``Input._set_time``                              thermca/input.py:135-146

Time integration is scipy's own ``RK45`` / ``RK23`` driven the way
``simple_solve_ivp`` drives it (thermca/solver.py:348-390) — scipy is the third-party
dependency the reference delegates stepping to (pinned 1.14.1; this image has 1.18.1).
The implicit theta step has no counterpart in the reference ("parity unpinned" for
that scheme): :func:`theta_step_reference` restates it with scipy's sparse direct
solver on the same operators.

Pinned against: golden fixtures in ``tests/golden`` produced by running the
unmodified reference (``oracle/make_golden.py``); see tests/test_oracle_golden.py.
"""
from __future__ import annotations

from bisect import bisect_right

import numpy as np
import scipy.sparse as sp

from program_oracle import eval_program, state_offsets, initial_state  # noqa: F401  (own evaluator: nothing is
#                                                     imported from the product package; initial_state re-exported)


class OracleRHS:
    """Stateful right-hand side: like the reference, ``net_temp``, ``heat_src`` and the
    conductance data persist between calls (bound temperatures and stationary nodes
    lag by one call, thermca/network.py:781-796)."""

    def __init__(self, spec):
        s = self.s = spec
        self.N, self.u = int(s["N"]), int(s["u"])
        self.net_temp = np.array(s["init_temp"], dtype=np.float64)
        self.heat_src = np.array(s["heat_src"], dtype=np.float64)
        self.capy = np.array(s["capy"], dtype=np.float64)
        self.vol_capy = np.array(s["vol_capy"], dtype=np.float64)
        self.condy = np.array(s["condy"], dtype=np.float64)
        self.cond_data = np.array(s["rc_data"], dtype=np.float64)
        self.clean_data = np.array(s["rc_clean_data"], dtype=np.float64)
        self.calc_data = self.cond_data.copy()
        self.indptr, self.indices = s["rc_indptr"], s["rc_indices"]
        self.tables = [(t["x"], t["y"]) for t in s["tables"]]
        self.input_vals = np.full(len(s["inputs"]), np.nan)
        self._in_old_t = [-np.inf] * len(s["inputs"])
        self._in_old_i = [0] * len(s["inputs"])
        self.lp_off, self.ffe_off, self.mor_off, self.n_state = state_offsets(s)
        self.time = None
        # per-part mutable state
        self.ffe = []
        for p in s["ffe"]:
            A = sp.csr_matrix((np.array(p["A_data"]), p["A_body"]["indices"], p["A_body"]["indptr"]),
                              shape=tuple(p["A_body"]["shape"]))
            self.ffe.append({"A": A, "A_body_data": np.array(p["A_body"]["data"]),
                             "films": np.array(p["films"]), "M_inv": np.array(p["M_inv"]),
                             "B_val": [np.array(v) for v in p["B_val"]]})
        self.mor = [{"films": np.array(p["films"]),
                     "A": p["A_body"] + sum(f * A for f, A in zip(p["films"], p["A_films"]))}
                    for p in s["mor"]]
        self.lp = []
        for p in s["lp"]:
            A = sp.csr_matrix((np.array(p["A_data"]), p["A_body"]["indices"], p["A_body"]["indptr"]),
                              shape=tuple(p["A_body"]["shape"]))
            self.lp.append({"A": A, "films": np.array(p["films"]),
                            "L_film_margs": np.array(p["L_film_margs"]),
                            "L_film_margs_sum": np.array(p["L_film_margs"]).sum(axis=0),
                            "M_inv": np.array(p["M_inv"]), "L_data": np.array(p["L_data"]),
                            "A_body_data": np.array(p["A_body"]["data"])})

    # -- helpers -------------------------------------------------------------------
    def _prog(self, k, args):
        p = self.s["programs"][int(k)]
        return eval_program(p["code"], p["consts"], int(p["result"]), args, self.input_vals, self.tables)

    def _interp(self, tab, x):
        xp, fp = self.tables[int(tab)]
        return np.interp(x, xp, fp)

    def _set_inputs(self, time):
        # thermca/input.py:135-146 (bisect_right - 1; index -1 wraps to the last row)
        for k, inp in enumerate(self.s["inputs"]):
            lo = self._in_old_i[k] if time > self._in_old_t[k] else 0
            lo = max(lo, 0)
            self._in_old_t[k] = time
            self._in_old_i[k] = bisect_right(list(inp["t"]), time, lo=lo) - 1
            self.input_vals[k] = inp["v"][self._in_old_i[k]]

    def _lp_series(self, p, st):
        # thermca/lpm/lp_io.py:83-103
        dof, F = int(p["dof"]), len(st["films"])
        L = np.zeros((dof, F))
        for i in range(F):
            film_conds = p["face_areas"][i] * st["films"][i]
            marg_conds = -st["L_data"][p["smx_didx"][i]]
            np.add.at(L[:, i], p["marg_idxs"][i], film_conds * marg_conds / (film_conds + marg_conds))
        return L

    def _lp_update_state(self, p, st):
        # thermca/lpm/lp_io.py:68-80
        st["L_film_margs"] = self._lp_series(p, st)
        st["L_film_margs_sum"] = st["L_film_margs"].sum(axis=0)
        diag = st["M_inv"] * st["L_film_margs"].sum(axis=1)
        data = np.array(st["A_body_data"])
        data[p["A_diag_didx"]] += diag
        st["A"].data = data

    # -- thermca/network.py:747-868 ------------------------------------------------------
    def set_time_dependent_properties(self, time, y):
        s, T = self.s, self.net_temp
        cond, clean = self.cond_data, self.clean_data
        self._set_inputs(time)
        for g in s["matl_groups"]:
            mt = T[g["idx"]]
            self.vol_capy[g["idx"]] = self._interp(g["vol_capy_tab"], mt)
            self.condy[g["idx"]] = self._interp(g["condy_tab"], mt)
        ci = s["capy_idxs"]
        self.capy[ci] = s["vol"][ci] * self.vol_capy[ci]
        fc = s["flow_const"]
        cond[fc["didx"]] = fc["vol_flow"] * (self.vol_capy[fc["i0"]] + self.vol_capy[fc["i1"]]) / 2
        for g in s["flow_funcs"]:
            cond[g["didx"]] = (self._prog(g["prog"], [T[g["i0"]], T[g["i1"]]])
                               * (self.vol_capy[g["i0"]] + self.vol_capy[g["i1"]]) / 2)
        for g in s["film_funcs"]:
            c = self._prog(g["prog"], [T[g["i0"]], T[g["i1"]]]) * g["areas"]
            clean[g["fwd"]] = c
            clean[g["bwd"]] = c
            cond[g["fwd"]] = c
            cond[g["bwd"]] = c
        for g in s["heat_funcs"]:
            for idxs, cnt in zip(g["idxs"], g["counts"]):
                self.heat_src[idxs] = self._prog(g["prog"], [np.array([np.mean(T[idxs])])])[0] / cnt
        for g in s["flux_funcs"]:
            self.heat_src[g["idxs"]] = self._prog(g["prog"], [T[g["idxs"]]]) * g["areas"]
        for g in s["bound_funcs"]:
            T[g["idxs"]] = self._prog(g["prog"], [])[0]
        # films of variable parts from the cleaned network conductance
        for p, st in zip(s["lp"], self.lp):
            if p["var_film_rc_idx"] is not None:
                st["films"] = clean[p["var_film_rc_idx"]] / p["dock_areas"]
        for p, st in zip(s["ffe"], self.ffe):
            if p["var_film_rc_idx"] is not None:
                st["films"] = clean[p["var_film_rc_idx"]] / p["dock_areas"]
        for p, st in zip(s["mor"], self.mor):
            if p["var_film_rc_idx"] is not None:
                st["films"] = clean[p["var_film_rc_idx"]] / p["dock_areas"]
        # body material updates (FFE only; thermca/fem/fe_system.py:185-193, ffe_io.py:67-74)
        for k, (p, st) in enumerate(zip(s["ffe"], self.ffe)):
            td = p["tdep"]
            if td is None:
                continue
            x = y[self.ffe_off[k]: self.ffe_off[k] + int(p["dof"])]
            cond_temp = (x[td["row"]] + x[td["col"]]) / 2.0
            n = int(p["dof"])
            L = sp.csr_matrix((self._interp(td["condy_tab"], cond_temp) * td["Lx_base_data"],
                               (td["row"], td["col"])), shape=(n, n))
            L.data[td["diag_didx"]] = -np.asarray(L.sum(axis=1)).ravel()
            M_diag = self._interp(td["vol_capy_tab"], x) * td["Mx_diag"]
            st["M_inv"] = 1.0 / M_diag
            st["B_val"] = [st["M_inv"][ix] * qv for ix, qv in zip(td["Qs_idx"], td["Qs_val"])]
            Minv = sp.diags(st["M_inv"], format="csr")
            # scipy's csr product drops exact zeros and leaves indices unsorted, exactly as in
            # the reference's own sparse_dot(M_inv, L) (thermca/fem/ffe_io.py:73)
            st["A_body_data"] = (Minv @ L).data
        # LP body material updates (thermca/lpm/lp_system.py:106-118, lp_io.py:59-66)
        for k, (p, st) in enumerate(zip(s["lp"], self.lp)):
            td = p.get("tdep")
            if td is None:
                continue
            dof = int(p["dof"])
            x = y[self.lp_off[k]: self.lp_off[k] + dof]
            T_all = np.empty(int(td["n_rows"]))
            T_all[:dof] = x
            for marg, surf in zip(td["marg_all"], td["surf_all"]):
                T_all[surf] = x[marg]
            cond_temp = (T_all[td["row"]] + T_all[td["col"]]) / 2.0
            L = sp.csr_matrix((self._interp(td["condy_tab"], cond_temp) * td["Lx_base_data"],
                               (td["row"], td["col"])), shape=(int(td["n_rows"]), dof))
            L.data[td["diag_didx"]] = -np.asarray(L[:dof].sum(axis=1)).ravel()
            st["L_data"] = L.data
            st["M_inv"] = 1.0 / (self._interp(td["vol_capy_tab"], x) * td["Mx_diag"])
            st["A_body_data"] = (sp.diags(st["M_inv"], format="csr") @ L[:dof]).data
        for p, st in zip(s["lp"], self.lp):
            if int(p["var"]):
                self._lp_update_state(p, st)
        for l2l in s["lp_to_lp"]:
            p0, p1 = int(l2l["p0"]), int(l2l["p1"])
            st0, st1 = self.lp[p0], self.lp[p1]
            st0["films"][int(l2l["f0"])] = (st1["L_film_margs_sum"][int(l2l["f1"])]
                                            / s["lp"][p0]["surf_areas"][int(l2l["s0"])])
            st1["films"][int(l2l["f1"])] = (st0["L_film_margs_sum"][int(l2l["f0"])]
                                            / s["lp"][p1]["surf_areas"][int(l2l["s1"])])
            self._lp_update_state(s["lp"][p0], st0)
            self._lp_update_state(s["lp"][p1], st1)
        for p, st in zip(s["lp"], self.lp):
            if int(p["var"]):
                cond[p["leak_fwd"]] = st["L_film_margs_sum"]
                cond[p["leak_bwd"]] = st["L_film_margs_sum"]
        for p, st in zip(s["ffe"], self.ffe):
            if p["var_film_rc_idx"] is not None or p["tdep"] is not None:
                data = st["A_body_data"].copy()          # thermca/fem/ffe_io.py:76-82
                for film, didx, adata in zip(st["films"], p["film_didx"], p["film_adata"]):
                    # A_films keep the M_inv of construction time even for temperature
                    # dependent bodies (thermca/fem/ffe_io.py:67-74 does not refresh them)
                    data[didx] += film * adata
                st["A"].data = data
        for p, st in zip(s["mor"], self.mor):
            if p["var_film_rc_idx"] is not None:      # thermca/fem/mor_io.py:90-92
                st["A"] = p["A_body"] + sum(f * A for f, A in zip(st["films"], p["A_films"]))

    # -- thermca/solver.py:193-302 -----------------------------------------------------------
    def __call__(self, time, y):
        s, T, u, N = self.s, self.net_temp, self.u, self.N
        y = np.asarray(y, dtype=np.float64)
        self.time = time
        T[:u] = y[:u]
        for k, (p, st) in enumerate(zip(s["lp"], self.lp)):
            x = y[self.lp_off[k]: self.lp_off[k] + int(p["dof"])]
            T[p["surf_net_idxs"]] = p["C_marg"].T @ x
        for k, p in enumerate(s["ffe"]):
            x = y[self.ffe_off[k]: self.ffe_off[k] + int(p["dof"])]
            T[p["surf_net_idxs"]] = [np.dot(cv, x[ci]) for ci, cv in zip(p["C_idx"], p["C_val"])]
        for k, p in enumerate(s["mor"]):
            S, r = int(p["S"]), int(p["r"])
            x = y[self.mor_off[k]: self.mor_off[k] + S * r].reshape(S, r)
            T[p["surf_net_idxs"]] = np.sum([xs @ C for xs, C in zip(x, p["C_surfs"])], axis=0) + p["ref_temp"]
        self.set_time_dependent_properties(time, y)
        # stationary nodes, Gauss-Seidel in index-list order (thermca/solver.py:30-42)
        ip, ix, cd = self.indptr, self.indices, self.cond_data
        for i in s["stat_idx"]:
            sl = slice(ip[i], ip[i + 1])
            sum_cond = cd[sl].sum()
            if sum_cond == 0:
                continue
            ap = self.heat_src[i] + np.dot(cd[sl], T[ix[sl]])
            T[i] -= T[i] - ap / sum_cond
        # differential-equation form (thermca/solver.py:45-58)
        calc = self.calc_data
        calc[:] = -cd
        for row in range(u):
            calc[s["diag_didx"][row]] = -calc[ip[row]: ip[row + 1]].sum()
        out = []
        if u > 0:
            Luu = np.zeros(u)
            Luk = np.zeros(u)
            for row in range(u):
                for k in range(ip[row], ip[row + 1]):
                    c = ix[k]
                    if c < u:
                        Luu[row] += calc[k] * T[c]
                    else:
                        Luk[row] += calc[k] * T[c]
            out.append((1.0 / self.capy[:u]) * (self.heat_src[:u] - Luk - Luu))
        for k, (p, st) in enumerate(zip(s["lp"], self.lp)):
            x = y[self.lp_off[k]: self.lp_off[k] + int(p["dof"])]
            surf_flux = self.heat_src[p["surf_net_idxs"]] / p["surf_areas"]
            flow = (p["Qs"] @ surf_flux).ravel()
            flow += (st["L_film_margs"] @ T[p["link_elem_net_idxs"]]).ravel()
            out.append(st["M_inv"] * flow - st["A"] @ x)
        for k, (p, st) in enumerate(zip(s["ffe"], self.ffe)):
            x = y[self.ffe_off[k]: self.ffe_off[k] + int(p["dof"])]
            surf_flux = self.heat_src[p["surf_net_idxs"]] / p["surf_areas"]
            np.add.at(surf_flux, p["film_surf_idxs"], st["films"] * T[p["link_elem_net_idxs"]])
            b = np.zeros(int(p["dof"]))
            for sidx in np.nonzero(surf_flux)[0]:
                b[p["B_idx"][sidx]] += st["B_val"][sidx] * surf_flux[sidx]
            out.append(b - st["A"] @ x)
        for k, (p, st) in enumerate(zip(s["mor"], self.mor)):
            S, r = int(p["S"]), int(p["r"])
            x = y[self.mor_off[k]: self.mor_off[k] + S * r].reshape(S, r)
            surf_flux = self.heat_src[p["surf_net_idxs"]] / p["surf_areas"]
            np.add.at(surf_flux, p["film_surf_idxs"],
                      st["films"] * (T[p["link_elem_net_idxs"]] - p["ref_temp"]))
            out.append((np.array([-a @ xs for a, xs in zip(st["A"], x)])
                        + [b * f for b, f in zip(p["B_T"], surf_flux)]).ravel())
        return np.hstack(out) if out else np.zeros(0)


def integrate(spec, time_span, rel_tol=1e-3, abs_tol=1e-6, method="RK45", num_skip=0, **options):
    """Drive scipy's stepper exactly like ``simple_solve_ivp`` + ``step_end_func``
    (thermca/solver.py:307-390).  Returns dict(times, net_temps, io_temps, nfev, nsteps)."""
    from scipy.integrate import RK23, RK45, LSODA
    rhs = OracleRHS(spec)
    y0 = initial_state(spec)
    t0, tf = float(time_span[0]), float(time_span[1])
    u = rhs.u
    rhs(t0, y0)
    times, net_temps, io_temps = [t0], [rhs.net_temp.copy()], [y0[u:].copy()]
    extra = {"net_capys": [rhs.capy.copy()], "net_heat_srcs": [rhs.heat_src.copy()],
             "cond_data": [rhs.cond_data.copy()], "clean_data": [rhs.clean_data.copy()]}
    nfev = 1
    nsteps = 0
    if tf > t0:
        cls = {"RK23": RK23, "RK45": RK45, "LSODA": LSODA}.get(method, method)
        solver = cls(rhs, t0, y0, tf, vectorized=False, rtol=rel_tol, atol=abs_tol, **options)
        result_step, next_coll = 0, num_skip
        status = None
        while status is None:
            msg = solver.step()
            if solver.status == "finished":
                status = 0
            elif solver.status == "failed":
                raise Exception("Solver failed" + str(msg))
            if next_coll == result_step or np.isclose(solver.t, tf):
                rhs.net_temp[:u] = solver.y[:u]
                times.append(solver.t)
                net_temps.append(rhs.net_temp.copy())
                io_temps.append(solver.y[u:].copy())
                extra["net_capys"].append(rhs.capy.copy()); extra["net_heat_srcs"].append(rhs.heat_src.copy())
                extra["cond_data"].append(rhs.cond_data.copy()); extra["clean_data"].append(rhs.clean_data.copy())
                next_coll += num_skip + 1
            result_step += 1
            nsteps += 1
        nfev += solver.nfev
    out = {"times": np.array(times), "net_temps": np.array(net_temps),
           "io_temps": np.array(io_temps), "nfev": nfev, "nsteps": nsteps}
    out.update({k: np.array(v) for k, v in extra.items()})
    return out
