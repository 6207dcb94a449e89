"""CPU restatement of the adaptive ROS2 W-method of csrc/solver.cu (tc_ros2_*).  TEST INFRASTRUCTURE ONLY.

The reference has no implicit scheme of its own (it hands LSODA requests to scipy), so this path is "unpinned" by
the reference's tests; this module pins the DEVICE implementation instead: the same right-hand side
(``rhs_oracle.OracleRHS`` = ``update_time_derivative``, thermca/solver.py:193-302), the same stage equations,
block structure, explicit/implicit switch and step controller, with sparse / dense direct solves in place of the
device's PCG.  Trajectories must agree to solver tolerance (tests/test_gpu_parity.py).

One attempt (Verwer et al. 1999, gamma = 1 + 1/sqrt(2), W = -blockdiag(A_p)):
    (I + g h A) k1 = f(t, y)
    (I + g h A) k2 = f(t + h, y + h k1) - 2 k1
    y+ = y + h (1.5 k1 + 0.5 k2),   err = h (0.5 k1 + 0.5 k2)
    f(t + h, y+) is evaluated at the end of every attempt: it is the next f(t, y) after an accepted step and leaves
    the network state (surface means, conductances, sources) the stored snapshot reports at the new state.
Step control: scipy's RK controller (scipy/integrate/_ivp/rk.py:111-165) with error exponent -1/2.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from rhs_oracle import OracleRHS
from program_oracle import initial_state

GAMMA = 1.0 + 1.0 / np.sqrt(2.0)


def _diag_first(A):
    """first stored diagonal entry per row (thermca/fem/ffe_io.py:52)"""
    A = sp.csr_matrix(A)
    out = np.zeros(A.shape[0])
    for i in range(A.shape[0]):
        for k in range(A.indptr[i], A.indptr[i + 1]):
            if A.indices[k] == i:
                out[i] = A.data[k]
                break
    return out


class Ros2Oracle:
    def __init__(self, spec, bound=2.0):
        self.s = spec
        self.rhs = OracleRHS(spec)
        self.bound = float(bound)
        self.u = int(spec["u"])
        r = self.rhs
        self.n = r.n_state
        # gate diagonals: what the device passes as `adiag` of each PCG block
        self.ffe_gate, self.ffe_frozen = [], []
        for p, st in zip(spec["ffe"], r.ffe):
            A0 = sp.csr_matrix((np.array(p["A_data"]), p["A_body"]["indices"], p["A_body"]["indptr"]),
                               shape=tuple(p["A_body"]["shape"]))
            self.ffe_gate.append(_diag_first(A0))
            self.ffe_frozen.append(A0 if p.get("tdep") is not None else None)
        self.solves = 0
        # films that lead to a STATIONARY node: rank-one Jacobian terms (csrc/solver.cu enqueue_jac_prepare)
        stat = set(int(i) for i in spec["stat_idx"])
        self.jac_films = []                              # (kind, part, film, C column)
        offs = {"lp": r.lp_off, "ffe": r.ffe_off, "mor": r.mor_off}
        for kind in ("lp", "ffe", "mor"):
            for k, p in enumerate(spec[kind]):
                for f in range(len(p["films"])):
                    if int(p["link_elem_net_idxs"][f]) not in stat:
                        continue
                    sidx = int(p["film_surf_idxs"][f])
                    c = np.zeros(self.n)
                    o = offs[kind][k]
                    if kind == "lp":
                        c[o:o + int(p["dof"])] = np.asarray(p["C_marg"])[:, sidx]
                    elif kind == "ffe":
                        c[o + np.asarray(p["C_idx"][sidx])] = np.asarray(p["C_val"][sidx])
                    else:
                        S, rr = int(p["S"]), int(p["r"])
                        c[o:o + S * rr] = np.asarray(p["C_surfs"]).reshape(S * rr, S)[:, sidx]
                    self.jac_films.append((kind, k, f, c))
        self._jac = None
        # a block with such a film is always solved implicitly: the rank-one term alone (without the film's own
        # diagonal term inside B) would be a destabilising Jacobian
        self.jac_parts = set((kind, k) for kind, k, _, _ in self.jac_films)

    def _jac_prepare(self, s):
        """Once per attempt.  The temperature of a stationary node n follows its neighbours instantly
        (thermca/solver.py:30-42): T_n = (q + sum_k g_k T_k) / G, so a film f between a part surface and n adds the
        rank-one term  u_f c_f^T  to the part's Jacobian (u_f: the film's load vector times theta = g_f / G, c_f:
        the surface-mean weights).  W = B - s sum_f u_f c_f^T is applied through the Woodbury identity around the
        block solves B^-1 the scheme already has:  W^-1 r = z + Y (I - s C^T Y)^-1 s C^T z,  z = B^-1 r,
        Y = B^-1 U."""
        self._jac = None
        if not self.jac_films:
            return
        r, spec = self.rhs, self.s
        ip, ix, cd = r.indptr, r.indices, r.cond_data
        K = len(self.jac_films)
        U = np.zeros((self.n, K))
        C = np.zeros((self.n, K))
        for j, (kind, k, f, c) in enumerate(self.jac_films):
            p = spec[kind][k]
            st = getattr(r, kind)[k]
            n = int(p["link_elem_net_idxs"][f])
            sn = int(np.asarray(p["surf_net_idxs"])[int(p["film_surf_idxs"][f])])
            G = g = 0.0
            for kk in range(ip[n], ip[n + 1]):
                G += cd[kk]
                if ix[kk] == sn:
                    g += cd[kk]
            theta = g / G if G > 0.0 else 0.0
            C[:, j] = c
            if kind == "lp":
                o, dof = r.lp_off[k], int(p["dof"])
                U[o:o + dof, j] = theta * st["M_inv"] * np.asarray(st["L_film_margs"]).reshape(dof, -1)[:, f]
            elif kind == "ffe":
                o = r.ffe_off[k]
                rows = np.searchsorted(np.asarray(p["A_body"]["indptr"]), np.asarray(p["film_didx"][f]), side="right") - 1
                np.add.at(U[:, j], o + rows, (theta * st["films"][f]) * np.asarray(p["film_adata"][f]))
            else:
                rr, q = int(p["r"]), int(p["film_surf_idxs"][f])
                o = r.mor_off[k] + q * rr
                U[o:o + rr, j] = (theta * st["films"][f]) * np.asarray(p["B_T"][q])
        Y = np.column_stack([self._block_solve_raw(s, U[:, j]) for j in range(K)])
        Ginv = np.linalg.inv(np.eye(K) - s * (C.T @ Y))
        self._jac = (C, Y, Ginv)

    def _block_solve(self, s, rhs):
        z = self._block_solve_raw(s, rhs)
        if self._jac is not None:
            C, Y, Ginv = self._jac
            z = z + Y @ (Ginv @ (s * (C.T @ z)))
        return z

    def _block_solve_raw(self, s, rhs):
        r, spec, u = self.rhs, self.s, self.u
        z = rhs.copy()
        if u > 0:                                    # unknown network nodes: (C + s L_uu) z = C rhs, always implicit
            ip, ix, calc = r.indptr, r.indices, r.calc_data
            L = np.zeros((u, u))
            for i in range(u):
                for k in range(ip[i], ip[i + 1]):
                    if ix[k] < u:
                        L[i, ix[k]] += calc[k]
            C = np.asarray(r.capy[:u], dtype=np.float64)
            z[:u] = np.linalg.solve(np.diag(C) + s * L, C * rhs[:u])
        for k, (p, st) in enumerate(zip(spec["lp"], r.lp)):
            o, n = r.lp_off[k], int(p["dof"])
            A = sp.csr_matrix(st["A"])
            if s * 2.0 * A.diagonal().max() < self.bound and ("lp", k) not in self.jac_parts:
                continue
            z[o:o + n] = spla.spsolve((sp.identity(n, format="csc") + s * A.tocsc()), rhs[o:o + n])
            self.solves += 1
        for k, (p, st) in enumerate(zip(spec["ffe"], r.ffe)):
            o, n = r.ffe_off[k], int(p["dof"])
            A = self.ffe_frozen[k] if self.ffe_frozen[k] is not None else sp.csr_matrix(st["A"])
            if s * 2.0 * self.ffe_gate[k].max() < self.bound and ("ffe", k) not in self.jac_parts:
                continue
            z[o:o + n] = spla.spsolve((sp.identity(n, format="csc") + s * A.tocsc()), rhs[o:o + n])
            self.solves += 1
        for k, (p, st) in enumerate(zip(spec["mor"], r.mor)):
            S, rr = int(p["S"]), int(p["r"])
            o = r.mor_off[k]
            for q in range(S):
                Aq = np.asarray(st["A"][q])
                if s * np.abs(Aq).sum(axis=1).max() < self.bound and ("mor", k) not in self.jac_parts:
                    continue
                sl = slice(o + q * rr, o + (q + 1) * rr)
                z[sl] = np.linalg.solve(np.eye(rr) + s * Aq, rhs[sl])
                self.solves += 1
        return z

    def integrate(self, t0, t1, rtol, atol, first_step=None, max_step=np.inf):
        rhs, u = self.rhs, self.u
        y = initial_state(self.s)
        t = float(t0)
        f1 = rhs(t, y)                                                # tc_ros2_begin: f(t0, y0) + initial snapshot
        times, net_temps, io_temps = [t], [rhs.net_temp.copy()], [y[u:].copy()]
        h_abs = float(first_step) if first_step else 1e-3 * abs(t1 - t0)
        acc_last, step_rejected = True, False          # FL_ACC_LAST starts at 1 (begin_common)
        accepted = rejected = 0
        if not (t1 > t0):
            return {"times": np.array(times), "net_temps": np.array(net_temps), "io_temps": np.array(io_temps),
                    "accepted": 0, "rejected": 0, "solves": 0}
        while True:
            min_step = 10.0 * abs(np.nextafter(t, np.inf) - t)
            if acc_last:
                if h_abs > max_step:
                    h_abs = max_step
                elif h_abs < min_step:
                    h_abs = min_step
                step_rejected = False
            if h_abs < min_step:
                raise RuntimeError("Required step size is less than spacing between numbers.")
            t_new = t + h_abs
            if t_new - t1 > 0.0:
                t_new = t1
            h = t_new - t
            h_abs = abs(h)
            self._jac_prepare(GAMMA * h)
            k1 = self._block_solve(GAMMA * h, f1)
            ystage = y + h * k1
            f2 = rhs(t + 1.0 * h, ystage)
            k2 = self._block_solve(GAMMA * h, f2 + (-2.0) * k1)
            ynew = y + h * (k1 * 1.5 + k2 * 0.5)
            f_end = rhs(t + 1.0 * h, ynew)       # next attempt's f(t, y) when accepted; state of the stored snapshot
            scale = atol + np.maximum(np.abs(y), np.abs(ynew)) * rtol
            e = (k1 * 0.5 + k2 * 0.5) * h / scale
            err = np.sqrt(np.sum(e * e)) / np.sqrt(float(self.n))
            if err < 1.0:
                factor = 10.0 if err == 0.0 else min(10.0, 0.9 * err ** -0.5)
                if step_rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                t = t_new
                y = ynew
                f1 = f_end
                accepted += 1
                acc_last = True
                # snapshot after the accepted step: the network state of the latest right-hand side call with the
                # unknown nodes of the new state (tc_snapshot_kernel / store_results, thermca/solver.py:393-424)
                net = rhs.net_temp.copy()
                net[:u] = y[:u]
                times.append(t); net_temps.append(net); io_temps.append(y[u:].copy())
                if t - t1 >= 0.0:
                    break
            else:
                h_abs *= max(0.2, 0.9 * err ** -0.5)
                rejected += 1
                step_rejected = True
                acc_last = False
        return {"times": np.array(times), "net_temps": np.array(net_temps), "io_temps": np.array(io_temps),
                "accepted": accepted, "rejected": rejected, "solves": self.solves}


def ros2_integrate(spec, t0, t1, rtol, atol, first_step=None, max_step=np.inf, bound=2.0):
    return Ros2Oracle(spec, bound).integrate(t0, t1, rtol, atol, first_step, max_step)
