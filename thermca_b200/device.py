"""Pack a lowered network for the device and drive the native solver.

Host side of the drop-in: turns the flat description of :mod:`thermca_b200.spec` into
device arenas + the network command stream (csrc/network.cuh), converts the part
operators CSR -> SELL-32 keeping the reference's in-row entry order (so the row sums are
bit-identical to scipy's csr_matvec, thermca/solver.py:264,281), and calls the C ABI
(include/thermca_b200.h).  torch only owns device memory and the stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .spec import state_offsets

# command opcodes, keep in sync with csrc/network.cuh
(C_END, C_SET_TU, C_SURF_MEAN, C_INPUTS, C_MATL_GROUP, C_CAPY, C_FLOW_CONST, C_FLOW_FUNC, C_FILM_FUNC,
 C_HEAT_FUNC, C_FLUX_FUNC, C_BOUND_FUNC, C_PART_FILMS, C_LP_UPDATE, C_LP2LP, C_LP_LEAK, C_STAT, C_DIAG,
 C_NET_DERIV, C_PART_FLUX, C_LP_LOAD, C_COPY_D, C_LP_TDEP, C_FUNCS, C_JTHETA) = range(25)

MEAN_ENTRIES_PER_BLOCK = 4096
MEAN_MAX_BLOCKS_PER_JOB = 64
from .trace import MAX_PROGRAM as VM_MAX      # one limit: tracer, packer and csrc/vm.cuh (TC_VM_MAX)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("thermca_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


# ---- CSR -> SELL-32 ----------------------------------------------------------------------------
def csr_to_sell(indptr, indices, data, n=None):
    """Returns dict(slice_ptr int64, col int32, val float64, pos int64[nnz]) where ``pos`` maps
    every CSR data index to its SELL position.  In-row order is preserved; padding is
    (value 0, column = own row)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = int(len(indptr) - 1 if n is None else n)
    lens = np.diff(indptr)
    nsl = (n + 31) // 32
    lens_pad = np.zeros(nsl * 32, dtype=np.int64)
    lens_pad[:n] = lens
    width = lens_pad.reshape(nsl, 32).max(axis=1) if nsl else np.zeros(0, dtype=np.int64)
    slice_ptr = np.zeros(nsl + 1, dtype=np.int64)
    np.cumsum(width * 32, out=slice_ptr[1:])
    total = int(slice_ptr[-1])
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    inrow = np.arange(len(indices), dtype=np.int64) - indptr[rows]
    pos = slice_ptr[rows >> 5] + inrow * 32 + (rows & 31)
    # padding: column = own row (rows past n point at row 0)
    col = np.empty(total, dtype=np.int32)
    val = np.zeros(total, dtype=np.float64)
    if total:
        sl_of = np.repeat(np.arange(nsl, dtype=np.int64), width * 32)
        lane = (np.arange(total, dtype=np.int64) - slice_ptr[sl_of]) & 31
        own = sl_of * 32 + lane
        own[own >= n] = 0
        col[:] = own
    col[pos] = np.asarray(indices, dtype=np.int32)
    val[pos] = np.asarray(data, dtype=np.float64)
    # compressed columns: col - row, when the band fits 16 bits (padding of rows past n points at row n-1)
    col16 = None
    if total:
        own_all = sl_of * 32 + lane
        ref_col = col.astype(np.int64)
        ref_col[own_all >= n] = n - 1
        rel = ref_col - own_all
        if rel.min() >= -32768 and rel.max() <= 32767:
            col16 = rel.astype(np.int16)
            col = ref_col.astype(np.int32)
    return {"slice_ptr": slice_ptr, "col": col, "col16": col16, "val": val, "pos": pos, "n": n, "nslices": nsl}


def csr_diag_positions(indptr, indices):
    """CSR data index of A_ii per row (first match, thermca/fem/ffe_io.py:52)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = len(indptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    hit = np.nonzero(rows == np.asarray(indices))[0]
    out = np.full(n, -1, dtype=np.int64)
    # first match per row: assign in reverse so the first wins
    out[rows[hit][::-1]] = hit[::-1]
    if (out < 0).any():
        raise ValueError("operator row without a stored diagonal entry")
    return out


def _group_lists(keys, order_keys, payload_arrays):
    """Group entries by ``keys`` (stable, secondary order ``order_keys``): returns
    (unique keys, start offsets, payloads reordered)."""
    keys = np.asarray(keys, dtype=np.int64)
    o = np.lexsort((np.asarray(order_keys), keys))
    ks = keys[o]
    uniq, first = np.unique(ks, return_index=True)
    start = np.append(first, len(ks)).astype(np.int32)
    return uniq, start, [np.asarray(p)[o] for p in payload_arrays]


class _Arena:
    def __init__(self):
        self.i = [np.zeros(2, dtype=np.int32)]       # offset 0 reserved
        self.d = [np.zeros(1, dtype=np.float64)]
        self.ni = 2
        self.nd = 1

    def add_i(self, a, align64=False):
        a = np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.int32)
        if align64 and (self.ni & 1):
            self.i.append(np.zeros(1, dtype=np.int32))
            self.ni += 1
        off = self.ni
        self.i.append(a)
        self.ni += a.size
        return off

    def add_i64(self, a):
        a = np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.int64)
        return self.add_i(a.view(np.int32), align64=True)

    def add_d(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
        off = self.nd
        self.d.append(a)
        self.nd += a.size
        return off

    def finish(self):
        return np.concatenate(self.i), np.concatenate(self.d)


class DeviceSolver:
    """Owns the device copy of one lowered network and the native solver handle."""

    def __init__(self, spec, device=None, history_bytes=1 << 30, use_col16=True):
        torch = _torch()
        self.torch = torch
        self.lib = _cabi.load_library()
        self.spec = spec
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # a dedicated non-default stream: the RK attempt is replayed as a CUDA graph and the
        # legacy default stream cannot be captured
        self.stream = torch.cuda.Stream(self.dev)
        self._keep = []                 # device tensors referenced by raw pointer
        self.N, self.u, self.nnz = int(spec["N"]), int(spec["u"]), int(len(spec["rc_data"]))
        self.lp_off, self.ffe_off, self.mor_off, self.n_state = state_offsets(spec)
        self.n_io = self.n_state - self.u
        self.history_bytes = history_bytes
        self.use_col16 = use_col16            # 16-bit relative column indices when the band allows
        self._handle = _cabi.vp()
        _cabi.check(self.lib.tc_create(C.byref(self._handle), C.c_void_p(self.stream.cuda_stream)))
        self._pack()
        torch.cuda.synchronize(self.dev)        # uploads were issued on torch's current stream
        self.set_state(self.initial_state())

    # -- helpers ------------------------------------------------------------------------------
    def _up(self, a, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype)))
        t = t.to(self.dev)
        self._keep.append(t)
        return t

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def initial_state(self):
        from .lowering import initial_state
        return initial_state(self.spec)

    # -- packing ------------------------------------------------------------------------------
    def _pack(self):
        s = self.spec
        A = _Arena()
        N, u = self.N, self.u
        o = {}
        o["T"] = A.add_d(s["init_temp"]); o["heat"] = A.add_d(s["heat_src"]); o["capy"] = A.add_d(s["capy"])
        o["vol"] = A.add_d(s["vol"]); o["volcapy"] = A.add_d(s["vol_capy"]); o["condy"] = A.add_d(s["condy"])
        o["cond"] = A.add_d(s["rc_data"]); o["clean"] = A.add_d(s["rc_clean_data"])
        o["calc"] = A.add_d(s["rc_data"])
        n_in = len(s["inputs"])
        o["inputs"] = A.add_d(np.full(max(n_in, 1), np.nan))
        o["indptr"] = A.add_i(s["rc_indptr"]); o["indices"] = A.add_i(s["rc_indices"])
        o["diag"] = A.add_i(s["diag_didx"] if u else np.zeros(1, dtype=np.int32))
        self.off = o

        # interpolation tables and link programs
        tab_dir = []
        for t in s["tables"]:
            off = A.add_d(np.concatenate([np.asarray(t["x"], dtype=np.float64), np.asarray(t["y"], dtype=np.float64)]))
            tab_dir += [off, len(t["x"])]
        prog_dir = []
        for p in s["programs"]:
            code = np.asarray(p["code"], dtype=np.int32).reshape(-1, 4)
            if len(code) > VM_MAX:
                raise NotImplementedError(f"link program with {len(code)} instructions (max {VM_MAX})")
            # int4 loads need 16-byte alignment
            while A.ni % 4:
                A.add_i(np.zeros(1, dtype=np.int32))
            coff = A.add_i(code)
            doff = A.add_d(p["consts"] if len(p["consts"]) else np.zeros(1))
            prog_dir += [coff, len(code), doff, int(p["result"])]
        tab_dir_off = A.add_i(tab_dir if tab_dir else [0, 0])
        prog_dir_off = A.add_i(prog_dir if prog_dir else [0, 0, 0, 0])

        cmds = []

        def cmd(op, *args):
            cmds.extend([op, len(args)])
            cmds.extend(int(a) for a in args)

        # ---- surface-mean jobs (thermca/solver.py:211-220) -----------------------------------
        jobs = []          # (entry idx array, weights, target net idx, ref_doff)
        job_of = {}        # (kind, part, surface) -> job index
        for k, p in enumerate(s["lp"]):
            Cm = np.asarray(p["C_marg"])
            for sidx, tgt in enumerate(p["surf_net_idxs"]):
                nz = np.nonzero(Cm[:, sidx])[0]
                job_of[(0, k, sidx)] = len(jobs)
                jobs.append((nz + self.lp_off[k], Cm[nz, sidx], int(tgt), -1))
        self._ffe_job_range = []           # (first job, number of jobs) of every FE part: tc_attach_dist
        for k, p in enumerate(s["ffe"]):
            self._ffe_job_range.append((len(jobs), len(p["surf_net_idxs"])))
            for sidx, tgt in enumerate(p["surf_net_idxs"]):
                job_of[(1, k, sidx)] = len(jobs)
                jobs.append((np.asarray(p["C_idx"][sidx]) + self.ffe_off[k], p["C_val"][sidx], int(tgt), -1))
        self._mor_dev = []
        for k, p in enumerate(s["mor"]):
            S, r = int(p["S"]), int(p["r"])
            ref_doff = A.add_d([float(p["ref_temp"])])
            Cs = np.asarray(p["C_surfs"]).reshape(S * r, S)
            for sidx, tgt in enumerate(p["surf_net_idxs"]):
                job_of[(2, k, sidx)] = len(jobs)
                jobs.append((np.arange(S * r) + self.mor_off[k], Cs[:, sidx], int(tgt), ref_doff))
            self._mor_dev.append({"ref_doff": ref_doff})
        job_of_block, first_block, nblk, ent_start = [], [], [], [0]
        ent_idx, ent_w, mean_args = [], [], [len(jobs)]
        for j, (idx, w, tgt, ref) in enumerate(jobs):
            nb = int(min(MEAN_MAX_BLOCKS_PER_JOB, max(1, -(-len(idx) // MEAN_ENTRIES_PER_BLOCK))))
            first_block.append(len(job_of_block))
            mean_args += [len(job_of_block), nb, tgt, ref]
            job_of_block += [j] * nb
            nblk.append(nb)
            ent_idx.append(np.asarray(idx, dtype=np.int32))
            ent_w.append(np.asarray(w, dtype=np.float64))
            ent_start.append(ent_start[-1] + len(idx))
        cmd(C_SET_TU)
        if jobs:
            cmd(C_SURF_MEAN, *mean_args)

        # ---- time dependent properties, reference statement order (network.py:757-796) ---------
        if n_in:
            args = [n_in]
            for inp in s["inputs"]:
                args += [A.add_d(inp["t"]), A.add_d(inp["v"]), len(inp["t"])]
            cmd(C_INPUTS, *args)
        for g in s["matl_groups"]:
            cmd(C_MATL_GROUP, len(g["idx"]), A.add_i(g["idx"]), int(g["vol_capy_tab"]), int(g["condy_tab"]))
        if len(s["capy_idxs"]):
            cmd(C_CAPY, len(s["capy_idxs"]), A.add_i(s["capy_idxs"]))
        fc = s["flow_const"]
        if len(fc["didx"]):
            cmd(C_FLOW_CONST, len(fc["didx"]), A.add_i(fc["didx"]), A.add_i(fc["i0"]), A.add_i(fc["i1"]),
                A.add_d(fc["vol_flow"]))
        # closures: one command; groups evaluated concurrently (a warp each), applied in the reference's order
        groups = []
        n_res = 0
        def group(kind, prog, n, *offs):
            nonlocal n_res
            groups.append([kind, int(prog), int(n), n_res] + [int(o) for o in offs] + [0] * (6 - len(offs)))
            n_res += int(n)
        for g in s["flow_funcs"]:
            group(0, g["prog"], len(g["didx"]), A.add_i(g["didx"]), A.add_i(g["i0"]), A.add_i(g["i1"]))
        for g in s["film_funcs"]:
            group(1, g["prog"], len(g["fwd"]), A.add_i(g["fwd"]), A.add_i(g["bwd"]), A.add_i(g["i0"]), A.add_i(g["i1"]),
                  A.add_d(g["areas"]))
        for g in s["heat_funcs"]:
            idxs = np.asarray(g["idxs"]).reshape(len(g["counts"]), -1)
            group(2, g["prog"], idxs.shape[0], idxs.shape[1], A.add_i(idxs), A.add_d(g["counts"]))
        for g in s["flux_funcs"]:
            group(3, g["prog"], len(g["idxs"]), A.add_i(g["idxs"]), A.add_d(g["areas"]))
        for g in s["bound_funcs"]:
            group(4, g["prog"], len(g["idxs"]), A.add_i(g["idxs"]))
        if groups:
            cmd(C_FUNCS, len(groups), A.add_d(np.zeros(max(n_res, 1))), *[w for g in groups for w in g])

        # ---- per-part arena data ------------------------------------------------------------------
        lp_dev = []
        for k, p in enumerate(s["lp"]):
            dof, F, S = int(p["dof"]), len(p["films"]), len(p["surf_areas"])
            ab = p["A_body"]
            sell = csr_to_sell(ab["indptr"], ab["indices"], p["A_data"])
            dpos_csr = np.asarray(p["A_diag_didx"], dtype=np.int64)
            d = {
                "sell": sell, "films": A.add_d(p["films"] if F else np.zeros(1)),
                "lfm": A.add_d(np.asarray(p["L_film_margs"]).reshape(dof, F) if F else np.zeros(1)),
                "lsum": A.add_d(np.asarray(p["L_film_margs"]).reshape(dof, F).sum(axis=0) if F else np.zeros(1)),
                "Ldata": A.add_d(p["L_data"]), "minv": A.add_d(p["M_inv"]),
                "abody_diag": A.add_d(np.asarray(ab["data"])[dpos_csr]),
                "val": A.add_d(sell["val"]), "diagpos": A.add_i(sell["pos"][dpos_csr]),
                "slice_ptr": A.add_i64(sell["slice_ptr"]), "col": A.add_i(sell["col"]),
                "b": A.add_d(np.zeros(dof)), "mass": A.add_d(1.0 / np.asarray(p["M_inv"])),
                "adiag": A.add_d(np.asarray(p["A_data"])[dpos_csr]),
                "qs": A.add_d(np.asarray(p["Qs"]).reshape(dof, S)), "areas": A.add_d(p["surf_areas"]),
                "surf_net": A.add_i(p["surf_net_idxs"]), "link": A.add_i(p["link_elem_net_idxs"] if F else [0]),
                "dock": A.add_d(p["dock_areas"] if F else np.zeros(1)), "dof": dof, "F": F, "S": S,
            }
            face_dir = []
            for f in range(F):
                face_dir += [A.add_d(p["face_areas"][f]), A.add_i(p["smx_didx"][f]), A.add_i(p["marg_idxs"][f]),
                             len(p["face_areas"][f])]
            d["face_dir"] = A.add_i(face_dir if face_dir else [0, 0, 0, 0])
            td = p.get("tdep")
            if td is not None:
                surfdir = []
                for marg, surf in zip(td["marg_all"], td["surf_all"]):
                    surfdir += [A.add_i(marg), A.add_i(surf), len(marg)]
                d["tdep_args"] = [
                    dof, int(td["n_rows"]), len(td["Lx_base_data"]), self.lp_off[k], A.add_d(td["Lx_base_data"]),
                    A.add_i(td["row"]), A.add_i(td["col"]), d["Ldata"], A.add_i(td["L_indptr"]), A.add_i(td["diag_didx"]),
                    A.add_d(td["Mx_diag"]), d["minv"], d["mass"], int(td["condy_tab"]), int(td["vol_capy_tab"]),
                    A.add_d(np.zeros(int(td["n_rows"]))), len(td["marg_all"]), A.add_i(surfdir if surfdir else [0, 0, 0]),
                    len(td["abody_from_L"]), A.add_i(td["abody_from_L"]), A.add_i(td["abody_rows"]),
                    A.add_i(sell["pos"]), d["val"], d["abody_diag"]]
            lp_dev.append(d)
        ffe_dev = []
        for k, p in enumerate(s["ffe"]):
            F, S = len(p["films"]), int(p["S"])
            ffe_dev.append({"films": A.add_d(p["films"] if F else np.zeros(1)), "flux": A.add_d(np.zeros(S)),
                            "areas": A.add_d(p["surf_areas"]), "surf_net": A.add_i(p["surf_net_idxs"]),
                            "film_surf": A.add_i(p["film_surf_idxs"] if F else [0]),
                            "link": A.add_i(p["link_elem_net_idxs"] if F else [0]),
                            "dock": A.add_d(p["dock_areas"] if F else np.zeros(1)), "F": F, "S": S})
        for k, p in enumerate(s["mor"]):
            F, S = int(p["F"]), int(p["S"])
            self._mor_dev[k].update({"films": A.add_d(p["films"] if F else np.zeros(1)),
                                     "flux": A.add_d(np.zeros(S)), "areas": A.add_d(p["surf_areas"]),
                                     "surf_net": A.add_i(p["surf_net_idxs"]),
                                     "film_surf": A.add_i(p["film_surf_idxs"] if F else [0]),
                                     "link": A.add_i(p["link_elem_net_idxs"] if F else [0]),
                                     "dock": A.add_d(p["dock_areas"] if F else np.zeros(1)), "F": F, "S": S})

        # ---- films between part surfaces and stationary nodes: rank-one terms of the W-method Jacobian
        #      (csrc/network.cuh C_JTHETA, csrc/solver.cu tc_set_jac_films); the right-hand side does not use them ----
        stat_nodes = set(int(i) for i in s["stat_idx"])
        self._jac_films = []
        for kind, (parts, devs) in enumerate(((s["lp"], lp_dev), (s["ffe"], ffe_dev), (s["mor"], self._mor_dev))):
            for k, (p, d) in enumerate(zip(parts, devs)):
                F = d["F"]
                flags = [1 if int(n) in stat_nodes else 0 for n in p["link_elem_net_idxs"]] if F else []
                d["jac_on"] = bool(any(flags)) and not (kind == 1 and (p.get("tdep") is not None or "sell" in p))
                if not d["jac_on"]:
                    continue
                fsurf = np.asarray(p["film_surf_idxs"], dtype=np.int64)
                d["theta"] = A.add_d(np.zeros(F))
                d["statflag"] = A.add_i(flags)
                d["surfnode"] = A.add_i(np.asarray(p["surf_net_idxs"])[fsurf])
                for f in range(F):
                    if flags[f]:
                        j = job_of[(kind, k, int(fsurf[f]))]
                        self._jac_films.append([kind, k, f, int(fsurf[f]), d["theta"] + f, d["films"] + f,
                                                first_block[j], nblk[j]])
        if len(self._jac_films) > 4:
            raise NotImplementedError("more than 4 films between part surfaces and stationary nodes")

        # ---- films of variable parts (network.py:803-817) -------------------------------------------
        for p, d in list(zip(s["lp"], lp_dev)) + list(zip(s["ffe"], ffe_dev)) + list(zip(s["mor"], self._mor_dev)):
            if p["var_film_rc_idx"] is not None:
                cmd(C_PART_FILMS, d["F"], d["films"], A.add_i(p["var_film_rc_idx"]), d["dock"])

        def lp_update(d):
            cmd(C_LP_UPDATE, d["dof"], d["F"], d["films"], d["lfm"], d["lsum"], d["Ldata"], d["face_dir"],
                d["minv"], d["abody_diag"], d["val"], d["diagpos"], d["adiag"])

        for p, d in zip(s["lp"], lp_dev):
            if "tdep_args" in d:
                cmd(C_LP_TDEP, *d["tdep_args"])
        for p, d in zip(s["lp"], lp_dev):
            if int(p["var"]):
                lp_update(d)
        for l in s["lp_to_lp"]:
            d0, d1 = lp_dev[int(l["p0"])], lp_dev[int(l["p1"])]
            cmd(C_LP2LP, d0["films"] + int(l["f0"]), d1["lsum"] + int(l["f1"]), d0["areas"] + int(l["s0"]),
                d1["films"] + int(l["f1"]), d0["lsum"] + int(l["f0"]), d1["areas"] + int(l["s1"]))
            lp_update(d0)
            lp_update(d1)
        for p, d in zip(s["lp"], lp_dev):
            if int(p["var"]):
                cmd(C_LP_LEAK, d["F"], A.add_i(p["leak_fwd"]), A.add_i(p["leak_bwd"]), d["lsum"])

        # ---- stationary nodes, differential form, network derivative (solver.py:224-255) --------------
        if len(s["stat_idx"]):
            cmd(C_STAT, len(s["stat_idx"]), A.add_i(s["stat_idx"]))
        for d in lp_dev + ffe_dev + list(self._mor_dev):
            if d.get("jac_on"):
                cmd(C_JTHETA, d["F"], d["theta"], d["link"], d["surfnode"], d["statflag"])
        cmd(C_DIAG)
        if u:
            cmd(C_NET_DERIV)
        # ---- part loads (solver.py:258-299) ------------------------------------------------------------
        for d in lp_dev:
            cmd(C_LP_LOAD, d["dof"], d["S"], d["F"], d["b"], d["qs"], d["surf_net"], d["areas"], d["lfm"],
                d["link"], d["minv"])
        for d in ffe_dev:
            cmd(C_PART_FLUX, d["S"], d["F"], d["flux"], d["surf_net"], d["areas"], d["film_surf"], d["films"],
                d["link"], -1)
        for d in self._mor_dev:
            cmd(C_PART_FLUX, d["S"], d["F"], d["flux"], d["surf_net"], d["areas"], d["film_surf"], d["films"],
                d["link"], d["ref_doff"])
        cmd(C_END)
        cmds_off = A.add_i(cmds)

        iarena, darena = A.finish()
        self.iarena = self._up(iarena)
        self.darena = self._up(darena)
        self._darena_host0 = darena
        t_job = self._up(np.asarray(job_of_block if job_of_block else [0], dtype=np.int32))
        t_first = self._up(np.asarray(first_block if first_block else [0], dtype=np.int32))
        t_nblk = self._up(np.asarray(nblk if nblk else [0], dtype=np.int32))
        t_estart = self._up(np.asarray(ent_start, dtype=np.int64))
        t_eidx = self._up(np.concatenate(ent_idx) if ent_idx else np.zeros(1, dtype=np.int32))
        t_ew = self._up(np.concatenate(ent_w) if ent_w else np.zeros(1))

        nd = _cabi.NetDesc()
        nd.iarena, nd.darena = self._p(self.iarena), self._p(self.darena)
        nd.cmds_off, nd.prog_dir_off, nd.tab_dir_off = cmds_off, prog_dir_off, tab_dir_off
        nd.N, nd.u, nd.nnz = N, u, self.nnz
        for name in ("T", "heat", "capy", "vol", "volcapy", "condy", "cond", "clean", "calc", "inputs",
                     "indptr", "indices", "diag"):
            setattr(nd, "o_" + name, o[name])
        nd.n_mean_jobs, nd.n_mean_blocks = len(jobs), len(job_of_block)
        nd.job_of_block, nd.first_block, nd.nblk = self._p(t_job), self._p(t_first), self._p(t_nblk)
        nd.ent_start, nd.ent_idx, nd.ent_w = self._p(t_estart), self._p(t_eidx), self._p(t_ew)
        nd.n_iarena, nd.n_darena = int(self.iarena.numel()), int(self.darena.numel())
        _cabi.check(self.lib.tc_set_network(self._handle, C.byref(nd)))

        for k, d in enumerate(lp_dev):
            ld = _cabi.LpDesc()
            ld.n, ld.nslices, ld.y_off = d["dof"], d["sell"]["nslices"], self.lp_off[k]
            ld.slice_ptr_ioff64, ld.col_ioff, ld.val_doff = d["slice_ptr"], d["col"], d["val"]
            ld.b_doff, ld.mass_doff, ld.adiag_doff = d["b"], d["mass"], d["adiag"]
            ld.F, ld.lfm_doff, ld.minv_doff = d["F"], d["lfm"], d["minv"]
            _cabi.check(self.lib.tc_add_lp_part(self._handle, C.byref(ld)))
        self.ffe_sell = []
        for k, (p, d) in enumerate(zip(s["ffe"], ffe_dev)):
            self._add_ffe(k, p, d)
        for k, (p, d) in enumerate(zip(s["mor"], self._mor_dev)):
            md = _cabi.MorDesc()
            md.S, md.r, md.F, md.y_off = int(p["S"]), int(p["r"]), int(p["F"]), self.mor_off[k]
            md.A_body = self._p(self._up(np.asarray(p["A_body"], dtype=np.float64)))
            af = np.asarray(p["A_films"], dtype=np.float64)
            md.A_films = self._p(self._up(af if af.size else np.zeros(1)))
            md.B_T = self._p(self._up(np.asarray(p["B_T"], dtype=np.float64)))
            md.films_doff, md.flux_doff = d["films"], d["flux"]
            _cabi.check(self.lib.tc_add_mor_part(self._handle, C.byref(md)))
        _cabi.check(self.lib.tc_finalize(self._handle, self.n_state))
        if self._jac_films:
            rec = np.ascontiguousarray(self._jac_films, dtype=np.int32)
            _cabi.check(self.lib.tc_set_jac_films(self._handle, len(self._jac_films), rec.ctypes.data_as(C.c_void_p)))

    def _add_ffe_tdep(self, k, p, d):
        """Temperature dependent body: ship Lx_base (geometry) + tables; the operator is formed edge
        by edge in the kernel (thermca/fem/fe_system.py:185-193)."""
        td = p["tdep"]
        n = int(p["dof"])
        row = np.asarray(td["row"], dtype=np.int64)
        indptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=n))])
        sell = csr_to_sell(indptr, td["col"], td["Lx_base_data"])
        t_sp, t_col, t_val = self._up(sell["slice_ptr"]), self._up(sell["col"]), self._up(sell["val"])
        t_mx = self._up(np.asarray(td["Mx_diag"], dtype=np.float64))
        t_zero = self._up(np.zeros(n))
        self.ffe_sell.append({"slice_ptr": t_sp, "col": t_col, "val": t_val, "mass": t_mx, "adiag": t_zero,
                              "n": n, "nslices": sell["nslices"]})
        # per surface/film row: Qs entries and film-diagonal entries
        q_rows = np.concatenate([np.asarray(ix, dtype=np.int64) for ix in td["Qs_idx"]])
        q_surf = np.concatenate([np.full(len(ix), sidx, dtype=np.int32) for sidx, ix in enumerate(td["Qs_idx"])])
        q_val = np.concatenate([np.asarray(v, dtype=np.float64) for v in td["Qs_val"]])
        ab_indptr = np.asarray(p["A_body"]["indptr"], dtype=np.int64)
        if len(p["films"]):
            f_cpos = np.concatenate([np.asarray(ix, dtype=np.int64) for ix in p["film_didx"]])
            f_rows = np.searchsorted(ab_indptr, f_cpos, side="right") - 1
            f_film = np.concatenate([np.full(len(ix), f, dtype=np.int32) for f, ix in enumerate(p["film_didx"])])
            f_adata = np.concatenate([np.asarray(a, dtype=np.float64) for a in p["film_adata"]])
        else:
            f_rows, f_film, f_adata = np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0)
        rows = np.unique(np.concatenate([q_rows, f_rows]))
        oq = np.lexsort((q_surf, q_rows))
        of = np.lexsort((f_film, f_rows))
        q_start = np.searchsorted(q_rows[oq], np.append(rows, rows[-1] + 1 if len(rows) else 0)).astype(np.int32)
        f_start = np.searchsorted(f_rows[of], np.append(rows, rows[-1] + 1 if len(rows) else 0)).astype(np.int32)
        fd = _cabi.FfeDesc()
        fd.n, fd.nslices, fd.y_off = n, sell["nslices"], self.ffe_off[k]
        fd.slice_ptr, fd.col, fd.val = self._p(t_sp), self._p(t_col), self._p(t_val)
        fd.col16 = self._p(None)
        fd.mass, fd.adiag = self._p(t_mx), self._p(t_zero)
        fd.load_nrows, fd.fd_nrows = 0, 0
        fd.films_doff, fd.flux_doff = d["films"], d["flux"]
        fd.tdep, fd.condy_tab, fd.volcapy_tab = 1, int(td["condy_tab"]), int(td["vol_capy_tab"])
        fd.ts_nrows = len(rows)
        fd.ts_row = self._p(self._up(rows.astype(np.int32)))
        fd.ts_start = self._p(self._up(q_start))
        fd.ts_surf = self._p(self._up(q_surf[oq].astype(np.int32) if len(oq) else np.zeros(1, np.int32)))
        fd.ts_qval = self._p(self._up(q_val[oq] if len(oq) else np.zeros(1)))
        fd.ts_fstart = self._p(self._up(f_start))
        fd.ts_film = self._p(self._up(f_film[of].astype(np.int32) if len(of) else np.zeros(1, np.int32)))
        fd.ts_adata = self._p(self._up(f_adata[of] if len(of) else np.zeros(1)))
        # Jacobian approximation of the implicit W-method: the operator as the reference built it at the initial
        # temperatures (FFEIO.A incl. the initial film diagonal, fem/ffe_io.py:24-65), frozen
        ab = p["A_body"]
        jsell = csr_to_sell(ab["indptr"], ab["indices"], p["A_data"])
        dpos = csr_diag_positions(ab["indptr"], ab["indices"])
        fd.j_nslices = jsell["nslices"]
        fd.j_slice_ptr, fd.j_col = self._p(self._up(jsell["slice_ptr"])), self._p(self._up(jsell["col"]))
        fd.j_val = self._p(self._up(jsell["val"]))
        fd.j_mass = self._p(self._up(1.0 / np.asarray(p["M_inv"], dtype=np.float64)))
        fd.j_adiag = self._p(self._up(np.asarray(p["A_data"], dtype=np.float64)[dpos]))
        _cabi.check(self.lib.tc_add_ffe_part(self._handle, C.byref(fd)))

    def _add_ffe(self, k, p, d):
        if p.get("tdep") is not None:
            return self._add_ffe_tdep(k, p, d)
        ab = p["A_body"]
        n = int(p["dof"])
        if "sell" in p:                  # operator already resident on the device (csrc/synth.cu)
            sell = p["sell"]
            t_sp, t_col, t_val = sell["slice_ptr"], sell["col"], sell["val"]
            t_col16 = sell.get("col16")
            t_mass, t_adiag = sell["mass"], sell["adiag"]
            nslices = int(sell["nslices"])
            pos = None
            self._keep += [t_sp, t_col, t_val, t_mass, t_adiag]
        else:
            sell = csr_to_sell(ab["indptr"], ab["indices"], p["A_data"])
            t_sp, t_col, t_val = self._up(sell["slice_ptr"]), self._up(sell["col"]), self._up(sell["val"])
            t_col16 = self._up(sell["col16"]) if sell["col16"] is not None else None
            t_mass = self._up(1.0 / np.asarray(p["M_inv"], dtype=np.float64))
            dpos = csr_diag_positions(ab["indptr"], ab["indices"])
            t_adiag = self._up(np.asarray(p["A_data"], dtype=np.float64)[dpos])
            nslices = sell["nslices"]
            pos = sell["pos"]
        self.ffe_sell.append({"slice_ptr": t_sp, "col": t_col, "val": t_val, "mass": t_mass, "adiag": t_adiag,
                              "n": n, "nslices": nslices, "col16": t_col16 if self.use_col16 else None})
        # surface loads grouped by row
        rows = np.concatenate([np.asarray(ix, dtype=np.int64) for ix in p["B_idx"]]) if p["B_idx"] else np.zeros(0, np.int64)
        surf = np.concatenate([np.full(len(ix), sidx, dtype=np.int32) for sidx, ix in enumerate(p["B_idx"])]) \
            if p["B_idx"] else np.zeros(0, np.int32)
        bval = np.concatenate([np.asarray(v, dtype=np.float64) for v in p["B_val"]]) if p["B_val"] else np.zeros(0)
        fd = _cabi.FfeDesc()
        fd.n, fd.nslices, fd.y_off = n, nslices, self.ffe_off[k]
        fd.slice_ptr, fd.col, fd.val = self._p(t_sp), self._p(t_col), self._p(t_val)
        fd.col16 = self._p(t_col16 if self.use_col16 else None)
        if t_col16 is not None:
            self._keep.append(t_col16)
        fd.mass, fd.adiag = self._p(t_mass), self._p(t_adiag)
        if len(rows):
            urow, start, (surf_s, bval_s) = _group_lists(rows, surf, [surf, bval])
            fd.load_nrows = len(urow)
            fd.load_row = self._p(self._up(urow.astype(np.int32)))
            fd.load_start = self._p(self._up(start))
            fd.load_surf = self._p(self._up(surf_s.astype(np.int32)))
            fd.load_bval = self._p(self._up(bval_s.astype(np.float64)))
        else:
            fd.load_nrows = 0
        # variable films -> diagonal update lists (thermca/fem/ffe_io.py:76-82)
        var_films = p["var_film_rc_idx"] is not None and len(p["films"]) > 0
        if var_films or d.get("jac_on"):
            if pos is None:
                raise NotImplementedError("variable films on a device-generated operator")
            cpos = np.concatenate([np.asarray(ix, dtype=np.int64) for ix in p["film_didx"]])
            film = np.concatenate([np.full(len(ix), f, dtype=np.int32) for f, ix in enumerate(p["film_didx"])])
            adata = np.concatenate([np.asarray(a, dtype=np.float64) for a in p["film_adata"]])
            ucpos, start, (film_s, adata_s) = _group_lists(cpos, film, [film, adata])
            indptr = np.asarray(ab["indptr"], dtype=np.int64)
            rows_of = np.searchsorted(indptr, ucpos, side="right") - 1
            fd.fd_nrows = len(ucpos)
            fd.fd_pos = self._p(self._up(pos[ucpos].astype(np.int64)))
            fd.fd_row = self._p(self._up(rows_of.astype(np.int32)))
            fd.fd_abody = self._p(self._up(np.asarray(ab["data"], dtype=np.float64)[ucpos]))
            fd.fd_start = self._p(self._up(start))
            fd.fd_film = self._p(self._up(film_s.astype(np.int32)))
            fd.fd_adata = self._p(self._up(adata_s.astype(np.float64)))
        else:
            fd.fd_nrows = 0
        fd.films_doff, fd.flux_doff = d["films"], d["flux"]
        fd.fd_rhs = int(var_films)
        fd.tdep, fd.ts_nrows = 0, 0
        _cabi.check(self.lib.tc_add_ffe_part(self._handle, C.byref(fd)))

    # -- state ---------------------------------------------------------------------------------------
    def set_state(self, y):
        t = self.torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64)).to(self.dev) \
            if not self.torch.is_tensor(y) else y
        self._y_in = t
        self.torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_set_state(self._handle, self._p(t)))
        self.stream.synchronize()

    def get_state(self):
        out = self.torch.empty(max(self.n_state, 1), dtype=self.torch.float64, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_get_state(self._handle, self._p(out)))
        return out[: self.n_state].cpu().numpy()

    def get_state_into(self, out):
        """Copy the state into a caller-owned device tensor (no host round trip)."""
        self.torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_get_state(self._handle, self._p(out)))
        return out

    def network_state(self):
        """Host copies of the mutable network arrays (T, heat_src, capy, cond, clean_cond)."""
        self.stream.synchronize()
        o, N, nnz = self.off, self.N, self.nnz
        get = lambda key, n: self.darena[o[key]: o[key] + n].cpu().numpy()       # only the five ranges, not the arena
        return {"T": get("T", N), "heat_src": get("heat", N), "capy": get("capy", N), "cond": get("cond", nnz),
                "clean": get("clean", nnz)}

    # -- right-hand side (parity hook) ------------------------------------------------------------------
    def rhs(self, t, y):
        torch = self.torch
        yd = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64)).to(self.dev)
        out = torch.zeros(max(self.n_state, 1), dtype=torch.float64, device=self.dev)
        torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_rhs(self._handle, float(t), self._p(yd), self._p(out)))
        return out[: self.n_state].cpu().numpy()

    # -- history ring --------------------------------------------------------------------------------------
    def set_probes(self, probe_dofs):
        """Results at scale: store only the listed part DOFs (indices into the part block of
        the state, the ``io_temp`` layout of thermca/solver.py:315-316) per accepted step instead of the full part
        state; network temperatures (which include every surface mean) are always stored.  ``None`` restores full
        rows.  The full field stays on the device and is fetched on demand with ``get_state()``."""
        if probe_dofs is None:
            self._probe_idx, self._probe_mode = None, False
            _cabi.check(self.lib.tc_set_probes(self._handle, C.c_void_p(0), 0))
            return
        idx = np.ascontiguousarray(probe_dofs, dtype=np.int64).ravel()
        if idx.size and (idx.min() < 0 or idx.max() >= self.n_io):
            raise IndexError("probe DOF outside the part block of the state")
        # an EMPTY list is a real zero-width probe mode (no part DOFs per step: the history ring gets no part buffer),
        # not a request for full rows
        self._probe_mode = True
        self._probe_idx = self._up(idx) if idx.size else None
        self._n_probe = int(idx.size)
        _cabi.check(self.lib.tc_set_probes(self._handle, self._p(self._probe_idx), int(idx.size)))

    def _alloc_history(self, num_skip):
        torch = self.torch
        w_net = 1 + 3 * self.N + 2 * self.nnz
        self._io_width = self._n_probe if getattr(self, "_probe_mode", False) else self.n_io
        # a row-partitioned run needs the SAME ring capacity on every rank (a full ring pauses the device loop):
        # partition.py sets _cap_width to the widest rank's row
        per = 8 * (max(self._io_width, getattr(self, "_cap_width", 0)) + w_net)
        cap = int(max(8, min(8192, self.history_bytes // max(per, 1))))
        self.capacity = cap
        self.hist_io = torch.zeros((cap, max(self._io_width, 1)), dtype=torch.float64, device=self.dev)
        self.hist_net = torch.zeros((cap, w_net), dtype=torch.float64, device=self.dev)
        torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_set_history(self._handle, self._p(self.hist_io) if self._io_width else C.c_void_p(0),
                                            self._p(self.hist_net), cap, int(num_skip)))

    def info(self):
        info = (_cabi.i64 * 8)()
        t, h = _cabi.f64(), _cabi.f64()
        _cabi.check(self.lib.tc_get_info(self._handle, info, C.byref(t), C.byref(h)))
        return {"done": int(info[0]), "accepted": int(info[1]), "rejected": int(info[2]), "status": int(info[3]),
                "stored": int(info[4]), "nfev": int(info[5]), "result_step": int(info[6]), "t": t.value,
                "h_next": h.value}

    def graph_info(self):
        """Which attempt graphs are instantiated: ``{"rk": bool, "w": 0 | 2 (ROS2) | 3 (ROW3w) | -1 (capture refused)}``."""
        out = (C.c_int32 * 2)()
        _cabi.check(self.lib.tc_graph_info(self._handle, out))
        return {"rk": bool(out[0]), "w": int(out[1])}

    def _drain(self, out):
        n = self.info()["stored"]
        self.stream.synchronize()
        if n:
            net = self.hist_net[:n].cpu().numpy()
            io = self.hist_io[:n, : self._io_width].cpu().numpy() if self._io_width else np.zeros((n, 0))
            out["net"].append(net)
            out["io"].append(io)
        _cabi.check(self.lib.tc_reset_history(self._handle))

    def _collect(self, out):
        N, nnz = self.N, self.nnz
        net = np.vstack(out["net"]) if out["net"] else np.zeros((0, 1 + 3 * N + 2 * nnz))
        io = np.vstack(out["io"]) if out["io"] else np.zeros((0, self._io_width))
        return {"times": net[:, 0].copy(), "net_temps": net[:, 1:1 + N].copy(),
                "net_capys": net[:, 1 + N:1 + 2 * N].copy(), "net_heat_srcs": net[:, 1 + 2 * N:1 + 3 * N].copy(),
                "cond_data": net[:, 1 + 3 * N:1 + 3 * N + nnz].copy(), "clean_data": net[:, 1 + 3 * N + nnz:].copy(),
                "io_temps": io, "probes": getattr(self, "_probe_mode", False)}

    # -- integrators -----------------------------------------------------------------------------------------
    def run_rk(self, t0, t1, rtol=1e-3, atol=1e-6, method="RK45", num_skip=0, max_step=np.inf,
               first_step=None, attempts_per_poll=32, progress=None):
        """scipy-equivalent adaptive RK45/RK23 with the step controller on the device."""
        mcode = {"RK45": 0, "RK23": 1}[method]
        self._alloc_history(num_skip)
        ms = 0.0 if not np.isfinite(max_step) else float(max_step)
        _cabi.check(self.lib.tc_rk_begin(self._handle, mcode, float(t0), float(t1), float(rtol), float(atol),
                                         ms, float(first_step or 0.0)))
        out = {"net": [], "io": []}
        while True:
            info = self.info()
            if info["done"] and info["status"] != 2:
                break
            if info["status"] == 2 or info["stored"] + attempts_per_poll > self.capacity:
                self._drain(out)
            burst = min(attempts_per_poll, self.capacity - self.info()["stored"])
            _cabi.check(self.lib.tc_rk_advance(self._handle, int(burst), int(burst)))
            if progress is not None:
                progress(self.info())
        info = self.info()
        self._drain(out)
        res = self._collect(out)
        res.update({"nfev": info["nfev"], "accepted": info["accepted"], "rejected": info["rejected"],
                    "status": info["status"]})
        if info["status"] == -1:
            raise Exception("Solver failed" + "Required step size is less than spacing between numbers.")
        return res

    def run_ros2(self, t0, t1, rtol=1e-3, atol=1e-6, num_skip=0, max_step=np.inf, first_step=None,
                 pcg_rtol=1e-8, pcg_maxit=2000, pcg_mode=0, attempts_per_poll=8, progress=None, scheme="ROS2"):
        """Adaptive linearly implicit W-methods (new methods; FE / LP parts implicit via Jacobi-PCG): ``scheme="ROS2"``
        (2 stages, order 2) or ``"ROW3"`` (4 stages, 3 right-hand sides, order 3 for any Jacobian approximation)."""
        begin, advance = {"ROS2": (self.lib.tc_ros2_begin, self.lib.tc_ros2_advance),
                          "ROW3": (self.lib.tc_row3_begin, self.lib.tc_row3_advance)}[scheme]
        self._alloc_history(num_skip)
        ms = 0.0 if not np.isfinite(max_step) else float(max_step)
        _cabi.check(begin(self._handle, float(t0), float(t1), float(rtol), float(atol), ms,
                                           float(first_step or 0.0), float(pcg_rtol), int(pcg_maxit), int(pcg_mode)))
        out = {"net": [], "io": []}
        while True:
            info = self.info()
            if info["done"] and info["status"] != 2:
                break
            if info["status"] == 2 or info["stored"] + attempts_per_poll > self.capacity:
                self._drain(out)
            _cabi.check(advance(self._handle, int(attempts_per_poll), int(attempts_per_poll)))
            if progress is not None:
                progress(self.info())
        info = self.info()
        self._drain(out)
        res = self._collect(out)
        res.update({"accepted": info["accepted"], "rejected": info["rejected"], "status": info["status"],
                    "pcg": self.pcg_stats()})
        if info["status"] == -1:
            raise Exception("Solver failed" + "Required step size is less than spacing between numbers.")
        bad = self.pcg_status()
        res["pcg_failed_solves"] = bad["failed_solves"]
        if bad["failed_solves"]:
            # a W-method tolerates inexact stage solves (the error estimator sees them), but never silently
            import warnings
            warnings.warn(f"{bad['failed_solves']} PCG stage solves ended without reaching pcg_rtol "
                          f"({bad['reason']}); raise pcg_maxit or loosen pcg_rtol", RuntimeWarning)
        return res

    def theta_steps(self, nsteps, h, theta=0.5, t0=None, pcg_rtol=1e-10, pcg_maxit=2000, pcg_mode=0, sync=True):
        """``nsteps`` fixed steps of the theta scheme (I + theta h A) d = h f(t + theta h, y)."""
        if t0 is not None:
            _cabi.check(self.lib.tc_set_time(self._handle, float(t0)))
        _cabi.check(self.lib.tc_theta_steps(self._handle, int(nsteps), float(theta), float(h), float(pcg_rtol),
                                            int(pcg_maxit), int(pcg_mode)))
        if sync:
            self.stream.synchronize()
            bad = self.pcg_status()
            if bad["failed_solves"]:       # fixed steps have no error estimator behind them: a bad solve is an error
                raise RuntimeError(f"PCG did not converge in {bad['failed_solves']} of {nsteps} theta steps "
                                   f"({bad['reason']}, pcg_maxit={pcg_maxit}, pcg_rtol={pcg_rtol})")

    def attach_partition(self, k, body):
        """FE part ``k`` of this (localised) network holds the local rows of a row-partitioned body: ``body`` is the
        ``thermca_b200.dist.DistBody`` built on this solver's SELL tensors and stream (thermca_b200/partition.py)."""
        j0, nj = self._ffe_job_range[k]
        _cabi.check(self.lib.tc_attach_dist(self._handle, int(k), body._h, int(j0), int(nj)))
        # global row count: the adaptive explicit pairs (run_rk) take their error / initial-step norms over ALL ranks
        _cabi.check(self.lib.tc_set_dist_global_rows(self._handle, int(body.n_global)))
        self._bodies = getattr(self, "_bodies", []) + [body]

    def set_block_jacobi(self, h, theta=0.5):
        """Build the block-Jacobi preconditioner (32 x 32 slice blocks of M + theta*h*K, csrc/pcg.cuh) of every
        full-order FE part for the CURRENT operator values (films folded in) and hand it to the solver; use with
        ``theta_steps(..., pcg_mode=2)``.  Rebuild after changing h or a film coefficient."""
        from .linsolve import slice_block_inverses
        self._binv = []
        self.stream.synchronize()
        for k, sell in enumerate(self.ffe_sell):
            if self.spec["ffe"][k].get("tdep") is not None:
                continue
            binv = slice_block_inverses(sell["slice_ptr"], sell["col"], sell["val"], sell["mass"], int(sell["n"]),
                                        float(theta) * float(h))
            self._binv.append(binv)
            _cabi.check(self.lib.tc_set_block_precond(self._handle, k, self._p(binv), float(theta) * float(h)))
        self.torch.cuda.synchronize(self.dev)

    def run_theta(self, t0, t1, h, theta=0.5, num_skip=0, pcg_rtol=1e-8, pcg_maxit=2000, progress=None):
        """Fixed-step theta scheme from t0 to t1 (new method; the reference has no implicit scheme).  Returns the same
        dictionary as ``run_rk``: the state is stored at t0, after every ``num_skip + 1``-th step and at t1; the
        network rows of a stored state (surface means, conductances, sources) come from one extra right-hand side at
        that state, as after the FSAL stage of the explicit pairs."""
        nsteps = int(round((t1 - t0) / h))
        if nsteps < 0 or abs(t0 + nsteps * h - t1) > 1e-9 * max(abs(t1), abs(h)):
            raise ValueError("THETA_PCG needs time_span[1] - time_span[0] to be a whole number of steps")
        if self.u > 0 or self.spec["mor"]:
            import warnings
            warnings.warn("THETA_PCG advances FE and LP parts implicitly; the unknown network nodes and MOR parts of this "
                          "model take the explicit Euler update with the same step (stable only while step * their "
                          "fastest rate < 2): prefer method='ROW3_PCG' for stiff lumped nodes", RuntimeWarning)
        u = self.u
        rows = {k: [] for k in ("times", "net_temps", "net_capys", "net_heat_srcs", "cond_data", "clean_data", "io_temps")}

        def store(t):
            y = self.get_state()
            self.rhs(t, y)                                   # network arrays at (t, y)
            ns = self.network_state()
            T = ns["T"].copy()
            T[:u] = y[:u]                                    # thermca/solver.py:314
            rows["times"].append(t); rows["net_temps"].append(T); rows["net_capys"].append(ns["capy"])
            rows["net_heat_srcs"].append(ns["heat_src"]); rows["cond_data"].append(ns["cond"])
            rows["clean_data"].append(ns["clean"]); rows["io_temps"].append(y[u:].copy())
        _cabi.check(self.lib.tc_set_time(self._handle, float(t0)))
        store(t0)
        for k in range(nsteps):
            self.theta_steps(1, h, theta, t0=t0 + k * h, pcg_rtol=pcg_rtol, pcg_maxit=pcg_maxit)
            if (k + 1) % (num_skip + 1) == 0 or k + 1 == nsteps:
                store(t0 + (k + 1) * h if k + 1 < nsteps else t1)
                if progress is not None:
                    progress({"t": rows["times"][-1]})
        res = {k: np.asarray(v) for k, v in rows.items()}
        res.update({"status": 0, "accepted": nsteps, "rejected": 0, "pcg": self.pcg_stats()})
        return res

    def theta_stream(self, host_in, host_out, h, theta=0.5, pcg_rtol=1e-8, pcg_maxit=2000, pcg_mode=0):
        """Host-buffer throughput API: one theta step for every pinned host state in ``host_in`` (list of 1-D float64
        tensors), results into the pinned tensors ``host_out``.  Double buffered: the upload of job j+1 and the
        download of job j-1 run on their own streams while job j is being stepped.  Returns when all results are
        in host memory."""
        torch = self.torch
        n = self.n_state
        st_in, st_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        if not hasattr(self, "_stage"):
            self._stage = [[torch.empty(n, dtype=torch.float64, device=self.dev) for _ in range(2)] for _ in range(2)]
        ins, outs = self._stage
        up = [torch.cuda.Event() for _ in host_in]
        done = [torch.cuda.Event() for _ in host_in]
        down = [torch.cuda.Event() for _ in host_in]
        torch.cuda.synchronize(self.dev)
        for j, (hi, ho) in enumerate(zip(host_in, host_out)):
            with torch.cuda.stream(st_in):
                if j >= 2:
                    st_in.wait_event(done[j - 2])                 # staging slot free again
                ins[j & 1].copy_(hi, non_blocking=True)
                up[j].record(st_in)
            self.stream.wait_event(up[j])
            if j >= 2:
                self.stream.wait_event(down[j - 2])               # result slot drained
            _cabi.check(self.lib.tc_set_state(self._handle, self._p(ins[j & 1])))
            _cabi.check(self.lib.tc_theta_steps(self._handle, 1, float(theta), float(h), float(pcg_rtol),
                                                int(pcg_maxit), int(pcg_mode)))
            _cabi.check(self.lib.tc_get_state_async(self._handle, self._p(outs[j & 1])))
            done[j].record(self.stream)
            with torch.cuda.stream(st_out):
                st_out.wait_event(done[j])
                ho.copy_(outs[j & 1], non_blocking=True)
                down[j].record(st_out)
        torch.cuda.synchronize(self.dev)
        bad = self.pcg_status()
        if bad["failed_solves"]:
            raise RuntimeError(f"PCG did not converge in {bad['failed_solves']} streamed steps ({bad['reason']})")

    def pcg_status(self):
        """Failed PCG solves since the last call (csrc/pcg.cuh status flag); cleared on read."""
        st = (_cabi.i64 * 2)()
        _cabi.check(self.lib.tc_pcg_status(self._handle, st))
        reason = {0: "none", 1: "iteration limit", 2: "breakdown (non-finite or non-positive curvature)"}.get(int(st[0]), "?")
        return {"status": int(st[0]), "failed_solves": int(st[1]), "reason": reason}

    def timing(self, enable=True):
        """Toggle CUDA-event timing of PCG launches; returns what accumulated since the last call."""
        ms, n, launches = _cabi.f64(), _cabi.i64(), _cabi.i64()
        _cabi.check(self.lib.tc_timing(self._handle, int(bool(enable)), C.byref(ms), C.byref(n), C.byref(launches)))
        return {"pcg_ms": ms.value, "pcg_launches": int(n.value), "launches": int(launches.value)}

    def pcg_profile(self):
        names = ["loop top", "spmv", "dot + reduce-barrier (p.q)", "update x,r", "reduce-barrier (r.z, r.r)",
                 "direction", "barrier"]
        buf = (C.c_uint64 * 16)()
        _cabi.check(self.lib.tc_pcg_profile(self._handle, buf))
        return {n: buf[i] / 1e6 for i, n in enumerate(names)}

    def pcg_stats(self):
        st = (_cabi.i64 * 4)()
        _cabi.check(self.lib.tc_pcg_stats(self._handle, st))
        return {"iterations": int(st[0]), "solves": int(st[1]), "last": int(st[2]), "coop_grid": int(st[3])}

    def close(self):
        if self._handle:
            self.lib.tc_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
