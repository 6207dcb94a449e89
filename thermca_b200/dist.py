"""Multi-GPU plumbing: row partition of one large FE body, halo plan, IPC exchange of peer buffers and
the driver of the fused peer-memory theta/PCG kernel (csrc/dist.cu).

One process per GPU (torchrun).  ``torch.distributed`` is used only for setup (all-gather of
CUDA IPC handles, sizes and ghost lists), barriers and the max-over-ranks timing; the data path — ghost
values and the PCG scalar reductions — moves by peer stores over NVLink inside the kernel.
The reference has no multi-process code; per rank the arithmetic is
the single-GPU path (thermca/solver.py:265-281 + the implicit solve) on the local rows.

Layout: every rank owns the contiguous global rows [bounds[rank], bounds[rank+1]) of a (bandwidth
reducing) ordering.  Ghosts = the sorted global columns outside the own range that the local rows
reference.  The SELL operator of a rank keeps its LOCAL columns; slices with couplings to ghosts are
*boundary slices* and carry those couplings in a short side list (ghost slot, value).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


# ---- host logic (numpy; covered by the world-size-2 gloo tests) ---------------------------------
def slab_partition(nx, ny, nz, world):
    """Contiguous row slabs cut between y-layers of the Kuhn cuboid (rows are ordered with y
    slowest, thermca_b200/synth.py:vertex_index).  Returns (bounds[world+1], halo_rows): a row of
    layer j couples to layers j-1 .. j+1 only, the largest column offset is one layer + one
    z-column + 1."""
    layer = (nx + 1) * (nz + 1)
    layers = ny + 1
    base, rem = divmod(layers, world)
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    if min(counts) < 2:
        raise ValueError("fewer than two y-layers per rank")
    bounds = np.concatenate([[0], np.cumsum(counts)]) * layer
    halo = layer + (nz + 1) + 1
    return bounds.astype(np.int64), int(halo)


def even_row_bounds(n, world, multiple=32):
    """Contiguous row blocks of (almost) equal size, cut at multiples of the SELL slice height."""
    per = -(-n // world)
    per = -(-per // multiple) * multiple
    bounds = np.minimum(np.arange(world + 1, dtype=np.int64) * per, n)
    if np.any(np.diff(bounds) <= 0):
        raise ValueError("more ranks than row blocks")
    return bounds


def ghost_columns(cols_global, r0, r1):
    """Sorted unique global columns outside [r0, r1) among ``cols_global`` (1-D integer array)."""
    c = np.asarray(cols_global)
    out = c[(c < r0) | (c >= r1)]
    return np.unique(out).astype(np.int64)


def remap_columns(cols_global, r0, r1, ghosts):
    """Global column -> index into the extended vector [local | ghosts]."""
    c = np.asarray(cols_global).astype(np.int64)
    inside = (c >= r0) & (c < r1)
    loc = np.where(inside, c - r0, (r1 - r0) + np.searchsorted(ghosts, c))
    return loc


def send_plan(all_ghosts, bounds, rank):
    """What this rank must push: ``all_ghosts[w]`` = sorted ghost columns of rank w.  Returns
    (send_row int32, send_peer int32, send_dst int64, neighbours int32): entry e sends local row
    send_row[e] into ghost slot send_dst[e] of rank send_peer[e]."""
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    rows, peers, dsts, nbrs = [], [], [], set()
    for w, g in enumerate(all_ghosts):
        if w == rank:
            continue
        g = np.asarray(g, dtype=np.int64)
        hit = np.nonzero((g >= r0) & (g < r1))[0]
        if hit.size:
            rows.append(g[hit] - r0)
            peers.append(np.full(hit.size, w))
            dsts.append(hit)
            nbrs.add(w)
    # the ranks that own my ghosts are neighbours as well (symmetric for a symmetric pattern; kept general)
    mine = np.asarray(all_ghosts[rank], dtype=np.int64)
    if mine.size:
        owners = np.searchsorted(bounds, mine, side="right") - 1
        nbrs.update(int(o) for o in np.unique(owners))
    cat = lambda v, dt: np.concatenate(v).astype(dt) if v else np.zeros(0, dtype=dt)
    return cat(rows, np.int32), cat(peers, np.int32), cat(dsts, np.int64), np.array(sorted(nbrs), dtype=np.int32)


def emulate_push(x_local, n_ghost, plan, rank, world, all_gather_object):
    """Host restatement of the in-kernel ghost exchange (used by the gloo tests): returns the extended vector."""
    send_row, send_peer, send_dst, _ = plan
    out = [None] * world
    all_gather_object(out, (send_peer, send_dst, np.asarray(x_local)[send_row]))
    ext = np.concatenate([np.asarray(x_local, dtype=np.float64), np.full(n_ghost, np.nan)])
    for w, (peer, dst, val) in enumerate(out):
        sel = peer == rank
        ext[len(x_local) + dst[sel]] = val[sel]
    return ext


def local_spmv_with_halo(A_rows, x_local, row_begin, halo, recv_lo, recv_hi, n_global):
    """Host restatement of a slab product with +-1 neighbours: ``A_rows`` holds the rows of this rank with
    GLOBAL columns; the gathered vector is [lower halo | local | upper halo]."""
    n_loc = len(x_local)
    ext = np.zeros(n_global)
    ext[row_begin: row_begin + n_loc] = x_local
    if recv_lo is not None:
        ext[row_begin - halo: row_begin] = recv_lo
    if recv_hi is not None:
        ext[row_begin + n_loc: row_begin + n_loc + halo] = recv_hi
    return A_rows @ ext


def csr_rows_to_sell(indptr, indices, data, n_rows, pad_col_offset=0):
    """SELL-32 arrays (slice_ptr int64, col int64, val f64) of CSR rows; short rows are padded with
    (0, own row + pad_col_offset) as csrc/parts.cuh expects.  Row i of the CSR is local row i; columns are
    kept as given (pass the global row offset as ``pad_col_offset`` when they are global)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    lens = np.diff(indptr)
    nsl = (n_rows + 31) // 32
    lens_p = np.zeros(nsl * 32, dtype=np.int64)
    lens_p[:n_rows] = lens
    width = lens_p.reshape(nsl, 32).max(axis=1)
    slice_ptr = np.concatenate([[0], np.cumsum(width * 32)]).astype(np.int64)
    total = int(slice_ptr[-1])
    pos = np.arange(total, dtype=np.int64)
    sl = np.searchsorted(slice_ptr, pos, side="right") - 1
    lane = (pos - slice_ptr[sl]) & 31
    col = np.minimum(32 * sl + lane, n_rows - 1) + pad_col_offset          # padding: own row
    val = np.zeros(total, dtype=np.float64)
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), lens)
    k = np.arange(int(indptr[-1]), dtype=np.int64) - np.repeat(indptr[:-1], lens)
    p = slice_ptr[rows >> 5] + 32 * k + (rows & 31)
    col[p] = np.asarray(indices, dtype=np.int64)
    val[p] = np.asarray(data, dtype=np.float64)
    return slice_ptr, col, val


# ---- device side ------------------------------------------------------------------------------------
class DistBody:
    """One FE body row-partitioned over the ranks of the default process group.

    ``op`` describes the LOCAL rows [row_begin, row_begin + n) on the device: ``slice_ptr`` (int64), ``col``
    (int32, GLOBAL columns), ``val``, optional ``col16`` (column - row), ``mass``, ``adiag`` (diagonal of
    A = M^-1 K incl. films), ``nslices``.  ``b``: load vector of the local rows (device)."""

    def __init__(self, op, b, bounds, init_temp=20.0, stream=None, presplit=None):
        """``presplit`` (host-side split done by ``split_local_rows``): dict with ``ghosts`` (sorted global columns),
        ``bidx`` / ``gptr`` / ``gslot`` / ``gval`` (numpy); then ``op`` already holds LOCAL columns only.
        ``stream``: run on this torch stream (a solver's) instead of an own one."""
        import torch
        import torch.distributed as dist
        from . import _cabi
        self.torch, self.dist, self.lib = torch, dist, _cabi.load_library()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.bounds = np.asarray(bounds, dtype=np.int64)
        r0, r1 = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.row_begin, self.n, self.n_global = r0, r1 - r0, int(self.bounds[-1])
        assert int(op["n"]) == self.n
        self.op, self.b = op, b
        nsl = int(op["nslices"])
        col = op["col"]
        if presplit is not None:
            up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev).to(dt)
            ghosts = up(presplit["ghosts"], torch.int64)
            self.n_ghost = int(ghosts.numel())
            self.bidx = up(presplit["bidx"], torch.int32)
            self.n_bnd = int((presplit["bidx"] >= 0).sum())
            self.gptr = up(presplit["gptr"], torch.int64)
            self.gslot = up(presplit["gslot"] if len(presplit["gslot"]) else np.zeros(1), torch.int32)
            self.gval = up(presplit["gval"] if len(presplit["gval"]) else np.zeros(1), torch.float64)
            self.col_loc = op["col"]
        else:
            ghosts = self._split_on_device(op, col, r0, r1, nsl)
        self.stream = torch.cuda.Stream(self.dev) if stream is None else stream
        self._finish_setup(op, ghosts, nsl, init_temp)

    def _split_on_device(self, op, col, r0, r1, nsl):
        """Ghosts, boundary slices and ghost side lists of an operator with GLOBAL columns (on the device: 10^8
        entries at 10 M rows); leaves LOCAL columns in ``op``."""
        torch = self.torch
        outside = (col < r0) | (col >= r1)
        gpos = torch.nonzero(outside).ravel()                                # entry positions of ghost couplings
        gcol = col[gpos].to(torch.int64)
        ghosts = torch.unique(gcol)                                          # sorted global columns
        self.n_ghost = int(ghosts.numel())
        sl_of = torch.searchsorted(op["slice_ptr"], gpos, right=True) - 1    # slice of every ghost coupling
        lane_of = (gpos - op["slice_ptr"][sl_of]) & 31
        bnd_mask = torch.zeros(nsl, dtype=torch.bool, device=self.dev)
        bnd_mask[sl_of] = True
        ids = torch.arange(nsl, device=self.dev, dtype=torch.int32)
        bnd_ids = ids[bnd_mask]
        self.n_bnd = int(bnd_ids.numel())
        # side lists: SELL-32 per boundary slice, entry k of lane l = k-th ghost coupling of that row (stored order)
        b_of_slice = torch.full((nsl,), -1, dtype=torch.int64, device=self.dev)
        b_of_slice[bnd_ids.to(torch.int64)] = torch.arange(self.n_bnd, device=self.dev)
        self.bidx = b_of_slice.to(torch.int32)
        key = sl_of * 32 + lane_of                                           # gpos ascending -> ascending k within a row
        skey, perm = torch.sort(key, stable=True)
        uniq, counts = torch.unique_consecutive(skey, return_counts=True)
        first = torch.cumsum(counts, 0) - counts
        k_idx = torch.arange(skey.numel(), device=self.dev) - torch.repeat_interleave(first, counts)
        gwidth = torch.zeros(max(self.n_bnd, 1), dtype=torch.int64, device=self.dev)
        if self.n_bnd:
            gwidth.scatter_reduce_(0, b_of_slice[uniq // 32], counts, reduce="amax", include_self=True)
        gptr = torch.zeros(self.n_bnd + 1, dtype=torch.int64, device=self.dev)
        gptr[1:] = torch.cumsum(gwidth[: self.n_bnd] * 32, 0)
        total = int(gptr[-1].item()) if self.n_bnd else 0
        self.gslot = torch.zeros(max(total, 1), dtype=torch.int32, device=self.dev)
        self.gval = torch.zeros(max(total, 1), dtype=torch.float64, device=self.dev)
        if total:
            sp_pos, sp_sl, sp_lane = gpos[perm], sl_of[perm], lane_of[perm]
            dst = gptr[b_of_slice[sp_sl]] + 32 * k_idx + sp_lane
            self.gslot[dst] = torch.searchsorted(ghosts, gcol[perm]).to(torch.int32)
            self.gval[dst] = op["val"][sp_pos]
        self.gptr = gptr
        # the main SELL keeps LOCAL columns only: ghost couplings become padding (0, own row)
        row_of = (sl_of * 32 + lane_of).clamp(max=self.n - 1)
        col_loc = col.to(torch.int64) - r0
        col_loc[gpos] = row_of
        op["val"][gpos] = 0.0
        if op.get("col16") is not None:
            op["col16"][gpos] = (row_of - (sl_of * 32 + lane_of)).to(torch.int16)
        self.col_loc = col_loc.to(torch.int32)
        op["col"] = self.col_loc                 # the global columns are not needed any more (0.6 GB at 10 M rows)
        del col_loc, outside, col, gpos, gcol, sl_of, lane_of, key, skey, perm
        return ghosts

    def _finish_setup(self, op, ghosts, nsl, init_temp):
        import ctypes as C
        from . import _cabi
        torch, dist = self.torch, self.dist
        self.nnz = int(torch.count_nonzero(op["val"]).item())
        # ---- halo plan (setup-time collectives) --------------------------------------------------------
        all_ghosts = [None] * self.world
        dist.all_gather_object(all_ghosts, ghosts.cpu().numpy())
        self.plan = send_plan(all_ghosts, self.bounds, self.rank)
        ext = [len(all_ghosts[w]) for w in range(self.world)]          # ghost count of every rank
        torch.cuda.synchronize(self.dev)
        self._h = _cabi.vp()
        _cabi.check(self.lib.tc_dist_create(C.byref(self._h), C.c_void_p(self.stream.cuda_stream), self.rank,
                                            self.world, self.n, self.n_ghost))
        hbuf = (C.c_ubyte * 64)()
        _cabi.check(self.lib.tc_dist_ipc_handle(self._h, hbuf))
        allh = [None] * self.world
        dist.all_gather_object(allh, bytes(hbuf))
        handles = b"".join(allh)
        self._handles = C.create_string_buffer(handles, len(handles))
        _cabi.check(self.lib.tc_dist_open_peers(self._h, self._handles, (_cabi.i64 * self.world)(*ext)))
        sr, sp_, sd, nb = self.plan
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        self._plan_keep = [np.ascontiguousarray(a) for a in (sr, sp_, sd, nb)]
        sr, sp_, sd, nb = self._plan_keep
        _cabi.check(self.lib.tc_dist_set_halo(self._h, len(sr), ptr(sr), ptr(sp_), ptr(sd), len(nb), ptr(nb)))
        p = lambda t: C.c_void_p(t.data_ptr())
        c16 = op.get("col16") if os.environ.get("TC_COL16", "1") != "0" else None
        _cabi.check(self.lib.tc_dist_set_operator(self._h, nsl, p(op["slice_ptr"]), p(self.col_loc),
                                                  p(c16) if c16 is not None else C.c_void_p(0),
                                                  p(op["val"]), p(op["mass"]), p(op["adiag"]),
                                                  p(self.b) if self.b is not None else C.c_void_p(0),
                                                  self.n_bnd, p(self.bidx), p(self.gptr), p(self.gslot), p(self.gval)))
        if init_temp is not None:
            self.set_state(torch.full((self.n,), float(init_temp), dtype=torch.float64, device=self.dev))
        self.set_pcg(os.environ.get("TC_DIST_PCG", "auto"))
        dist.barrier()

    def set_pcg(self, mode="auto"):
        """PCG formulation of ``theta_steps`` (csrc/dist.cu), the same on every rank: 'classic' (two all-reduces and a
        barrier per iteration: the default, fastest at every measured size), 'sr' (single-reduction: one all-reduce of
        seven sums + a local barrier per iteration) or 'pipe' (pipelined recurrences fused into the product sweep: one
        sweep + one all-reduce per iteration).  The alternatives are measured in DESIGN.md section 6."""
        from . import _cabi
        if mode == "auto":
            mode = "classic"
        want = {"classic": 0, "pipe": 1, "sr": 2}[mode]
        rc = int(self.lib.tc_dist_set_pcg(self._h, want))
        if want:
            # a body the alternative kernels cannot serve on ANY rank falls back to the classic kernel on ALL ranks
            flag = self.torch.tensor([rc], dtype=self.torch.int64, device=self.dev)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
            if int(flag.item()) != 0:
                _cabi.check(self.lib.tc_dist_set_pcg(self._h, 0))
        else:
            _cabi.check(rc)
        self.pcg = ("classic", "pipe", "sr")[int(self.lib.tc_dist_get_pcg(self._h))]
        return self.pcg

    # ---- construction from a replicated host operator (a real, lowered FE body) -------------------------
    @classmethod
    def from_csr(cls, A, mass, b, init_temp=20.0, reorder=True):
        """Partition the body whose operator ``A = M^-1 K`` (scipy CSR, replicated on every rank, e.g.
        ``FFEIO.A`` of a lowered part, thermca/fem/ffe_io.py:40-65), lumped mass and load vector are given on the
        host.  ``reorder``: reverse Cuthill-McKee (scipy) so that contiguous row blocks have few ghosts; the
        user-visible DOF order is kept (``gather_state`` / ``set_global_state`` permute back)."""
        import scipy.sparse as sp
        import torch
        import torch.distributed as dist
        A = sp.csr_matrix(A)
        n = A.shape[0]
        if reorder:
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            perm = np.asarray(reverse_cuthill_mckee(sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr),
                                                                   shape=A.shape), symmetric_mode=True), dtype=np.int64)
        else:
            perm = np.arange(n, dtype=np.int64)
        inv = np.empty(n, dtype=np.int64)
        inv[perm] = np.arange(n)
        Ap = A[perm][:, perm].tocsr()
        Ap.sort_indices()
        rank, world = dist.get_rank(), dist.get_world_size()
        bounds = even_row_bounds(n, world)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        rows = Ap[r0:r1]
        slice_ptr, col, val = csr_rows_to_sell(rows.indptr, rows.indices, rows.data, r1 - r0, pad_col_offset=r0)
        dev = torch.device("cuda", torch.cuda.current_device())
        nsl = len(slice_ptr) - 1
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
        mass_p = np.asarray(mass, dtype=np.float64)[perm]
        # 16-bit relative columns when the band allows it (interior slices only use them)
        pos = np.arange(col.size, dtype=np.int64)
        sl = np.searchsorted(slice_ptr, pos, side="right") - 1
        row_g = r0 + 32 * sl + ((pos - slice_ptr[sl]) & 31)
        rel = col - np.minimum(row_g, r1 - 1)
        inside = (col >= r0) & (col < r1)
        c16 = None
        if np.all(np.abs(rel[inside]) < 32768):
            c16 = t(np.where(inside, rel, 0), torch.int16)
        op = {"n": r1 - r0, "nslices": nsl, "slice_ptr": t(slice_ptr, torch.int64), "col": t(col, torch.int32),
              "val": t(val, torch.float64), "mass": t(mass_p[r0:r1], torch.float64),
              "adiag": t(Ap.diagonal()[r0:r1], torch.float64), "col16": c16}
        body = cls(op, t(np.asarray(b, dtype=np.float64)[perm][r0:r1], torch.float64), bounds,
                   init_temp=init_temp)
        body.perm, body.inv_perm = perm, inv
        return body

    # ---- state ------------------------------------------------------------------------------------------
    def set_state(self, y):
        from . import _cabi
        self.torch.cuda.synchronize(self.dev)
        self._y_keep = y
        _cabi.check(self.lib.tc_dist_set_state(self._h, C.c_void_p(y.data_ptr())))

    def get_state(self, out=None):
        from . import _cabi
        if out is None:
            out = self.torch.empty(self.n, dtype=self.torch.float64, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_dist_get_state(self._h, C.c_void_p(out.data_ptr())))
        return out

    def set_load(self, b):
        """Swap the load vector of the local rows (kept by reference)."""
        from . import _cabi
        self.b = b
        _cabi.check(self.lib.tc_dist_set_load(self._h, C.c_void_p(b.data_ptr())))

    def gather_state(self):
        """Global state in the caller's DOF order on every rank (setup / result path, NCCL all-gather)."""
        torch, dist = self.torch, self.dist
        mine = self.get_state()
        sizes = [int(self.bounds[w + 1] - self.bounds[w]) for w in range(self.world)]
        parts = [torch.empty(s, dtype=torch.float64, device=self.dev) for s in sizes]
        dist.all_gather(parts, mine)
        y = torch.cat(parts)
        inv = getattr(self, "inv_perm", None)
        return y if inv is None else y[torch.from_numpy(inv).to(self.dev)]

    def set_global_state(self, y_global):
        perm = getattr(self, "perm", None)
        y = self.torch.as_tensor(y_global, dtype=self.torch.float64, device=self.dev)
        if perm is not None:
            y = y[self.torch.from_numpy(perm).to(self.dev)]
        self.set_state(y[self.row_begin: self.row_begin + self.n].contiguous())

    # ---- stepping ---------------------------------------------------------------------------------------
    def theta_steps(self, nsteps, h, theta=0.5, rtol=1e-8, maxit=2000, sync=False):
        from . import _cabi
        _cabi.check(self.lib.tc_dist_theta_steps(self._h, int(nsteps), float(theta), float(h), float(rtol), int(maxit)))
        if sync:
            self.check()

    def theta_stream(self, host_in, host_out, h, theta=0.5, rtol=1e-8, maxit=2000):
        """Host-buffer throughput API (per rank: the local rows): one theta step for every pinned host state in
        ``host_in``, results into the pinned ``host_out``; uploads / downloads are double buffered against the steps
        (see DeviceSolver.theta_stream)."""
        from . import _cabi
        torch = self.torch
        st_in, st_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        if not hasattr(self, "_stage"):
            self._stage = [[torch.empty(self.n, dtype=torch.float64, device=self.dev) for _ in range(2)] for _ in range(2)]
        ins, outs = self._stage
        up = [torch.cuda.Event() for _ in host_in]
        done = [torch.cuda.Event() for _ in host_in]
        down = [torch.cuda.Event() for _ in host_in]
        torch.cuda.synchronize(self.dev)
        for j, (hi, ho) in enumerate(zip(host_in, host_out)):
            with torch.cuda.stream(st_in):
                if j >= 2:
                    st_in.wait_event(done[j - 2])
                ins[j & 1].copy_(hi, non_blocking=True)
                up[j].record(st_in)
            self.stream.wait_event(up[j])
            if j >= 2:
                self.stream.wait_event(down[j - 2])
            _cabi.check(self.lib.tc_dist_copy_state_async(self._h, C.c_void_p(ins[j & 1].data_ptr()), 0))
            _cabi.check(self.lib.tc_dist_theta_steps(self._h, 1, float(theta), float(h), float(rtol), int(maxit)))
            _cabi.check(self.lib.tc_dist_copy_state_async(self._h, C.c_void_p(outs[j & 1].data_ptr()), 1))
            done[j].record(self.stream)
            with torch.cuda.stream(st_out):
                st_out.wait_event(done[j])
                ho.copy_(outs[j & 1], non_blocking=True)
                down[j].record(st_out)
        torch.cuda.synchronize(self.dev)
        self.check()

    def check(self):
        """Synchronise and raise on EVERY rank if any rank saw a communication timeout or a failed solve."""
        torch, dist = self.torch, self.dist
        rc = int(self.lib.tc_dist_check(self._h))
        msg = self.lib.tc_last_error().decode() if rc else ""
        flag = torch.tensor([rc], dtype=torch.int64, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        worst = int(flag.item())
        if worst:
            raise RuntimeError(f"row-partitioned step failed (worst code {worst} over ranks; this rank: {rc} {msg})")

    def reset_error(self):
        from . import _cabi
        _cabi.check(self.lib.tc_dist_reset_error(self._h))

    def stats(self):
        from . import _cabi
        st = (_cabi.i64 * 6)()
        _cabi.check(self.lib.tc_dist_stats(self._h, st))
        return {"iters_last": int(st[0]), "iters_total": int(st[1]), "error": int(st[2]),
                "solves": int(st[3]) % 1000000, "grid": int(st[3]) // 1000000, "failed_solves": int(st[4]),
                "late_ghosts": int(st[5])}

    PHASES = ["loop top", "spmv", "dot", "reduce-barrier p.q (+allreduce)", "update x,r",
              "reduce-barrier r.z,r.r (+allreduce)", "direction + halo push", "barrier"]

    def profile(self):
        """Accumulated per-phase wall time (ms) seen by block 0 and by the last block; cleared on read."""
        from . import _cabi
        buf = (C.c_uint64 * 32)()
        _cabi.check(self.lib.tc_dist_profile(self._h, buf))
        return {"block0": {n: buf[i] / 1e6 for i, n in enumerate(self.PHASES)},
                "last": {n: buf[16 + i] / 1e6 for i, n in enumerate(self.PHASES)}}

    def global_sums(self, y=None):
        """(sum_i M_i y_i, sum_i y_i^2) over all ranks: thermal energy / capacity-weighted mean and the 2-norm."""
        torch, dist = self.torch, self.dist
        y = self.get_state() if y is None else y
        v = torch.stack([(self.op["mass"] * y).sum(), (y * y).sum()])
        dist.all_reduce(v)
        return float(v[0].item()), float(v[1].item())

    def close(self):
        if self._h:
            self.lib.tc_dist_close_peers(self._h)     # drain + unmap the peers' buffers
            self.dist.barrier()                        # nobody still maps my buffer
            self.lib.tc_dist_destroy(self._h)
            self._h = None


class DistCuboid(DistBody):
    """BASELINE config #5: the synthetic FE cuboid row-partitioned over ``world`` GPUs (slabs along y)."""

    def __init__(self, nx, ny, nz, lx=0.5, ly=1.0, lz=0.05, heat=1000.0, film=10.0, env_temp=20.0,
                 init_temp=20.0, jitter=0.15, condy=50.0, vol_capy=8000.0 * 500.0):
        import torch
        import torch.distributed as dist
        from .workloads import device_cuboid_operator
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", torch.cuda.current_device())
        bounds, halo = slab_partition(nx, ny, nz, world)
        self.halo = halo
        self.dims = (nx, ny, nz)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        op = device_cuboid_operator(nx, ny, nz, lx, ly, lz, jitter, condy, vol_capy, r0, r1, dev)
        # surface areas are global sums of the local load columns
        areas = torch.stack([op["q_bottom"].sum(), op["q_upper"].sum()])
        dist.all_reduce(areas)
        self.areas = areas.cpu().numpy()
        minv = 1.0 / op["mass"]
        # constant film folded into the diagonal (thermca/fem/ffe_io.py:57-60); diagonal = slot 7 of 15
        rows = torch.nonzero(op["q_upper"]).ravel()
        add = film * minv[rows] * op["q_upper"][rows]
        pos = (rows >> 5) * (32 * 15) + 7 * 32 + (rows & 31)
        op["val"][pos] += add
        op["adiag"][rows] += add
        # load vector b = M^-1 (Qs flux), flux = [heat/area_bottom, film*T_env]  (solver.py:267-281)
        b = minv * (op["q_bottom"] * (heat / self.areas[0]) + op["q_upper"] * (film * env_temp))
        super().__init__(op, b, bounds, init_temp=init_temp)
