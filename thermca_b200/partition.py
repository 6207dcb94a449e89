"""General networks on several GPUs: ONE full-order FE body of a lowered network is row-partitioned, everything else
(network nodes, link programs, LP / MOR parts, other FE bodies) is replicated on every rank.

Per rank the lowered network is *localised*: the partitioned part keeps the rows this rank owns (reverse Cuthill-McKee
order, contiguous even blocks) with their LOCAL couplings, surface load / mean / film-diagonal lists restricted to
those rows; couplings to other ranks' rows become the ghost side lists of ``thermca_b200.dist.DistBody``.  The device
solver of the localised network (``DeviceSolver``) then gets the partition attached (``tc_attach_dist``): surface
means are completed across ranks (thermca/solver.py:211-220 is a sum over all rows of the body) before the replicated
network kernel runs, and the body's product / implicit solve happens in the fused kernel of csrc/dist.cu.

Host logic only (numpy / scipy); covered by CPU tests that re-assemble the global product from the local pieces.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .dist import even_row_bounds


def rcm_permutation(indptr, indices, n):
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    pat = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(n, n))
    return np.asarray(reverse_cuthill_mckee(pat, symmetric_mode=True), dtype=np.int64)


def permuted_entries(indptr, indices, n, perm):
    """Entry permutation of ``A[perm][:, perm]`` with sorted indices: returns (new_indptr, new_indices, eperm) with
    new_data = data[eperm] for ANY data array stored on the old pattern."""
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    indptr = np.asarray(indptr, dtype=np.int64)
    rows_old = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    r_new, c_new = inv[rows_old], inv[np.asarray(indices, dtype=np.int64)]
    eperm = np.lexsort((c_new, r_new))
    new_indptr = np.concatenate([[0], np.cumsum(np.bincount(r_new, minlength=n))]).astype(np.int64)
    return new_indptr, c_new[eperm], eperm, inv


def split_local_rows(indptr, indices, data, r0, r1):
    """Rows [r0, r1) of a CSR with sorted indices (global columns) -> local CSR (columns - r0, ghost couplings
    removed) + ghost side lists in the boundary-slice SELL layout of csrc/dist.cu.

    Returns (loc_indptr, loc_indices, keep_mask, presplit): ``keep_mask`` marks the entries of the row range that
    stay in the local CSR (apply it to any data array of those rows); ``presplit`` = dict(ghosts, bidx, gptr, gslot,
    gval)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    e0, e1 = int(indptr[r0]), int(indptr[r1])
    cols = np.asarray(indices[e0:e1], dtype=np.int64)
    vals = np.asarray(data[e0:e1], dtype=np.float64)
    n_loc = r1 - r0
    row_of = np.repeat(np.arange(n_loc, dtype=np.int64), np.diff(indptr[r0:r1 + 1]))
    keep = (cols >= r0) & (cols < r1)
    loc_indptr = np.concatenate([[0], np.cumsum(np.bincount(row_of[keep], minlength=n_loc))]).astype(np.int64)
    loc_indices = (cols[keep] - r0).astype(np.int32)
    # ghost couplings
    g_rows, g_cols, g_vals = row_of[~keep], cols[~keep], vals[~keep]
    ghosts = np.unique(g_cols)
    nsl = (n_loc + 31) // 32
    bidx = np.full(nsl, -1, dtype=np.int32)
    bsl = np.unique(g_rows >> 5)
    bidx[bsl] = np.arange(len(bsl), dtype=np.int32)
    if len(g_rows):
        # k-th ghost coupling of a row (stored order: entries are row-major with sorted columns)
        first = np.concatenate([[True], g_rows[1:] != g_rows[:-1]])
        start = np.maximum.accumulate(np.where(first, np.arange(len(g_rows)), 0))
        k_idx = np.arange(len(g_rows)) - start
        cnt = np.bincount(g_rows, minlength=nsl * 32)
        width = cnt.reshape(nsl, 32).max(axis=1)[bsl]
        gptr = np.concatenate([[0], np.cumsum(width * 32)]).astype(np.int64)
        gslot = np.zeros(int(gptr[-1]), dtype=np.int32)
        gval = np.zeros(int(gptr[-1]), dtype=np.float64)
        dst = gptr[bidx[g_rows >> 5]] + 32 * k_idx + (g_rows & 31)
        gslot[dst] = np.searchsorted(ghosts, g_cols).astype(np.int32)
        gval[dst] = g_vals
    else:
        gptr, gslot, gval = np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0)
    return loc_indptr, loc_indices, keep, {"ghosts": ghosts.astype(np.int64), "bidx": bidx, "gptr": gptr, "gslot": gslot,
                                           "gval": gval}


def localize_ffe_part(part, rank, world, perm=None):
    """The rows rank ``rank`` of ``world`` owns of a lowered full-order FE part (``spec["ffe"][k]``).

    Returns (local_part, presplit, info): ``local_part`` has the layout of a lowered part (thermca_b200/lowering.py),
    so ``DeviceSolver`` packs it like any other; ``info`` = dict(perm, inv, bounds, r0, r1)."""
    if part.get("tdep") is not None:
        raise NotImplementedError("temperature dependent bodies cannot be row-partitioned")
    n = int(part["dof"])
    ab = part["A_body"]
    indptr, indices = np.asarray(ab["indptr"], dtype=np.int64), np.asarray(ab["indices"], dtype=np.int64)
    if perm is None:
        perm = rcm_permutation(indptr, indices, n)
    new_indptr, new_indices, eperm, inv = permuted_entries(indptr, indices, n, perm)
    inv_eperm = np.empty(len(eperm), dtype=np.int64)
    inv_eperm[eperm] = np.arange(len(eperm))
    data_full = np.asarray(part["A_data"], dtype=np.float64)[eperm]          # with the initial films
    data_body = np.asarray(ab["data"], dtype=np.float64)[eperm]              # body only
    bounds = even_row_bounds(n, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    loc_indptr, loc_indices, keep, presplit = split_local_rows(new_indptr, new_indices, data_full, r0, r1)
    e0, e1 = int(new_indptr[r0]), int(new_indptr[r1])
    # position of a kept entry of the row range inside the local CSR
    loc_pos = np.cumsum(keep) - 1
    minv = np.asarray(part["M_inv"], dtype=np.float64)[perm][r0:r1]

    def rows_local(idx):
        """global DOF indices -> (mask of those owned here, local row indices of the owned ones)"""
        new = inv[np.asarray(idx, dtype=np.int64)]
        m = (new >= r0) & (new < r1)
        return m, new[m] - r0

    B_idx, B_val, C_idx, C_val = [], [], [], []
    for ix, v in zip(part["B_idx"], part["B_val"]):
        m, loc = rows_local(ix)
        B_idx.append(loc); B_val.append(np.asarray(v, dtype=np.float64)[m])
    for ix, v in zip(part["C_idx"], part["C_val"]):
        m, loc = rows_local(ix)
        C_idx.append(loc); C_val.append(np.asarray(v, dtype=np.float64)[m])
    film_didx, film_adata = [], []
    for ix, a in zip(part.get("film_didx", []), part.get("film_adata", [])):
        newpos = inv_eperm[np.asarray(ix, dtype=np.int64)]                  # positions in the permuted CSR
        m = (newpos >= e0) & (newpos < e1)
        rel = newpos[m] - e0
        if not keep[rel].all():
            raise AssertionError("a film diagonal entry is not a local coupling")
        film_didx.append(loc_pos[rel].astype(np.int64))
        film_adata.append(np.asarray(a, dtype=np.float64)[m])
    local = dict(part)
    local.update({
        "dof": np.int64(r1 - r0),
        "A_body": {"indptr": loc_indptr, "indices": loc_indices, "data": data_body[e0:e1][keep],
                   "shape": np.array([r1 - r0, r1 - r0])},
        "A_data": data_full[e0:e1][keep], "M_inv": minv,
        "B_idx": B_idx, "B_val": B_val, "C_idx": C_idx, "C_val": C_val,
        "film_didx": film_didx, "film_adata": film_adata,
    })
    return local, presplit, {"perm": perm, "inv": inv, "bounds": bounds, "r0": r0, "r1": r1}


def choose_partitioned_part(spec, min_dof=0):
    """Index of the FE part to partition: the largest full-order body without temperature dependent material."""
    best, best_n = -1, min_dof
    for k, p in enumerate(spec["ffe"]):
        if p.get("tdep") is None and int(p["dof"]) > best_n:
            best, best_n = k, int(p["dof"])
    return best


def gather_stored_rows(io_loc, a, n_loc, sizes, inv, device=None):
    """Stored part states of a partitioned adaptive run, rank-local -> global.  ``io_loc`` (stored steps x local part
    state) holds this rank's rows of the partitioned body in columns ``[a, a + n_loc)`` (partition order) between the
    replicated parts; returns the rows with the body completed over all ranks (``sizes`` rows per rank) and put back
    into the caller's DOF order (``inv``: global partition order -> caller's order, partition.localize_ffe_part).  The
    number of stored steps is the same on every rank (all ranks take bit-identical step decisions)."""
    import torch
    import torch.distributed as dist
    m = io_loc.shape[0]
    mine = torch.from_numpy(np.ascontiguousarray(io_loc[:, a: a + n_loc].T))          # rows of the body x steps
    if device is not None:
        mine = mine.to(device)
    widest = int(max(sizes))                                    # equal-sized pieces: every backend gathers those
    padded = torch.zeros((widest, m), dtype=torch.float64, device=mine.device)
    padded[:n_loc] = mine
    parts = [torch.empty((widest, m), dtype=torch.float64, device=mine.device) for _ in sizes]
    dist.all_gather(parts, padded)
    order = torch.from_numpy(np.asarray(inv, dtype=np.int64)).to(mine.device)
    body = torch.cat([p[:int(sz)] for p, sz in zip(parts, sizes)], dim=0)[order].T.cpu().numpy()
    return np.concatenate([io_loc[:, :a], body, io_loc[:, a + n_loc:]], axis=1)
