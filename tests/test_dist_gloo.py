"""Host-side logic of the row partition on CPU: world_size 2 over gloo (no GPU)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from thermca_b200.dist import (slab_partition, local_spmv_with_halo, even_row_bounds, ghost_columns, remap_columns,
                               send_plan, emulate_push, csr_rows_to_sell)
from thermca_b200.synth import cuboid_spec


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nx, ny, nz, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = cuboid_spec(nx, ny, nz, jitter=0.15)
    p = spec["ffe"][0]
    n = int(p["dof"])
    A = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    mass = 1.0 / p["M_inv"]
    bounds, halo = slab_partition(nx, ny, nz, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    A_loc = A[r0:r1]
    # every column this slab touches lies inside [r0 - halo, r1 + halo)
    cols = A_loc.indices
    assert cols.min() >= max(r0 - halo, 0) and cols.max() < min(r1 + halo, n)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    x_loc = x[r0:r1].copy()

    def exchange(v_loc):
        lo = hi = None
        reqs = []
        if rank > 0:
            lo = torch.empty(halo, dtype=torch.float64)
            reqs += [dist.isend(torch.from_numpy(v_loc[:halo].copy()), rank - 1), dist.irecv(lo, rank - 1)]
        if rank < world - 1:
            hi = torch.empty(halo, dtype=torch.float64)
            reqs += [dist.isend(torch.from_numpy(v_loc[-halo:].copy()), rank + 1), dist.irecv(hi, rank + 1)]
        for r in reqs:
            r.wait()
        return (None if lo is None else lo.numpy()), (None if hi is None else hi.numpy())

    lo, hi = exchange(x_loc)
    y_loc = local_spmv_with_halo(A_loc, x_loc, r0, halo, lo, hi, n)
    assert np.array_equal(y_loc, (A @ x)[r0:r1])

    # distributed Jacobi-PCG on (M + s K) with one allreduce per dot product == serial CG
    shift = 0.5
    b = mass[r0:r1] * rng.standard_normal(n)[r0:r1]
    dinv = 1.0 / (mass[r0:r1] * (1.0 + shift * A.diagonal()[r0:r1]))
    xs = np.zeros(r1 - r0); r = b.copy(); z = dinv * r; pvec = z.copy()

    def gsum(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())
    rho = gsum(r @ z)
    for _ in range(25):
        lo, hi = exchange(pvec)
        q = mass[r0:r1] * (pvec + shift * local_spmv_with_halo(A_loc, pvec, r0, halo, lo, hi, n))
        alpha = rho / gsum(pvec @ q)
        xs += alpha * pvec; r -= alpha * q
        z = dinv * r
        rho_new = gsum(r @ z)
        pvec = z + (rho_new / rho) * pvec
        rho = rho_new
    res = gsum(r @ r) ** 0.5
    np.save(os.path.join(out_dir, f"x{rank}.npy"), xs)
    np.save(os.path.join(out_dir, f"b{rank}.npy"), b)
    assert res < 1e-8 * gsum(b @ b) ** 0.5
    dist.destroy_process_group()


def test_slab_partition_layers():
    bounds, halo = slab_partition(4, 9, 3, 2)
    layer = 5 * 4
    assert list(bounds) == [0, 5 * layer, 10 * layer] and halo == layer + 4 + 1
    bounds, _ = slab_partition(4, 10, 3, 4)
    assert list(np.diff(bounds) // layer) == [3, 3, 3, 2]
    with pytest.raises(ValueError):
        slab_partition(4, 2, 3, 4)


def test_partitioned_spmv_and_pcg_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 5, 9, 3, str(tmp_path)), nprocs=2, join=True)
    # assembled solution solves the global system
    spec = cuboid_spec(5, 9, 3, jitter=0.15)
    p = spec["ffe"][0]
    n = int(p["dof"])
    A = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    mass = 1.0 / p["M_inv"]
    x = np.concatenate([np.load(tmp_path / "x0.npy"), np.load(tmp_path / "x1.npy")])
    b = np.concatenate([np.load(tmp_path / "b0.npy"), np.load(tmp_path / "b1.npy")])
    assert np.linalg.norm(mass * (x + 0.5 * (A @ x)) - b) < 1e-8 * np.linalg.norm(b)


def _sell_matvec(slice_ptr, col, val, x_ext, n_rows):
    """What csrc/dist.cu:slab_spmv computes for the local rows, on the host."""
    y = np.zeros(n_rows)
    for s_ in range(len(slice_ptr) - 1):
        w = (slice_ptr[s_ + 1] - slice_ptr[s_]) // 32
        for lane in range(32):
            r = 32 * s_ + lane
            if r < n_rows:
                acc = 0.0
                for k in range(w):
                    p_ = slice_ptr[s_] + 32 * k + lane
                    acc += val[p_] * x_ext[col[p_]]
                y[r] = acc
    return y


def _worker_general(rank, world, port, out_dir):
    """General partition: scrambled DOF order -> RCM -> even row blocks -> ghost lists, column remap, send plan,
    emulated push: the partitioned product and a distributed PCG equal the serial ones."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    spec = cuboid_spec(4, 11, 3, jitter=0.15)
    p = spec["ffe"][0]
    n = int(p["dof"])
    A0 = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    scr = np.random.default_rng(11).permutation(n)
    A = A0[scr][:, scr].tocsr()
    mass = (1.0 / p["M_inv"])[scr]
    perm = np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True), dtype=np.int64)
    Ap = A[perm][:, perm].tocsr(); Ap.sort_indices()
    mass_p = mass[perm]
    bounds = even_row_bounds(n, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    rows = Ap[r0:r1]
    slice_ptr, col_g, val = csr_rows_to_sell(rows.indptr, rows.indices, rows.data, r1 - r0, pad_col_offset=r0)
    ghosts = ghost_columns(col_g, r0, r1)
    col = remap_columns(col_g, r0, r1, ghosts)
    assert col.min() >= 0 and col.max() < (r1 - r0) + len(ghosts)
    all_ghosts = [None] * world
    dist.all_gather_object(all_ghosts, ghosts)
    plan = send_plan(all_ghosts, bounds, rank)
    # every ghost slot is written exactly once
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    ext = emulate_push(x[r0:r1], len(ghosts), plan, rank, world, dist.all_gather_object)
    assert not np.isnan(ext).any()
    assert np.array_equal(ext[r1 - r0:], x[ghosts])
    y_loc = _sell_matvec(slice_ptr, col, val, ext, r1 - r0)
    assert np.allclose(y_loc, (Ap @ x)[r0:r1], rtol=1e-13, atol=1e-13)
    # neighbours are symmetric here
    nb = [None] * world
    dist.all_gather_object(nb, list(plan[3]))
    for w in plan[3]:
        assert rank in nb[int(w)]
    np.save(os.path.join(out_dir, f"nbr{rank}.npy"), plan[3])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_general_partition_ghost_plan(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker_general, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(len(np.load(tmp_path / f"nbr{r}.npy")) >= 1 for r in range(world))


def test_even_row_bounds():
    b = even_row_bounds(1000, 3)
    assert b[0] == 0 and b[-1] == 1000 and all(v % 32 == 0 for v in b[:-1])
    with pytest.raises(ValueError):
        even_row_bounds(40, 4)


def _worker_adaptive(rank, world, port, out_dir):
    """Host-side bookkeeping of the adaptive schemes on a partitioned body (csrc/solver.cu tc_dist_norm_* +
    csrc/dist.cu tc_dist_sum_exchange, thermca_b200.partition.gather_stored_rows): (a) the class-split, rank-ordered
    error norm equals scipy's rms norm over the GLOBAL state (scipy rk.py:145-165) and is bit-identical on every rank;
    (b) stored rank-local rows are completed over the ranks and returned in the caller's DOF order."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from thermca_b200.partition import gather_stored_rows
    rng = np.random.default_rng(3)
    n_a, n_b, n_body = 5, 7, 32 * 9 + 11                    # replicated entries before / after the body
    bounds = even_row_bounds(n_body, world)
    perm = rng.permutation(n_body)                           # caller's order -> partition order
    inv = np.empty(n_body, dtype=np.int64)
    inv[perm] = np.arange(n_body)
    m = 4                                                    # stored steps
    glob = rng.standard_normal((m, n_a + n_body + n_b))      # caller's layout [repl | body | repl]
    body_part = glob[:, n_a:n_a + n_body][:, perm]           # partition order
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    io_loc = np.concatenate([glob[:, :n_a], body_part[:, r0:r1], glob[:, n_a + n_body:]], axis=1)
    sizes = [int(bounds[w + 1] - bounds[w]) for w in range(world)]
    back = gather_stored_rows(io_loc, n_a, r1 - r0, sizes, inv)
    assert np.array_equal(back, glob)
    # (a) error norm of one attempt: e = err / scale per entry
    e_glob = glob[0]
    e_loc = io_loc[0]
    lo, hi = n_a, n_a + (r1 - r0)
    cls_body = float(np.sum(e_loc[lo:hi] ** 2))
    cls_repl = float(np.sum(e_loc[:lo] ** 2) + np.sum(e_loc[hi:] ** 2))
    mine = torch.tensor([cls_body + (cls_repl if rank == 0 else 0.0),
                         float((r1 - r0) + (n_a + n_b if rank == 0 else 0))], dtype=torch.float64)
    allv = [torch.empty(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allv, mine)
    tot, cnt = 0.0, 0.0
    for w in range(world):                                   # rank order, as the device exchange adds them
        tot += float(allv[w][0]); cnt += float(allv[w][1])
    err = np.sqrt(tot) / np.sqrt(cnt)
    assert cnt == e_glob.size
    assert abs(err - np.linalg.norm(e_glob) / np.sqrt(e_glob.size)) <= 1e-15 * err
    errs = [None] * world
    dist.all_gather_object(errs, err.hex() if hasattr(err, "hex") else float(err).hex())
    assert len(set(errs)) == 1                               # identical decisions on every rank
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_adaptive_bookkeeping(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker_adaptive, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}.npy").exists() for r in range(world))
