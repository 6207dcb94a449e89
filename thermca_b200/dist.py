"""Multi-GPU plumbing: slab partition of one large FE body, IPC exchange of peer buffers and
the driver of the fused peer-memory theta/PCG kernel (csrc/dist.cu).

One process per GPU (torchrun).  ``torch.distributed`` is used only for setup (all-gather of
CUDA IPC handles and sizes), barriers and the max-over-ranks timing; the data path — halo
layers and the PCG scalar reductions — moves by peer stores over NVLink inside the kernel.
The reference has no multi-process code; per rank the arithmetic is
the single-GPU path on rows [row_begin, row_end).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


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


def local_spmv_with_halo(A_rows, x_local, row_begin, halo, recv_lo, recv_hi, n_global):
    """Host restatement of the slab product used by the gloo tests: ``A_rows`` holds the rows of
    this rank with GLOBAL columns; the gathered vector is [lower halo | local | upper halo]."""
    n_loc = len(x_local)
    ext = np.zeros(n_global)
    ext[row_begin: row_begin + n_loc] = x_local
    if recv_lo is not None:
        ext[row_begin - halo: row_begin] = recv_lo
    if recv_hi is not None:
        ext[row_begin + n_loc: row_begin + n_loc + halo] = recv_hi
    return A_rows @ ext


class DistCuboid:
    """BASELINE config #5: the synthetic FE cuboid row-partitioned over ``world`` GPUs."""

    def __init__(self, nx, ny, nz, lx=0.5, ly=1.0, lz=0.05, heat=1000.0, film=10.0, env_temp=20.0,
                 init_temp=20.0, jitter=0.15, condy=50.0, vol_capy=8000.0 * 500.0):
        import torch
        import torch.distributed as dist
        from . import _cabi
        from .workloads import device_cuboid_operator
        self.torch, self.dist, self.lib = torch, dist, _cabi.load_library()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        bounds, halo = slab_partition(nx, ny, nz, self.world)
        self.bounds, self.halo = bounds, halo
        r0, r1 = int(bounds[self.rank]), int(bounds[self.rank + 1])
        self.row_begin, self.n = r0, r1 - r0
        self.n_global = int(bounds[-1])
        op = device_cuboid_operator(nx, ny, nz, lx, ly, lz, jitter, condy, vol_capy, r0, r1, self.dev)
        self.op = op
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
        self.b = minv * (op["q_bottom"] * (heat / self.areas[0]) + op["q_upper"] * (film * env_temp))
        self.stream = torch.cuda.Stream(self.dev)
        self.nnz = int(torch.count_nonzero(op["val"]).item())
        torch.cuda.synchronize(self.dev)
        self._h = _cabi.vp()
        _cabi.check(self.lib.tc_dist_create(C.byref(self._h), C.c_void_p(self.stream.cuda_stream), self.rank,
                                            self.world, self.n, halo, r0))
        # exchange IPC handles and slab sizes (setup only)
        hbuf = (C.c_ubyte * 64)()
        _cabi.check(self.lib.tc_dist_ipc_handle(self._h, hbuf))
        mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=self.dev)
        allh = [torch.empty(64, dtype=torch.uint8, device=self.dev) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        handles = bytes(torch.cat(allh).cpu().numpy().tobytes())
        peer_n = (_cabi.i64 * self.world)(*[int(bounds[w + 1] - bounds[w]) for w in range(self.world)])
        self._handles = C.create_string_buffer(handles, len(handles))
        _cabi.check(self.lib.tc_dist_open_peers(self._h, self._handles, peer_n))
        p = lambda t: C.c_void_p(t.data_ptr())
        c16 = op.get("col16") if os.environ.get("TC_COL16", "1") != "0" else None
        _cabi.check(self.lib.tc_dist_set_operator(self._h, op["nslices"], p(op["slice_ptr"]), p(op["col"]),
                                                  p(c16) if c16 is not None else C.c_void_p(0),
                                                  p(op["val"]), p(op["mass"]), p(op["adiag"]), p(self.b)))
        self.set_state(torch.full((self.n,), float(init_temp), dtype=torch.float64, device=self.dev))
        dist.barrier()

    def set_state(self, y):
        self.torch.cuda.synchronize(self.dev)
        self._y_keep = y
        from . import _cabi
        _cabi.check(self.lib.tc_dist_set_state(self._h, C.c_void_p(y.data_ptr())))

    def get_state(self):
        from . import _cabi
        out = self.torch.empty(self.n, dtype=self.torch.float64, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        _cabi.check(self.lib.tc_dist_get_state(self._h, C.c_void_p(out.data_ptr())))
        return out

    def theta_steps(self, nsteps, h, theta=0.5, rtol=1e-8, maxit=2000):
        from . import _cabi
        _cabi.check(self.lib.tc_dist_theta_steps(self._h, int(nsteps), float(theta), float(h), float(rtol), int(maxit)))

    def stats(self):
        from . import _cabi
        st = (_cabi.i64 * 4)()
        _cabi.check(self.lib.tc_dist_stats(self._h, st))
        return {"iters_last": int(st[0]), "iters_total": int(st[1]), "error": int(st[2]),
                "solves": int(st[3]) % 1000000, "grid": int(st[3]) // 1000000}

    PHASES = ["rhs+init", "halo flag", "spmv", "-", "dot", "sync1", "allreduce pq",
              "update x,r", "sync2", "allreduce rho", "direction", "sync3"]

    def profile(self):
        """Accumulated per-phase wall time (ms) seen by block 0 and by the last block; cleared on read."""
        from . import _cabi
        buf = (C.c_uint64 * 32)()
        _cabi.check(self.lib.tc_dist_profile(self._h, buf))
        return {"block0": {n: buf[i] / 1e6 for i, n in enumerate(self.PHASES)},
                "last": {n: buf[16 + i] / 1e6 for i, n in enumerate(self.PHASES)}}

    def close(self):
        if self._h:
            self.dist.barrier()
            self.lib.tc_dist_destroy(self._h)
            self._h = None


def run_partitioned_bench(args, world, rank, local_rank):
    """bench.py leg for N > 1: weak scaling, args.dof rows per GPU, slabs along y."""
    import torch
    import torch.distributed as dist
    from .workloads import cuboid_dims
    import bench as B
    nx, ny, nz = cuboid_dims(args.dof)
    ny_tot = (ny + 1) * world - 1                     # the same number of y-layers per rank
    dc = DistCuboid(nx, ny_tot, nz, ly=1.0 * world, heat=B.HEAT * world, film=B.FILM, env_temp=B.T_ENV)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with B.ClockSampler(local_rank) as clocks:
        dc.theta_steps(args.warmup, B.H_STEP, B.THETA, B.PCG_RTOL)
        it0 = dc.stats()["iters_total"]
        dc.profile()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(dc.stream)
        dc.theta_steps(args.steps, B.H_STEP, B.THETA, B.PCG_RTOL)
        e1.record(dc.stream)
        torch.cuda.synchronize()
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    st = dc.stats()
    iters = st["iters_total"] - it0
    prof = dc.profile()
    err = torch.tensor([st["error"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    nnz = torch.tensor([dc.nnz], dtype=torch.int64, device="cuda")
    dist.all_reduce(nnz)
    n_tot = dc.n_global
    # e2e: state of every slab up from pinned host memory, one step, state down
    host_y = dc.get_state().cpu().pin_memory()
    dev_y = torch.empty(dc.n, dtype=torch.float64, device="cuda")
    e2e_steps = 2
    torch.cuda.synchronize(); dist.barrier()
    import time
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dev_y.copy_(host_y, non_blocking=True)
        dc.set_state(dev_y)
        dc.theta_steps(1, B.H_STEP, B.THETA, B.PCG_RTOL)
        host_y.copy_(dc.get_state())
    torch.cuda.synchronize(); dist.barrier()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    # per-rank busy / wait split (block 0 of every rank): exposes systematic skew between GPUs
    pl = prof["last"]
    mine = torch.tensor([pl["spmv"] + pl["update x,r"] + pl["direction"] + pl["dot"],
                         pl["allreduce pq"] + pl["allreduce rho"] + pl["sync1"] + pl["sync2"] + pl["sync3"]],
                        dtype=torch.float64, device="cuda")
    allp = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(allp, mine)
    per_rank = [[round(float(v[0]), 1), round(float(v[1]), 1)] for v in allp]
    peak, peak_kind = B.measured_peak()
    bytes_step = B.algorithmic_bytes_per_iteration(dc.n, dc.nnz) * (iters / args.steps) + dc.n * (12 * 15 + 80)
    achieved = bytes_step / (ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": "DOF*steps/s (implicit theta-PCG step, full FE cuboid)", "value": n_tot * args.steps / (ms * 1e-3),
        "unit": "DOF*steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg#5 row-partitioned FE cuboid {nx}x{ny_tot}x{nz} cells = {n_tot} DOF, "
                               f"{world} slabs, theta=0.5 Jacobi-PCG step, halo+reductions by peer stores in-kernel",
                   "dof": n_tot, "dof_per_gpu": dc.n, "nnz": int(nnz.item()), "pcg_rtol": B.PCG_RTOL,
                   "h_step_s": B.H_STEP, "iters_per_step": iters / args.steps, "halo_rows": dc.halo,
                   "l2": "per-GPU inputs exceed the 126 MB L2", "comm_error": int(err.item()),
                   "phase_ms_rank0": prof, "grid": st["grid"],
                   "busy_wait_ms_per_rank": per_rank},
        "clocks": clocks.summary(),
        "e2e": {"value": n_tot * e2e_steps / float(e2e.item()), "unit": "DOF*steps/s",
                "h2d_bytes_per_step": n_tot * 8, "d2h_bytes_per_step": n_tot * 8, "steps": e2e_steps},
        "gpu_launches": args.steps * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "tc_dist_theta_kernel (per GPU)", "peak_kind": peak_kind},
    }
    dc.close()
    return line
