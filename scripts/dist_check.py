"""Multi-GPU check (torchrun, N >= 1): the fused peer-memory theta/PCG path on row blocks must reproduce the
single-GPU path on the same global problem.  Two bodies:
  * the synthetic Kuhn cuboid in y-slabs (rank +-1 neighbours),
  * the same operator as a "real" body handed over as a host CSR in a scrambled DOF order: reverse
    Cuthill-McKee reordering, even row blocks, general neighbour sets (DistBody.from_csr).
Prints one line per rank and `dist_check ok`; `--json PATH` writes the numbers (kept under profiles/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from thermca_b200.dist import DistBody, DistCuboid
    from thermca_b200.device import DeviceSolver
    from thermca_b200.workloads import device_cuboid_spec, sell_to_csr_host
    steps, rtol = 4, 1e-12
    nx, ny, nz = 20, 16 * world - 1, 6
    # ---- (a) cuboid slabs ------------------------------------------------------------------------
    dc = DistCuboid(nx, ny, nz)
    spec = device_cuboid_spec(nx, ny, nz)
    ref = DeviceSolver(spec)
    ref.theta_steps(steps, 1.0, 0.5, t0=0.0, pcg_rtol=rtol)
    y_full = ref.get_state()
    pst = ref.pcg_stats()
    ok, err_by_mode, iters_by_mode = True, {}, {}
    # the PCG formulations of the fused kernel: classic (same arithmetic as the single-GPU kernel: identical iteration
    # counts), pipelined (one sweep + one reduction per iteration) and single-reduction (one all-reduce of seven sums +
    # a local barrier per iteration); the latter two round differently: counts within one per step
    for mode in ("classic", "pipe", "sr"):
        assert dc.set_pcg(mode) == mode
        dc.set_state(torch.full((dc.n,), 20.0, dtype=torch.float64, device="cuda"))
        it0 = dc.stats()["iters_total"]
        dc.theta_steps(steps, 1.0, 0.5, rtol=rtol, sync=True)
        st = dc.stats()
        iters = st["iters_total"] - it0
        y_loc = dc.get_state().cpu().numpy()
        y_ref = y_full[dc.row_begin: dc.row_begin + dc.n]
        err_a = float(np.max(np.abs(y_loc - y_ref) / np.abs(y_ref)))
        e_d, n_d = dc.global_sums()
        slack = 0 if mode == "classic" else steps
        ok = ok and err_a < 1e-9 and st["error"] == 0 and st["failed_solves"] == 0 and abs(iters - pst["iterations"]) <= slack
        err_by_mode[mode], iters_by_mode[mode] = err_a, iters
        if not ok:
            print(f"[cuboid/{mode}] rank {rank}: CHECK FAILED err {err_a:.3e} stats {st} iters {iters} vs {pst['iterations']}", flush=True)
        print(f"[cuboid/{mode}] rank {rank}/{world}: rows [{dc.row_begin},{dc.row_begin + dc.n}) ghosts {dc.n_ghost} "
              f"boundary slices {dc.n_bnd} neighbours {list(dc.plan[3])} iters {iters} (single GPU "
              f"{pst['iterations']}) max rel err {err_a:.3e}", flush=True)
    err_a = max(err_by_mode.values())
    # ---- (b) a body handed over as host CSR in a scrambled order -----------------------------------
    op = spec["ffe"][0]["sell"]
    A = sell_to_csr_host(op).tocsr()
    n = A.shape[0]
    mass = op["mass"].cpu().numpy()
    part = spec["ffe"][0]
    b = np.zeros(n)
    flux = [1000.0 / part["surf_areas"][0], 10.0 * 20.0]
    for s in range(2):
        np.add.at(b, part["B_idx"][s], part["B_val"][s] * flux[s])
    scr = np.random.default_rng(3).permutation(n)                    # the caller's DOF order
    A_s = A[scr][:, scr].tocsr()
    body = DistBody.from_csr(A_s, mass[scr], b[scr], init_temp=20.0)
    err_b = 0.0
    for mode in ("classic", "pipe", "sr"):
        assert body.set_pcg(mode) == mode
        body.set_global_state(torch.full((n,), 20.0, dtype=torch.float64, device="cuda"))
        it0 = body.stats()["iters_total"]
        body.theta_steps(steps, 1.0, 0.5, rtol=rtol, sync=True)
        y_body = body.gather_state().cpu().numpy()                       # caller's order
        err_m = float(np.max(np.abs(y_body - y_full[scr]) / np.abs(y_full[scr])))
        stb = body.stats()
        ok = ok and err_m < 1e-9 and stb["error"] == 0 and stb["failed_solves"] == 0
        err_b = max(err_b, err_m)
        if not ok:
            print(f"[csr body/{mode}] rank {rank}: CHECK FAILED err {err_m:.3e} stats {stb}", flush=True)
        print(f"[csr body/{mode}] rank {rank}/{world}: rows {body.n} ghosts {body.n_ghost} boundary slices {body.n_bnd} "
              f"neighbours {list(body.plan[3])} iters {stb['iters_total'] - it0} max rel err {err_m:.3e}", flush=True)
    ref.close()
    body.close()
    dc.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    errs = torch.tensor([err_a, err_b], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    if rank == 0 and "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump({"world": world, "steps": steps, "pcg_rtol": rtol, "cuboid_cells": [nx, ny, nz],
                       "max_rel_err_cuboid_slabs": float(errs[0]), "max_rel_err_csr_body_rcm": float(errs[1]),
                       "iters_dist": iters_by_mode, "iters_single_gpu": pst["iterations"],
                       "max_rel_err_by_pcg_mode": err_by_mode,
                       "energy_sum_M_y": e_d, "ok": not bool(flag.item())}, f)
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("dist_check ok")


if __name__ == "__main__":
    main()
