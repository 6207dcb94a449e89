"""Multi-GPU check (torchrun, N >= 2): the fused peer-memory theta/PCG path on row slabs must
reproduce the single-GPU path on the same global mesh."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from thermca_b200.dist import DistCuboid
    from thermca_b200.device import DeviceSolver
    from thermca_b200.workloads import device_cuboid_spec
    nx, ny, nz = 20, 16 * world - 1, 6
    dc = DistCuboid(nx, ny, nz)
    dc.theta_steps(4, 1.0, 0.5, rtol=1e-12)
    st = dc.stats()
    y_loc = dc.get_state().cpu().numpy()
    ok = True
    # single-GPU reference of the same global problem on every rank (small)
    ref = DeviceSolver(device_cuboid_spec(nx, ny, nz))
    ref.theta_steps(4, 1.0, 0.5, t0=0.0, pcg_rtol=1e-12)
    y_ref = ref.get_state()[dc.row_begin: dc.row_begin + dc.n]
    pst = ref.pcg_stats()
    ref.close()
    err = np.max(np.abs(y_loc - y_ref) / np.abs(y_ref))
    print(f"rank {rank}/{world}: rows [{dc.row_begin},{dc.row_begin + dc.n}) halo {dc.halo} iters {st} ref {pst} "
          f"max rel err {err:.3e} range [{y_loc.min():.6f},{y_loc.max():.6f}]", flush=True)
    ok = ok and err < 1e-9 and st["error"] == 0
    dc.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("dist_check ok")

if __name__ == "__main__":
    main()
