"""Classic vs pipelined PCG of the row-partitioned theta step (csrc/dist.cu) on the same mesh, under torchrun (N >= 1):
ms per step, us per iteration, iteration counts, phase profile, and the difference between the two final states.
  torchrun --nproc-per-node N scripts/pipe_bench.py --rows-per-gpu 1300000 [--json PATH] [--variants a,b,...]
A variant is  mode[:shape[:hint[:slack[:grid]]]]  e.g. classic, sr:232, pipe:216:0, classic:::0, classic::::296
(TC_PIPE_SHAPE / TC_PIPE_L2_HINT / TC_DIST_BND_SLACK / TC_DIST_GRID)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=1300000)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--h", type=float, default=1.0)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--variants", default="classic,pipe")
    ap.add_argument("--json", default=None)
    ap.add_argument("--blocks", action="store_true", help="per-block phase profile (TC_DIST_DIAG=3) of rank 0")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from thermca_b200.dist import DistCuboid
    from thermca_b200.workloads import cuboid_dims
    nx, ny, nz = cuboid_dims(args.rows_per_gpu * world)
    results, states = [], {}
    for var in args.variants.split(","):
        parts = var.split(":")
        mode = parts[0]
        for k in ("TC_PIPE_SHAPE", "TC_PIPE_L2_HINT", "TC_DIST_BND_SLACK", "TC_DIST_GRID"):
            os.environ.pop(k, None)
        if len(parts) > 3 and parts[3]:
            os.environ["TC_DIST_BND_SLACK"] = parts[3]
        if len(parts) > 4 and parts[4]:
            os.environ["TC_DIST_GRID"] = parts[4]
        if len(parts) > 1 and parts[1]:
            os.environ["TC_PIPE_SHAPE"] = parts[1]
        if len(parts) > 2 and parts[2]:
            os.environ["TC_PIPE_L2_HINT"] = parts[2]
        if args.blocks:
            os.environ["TC_DIST_DIAG"] = "3"
        dc = DistCuboid(nx, ny, nz)
        assert dc.set_pcg(mode) == mode
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dc.theta_steps(args.warmup, args.h, 0.5, args.rtol, sync=True)
        it0 = dc.stats()["iters_total"]
        dc.profile()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record(dc.stream)
        dc.theta_steps(args.steps, args.h, 0.5, args.rtol)
        e1.record(dc.stream)
        torch.cuda.synchronize(); dist.barrier()
        dc.check()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        st = dc.stats()
        iters = st["iters_total"] - it0
        prof = dc.profile()
        blocks = None
        if args.blocks:
            import ctypes as C
            import numpy as np
            buf = (C.c_uint64 * (10 * 2048))()
            g = int(dc.lib.tc_dist_block_profile(dc._h, buf, 2048))
            a = np.frombuffer(buf, dtype=np.uint64).reshape(2048, 10)[:g].astype(np.float64)
            blocks = {"grid": g, "sm": a[:, 8].astype(int).tolist(),
                      "us_per_iter": (a[:, :8] / 1e3 / max(iters + args.warmup * iters // args.steps, 1)).round(2).tolist()}
        states[var] = dc.get_state().clone()
        r = {"variant": var, "world": world, "dof": dc.n_global, "rows_per_gpu": dc.n, "ms_per_step": float(ms) / args.steps,
             "iters_per_step": iters / args.steps, "us_per_iter": 1e3 * float(ms) / max(iters, 1),
             "mdof_steps_per_s": dc.n_global * args.steps / float(ms) / 1e3, "grid": st["grid"],
             "us_per_iter_by_phase_block0": {k: round(1e3 * v / max(iters, 1), 2) for k, v in prof["block0"].items() if v},
             "us_per_iter_by_phase_last": {k: round(1e3 * v / max(iters, 1), 2) for k, v in prof["last"].items() if v}}
        if blocks is not None:
            r["blocks"] = blocks
        first = next(iter(states.values()))
        d = torch.max(torch.abs(states[var] - first) / torch.abs(first)).reshape(1)
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        r["max_rel_vs_first_variant"] = float(d)
        results.append(r)
        if rank == 0:
            print(json.dumps({k: v for k, v in r.items() if k != "blocks"}), flush=True)
        dc.close()
        del dc
        torch.cuda.empty_cache()
    dist.destroy_process_group()
    if rank == 0 and args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)
    if rank == 0:
        print("pipe_bench ok")


if __name__ == "__main__":
    main()
