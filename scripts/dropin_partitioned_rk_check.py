"""The reference's own adaptive pairs on a row-partitioned body: `Network.sim(..., method='RK45' | 'RK23',
partition=True)` of the UNMODIFIED reference under torchrun (thermca_b200.solver._rk_partitioned) must reproduce the same
call on one GPU — same accepted steps, times, temperatures — on (a) a reference-built FEPart cuboid and (b) the general
network `plate_ffe_var` (nonlinear films, unknown + stationary node, Input, flux source).  World size 1 exercises the same
kernels without neighbours (tests/test_gpu_dist_single.py).  Needs oracle/_ref (or /root/reference).
  torchrun --nproc-per-node N scripts/dropin_partitioned_rk_check.py [--json out.json]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    json_path = os.path.abspath(sys.argv[sys.argv.index("--json") + 1]) if "--json" in sys.argv else None
    os.chdir(tempfile.mkdtemp())                   # the reference's disk_cache writes ./.thermca_cache
    import refenv
    refenv.activate()
    import thermca_b200
    from thermca import BoundNode, FEPart, FilmLink, HeatSource, Mesh, Model, Network, Solid
    from thermca_b200.synth import kuhn_cuboid_mesh
    import ref_models
    steel = Solid(dens=8000.0, spec_heat=500.0, condy=50.0, name="test_steel")

    def cuboid():
        pts, tets, tri_bottom, tri_upper = kuhn_cuboid_mesh(16, 16 * max(world, 2), 5, 0.5, 1.0, 0.05, jitter=0.15)
        mesh = Mesh(pts, [tets, tri_bottom, tri_upper], ["tetra", "triangle", "triangle"], ["cuboid", "bottom", "upper_faces"])
        with Model() as model:
            body = FEPart(mesh, steel, init_temp=20.0, lump_cond_meth=None, name="cuboid")
            env = BoundNode(temp=20.0, name="environment")
            HeatSource(body.surf.bottom, 1000.0)
            FilmLink(body.surf.upper_faces, env, 10.0)
        return Network(model)
    thermca_b200.install(stationary=False, assembly=False, mor_basis=False)
    cases = [("cuboid", cuboid, 3.0, "RK45", 1e-4, 1), ("cuboid", cuboid, 3.0, "RK23", 1e-3, 2),
             ("plate_ffe_var", lambda: ref_models.plate_ffe_var()[0], 60.0, "RK45", 1e-5, 0),
             # the W-methods (stage solves across the GPUs); 'LSODA' requests are served by ROW3w
             ("cuboid", cuboid, 300.0, "LSODA", 1e-4, 3), ("cuboid", cuboid, 100.0, "ROS2_PCG", 1e-3, 3),
             ("plate_ffe_var", lambda: ref_models.plate_ffe_var()[0], 200.0, "ROW3_PCG", 1e-5, 1)]
    report, ok = {}, True
    for name, build, t1, method, rtol, skip in cases:
        kw = dict(method=method, rel_tol=rtol, abs_tol=1e-7, progress_bar=False, num_res_skip=skip)
        if method not in ("RK45", "RK23"):
            kw["pcg_rtol"] = 1e-12
        # plate_ffe_var has a film towards a stationary node: its rank-one Jacobian term is dropped on a partitioned body
        # (csrc/solver.cu:tc_attach_dist), so the W-methods take other (not fewer accurate) steps than on one GPU there
        same_scheme = not (name == "plate_ffe_var" and method not in ("RK45", "RK23"))
        a0 = time.perf_counter()
        d = build().sim([0.0, t1], partition=True, **kw)._data
        a1 = time.perf_counter()
        d1 = build().sim([0.0, t1], **kw)._data
        same_steps = len(d.times) == len(d1.times)
        if same_scheme:
            e_t = float(np.max(np.abs(np.asarray(d.times) - np.asarray(d1.times)) / np.maximum(np.abs(d1.times), 1e-30))) if same_steps else np.inf
            e_io = float(np.max(np.abs(d.io_temps - d1.io_temps)) / np.max(np.abs(d1.io_temps))) if same_steps else np.inf
            e_net = float(np.max(np.abs(d.net_temps - d1.net_temps)) / np.max(np.abs(d1.net_temps))) if same_steps else np.inf
            good = same_steps and e_t < 1e-8 and e_io < 1e-8 and e_net < 1e-8
        else:                                       # different step sequences: compare the end states within the tolerance
            e_t = float(abs(d.times[-1] - d1.times[-1]))
            e_io = float(np.max(np.abs(d.io_temps[-1] - d1.io_temps[-1])) / np.max(np.abs(d1.io_temps[-1])))
            e_net = float(np.max(np.abs(d.net_temps[-1] - d1.net_temps[-1])) / np.max(np.abs(d1.net_temps[-1])))
            good = e_t == 0.0 and e_io < 20 * rtol and e_net < 20 * rtol
        ok = ok and good
        report[f"{name}/{method}"] = {"stored_states": len(d.times), "stored_states_one_gpu": len(d1.times),
                                      "dof": int(d.io_temps.shape[1]), "max_rel_times": e_t, "max_rel_part_states": e_io,
                                      "max_rel_network_temps": e_net, "ok": bool(good), "wall_s_partitioned": a1 - a0}
        print(f"rank {rank}/{world}: {name} {method}: {len(d.times)} stored states (one GPU {len(d1.times)}), times "
              f"{e_t:.2e}, part states {e_io:.2e}, network temperatures {e_net:.2e}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if rank == 0 and json_path:
        with open(json_path, "w") as f:
            json.dump({"world": world, "cases": report, "ok": not bool(flag.item())}, f, indent=1)
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("dropin_partitioned_rk_check ok")


if __name__ == "__main__":
    main()
