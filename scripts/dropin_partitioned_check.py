"""`Network.sim(..., method='THETA_PCG')` of the UNMODIFIED reference under torchrun: a reference-built FEPart is
row-partitioned over the ranks (thermca_b200.solver._theta_partitioned -> DistBody.from_csr, reverse Cuthill-McKee,
caller's DOF order kept) and must reproduce (a) the same call on one GPU (partition=False) and (b) a CPU restatement
with a sparse direct solver on the reference's own FFEIO operators.  Needs oracle/_ref (or /root/reference).
  torchrun --nproc-per-node N scripts/dropin_partitioned_check.py [--json out.json]"""
import json
import os
import sys
import tempfile

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
    import scipy.sparse as sp
    import thermca_b200
    from thermca import BoundNode, FEPart, FilmLink, HeatSource, Mesh, Model, Network, Solid
    from thermca_b200.synth import kuhn_cuboid_mesh
    from theta_oracle import theta_step_direct
    steel = Solid(dens=8000.0, spec_heat=500.0, condy=50.0, name="test_steel")

    def build():
        pts, tets, tri_bottom, tri_upper = kuhn_cuboid_mesh(16, 16 * max(world, 2), 5, 0.5, 1.0, 0.05, jitter=0.15)
        mesh = Mesh(pts, [tets, tri_bottom, tri_upper], ["tetra", "triangle", "triangle"], ["cuboid", "bottom", "upper_faces"])
        with Model() as model:
            cuboid = FEPart(mesh, steel, init_temp=20.0, lump_cond_meth=None, name="cuboid")
            env = BoundNode(temp=20.0, name="environment")
            HeatSource(cuboid.surf.bottom, 1000.0)
            FilmLink(cuboid.surf.upper_faces, env, 10.0)
        return Network(model)
    thermca_b200.install(stationary=False, assembly=False, mor_basis=False)
    h, nsteps = 1.0, 6
    net = build()
    res = net.sim([0.0, h * nsteps], method="THETA_PCG", step=h, pcg_rtol=1e-12, progress_bar=False, num_res_skip=2)
    net1 = build()
    res1 = net1.sim([0.0, h * nsteps], method="THETA_PCG", step=h, pcg_rtol=1e-12, progress_bar=False, num_res_skip=2,
                    partition=False)
    d, d1 = res._data, res1._data
    io = net1._ffe_ios[0]
    A = sp.csr_matrix(io.A)
    n = A.shape[0]
    b = np.asarray(io.B @ np.array([1000.0 / io.fe_sys.surf_areas[0], 10.0 * 20.0])).ravel()
    y = np.full(n, 20.0)
    for _ in range(nsteps):
        y = theta_step_direct(A, b - A @ y, y, h, 0.5)
    e_single = float(np.max(np.abs(d.io_temps - d1.io_temps)) / np.max(np.abs(d1.io_temps)))
    e_direct = float(np.max(np.abs(d.io_temps[-1] - y)) / np.max(np.abs(y)))
    e_net = float(np.max(np.abs(d.net_temps - d1.net_temps)) / np.max(np.abs(d1.net_temps)))
    part = net._fe_parts[0]
    t_surf = float(np.asarray(res[part.surf.bottom].temp(h * nsteps)).ravel()[0])
    ok = e_single < 1e-9 and e_direct < 1e-9 and e_net < 1e-9 and list(d.times) == list(d1.times)
    print(f"rank {rank}/{world}: {n} DOF, stored times {list(d.times)}, partitioned vs one GPU {e_single:.2e} "
          f"(network rows {e_net:.2e}), vs direct solver {e_direct:.2e}, bottom surface mean {t_surf:.6f}", flush=True)
    # ---- a general network: nonlinear films, an unknown network node, a stationary node, an Input driven source, two
    # films on one surface, a temperature dependent flux source (oracle/ref_models.py:plate_ffe_var); the FE body is
    # partitioned, the network replicated, surface means completed across the ranks every right-hand side
    import ref_models
    netv, _ = ref_models.plate_ffe_var()
    resv = netv.sim([0.0, 24.0], method="THETA_PCG", step=2.0, pcg_rtol=1e-12, progress_bar=False, num_res_skip=3)
    netv1, _ = ref_models.plate_ffe_var()
    resv1 = netv1.sim([0.0, 24.0], method="THETA_PCG", step=2.0, pcg_rtol=1e-12, progress_bar=False, num_res_skip=3,
                      partition=False)
    dv, dv1 = resv._data, resv1._data
    e_gen_io = float(np.max(np.abs(dv.io_temps - dv1.io_temps)) / np.max(np.abs(dv1.io_temps)))
    e_gen_net = float(np.max(np.abs(dv.net_temps - dv1.net_temps)) / np.max(np.abs(dv1.net_temps)))
    e_gen_cond = float(np.max(np.abs(np.asarray(dv.net_clean_conds) - np.asarray(dv1.net_clean_conds))))
    ok = ok and e_gen_io < 1e-9 and e_gen_net < 1e-9 and e_gen_cond < 1e-6 and list(dv.times) == list(dv1.times)
    print(f"rank {rank}/{world}: general network ({netv._ffe_ios[0].fe_sys.dof} DOF body, {len(dv.times)} stored states): "
          f"partitioned vs one GPU part states {e_gen_io:.2e}, network temperatures {e_gen_net:.2e}, "
          f"conductances {e_gen_cond:.2e}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if rank == 0 and json_path:
        with open(json_path, "w") as f:
            json.dump({"world": world, "dof": n, "steps": nsteps, "max_rel_vs_single_gpu": e_single,
                       "max_rel_vs_direct_solver": e_direct, "max_rel_network_rows": e_net,
                       "general_network_part_states": e_gen_io, "general_network_net_temps": e_gen_net,
                       "general_network_conductances_abs": e_gen_cond, "ok": not bool(flag.item())}, f)
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("dropin_partitioned_check ok")


if __name__ == "__main__":
    main()
