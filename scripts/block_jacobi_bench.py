"""Point Jacobi vs block-Jacobi (32 x 32 slice blocks, FP32-stored inverses) at equal final accuracy:
iterations/step and DOF*steps/s of the theta step for h = 1, 10, 100 s, and the steady-state solve.
  python scripts/block_jacobi_bench.py [dof=1e6] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from thermca_b200.device import DeviceSolver
from thermca_b200.workloads import cuboid_dims, device_cuboid_spec, sell_to_csr_host

dof = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
nx, ny, nz = cuboid_dims(dof)
rows = []
for h in (1.0, 10.0, 100.0):
    rec = {"h": h}
    for name, mode in (("jacobi", 0), ("block_jacobi", 2)):
        dev = DeviceSolver(device_cuboid_spec(nx, ny, nz))
        t_build = 0.0
        if mode == 2:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            dev.set_block_jacobi(h, 0.5)
            t_build = time.perf_counter() - t0
        dev.theta_steps(2, h, 0.5, t0=0.0, pcg_rtol=1e-8, pcg_mode=mode)
        it0 = dev.pcg_stats()["iterations"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 6
        torch.cuda.synchronize()
        e0.record(dev.stream)
        dev.theta_steps(steps, h, 0.5, pcg_rtol=1e-8, pcg_mode=mode, sync=False)
        e1.record(dev.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        its = (dev.pcg_stats()["iterations"] - it0) / steps
        y = dev.get_state()
        rec[name] = {"iters_per_step": its, "ms_per_step": ms, "mdof_steps_per_s": dev.n_state / ms / 1e3,
                     "us_per_iter": 1e3 * ms / its, "precond_build_s": t_build, "t_max": float(y.max())}
        n = dev.n_state
        dev.close()
    rec["temps_agree_rel"] = abs(rec["jacobi"]["t_max"] - rec["block_jacobi"]["t_max"]) / rec["jacobi"]["t_max"]
    rec["speedup_block_over_point"] = rec["jacobi"]["ms_per_step"] / rec["block_jacobi"]["ms_per_step"]
    print(json.dumps(rec), flush=True)
    rows.append(rec)
# steady state K T = q on a smaller body (host CSR path of thermca_b200.stationary)
from thermca_b200.stationary import spd_solve
import scipy.sparse as sp
sx, sy, sz = cuboid_dims(min(dof, 300000))
spec = device_cuboid_spec(sx, sy, sz)
op = spec["ffe"][0]["sell"]
A = sell_to_csr_host(op).tocsr()
K = (sp.diags(op["mass"].cpu().numpy()) @ A).tocsr()
b = np.zeros(K.shape[0]); b[spec["ffe"][0]["B_idx"][0]] = 1.0
st = {"dof": K.shape[0]}
for name in ("jacobi", "block"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, info = spd_solve(K, b, rtol=1e-10, return_info=True, precond=name)
    st[name] = {"iterations": info["iterations"], "relres": info["relres"], "seconds_incl_setup": time.perf_counter() - t0}
print(json.dumps({"steady_state": st}), flush=True)
if len(sys.argv) > 2:
    json.dump({"dof": n, "theta_steps": rows, "steady_state": st}, open(sys.argv[2], "w"), indent=1)
