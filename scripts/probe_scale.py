"""Results at scale on the headline mesh: adaptive RK45 on the 10.35 M-DOF cuboid with the probe ring
(network vector + 4 probe DOFs per accepted step) vs full part states per step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from thermca_b200.device import DeviceSolver
from thermca_b200.workloads import cuboid_dims, device_cuboid_spec

nx, ny, nz = cuboid_dims(float(sys.argv[1]) if len(sys.argv) > 1 else 10.0e6)
spec = device_cuboid_spec(nx, ny, nz)
n = int(spec["ffe"][0]["dof"])
t_end = 40.0
for mode in ("probes", "full"):
    dev = DeviceSolver(spec)
    if mode == "probes":
        dev.set_probes(np.array([0, n // 3, n // 2, n - 1]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = dev.run_rk(0.0, t_end, 1e-3, 1e-6, "RK45")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    att = out["accepted"] + out["rejected"]
    print(f"{mode}: n = {n}, {out['accepted']} accepted / {out['rejected']} rejected attempts in {dt:.3f} s "
          f"({1e3 * dt / att:.2f} ms per attempt), host result arrays {out['io_temps'].nbytes / 1e6:.1f} MB, "
          f"surface means at t_end {out['net_temps'][-1][1:]}", flush=True)
    dev.close()
